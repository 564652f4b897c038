set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 300 python -m pytest tests -m gpu -x -q -k "sharded" 2>&1 | tail -5
python bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r5_ref.json 2>gpurun_out/r5_ref.err; cat gpurun_out/r5_ref.json | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r5_bench_n2.json 2> gpurun_out/r5_bench_n2.err; tail -5 gpurun_out/r5_bench_n2.err; cat gpurun_out/r5_bench_n2.json
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r5_bench_n1.json 2> gpurun_out/r5_bench_n1.err; cat gpurun_out/r5_bench_n1.json
