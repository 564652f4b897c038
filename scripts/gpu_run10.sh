set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
free -g | head -2
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r10_bench_n$N.json 2> gpurun_out/r10_bench_n$N.err; tail -3 gpurun_out/r10_bench_n$N.err; cat gpurun_out/r10_bench_n$N.json | cut -c1-700
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > gpurun_out/r10_ref_n8.json 2> gpurun_out/r10_ref_n8.err; cat gpurun_out/r10_ref_n8.json | cut -c1-300
