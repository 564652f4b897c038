set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r1_pytest.log; cat gpurun_out/r1_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -5
python bench.py --steps 5 --warmup 3 > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; tail -3 gpurun_out/r1_bench.err; cat gpurun_out/r1_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1_ref.json 2>&1; cat gpurun_out/r1_ref.json
CMD="python bench.py --steps 2 --warmup 1 --kpts-per-step 2 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r1_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r1_launches.csv $CMD > gpurun_out/r1_ncu1.log 2>&1
$CMD > gpurun_out/r1_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 2 -c 2 -o gpurun_out/r1_k2 $CMD > gpurun_out/r1_ncu2.log 2>&1
tail -3 gpurun_out/r1_ncu2.log
