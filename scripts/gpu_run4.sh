set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r4_pytest.log; cat gpurun_out/r4_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r4_bench.json 2> gpurun_out/r4_bench.err; tail -3 gpurun_out/r4_bench.err; cat gpurun_out/r4_bench.json
CMD="python bench.py --steps 2 --warmup 1 --kpts-per-step 2 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r4_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r4_launches.csv $CMD > gpurun_out/r4_ncu1.log 2>&1
