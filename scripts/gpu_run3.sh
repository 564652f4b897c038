set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r3_pytest.log; cat gpurun_out/r3_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r3_bench.json 2> gpurun_out/r3_bench.err; tail -3 gpurun_out/r3_bench.err; cat gpurun_out/r3_bench.json
CMD="python bench.py --steps 2 --warmup 1 --kpts-per-step 2 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r3_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r3_launches.csv $CMD > gpurun_out/r3_ncu1.log 2>&1
$CMD > gpurun_out/r3_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 2 -c 2 -o gpurun_out/r3_k2 $CMD > gpurun_out/r3_ncu2.log 2>&1
tail -3 gpurun_out/r3_ncu2.log
