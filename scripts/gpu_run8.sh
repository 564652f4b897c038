set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r8_pytest.log; cat gpurun_out/r8_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r8_bench.json 2> gpurun_out/r8_bench.err; tail -3 gpurun_out/r8_bench.err; cat gpurun_out/r8_bench.json
for wl in graphene4x4 si333_vacancy mos2_8x8_spinor; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload $wl --kpts-per-step 8 > gpurun_out/r8_bench_$wl.json 2>gpurun_out/r8_bench_$wl.err; cat gpurun_out/r8_bench_$wl.json | cut -c1-1500; done
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r8_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r8_launches.csv $CMD > gpurun_out/r8_ncu1.log 2>&1
$CMD > gpurun_out/r8_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 1 -c 2 -o gpurun_out/r8_k2 $CMD > gpurun_out/r8_ncu2.log 2>&1
tail -3 gpurun_out/r8_ncu2.log
