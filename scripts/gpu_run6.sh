set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r6_pytest.log; cat gpurun_out/r6_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r6_bench.json 2> gpurun_out/r6_bench.err; tail -3 gpurun_out/r6_bench.err; cat gpurun_out/r6_bench.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --per-kpt-launch > gpurun_out/r6_bench_perk.json 2>&1; cat gpurun_out/r6_bench_perk.json | cut -c1-400
for wl in graphene4x4 si333_vacancy mos2_8x8_spinor; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl --kpts-per-step 8 > gpurun_out/r6_bench_$wl.json 2>gpurun_out/r6_bench_$wl.err; cat gpurun_out/r6_bench_$wl.json | cut -c1-1800; done
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r6_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r6_launches.csv $CMD > gpurun_out/r6_ncu1.log 2>&1
$CMD > gpurun_out/r6_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 1 -c 2 -o gpurun_out/r6_k2 $CMD > gpurun_out/r6_ncu2.log 2>&1
tail -3 gpurun_out/r6_ncu2.log
