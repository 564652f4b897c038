set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r13_pytest.log; cat gpurun_out/r13_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r13_bench.json 2> gpurun_out/r13_bench.err; tail -3 gpurun_out/r13_bench.err; cat gpurun_out/r13_bench.json | cut -c1-1600
for wl in graphene4x4 si333_vacancy mos2_8x8_spinor; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl --kpts-per-step 8 > gpurun_out/r13_bench_$wl.json 2>gpurun_out/r13_bench_$wl.err; cat gpurun_out/r13_bench_$wl.json | cut -c1-1400; done
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r13_ref.json 2>&1; cat gpurun_out/r13_ref.json | cut -c1-400
