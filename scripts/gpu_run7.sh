set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/r7_pytest.log; cat gpurun_out/r7_pytest.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r7_bench.json 2> gpurun_out/r7_bench.err; tail -3 gpurun_out/r7_bench.err; cat gpurun_out/r7_bench.json
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --per-kpt-launch > gpurun_out/r7_bench_perk.json 2>&1; cat gpurun_out/r7_bench_perk.json | cut -c1-400
timeout 600 python scripts/tune_k2.py sige555 2>&1 | tail -70
