set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_pytest.log; cat gpurun_out/r2_pytest.log
timeout 600 python scripts/tune_k2.py sige555 2>&1 | tail -70
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -3 gpurun_out/r2_bench.err; cat gpurun_out/r2_bench.json
