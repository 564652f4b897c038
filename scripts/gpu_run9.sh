set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -30 > gpurun_out/r9_pytest.log; cat gpurun_out/r9_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -3
