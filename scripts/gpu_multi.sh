# Multi-GPU capture: bash scripts/gpu_multi.sh <tag> <N> [pytest]   (run under gpurun --gpus N)
set -x
mkdir -p gpurun_out
T=${1:-r02}; N=${2:-2}
if [ "${3:-}" = pytest ]; then
  timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q -rf 2>&1 | tail -15 > gpurun_out/${T}_pytest_multigpu_${N}gpu.log; tail -3 gpurun_out/${T}_pytest_multigpu_${N}gpu.log
fi
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
  bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_n$N.json 2> gpurun_out/${T}_n$N.err
tail -3 gpurun_out/${T}_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/${T}_n$N.json") if l.startswith("{")][-1]); e=d["e2e"]
print("N", d["n_gpus"], "value", round(d["value"]), "ms/step", round(d["ms_per_step"],4), "e2e", round(e["value"],1), e["kpts_per_rank"], e["h2d_GBps_per_rank"], e["gpu_of_rank"], e["e2e_checked"], d["config"]["parallelism"][:120], d["clocks"])
PY
