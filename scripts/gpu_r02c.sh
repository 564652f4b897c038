# round-2 third capture (2 GPUs): the whole GPU suite incl. tests/test_multigpu.py, bench at N=1 and N=2.
set -x
mkdir -p gpurun_out
T=${1:-r02c}
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --maxfail=15 -rf -rs 2>&1 | tail -60 > gpurun_out/${T}_pytest.log; tail -8 gpurun_out/${T}_pytest.log
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -v 2>&1 | tail -30 > gpurun_out/${T}_pytest_multigpu.log; tail -12 gpurun_out/${T}_pytest_multigpu.log
python -m bandup_b200.numa --probe > gpurun_out/${T}_probe.json 2>gpurun_out/${T}_probe.err; cat gpurun_out/${T}_probe.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e-file > gpurun_out/${T}_n1.json 2> gpurun_out/${T}_n1.err
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/${T}_n2.json 2> gpurun_out/${T}_n2.err
grep -i "nranks\|Init COMPLETE" gpurun_out/${T}_n2.err gpurun_out/${T}_n2.json | head -8
for wl in graphene4x4 si333_vacancy; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload $wl > gpurun_out/${T}_n1_$wl.json 2>gpurun_out/${T}_n1_$wl.err; done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_n*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f, "NO JSON", e); continue
    r=d.get("roofline") or {}; e=d.get("e2e") or {}
    print(f.split("/")[-1], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "K2", r.get("achieved") and round(r["achieved"]), "e2e", e.get("value"), e.get("h2d_GBps_per_rank"), e.get("gpu_of_rank"), e.get("e2e_checked"), e.get("gpu_order_source"), r.get("kernel_ms_per_step"))
PY
