# round-2 sixth capture (8 GPUs): greedy host->device order probe, bench at N=4 (and N=2) with it.
set -x
mkdir -p gpurun_out
T=${1:-r02f}
python -m bandup_b200.numa --probe > gpurun_out/${T}_probe.json 2>gpurun_out/${T}_probe.err; cat gpurun_out/${T}_probe.json
for N in 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_n$N.json 2> gpurun_out/${T}_n$N.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_n*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{\"metric")][-1])
    except Exception as e:
        print(f, "NO JSON", e); continue
    r=d.get("roofline") or {}; e=d.get("e2e") or {}
    print(f.split("/")[-1], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "K2", r.get("achieved") and round(r["achieved"]), "e2e", e.get("value"), e.get("h2d_GBps_per_rank"), e.get("gpu_of_rank"), e.get("kpts_per_rank"), e.get("e2e_checked"), e.get("gpu_order_source"), e.get("h2d_probe_prefix_total_GBps"))
PY
