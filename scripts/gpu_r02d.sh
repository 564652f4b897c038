# round-2 fourth capture (8 GPUs): the uplink-sharing probe, bench at N=4 and N=8 as the driver launches it.
set -x
mkdir -p gpurun_out
T=${1:-r02d}
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
nvidia-smi --query-gpu=index,pci.bus_id --format=csv,noheader >> gpurun_out/${T}_topo.txt
nproc >> gpurun_out/${T}_topo.txt
python -m bandup_b200.numa --probe > gpurun_out/${T}_probe.json 2>gpurun_out/${T}_probe.err; cat gpurun_out/${T}_probe.json
for N in 4 8; do
  NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${T}_n$N.json 2> gpurun_out/${T}_n$N.err
  grep -h "nranks" gpurun_out/${T}_n$N.err gpurun_out/${T}_n$N.json | grep -o "comm 0x[0-9a-f]* rank [0-9]* nranks [0-9]*.*" | sort -u | head -20 > gpurun_out/${T}_n${N}_comms.txt
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29549 bench.py --gpus 4 --steps 10 --warmup 3 --no-gpu-probe --gpu-order 0,1,2,3,4,5,6,7 > gpurun_out/${T}_n4_indexorder.json 2> gpurun_out/${T}_n4_indexorder.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_n*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{\"metric")][-1])
    except Exception as e:
        print(f, "NO JSON", e); continue
    r=d.get("roofline") or {}; e=d.get("e2e") or {}
    print(f.split("/")[-1], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "K2", r.get("achieved") and round(r["achieved"]), "e2e", e.get("value"), e.get("h2d_GBps_per_rank"), e.get("gpu_of_rank"), e.get("kpts_per_rank"), e.get("e2e_checked"), e.get("gpu_order_source"), e.get("h2d_groups"))
PY
