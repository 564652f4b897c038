# step rate vs SC k-points per launch on the two small cells (1 GPU): graphene 12..184, Si 3x3x3 25..198 per batch.
set -x
mkdir -p gpurun_out
T=${1:-r02i}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for k in 12 23 46 92 184; do timeout 200 $B --workload graphene4x4 --kpts-per-step $k > gpurun_out/${T}_c1_k$k.json 2>gpurun_out/${T}_c1_k$k.err; done
for k in 25 50 99 198; do timeout 200 $B --workload si333_vacancy --kpts-per-step $k > gpurun_out/${T}_c2_k$k.json 2>gpurun_out/${T}_c2_k$k.err; done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{\"metric")][-1])
    r=d["roofline"]
    print(f.split("/")[-1][5:-5], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "K2 ms", round(r["ms_per_launch"],4), "K2", round(r["achieved"]), r["kernel_ms_per_step"])
PY
