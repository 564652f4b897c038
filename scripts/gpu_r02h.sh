# K2 ring shape under the power cap (1 GPU): CTAs per SM x stages, C5 and the small cells.
set -x
mkdir -p gpurun_out
T=${1:-r02h}
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for tu in 0,2,3 0,2,4 0,2,5 0,1,6 0,1,8 0,1,10 0,1,12; do timeout 200 $B --workload sige555 --tuning $tu > gpurun_out/${T}_ring_sige555_$tu.json 2>gpurun_out/${T}_ring_sige555_$tu.err; done
for tu in 0,2,3 0,2,4 0,1,8; do for wl in si333_vacancy mos2_8x8_spinor graphene4x4; do timeout 200 $B --workload $wl --tuning $tu > gpurun_out/${T}_ring_${wl}_$tu.json 2>gpurun_out/${T}_ring_${wl}_$tu.err; done; done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{\"metric")][-1])
    except Exception as e:
        print(f, "NO JSON", e); continue
    r=d.get("roofline") or {}; c=d.get("clocks") or {}
    print(f.split("/")[-1][5:-5], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "K2", r.get("achieved") and round(r["achieved"]), "mhz", c.get("sm_mhz"), "W", c.get("power_w_max"))
PY
