# round-2 fifth capture (1 GPU): tests after the K1/K3 changes, tile-size sweep on the small cells, complex128 bench.
set -x
mkdir -p gpurun_out
T=${1:-r02e}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=15 -rf -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/${T}_pytest.log; tail -8 gpurun_out/${T}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for wl in sige555 graphene4x4 si333_vacancy mos2_8x8_spinor; do timeout 200 $B --workload $wl > gpurun_out/${T}_n1_$wl.json 2>gpurun_out/${T}_n1_$wl.err; done
for cv in 1024 2048 4096 8192 16384; do timeout 200 $B --workload graphene4x4 --tuning $cv,0,0 > gpurun_out/${T}_tile_graphene4x4_$cv.json 2>gpurun_out/${T}_tile_graphene4x4_$cv.err; done
for cv in 2048 4096 8192; do timeout 200 $B --workload si333_vacancy --tuning $cv,0,0 > gpurun_out/${T}_tile_si333_vacancy_$cv.json 2>gpurun_out/${T}_tile_si333_vacancy_$cv.err; done
for st in 4 5; do timeout 200 $B --workload graphene4x4 --tuning 0,0,$st > gpurun_out/${T}_stages_graphene4x4_$st.json 2>gpurun_out/${T}_stages_graphene4x4_$st.err; done
timeout 300 python bench.py --steps 10 --warmup 3 --complex128 > gpurun_out/${T}_c128.json 2>gpurun_out/${T}_c128.err
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{\"metric")][-1])
    except Exception as e:
        print(f, "NO JSON", e); continue
    r=d.get("roofline") or {}; e=d.get("e2e") or {}; c=d.get("clocks") or {}
    print(f.split("/")[-1][5:-5], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "K2", r.get("achieved") and round(r["achieved"]), "e2e", e.get("value"), e.get("e2e_checked"), "mhz", c.get("sm_mhz"), r.get("kernel_ms_per_step"))
PY
