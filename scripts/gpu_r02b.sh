# round-2 second capture (1 GPU): tests, A/B matrix of K2 accumulator layout x dependent-launch bits.
set -x
mkdir -p gpurun_out
T=${1:-r02b}
timeout 1500 python -m pytest tests -m gpu -q --maxfail=15 -rf 2>&1 | tail -60 > gpurun_out/${T}_pytest.log; tail -8 gpurun_out/${T}_pytest.log
B="python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e"
for wl in sige555 graphene4x4; do for v in 0 1 2; do for p in 0 3 1 2; do
  timeout 200 $B --workload $wl --k2-variant $v --dependent-launch $p > gpurun_out/${T}_ab_${wl}_v${v}_p${p}.json 2>gpurun_out/${T}_ab_${wl}_v${v}_p${p}.err
done; done; done
for wl in si333_vacancy mos2_8x8_spinor; do for v in 0 1 2; do for p in 0 3; do
  timeout 200 $B --workload $wl --k2-variant $v --dependent-launch $p > gpurun_out/${T}_ab_${wl}_v${v}_p${p}.json 2>gpurun_out/${T}_ab_${wl}_v${v}_p${p}.err
done; done; done
for v in 1 2; do timeout 200 $B --workload mos2_8x8_spinor --kpt-offset 4 --k2-variant $v --dependent-launch 0 > gpurun_out/${T}_ab_mos2off4_v${v}_p0.json 2>gpurun_out/${T}_ab_mos2off4_v${v}_p0.err; done
for v in 1 2; do timeout 200 $B --workload mos2_8x8_spinor --kpt-offset 8 --k2-variant $v --dependent-launch 0 > gpurun_out/${T}_ab_mos2off8_v${v}_p0.json 2>gpurun_out/${T}_ab_mos2off8_v${v}_p0.err; done
CMD4="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --min-seconds 0 --workload si333_vacancy"
$CMD4 > gpurun_out/${T}_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k3_final" -s 2 -c 1 -o gpurun_out/${T}_c2_k3 $CMD4 > gpurun_out/${T}_ncu4.log 2>&1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f, "NO JSON", e); continue
    r=d.get("roofline") or {}
    c=d.get("clocks") or {}
    print(f.split("/")[-1][5:-5], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "K2", r.get("achieved") and round(r["achieved"]), "mhz", c.get("sm_mhz"), "W", c.get("power_w_max"), c.get("reasons"))
PY
