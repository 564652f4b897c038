# Bench lines only (both arms, every workload) with the traffic stamp of the current build, then a
# time-bounded fuzz run:  gpurun -- 'bash scripts/gpu_lines.sh <tag> [fuzz seconds] [fuzz seed]'
set -x
mkdir -p gpurun_out
T=${1:-rl}; FS=${2:-0}; SEED=${3:-7}
cp bandup_b200/lib/libbandup_gpu.digest gpurun_out/${T}_lib.digest
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref.json 2>gpurun_out/${T}_ref.err
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
for wl in graphene4x4 si333_vacancy mos2_8x8_spinor; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/${T}_bench_$wl.json 2>gpurun_out/${T}_bench_$wl.err; done
timeout 600 python bench.py --steps 10 --warmup 3 --complex128 > gpurun_out/${T}_bench_c128.json 2>gpurun_out/${T}_bench_c128.err
if [ "$FS" != 0 ]; then timeout $((FS + 120)) python scripts/fuzz_gpu.py --seconds $FS --seed $SEED > gpurun_out/${T}_fuzz.log 2>&1; tail -3 gpurun_out/${T}_fuzz.log; fi
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{\"metric")][-1])
    except Exception as e:
        print(f, "NO JSON", e); continue
    r=d.get("roofline") or {}; e=d.get("e2e") or {}; c=d.get("clocks") or {}; ef=d.get("e2e_file") or {}
    print(f.split("/")[-1], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "K2", r.get("achieved") and round(r["achieved"]), "traffic", r.get("traffic"), "e2e", e.get("value"), e.get("e2e_checked"), "file", ef.get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "mhz", c.get("sm_mhz"))
PY
