set -x
mkdir -p gpurun_out
T=${1:-rf}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_pytest.log
python __graft_entry__.py --smoke 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref.json 2>gpurun_out/${T}_ref.err
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
for wl in graphene4x4 si333_vacancy mos2_8x8_spinor; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/${T}_bench_$wl.json 2>gpurun_out/${T}_bench_$wl.err; done
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/${T}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
$CMD > gpurun_out/${T}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 1 -c 2 -o gpurun_out/${T}_k2 $CMD > gpurun_out/${T}_ncu2.log 2>&1
tail -2 gpurun_out/${T}_ncu2.log
python - <<PY
import json
for f in ["${T}_ref","${T}_bench","${T}_bench_graphene4x4","${T}_bench_si333_vacancy","${T}_bench_mos2_8x8_spinor"]:
    d=json.loads([l for l in open("gpurun_out/"+f+".json").read().splitlines() if l.startswith("{")][-1])
    print(f, round(d["value"],1), d.get("roofline",{}).get("achieved"), (d.get("e2e") or {}).get("value"), (d.get("cpu_baseline") or {}).get("value"))
PY
