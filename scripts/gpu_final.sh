# The round's capture on a 1-GPU box: pytest -m gpu, smoke(), both bench arms, the other workloads, the ncu
# launch list and the --set full captures of K2 (scalar, spinor) and of the small-cell chain.
#   gpurun -- 'bash scripts/gpu_final.sh r02'   then   python scripts/summarize_ncu.py r02 r02
set -x
mkdir -p gpurun_out
T=${1:-rf}
cp bandup_b200/lib/libbandup_gpu.digest gpurun_out/${T}_lib.digest
python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=15 -rf -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/${T}_pytest.log; tail -6 gpurun_out/${T}_pytest.log
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --min-seconds 0"
$CMD > gpurun_out/${T}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
$CMD > gpurun_out/${T}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 1 -c 2 -o gpurun_out/${T}_k2 $CMD > gpurun_out/${T}_ncu2.log 2>&1
# the K2 capture stamps profiles/k2_traffic.json (on this box's copy of the repo) BEFORE the bench lines are
# taken, so that they carry roofline.traffic of this very build; the same summary is re-made at home
python scripts/summarize_ncu.py ${T} ${T}box > gpurun_out/${T}_summarize.log 2>&1; cp profiles/k2_traffic.json gpurun_out/${T}_k2_traffic.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref.json 2>gpurun_out/${T}_ref.err
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
for wl in graphene4x4 si333_vacancy mos2_8x8_spinor; do timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/${T}_bench_$wl.json 2>gpurun_out/${T}_bench_$wl.err; done
timeout 600 python bench.py --steps 10 --warmup 3 --complex128 > gpurun_out/${T}_bench_c128.json 2>gpurun_out/${T}_bench_c128.err
CMD3="$CMD --workload mos2_8x8_spinor"
$CMD3 > gpurun_out/${T}_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 1 -c 1 -o gpurun_out/${T}_k2_spinor $CMD3 > gpurun_out/${T}_ncu3.log 2>&1
CMD4="$CMD --workload graphene4x4"
$CMD4 > gpurun_out/${T}_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k1_coset|k2_weights|k3_final" -s 3 -c 3 -o gpurun_out/${T}_c1 $CMD4 > gpurun_out/${T}_ncu4.log 2>&1
CMD5="$CMD --workload si333_vacancy"
$CMD5 > gpurun_out/${T}_plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k1_coset|k2_weights|k3_final" -s 3 -c 3 -o gpurun_out/${T}_c2 $CMD5 > gpurun_out/${T}_ncu5.log 2>&1
CMD6="$CMD --complex128"
$CMD6 > gpurun_out/${T}_plain6.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 1 -c 1 -o gpurun_out/${T}_k2_c128 $CMD6 > gpurun_out/${T}_ncu6.log 2>&1
tail -2 gpurun_out/${T}_ncu2.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{\"metric")][-1])
    except Exception as e:
        print(f, "NO JSON", e); continue
    r=d.get("roofline") or {}; e=d.get("e2e") or {}; c=d.get("clocks") or {}; ef=d.get("e2e_file") or {}
    print(f.split("/")[-1], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "K2", r.get("achieved") and round(r["achieved"]), "e2e", e.get("value"), e.get("e2e_checked"), "file", ef.get("value"), ef.get("by_reader"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "mhz", c.get("sm_mhz"))
PY
