# round-2 first capture on a 1-GPU box: tests, smoke, both bench arms, A/B of the K2 variants and of
# the dependent-launch chain, ncu launch list + full capture of K2.   gpurun -- bash scripts/gpu_r02a.sh [tag]
set -x
mkdir -p gpurun_out
T=${1:-r02a}
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv,noheader > gpurun_out/${T}_topo.txt 2>&1
nvidia-smi topo -m >> gpurun_out/${T}_topo.txt 2>&1
nproc >> gpurun_out/${T}_topo.txt; free -g | head -2 >> gpurun_out/${T}_topo.txt
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ]; then echo "$(basename $d) class=$(cat $d/class) numa=$(cat $d/numa_node 2>/dev/null) -> $(readlink -f $d)"; fi; done >> gpurun_out/${T}_topo.txt 2>&1
python __graft_entry__.py --smoke > gpurun_out/${T}_smoke.log 2>&1; tail -3 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -q --maxfail=15 -rf 2>&1 | tail -40 > gpurun_out/${T}_pytest.log; tail -15 gpurun_out/${T}_pytest.log
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_ref.json 2>gpurun_out/${T}_ref.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
for wl in graphene4x4 si333_vacancy mos2_8x8_spinor; do timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload $wl > gpurun_out/${T}_bench_$wl.json 2>gpurun_out/${T}_bench_$wl.err; done
# A/B: accumulator layout of K2 and the dependent-launch chain, device-resident numbers only
for wl in sige555 graphene4x4 si333_vacancy mos2_8x8_spinor; do
  for v in 1 2; do timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload $wl --k2-variant $v > gpurun_out/${T}_ab_${wl}_v$v.json 2>gpurun_out/${T}_ab_${wl}_v$v.err; done
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --workload $wl --no-dependent-launch > gpurun_out/${T}_ab_${wl}_nopdl.json 2>gpurun_out/${T}_ab_${wl}_nopdl.err
done
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --min-seconds 0"
$CMD > gpurun_out/${T}_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches.csv $CMD > gpurun_out/${T}_ncu1.log 2>&1
$CMD > gpurun_out/${T}_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 1 -c 2 -o gpurun_out/${T}_k2 $CMD > gpurun_out/${T}_ncu2.log 2>&1
CMD3="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --min-seconds 0 --workload mos2_8x8_spinor"
$CMD3 > gpurun_out/${T}_plain3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k2_weights_reduce -s 1 -c 1 -o gpurun_out/${T}_k2_spinor $CMD3 > gpurun_out/${T}_ncu3.log 2>&1
CMD4="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --min-seconds 0 --workload graphene4x4"
$CMD4 > gpurun_out/${T}_plain4.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k1_coset|k2_weights|k3_final" -s 3 -c 3 -o gpurun_out/${T}_c1 $CMD4 > gpurun_out/${T}_ncu4.log 2>&1
tail -2 gpurun_out/${T}_ncu2.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    try:
        d=json.loads([l for l in open(f).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(f, "NO JSON", e); continue
    r=d.get("roofline") or {}
    print(f.split("/")[-1], "value", round(d["value"],1), "ms/step", round(d["ms_per_step"],4), "steps", d["steps"], "K2", r.get("achieved") and round(r["achieved"]), "e2e", (d.get("e2e") or {}).get("value"), "file", (d.get("e2e_file") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"), "kern", r.get("kernel_ms_per_step"))
PY
