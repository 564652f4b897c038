set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -30
lscpu | grep -E "NUMA|Socket|Model name|^CPU\(s\)"
for i in 0 1 2 3 4 5 6 7; do b=$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $i | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "gpu $i $b numa $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null)"; done
for flag in "" "--no-numa-bind"; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 $flag > gpurun_out/r12_bench_n8$flag.json 2> gpurun_out/r12_bench_n8$flag.err; tail -2 gpurun_out/r12_bench_n8$flag.err
python - <<PY
import json
txt=open("gpurun_out/r12_bench_n8$flag.json").read()
d=json.loads([l for l in txt.splitlines() if l.startswith("{")][-1])
print("$flag", d["value"], d["e2e"])
PY
done
