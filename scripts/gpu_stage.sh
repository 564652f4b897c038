# The staging pipeline on a GPU box: the GPU suite, the pageable-vs-page-locked probe at several
# thread counts (twice each), then one bench line (e2e.pageable_host_buffers, e2e_file).
set -x
mkdir -p gpurun_out
T=${1:-st}
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -4 > gpurun_out/${T}_pytest.log; tail -2 gpurun_out/${T}_pytest.log
for t in 4 8 12 16 8 12; do echo threads $t; BANDUP_GPU_STAGE_THREADS=$t timeout 120 python scripts/probe_pageable.py 2>&1 | grep -E "pageable|pinned_alloc"; done > gpurun_out/${T}_threads.log 2>&1
cat gpurun_out/${T}_threads.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -3 gpurun_out/${T}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${T}_bench.json").read().strip().splitlines()[-1]); e=d["e2e"]
print("value", d["value"], "e2e", e["value"], "pageable", e["pageable_host_buffers"], "file", d["e2e_file"]["value"], d["e2e_file"].get("by_reader"), e["pinned_block_probe"]["candidates_GBps"])
PY
