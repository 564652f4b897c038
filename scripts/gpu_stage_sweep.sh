# Pageable upload rate vs chunk size / store kind / threads (scripts/probe_pageable.py, 1.2 GB block)
mkdir -p gpurun_out
T=${1:-sw}
for cfg in "16384 1 8" "16384 0 8" "4096 1 8" "4096 0 8" "2048 0 8" "1024 0 8" "1024 1 8" "512 0 8" "2048 0 4" "2048 0 12" "1024 0 4" "16384 1 8"; do
  set -- $cfg
  echo "chunk_kb $1 nt $2 threads $3: $(BANDUP_GPU_STAGE_CHUNK_KB=$1 BANDUP_GPU_STAGE_NT=$2 BANDUP_GPU_STAGE_THREADS=$3 timeout 120 python scripts/probe_pageable.py 2>&1 | grep -E 'pageable' | sed 's/.*ring): //')"
done > gpurun_out/${T}_sweep.log 2>&1
cat gpurun_out/${T}_sweep.log
