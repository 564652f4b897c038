# K2 tile size sweep (batch launch): chunk_vecs 0 = the library's own choice
for spec in "graphene4x4 8" "si333_vacancy 8" "mos2_8x8_spinor 4" "sige555 4"; do
  set -- $spec; wl=$1; kps=$2
  for cv in ${CVS:-0 1024 2048 4096 8192 16384}; do
    python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --workload $wl --kpts-per-step $kps --tuning $cv,0,0 2>/dev/null | tail -1 | \
      python -c "import sys,json; j=json.loads(sys.stdin.read()); print('$wl', 'chunk_vecs', $cv, 'value %.0f GB/s' % j['value'], 'step %.1f us' % (j['ms_per_step']*1e3), 'K2 %.1f us' % (j['roofline']['ms_per_launch']*1e3))"
  done
done
