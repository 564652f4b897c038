set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/r15_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k3_delta_n|k2b_weights|k1_coset" -s 3 -c 3 -o gpurun_out/r15_small $CMD > gpurun_out/r15_ncu.log 2>&1
tail -3 gpurun_out/r15_ncu.log
