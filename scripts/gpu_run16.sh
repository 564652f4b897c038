for t in 16384,2,3 32768,2,3 65536,2,3 16384,2,4 32768,3,3 32768,2,2 16384,3,3; do
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e --tuning $t | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$t', round(d['value']), round(d['roofline']['achieved']), d['ms_per_step'])"
done
