# throughput of the ensemble kernel against samples per SM (latency-bound vs throughput-bound)
for s in 1036 2072 4144 8288 16576; do
python bench.py --samples $s --steps 3 --warmup 2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print($s, 'samples', round(d['ms_per_step'],3), 'ms', '%.3e'%d['value'], 'units/s', d['roofline']['system']['threads'], 'thr/CTA', d['roofline']['system']['ctas'], 'CTAs')"
done
