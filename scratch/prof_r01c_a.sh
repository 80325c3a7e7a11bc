set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_r01c.log 2>&1; tail -3 gpurun_out/pytest_r01c.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01c_bilinear.json 2> gpurun_out/bench_r01c_bilinear.err
for m in nearestneighbor bicubic; do python bench.py --method $m --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r01c_$m.json 2>/dev/null; done
for v in fill short; do python bench.py --variant $v --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_r01c_$v.json 2>/dev/null; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01c_reference.json 2>gpurun_out/bench_r01c_reference.err
PYTHONPATH=. python scratch/vec_time.py > gpurun_out/vector_timings_r01c.txt 2>&1
for f in gpurun_out/bench_r01c_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d['ms_per_step'],3), 'ms', d['value'], 'frac', d.get('roofline',{}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('clocks'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
cat gpurun_out/vector_timings_r01c.txt
