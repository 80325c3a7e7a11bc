set -x
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e"
S="--times 2 --steps 2 --warmup 3 --no-e2e --no-cpu"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01b_bilinear.json 2> gpurun_out/bench_r01b_bilinear.err
$B > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01b.csv $B > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_gather_bilinear_staged -s 3 -c 1 --csv --log-file gpurun_out/traffic_r01b_bilinear.csv $B > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"
for m in bilinear nearestneighbor; do
python bench.py --method $m $S > gpurun_out/b_small.json 2>gpurun_out/b_small.err && ncu --set full --clock-control none --import-source on -k regex:k_gather_bilinear_staged -s 3 -c 1 -o gpurun_out/prof_r01b_$m python bench.py --method $m $S > gpurun_out/ncu.log 2>&1
echo "full $m rc=$?"
done
python bench.py --method bicubic $S > gpurun_out/b_small.json 2>gpurun_out/b_small.err && ncu --set full --clock-control none --import-source on -k regex:k_gather_bicubic_staged -s 3 -c 1 -o gpurun_out/prof_r01b_bicubic python bench.py --method bicubic $S > gpurun_out/ncu.log 2>&1
echo "full bicubic rc=$?"
python bench.py --variant fill $S > gpurun_out/b_small.json 2>gpurun_out/b_small.err && ncu --set full --clock-control none --import-source on -k regex:k_gather_bilinear_staged -s 3 -c 1 -o gpurun_out/prof_r01b_fill python bench.py --variant fill $S > gpurun_out/ncu.log 2>&1
echo "full fill rc=$?"
for m in nearestneighbor bicubic; do python bench.py --method $m --steps 10 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r01b_$m.json 2>/dev/null; done
for v in fill short; do python bench.py --variant $v --steps 10 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_r01b_$v.json 2>/dev/null; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01b_reference.json 2>gpurun_out/bench_r01b_reference.err
timeout 900 python -m pytest tests/test_gpu_full_size.py -m gpu -q > gpurun_out/pytest_full.log 2>&1; tail -2 gpurun_out/pytest_full.log
PYTHONPATH=. python scratch/vec_time.py > gpurun_out/vector_timings_r01b.txt 2>&1
for f in gpurun_out/bench_r01b_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1].split('/')[-1], round(d['ms_per_step'],3), 'ms', d['value'], 'frac', d.get('roofline',{}).get('frac'), 'e2e', (d.get('e2e') or {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
