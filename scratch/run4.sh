timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log | cut -c1-700
PYTHONPATH=. python scratch/setup_time.py
PYTHONPATH=. FIMEX_B200_DIRECT_GATHER=1 python scratch/setup_time.py
for m in bilinear nearestneighbor; do python bench.py --method $m --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$m','ms',round(d['ms_per_step'],3),'values/s',d['value'],'frac',round(d['roofline']['frac'],4), d['clocks'])"; done
