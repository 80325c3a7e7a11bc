timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-400
cat gpurun_out/config_timings.jsonl | grep bicubic
python bench.py --method bicubic --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('bicubic','ms',round(d['ms_per_step'],3),'values/s',d['value'],'frac',round(d['roofline']['frac'],4), d['clocks'])"
