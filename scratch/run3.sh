python scratch/setup_time.py
FIMEX_B200_DIRECT_GATHER=1 python scratch/setup_time.py
python bench.py --method bicubic --steps 5 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('bicubic','ms',round(d['ms_per_step'],3),'values/s',d['value'],'frac',round(d['roofline']['frac'],4), d['clocks'])"
