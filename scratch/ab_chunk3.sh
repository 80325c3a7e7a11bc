for v in "" "--variant fill" "--variant short"; do
for c in 64 128; do
FIMEX_B200_ZCHUNK=$c python bench.py $v --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('[$v] chunk $c', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('default', round(d['ms_per_step'],3), 'ms', round(d['roofline']['frac'],4))"
python -m pytest tests/test_gpu_parity.py tests/test_gpu_full_size.py -m gpu -x -q -k "not exhaustive" 2>&1 | tail -2
