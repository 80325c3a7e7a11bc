for rep in 1 2; do
for c in 64 80 96 112 128 160; do
FIMEX_B200_ZCHUNK=$c python bench.py --method bilinear --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('bilinear chunk $c', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
