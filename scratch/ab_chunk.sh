for m in bilinear nearestneighbor; do
for c in 64 32 48 96; do
FIMEX_B200_ZCHUNK=$c python bench.py --method $m --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$m chunk $c', round(d['ms_per_step'],3), 'ms', d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
