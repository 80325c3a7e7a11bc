# usage: ab_all.sh "<variant flags>" lib1 lib2 ...
flags="$1"; shift
for rep in 1 2; do
for v in "$@"; do
FIMEX_B200_LIB=$PWD/scratch/lib_$v.so python bench.py $flags --steps 10 --warmup 3 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('$v [$flags]', round(d['ms_per_step'],3), 'ms  frac', round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])"
done; done
