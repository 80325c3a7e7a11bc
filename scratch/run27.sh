FIMEX_B200_LIB=$PWD/scratch/lib_b3.so timeout 900 python -m pytest tests -m gpu -q -x -k "structured or cached_interpolation or fuzz_scalar" 2>&1 | tail -2
bash scratch/ab.sh b2 b3
bash scratch/ab_nn.sh b2 b3
