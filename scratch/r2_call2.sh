#!/bin/bash
# round 2, GPU call 2: bulk-store bilinear gather -- parity in both forms, then A/B timing against the STG kernel
set -x
for mode in 1 2; do
  for lz in 8 4; do
    FIMEX_B200_BULK_STORE=$mode FIMEX_B200_BULK_LEVELS=$lz timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py "tests/test_gpu_full_field.py::test_full_field_rotated_pole" tests/test_gpu_full_size.py -m gpu -q -x -k "not bicubic and not kdtree and not fill2d and not forward and not config3 and not config4 and not config5" 2>&1 | tail -5 > gpurun_out/r2_call2_tests_m${mode}_l${lz}.log
    tail -3 gpurun_out/r2_call2_tests_m${mode}_l${lz}.log
  done
done
B="python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu"
FIMEX_B200_BULK_STORE=0 $B > gpurun_out/r2_ab_stg.json 2> gpurun_out/r2_ab.err
for chunk in 64 128; do
for mode in 1 2; do for lz in 8 4; do
  FIMEX_B200_BULK_STORE=$mode FIMEX_B200_BULK_LEVELS=$lz FIMEX_B200_BULK_CHUNK=$chunk $B > gpurun_out/r2_ab_m${mode}_l${lz}_c${chunk}.json 2>> gpurun_out/r2_ab.err
done; done; done
FIMEX_B200_BULK_STORE=0 $B > gpurun_out/r2_ab_stg2.json 2>> gpurun_out/r2_ab.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_ab_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d['ms_per_step'],3), round(d['roofline']['frac'],4), d['gpu_launches'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
    except Exception as e: print(f,'ERR',e)
PY
