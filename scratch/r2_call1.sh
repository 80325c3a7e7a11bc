#!/bin/bash
# round 2, GPU call 1: the new parity tests + store / shared-load microbenchmarks
set -x
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv
timeout 1500 python -m pytest tests/test_proj_known_answers.py tests/test_gpu_full_field.py "tests/test_gpu_interpolator_cases.py::test_config1_hirlam12_real_file" "tests/test_gpu_parity.py::test_mifi_interpolate_f_emep" -m gpu -q -x -s 2>&1 | tail -40 > gpurun_out/r2_call1_tests.log
cat gpurun_out/r2_call1_tests.log | tail -15
timeout 300 ./scratch/ubench/tma_store_bw 2000 > gpurun_out/r2_tma_store_2000.txt 2>&1
timeout 300 ./scratch/ubench/tma_store_bw 2048 > gpurun_out/r2_tma_store_2048.txt 2>&1
cat gpurun_out/r2_tma_store_2000.txt
