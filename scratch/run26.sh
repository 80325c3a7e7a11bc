timeout 1500 python -m pytest tests/test_gpu_fuzz.py -m gpu -q > gpurun_out/pytest_fuzz.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_fuzz.log | cut -c1-700
