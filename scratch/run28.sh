timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-400
PYTHONPATH=. python scratch/small_calls.py 2>&1 | tail -6
PYTHONPATH=. python scratch/e2e_probe.py 2>&1 | tail -1
