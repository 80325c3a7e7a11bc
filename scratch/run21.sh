timeout 1500 python -m pytest tests -m gpu -q -x -k "fill2d" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log | cut -c1-500
PYTHONPATH=. timeout 600 python scratch/fill_time.py
