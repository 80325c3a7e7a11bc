PYTHONPATH=. FIMEX_B200_TRACE=1 python scratch/e2e_probe.py 2>&1 | tail -3
timeout 1500 python -m pytest tests -m gpu -q -x -k "host or concurrent or typed or cached_interpolation or fused or cpp or edge" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-400
