timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-400
bash scratch/ab.sh base new
PYTHONPATH=. python scratch/vec_time.py
PYTHONPATH=. FIMEX_B200_DIRECT_GATHER=1 python scratch/vec_time.py
