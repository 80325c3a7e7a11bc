PYTHONPATH=. timeout 900 python scratch/setup_bench.py 2>&1 | tee gpurun_out/setup_timings_r01b.txt
