PYTHONPATH=. timeout 600 python scratch/fill_time.py
PYTHONPATH=. FIMEX_B200_FILL_SIMPLE=1 timeout 600 python scratch/fill_time.py
