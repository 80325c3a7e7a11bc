timeout 900 python -m pytest tests -m gpu -q -x -k "pinned or cpp" 2>&1 | tail -2
PYTHONPATH=. python scratch/pageable.py 2>&1 | tail -6
