timeout 900 python -m pytest tests -m gpu -q -x -k "forward or config5 or smoke" 2>&1 | tail -2
grep "5 forward" gpurun_out/config_timings.jsonl | tail -2
