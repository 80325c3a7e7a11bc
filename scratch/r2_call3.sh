#!/bin/bash
# round 2, GPU call 3 (2 GPUs): the rewritten bench.py -- N = 1 with extras, N = 2 strong and weak scaling, reference arm
set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -3 gpurun_out/r2_bench_n1.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2_strong.json 2> gpurun_out/r2_bench_n2_strong.err
tail -3 gpurun_out/r2_bench_n2_strong.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --scaling weak > gpurun_out/r2_bench_n2_weak.json 2> gpurun_out/r2_bench_n2_weak.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python - <<'PY'
import json
for f in ('n1','n2_strong','n2_weak','reference'):
    try:
        d=json.loads(open(f'gpurun_out/r2_bench_{f}.json').read().strip().splitlines()[-1])
        print(f, d['value'], d['ms_per_step'], d.get('roofline',{}).get('frac'), d.get('ranks_verified'), d.get('setup'), (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('frac_of_link'))
        if d.get('extra'): print(json.dumps(d['extra'], indent=1)[:6000])
    except Exception as e: print(f, 'ERR', e)
PY
