python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r01b_n8.json 2> gpurun_out/bench_r01b_n8.err; echo "n8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01b_n8.json'))
print('n8', d['n_gpus'], round(d['ms_per_step'],3), d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 2>/dev/null | cut -c1-200
