python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r01b_n2.json 2> gpurun_out/bench_r01b_n2.err; echo "n2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01b_n2.json'))
print('n2', d['n_gpus'], round(d['ms_per_step'],3), d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['clocks'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
