nvidia-smi topo -m 2>&1 | head -14
cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; nproc; lscpu | grep -i "numa\|socket\|model name" | head -6
for f in /sys/bus/pci/devices/*/numa_node; do d=$(dirname $f); if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class)" = "0x030200" ]; then echo "$d numa $(cat $f)"; fi; done | head -10
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r01b_n8b.json 2> gpurun_out/bench_r01b_n8b.err; echo "n8 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r01b_n8b.json'))
print('n8', round(d['ms_per_step'],3), d['value'], 'e2e', d['e2e']['value'], d['e2e']['how'][-60:])
PY
