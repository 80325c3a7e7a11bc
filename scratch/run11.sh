python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r01b_n2.log 2>&1; echo "n2 rc=$?"
tail -c 3000 gpurun_out/bench_r01b_n2.log
