set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r01_bilinear.json 2> gpurun_out/bench_r01_bilinear.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_gather_bilinear_staged -s 3 -c 1 --csv --log-file gpurun_out/traffic_r01_full.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_traffic.log 2>&1
echo "traffic rc=$?"
python bench.py --times 2 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/b_small.json 2>gpurun_out/b_small.err && ncu --set full --clock-control none --import-source on -k regex:k_gather_bilinear_staged -s 3 -c 1 -o gpurun_out/prof_bilinear_r01 python bench.py --times 2 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu.log 2>&1
echo "full rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r01_reference.json 2>gpurun_out/bench_r01_reference.err
cat gpurun_out/bench_r01_reference.json | cut -c1-400
