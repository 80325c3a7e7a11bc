S="--variant short --times 2 --steps 2 --warmup 3 --no-e2e --no-cpu"
python bench.py $S > gpurun_out/b_small.json 2>gpurun_out/b_small.err && ncu --set full --clock-control none --import-source on -k regex:k_gather_bilinear_staged -s 3 -c 1 -o gpurun_out/prof_r01b_short python bench.py $S > gpurun_out/ncu.log 2>&1
echo "rc=$?"
