PYTHONPATH=. python scratch/fill_prof.py > gpurun_out/fp_plain.log 2>&1 && PYTHONPATH=. ncu --set full --clock-control none --import-source on -k regex:k_fill2d_skewed -s 1 -c 1 -o gpurun_out/prof_fill2d python scratch/fill_prof.py > gpurun_out/ncu.log 2>&1
echo rc=$?
