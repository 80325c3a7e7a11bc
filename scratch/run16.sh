timeout 1500 python -m pytest tests -m gpu -q -x -k "bilinear or structured or cached_interpolation or typed or edge or seam or full_size_bilinear or fused or vector or concurrent or smoke" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-400
bash scratch/ab.sh base new
