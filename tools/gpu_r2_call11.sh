mkdir -p gpurun_out
bash tools/gpu_mesh_variants.sh > /dev/null 2>&1; grep -E "default|variants|ms_per_frame|passed|failed" gpurun_out/mesh_variants.log | cut -c1-150
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --no-cpu > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err; tail -1 gpurun_out/bench_b.json | cut -c1-300; tail -3 gpurun_out/bench_b.err
