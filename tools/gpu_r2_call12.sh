# 1 GPU: the GPU suite and the default bench on the build with the fused-reduce instantiations (the default kernels
# changed only in how a finished tile finds its slice), then compute-sanitizer memcheck over a small batched run.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 400 python bench.py --no-cpu --no-extra > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -1 gpurun_out/bench_c.json | cut -c1-400; tail -3 gpurun_out/bench_c.err
