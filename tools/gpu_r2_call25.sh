# 1 GPU: compute-sanitizer over every kernel (incl. the wavefront mesh kernels), then the GPU suite, smoke() and the
# default bench + reference arm on the final binaries.
mkdir -p gpurun_out
bash tools/gpu_sanitize.sh 2>&1 | tail -6
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json | cut -c1-300; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_reference.json | cut -c1-300
