# Parity tests and the headline bench line: the quick check after a kernel change.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
if [ $rc -ne 0 ]; then grep -E "Error|assert|FAIL|^E " gpurun_out/pytest_gpu.log | head -20; exit 1; fi
timeout 200 python bench.py --cpu-budget 8 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json
QUASI_CUDA_SCAN=boxes timeout 100 python bench.py --no-cpu --steps 10 2>/dev/null | cut -c60-200 | sed "s/^/boxes: /"
