mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_cuda_mesh.py -m gpu -x -q > gpurun_out/pytest_mesh.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_mesh.log
tail -3 gpurun_out/pytest_mesh.log
if [ $rc -ne 0 ]; then grep -E "Error|assert|FAIL|^E " gpurun_out/pytest_mesh.log | head -30; fi
timeout 120 python tools/mesh_bench.py --builder lbvh --json gpurun_out/mesh_bench_lbvh.json 2>&1 | tail -1 | cut -c1-600
timeout 120 python tools/mesh_bench.py --json gpurun_out/mesh_bench.json 2>&1 | tail -1 | cut -c1-400
