# The 4-wide form of the mesh kernel on a GPU: parity against the oracle, then config 3 timed
# on both trees. Every step under its own timeout.
mkdir -p gpurun_out
timeout -k 3 ${WIDE_PYTEST_S:-40} python -m pytest tests/test_cuda_mesh.py -k wide -x -q > gpurun_out/wide_pytest.log 2>&1
echo "pytest exit $?" >> gpurun_out/wide_pytest.log
tail -4 gpurun_out/wide_pytest.log
timeout -k 3 ${WIDE_BENCH_S:-25} python tools/mesh_bench.py --steps 3 --warmup 1 --wide 2 --json gpurun_out/mesh_bench_wide.json > gpurun_out/mesh_bench_wide.log 2>&1
echo "bench exit $?" >> gpurun_out/mesh_bench_wide.log
tail -2 gpurun_out/mesh_bench_wide.log | cut -c1-600
