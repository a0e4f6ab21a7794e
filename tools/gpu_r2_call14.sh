# 1 GPU: config 3 with the L2 hints (persisting window over the nodes; layers streamed with evict-first).
mkdir -p gpurun_out; log=gpurun_out/mesh_l2.log; : > $log
run() { echo "== $*" | tee -a $log; timeout 150 python tools/mesh_bench.py --steps 5 --warmup 2 "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  ', d['ms_per_frame_median'], 'ms', d['mrays_per_s'], 'Mrays/s', d.get('messages'))" 2>&1 | tail -1 | cut -c1-400 | tee -a $log; }
run
run --env QUASI_CUDA_MESH_L2_KEEP=100
run --env QUASI_CUDA_MESH_L2_KEEP=60
run --plugin quasi_b200/lib/variants/libquasi_cuda_stream.so
run --plugin quasi_b200/lib/variants/libquasi_cuda_stream.so --env QUASI_CUDA_MESH_L2_KEEP=100
run --builder lbvh --env QUASI_CUDA_MESH_L2_KEEP=100
run
QUASI_PLUGIN_OVERRIDE=quasi_b200/lib/variants/libquasi_cuda_stream.so QUASI_CUDA_MESH_L2_KEEP=100 timeout 300 python -m pytest tests/test_cuda_mesh.py -m gpu -q 2>&1 | tail -1 | tee -a $log
