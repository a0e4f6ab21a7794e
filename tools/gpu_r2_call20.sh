# 1 GPU: the trace kernel with the triangle tests spread over the warp's lanes: parity, config 3, thresholds.
mkdir -p gpurun_out; log=gpurun_out/mesh_wave3.log; : > $log
echo skip-parity
run() { echo "== $*" | tee -a $log; timeout 150 python tools/mesh_bench.py --steps 5 --warmup 2 "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  ', d['ms_per_frame_median'], 'ms', d['mrays_per_s'], 'Mrays/s nodes/ray', d['nodes_per_ray'], 'tris/ray', d['triangles_per_ray'])" 2>&1 | tail -1 | cut -c1-300 | tee -a $log; }
run
for v in quasi_b200/lib/variants/libquasi_cuda_w*.so; do [ -f $v ] && run --plugin $v; done
run --builder lbvh
run
