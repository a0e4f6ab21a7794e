# 1 GPU: the mesh scene's readback mirror by copy engine (first-hit layers under the later waves, beauty after the last
# launch): the GPU suite, then config 3 device-timed and end to end, with the mirror and without (QUASI_CUDA_MIRROR=0).
mkdir -p gpurun_out; log=gpurun_out/mesh_mirror.log; : > $log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee -a $log
for m in 1 0; do
  echo "== bench.py --workload mesh, QUASI_CUDA_MIRROR=$m" | tee -a $log
  QUASI_CUDA_MIRROR=$m timeout 300 python bench.py --workload mesh --steps 5 --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'], '| launches', d['gpu_launches'])" | tee -a $log
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a $log
