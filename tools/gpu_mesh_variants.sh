# Config 3 (the million-triangle scene) on the default library and on every library under lib/variants/:
# frame time, then the mesh parity tests on each variant.
mkdir -p gpurun_out; log=gpurun_out/mesh_variants.log; : > $log
echo "default" | tee -a $log; timeout 120 python tools/mesh_bench.py --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-260 | tee -a $log
for v in quasi_b200/lib/variants/*.so; do echo $v | tee -a $log; timeout 120 python tools/mesh_bench.py --steps 5 --warmup 2 --plugin $v 2>&1 | tail -1 | cut -c1-260 | tee -a $log; done
echo "default again" | tee -a $log; timeout 120 python tools/mesh_bench.py --steps 5 --warmup 2 2>&1 | tail -1 | cut -c1-260 | tee -a $log
for v in quasi_b200/lib/variants/*.so; do echo "parity $v" | tee -a $log; QUASI_PLUGIN_OVERRIDE=$v timeout 300 python -m pytest tests/test_cuda_mesh.py -m gpu -q 2>&1 | tail -1 | tee -a $log; done
