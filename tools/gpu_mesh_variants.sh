mkdir -p gpurun_out
for v in quasi_b200/lib/variants/*.so; do echo $v; QUASI_PLUGIN_OVERRIDE=$v timeout 300 python -m pytest tests/test_cuda_mesh.py -m gpu -x -q 2>&1 | tail -2; timeout 90 python tools/mesh_bench.py --steps 5 --warmup 2 --plugin $v 2>&1 | tail -1 | cut -c1-260; done
