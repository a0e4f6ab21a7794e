mkdir -p gpurun_out
rm -f gpurun_out/mesh_variants.log
timeout 240 python -m pytest tests/test_cuda_mesh.py -m gpu -x -q > gpurun_out/pytest_mesh.log 2>&1; rc=$?; echo "pytest exit $rc" >> gpurun_out/pytest_mesh.log
tail -3 gpurun_out/pytest_mesh.log
if [ $rc -ne 0 ]; then exit 1; fi
for v in quasi_b200/lib/variants/*.so; do echo $v; timeout 90 python tools/mesh_bench.py --steps 3 --warmup 1 --plugin $v 2>&1 | tail -1 | cut -c1-260 | tee -a gpurun_out/mesh_variants.log; done
timeout 120 python tools/mesh_bench.py --json gpurun_out/mesh_bench.json > gpurun_out/mesh_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mesh_pathtrace -s 20 -c 1 -o gpurun_out/mesh_full -f python tools/mesh_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_mesh.log 2>&1
cat gpurun_out/mesh_bench.log
