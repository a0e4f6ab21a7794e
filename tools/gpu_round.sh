mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 600 python tools/mesh_bench.py --json gpurun_out/mesh_bench.json > gpurun_out/mesh_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mesh_pathtrace -s 20 -c 1 -o gpurun_out/mesh_full -f python tools/mesh_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_mesh.log 2>&1
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.json; cat gpurun_out/mesh_bench.log
