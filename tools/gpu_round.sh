# Full GPU pass: tests, headline bench, mesh bench, ncu launch list + full capture of both kernels.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 40 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 120 python tools/mesh_bench.py --json gpurun_out/mesh_bench.json > gpurun_out/mesh_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pathtrace_accumulate -s 70 -c 1 -o gpurun_out/quad_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_quad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mesh_pathtrace -s 20 -c 1 -o gpurun_out/mesh_full -f python tools/mesh_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_mesh.log 2>&1
cat gpurun_out/bench.json; cat gpurun_out/mesh_bench.log
