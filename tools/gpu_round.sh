# Full GPU pass: tests, headline bench (+ reference arm), config 3 bench, configs 4 and 5, ncu launch list
# and one full capture of each path-tracing kernel. Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 40 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err
timeout 200 python bench.py --workload mesh --steps 5 > gpurun_out/bench_mesh.json 2>> gpurun_out/bench.err
timeout 200 python bench.py --workload mesh --mesh-builder lbvh --steps 5 --no-cpu > gpurun_out/bench_mesh_lbvh.json 2>> gpurun_out/bench.err
bash tools/gpu_configs.sh 1 > gpurun_out/configs.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pathtrace_accumulate -s 70 -c 1 -o gpurun_out/quad_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_quad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mesh_pathtrace -s 20 -c 1 -o gpurun_out/mesh_full -f python tools/mesh_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_mesh.log 2>&1
cut -c1-300 gpurun_out/bench.json; cut -c1-300 gpurun_out/bench_mesh.json; cat gpurun_out/configs.log | cut -c1-250
