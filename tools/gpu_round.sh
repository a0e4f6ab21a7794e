# Full single-GPU pass of a round: the GPU suite, smoke(), the headline bench exactly as the driver runs it (with its
# C++ host, mesh and CPU legs) and its reference arm, the other workloads, the ncu launch list of the bench.
# Everything lands in gpurun_out/.   gpurun --timeout 1200 -- 'bash tools/gpu_round.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json | cut -c1-600
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 --cpu-budget 40 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_reference.json | cut -c1-300
timeout 300 python bench.py --workload mesh --steps 5 > gpurun_out/bench_mesh.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_mesh.json | cut -c1-300
timeout 300 python bench.py --workload mesh --mesh-builder lbvh --steps 5 --no-cpu > gpurun_out/bench_mesh_lbvh.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_mesh_lbvh.json | cut -c1-200
timeout 300 python bench.py --workload demo4k --steps 2 --no-cpu --no-extra > gpurun_out/bench_4k.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_4k.json | cut -c1-300
bash tools/gpu_configs.sh 1 > gpurun_out/configs.log 2>&1; cut -c1-250 gpurun_out/configs.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-parity > gpurun_out/ncu_launches.log 2>&1
QUASI_CUDA_BATCH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pathtrace_accumulate -s 70 -c 1 -o gpurun_out/r2_final_single_full -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --no-parity > gpurun_out/ncu_single.log 2>&1
grep -c pathtrace gpurun_out/launches.csv; tail -5 gpurun_out/bench.err
