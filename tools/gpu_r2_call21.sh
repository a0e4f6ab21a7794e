# 1 GPU, the final binaries of this round: the whole GPU suite, smoke(), the default bench (as the driver runs it), the
# launch lists, and ncu --set full of the dominant kernels: the batched quad kernel (launch 4 of a short bench run) and
# the ten launches of one 16-spp frame of the mesh scene (raygen, 4 x trace, 4 x shade, resolve).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json | cut -c1-400; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --workload mesh --steps 5 --no-cpu > gpurun_out/bench_mesh.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_mesh.json | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-parity > gpurun_out/ncu_launches.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/mesh_launches.csv python tools/mesh_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_mesh_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pathtrace_accumulate -s 3 -c 1 -o gpurun_out/r2b_quad_batch_full -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --no-parity > gpurun_out/ncu_quad.log 2>&1; tail -1 gpurun_out/ncu_quad.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mesh_wave -s 10 -c 10 -o gpurun_out/r2b_mesh_wave_full -f python tools/mesh_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_mesh_wave.log 2>&1; tail -1 gpurun_out/ncu_mesh_wave.log
