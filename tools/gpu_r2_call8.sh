# Round 2, eighth GPU call: 5 resident CTAs per SM (48 registers) against 4, a full ncu capture of the mesh kernel of
# this build, the compute-sanitizer passes over every kernel.
mkdir -p gpurun_out; log=gpurun_out/call8.log; : > $log
bench() {
    echo "== bench $*" | tee -a $log
    env "$@" timeout 120 python bench.py --no-cpu --no-extra --no-parity --steps 20 2>>gpurun_out/call8.err | tail -1 > gpurun_out/_bench_line.json
    python -c "import json,sys; d=json.load(open('gpurun_out/_bench_line.json')); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'], '| launches/step', d['config'].get('launches_per_step'))" 2>&1 | tail -1 | tee -a $log
}
bench QUASI_NOTHING=1
bench QUASI_PLUGIN_OVERRIDE=quasi_b200/lib/variants/libquasi_cuda_blocks5.so
bench QUASI_PLUGIN_OVERRIDE=quasi_b200/lib/variants/libquasi_cuda_blocks5.so QUASI_CUDA_BATCH=1
bench QUASI_CUDA_BATCH=1
bench QUASI_NOTHING=2
QUASI_PLUGIN_OVERRIDE=quasi_b200/lib/variants/libquasi_cuda_blocks5.so timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -q 2>&1 | tail -2 | tee -a $log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mesh_pathtrace -s 20 -c 1 -o gpurun_out/r2_mesh_full -f python tools/mesh_bench.py --steps 1 --warmup 1 > gpurun_out/ncu_mesh.log 2>&1
bash tools/gpu_sanitize.sh 2>&1 | tee -a $log
ls -la gpurun_out | tail -6
