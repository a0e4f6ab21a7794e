# Round 2, fourth GPU call: the batched launch in its pixel-major form with the tile export (the readback mirror).
mkdir -p gpurun_out; log=gpurun_out/call4.log; : > $log; : > gpurun_out/call4_bench.jsonl
bench() {
    echo "== bench $*" | tee -a $log
    env "$@" timeout 120 python bench.py --no-cpu --steps 20 2>>gpurun_out/call4.err | tail -1 > gpurun_out/_bench_line.json
    echo "{\"options\": \"$*\", \"line\": $(cat gpurun_out/_bench_line.json || echo null)}" >> gpurun_out/call4_bench.jsonl
    python -c "import json,sys; d=json.load(open('gpurun_out/_bench_line.json')); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'], '| launches/step', d['config'].get('launches_per_step'))" 2>&1 | tail -1 | tee -a $log
}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/call4_pytest.log 2>&1; tail -25 gpurun_out/call4_pytest.log | cut -c1-300 | tee -a $log
bench QUASI_NOTHING=1
bench QUASI_CUDA_MIRROR=0
bench QUASI_CUDA_BATCH=1
bench QUASI_CUDA_BATCH=16
bench QUASI_CUDA_BATCH=32
bench QUASI_CUDA_BATCH=32 QUASI_CUDA_MIRROR=0
bench QUASI_NOTHING=2
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pathtrace_accumulate -s 3 -c 1 \
    -o gpurun_out/r2_v3_batch_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_v3_batch.log 2>&1
QUASI_CUDA_MIRROR=0 timeout 400 ncu --set full --clock-control none --import-source on -k regex:pathtrace_accumulate -s 3 -c 1 \
    -o gpurun_out/r2_v3_batch_nomirror_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_v3_batch_nm.log 2>&1
ls -la gpurun_out | tail -8
