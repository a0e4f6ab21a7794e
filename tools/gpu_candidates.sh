# First GPU call of the next round: everything that was built and checked on the CPU after the
# last GPU minute of round 1, parity first, then the headline bench of each.
#   gpurun --timeout 900 -- 'bash tools/gpu_candidates.sh'      (run tools/build_candidates.sh before, here)
mkdir -p gpurun_out; log=gpurun_out/candidates.log; : > $log; : > gpurun_out/candidates_bench.jsonl; : > gpurun_out/candidates.err
run() { echo "== $*" | tee -a $log; "$@" 2>&1 | tail -${TAIL:-3} | cut -c1-400 | tee -a $log; }

# 1. the run-time options of the default library: batched launches (both orders), the mirror
QUASI_TEST_BATCH=1 TAIL=6 run timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -x -q -k "batched"
# one line per bench: the options, device Mrays/s and ms per frame, e2e Mrays/s and ms per frame; the full JSON lines in candidates_bench.jsonl
bench() {
    echo "== bench $*" | tee -a $log
    env "$@" timeout 120 python bench.py --no-cpu --steps 20 2>>gpurun_out/candidates.err | tail -1 > gpurun_out/_bench_line.json
    echo "{\"options\": \"$*\", \"line\": $(cat gpurun_out/_bench_line.json || echo null)}" >> gpurun_out/candidates_bench.jsonl
    python -c "import json,sys; d=json.load(open('gpurun_out/_bench_line.json')); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'], '| launches/step', d['config'].get('launches_per_step'))" 2>&1 | tail -1 | tee -a $log
}
bench QUASI_NOTHING=1
bench QUASI_CUDA_BATCH=8
bench QUASI_CUDA_BATCH=16
bench QUASI_CUDA_BATCH=16 QUASI_CUDA_MIRROR=1
bench QUASI_CUDA_BATCH=64
bench QUASI_CUDA_BATCH=64 QUASI_CUDA_MIRROR=1
bench QUASI_CUDA_BATCH=64 QUASI_CUDA_BATCH_ORDER=sample
bench QUASI_CUDA_BATCH=64 QUASI_CUDA_BATCH_ORDER=sample QUASI_CUDA_MIRROR=1
TAIL=1 run env QUASI_CUDA_BATCH=16 timeout 120 python bench.py --workload mesh --no-cpu --steps 5
TAIL=1 run timeout 120 python bench.py --workload mesh --no-cpu --steps 5

# 2. the build-time options: parity suite and bench of every library under lib/variants/
for v in quasi_b200/lib/variants/*.so; do
    QUASI_PLUGIN_OVERRIDE=$v TAIL=4 run timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -x -q
    bench QUASI_PLUGIN_OVERRIDE=$v
    case $v in *pdl*) ;; *) bench QUASI_PLUGIN_OVERRIDE=$v QUASI_CUDA_BATCH=64 QUASI_CUDA_MIRROR=1;; esac
done
