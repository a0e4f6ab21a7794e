# Round 2, third GPU call: the product as it now stands (one quad kernel: per-warp path queues, packed scan records,
# table sin/cos, batched launches of 64 with per-tile epochs and the tile export as the readback mirror).
#   the whole GPU suite, the headline bench with the options one at a time, config 3, ncu of the final kernels.
mkdir -p gpurun_out; log=gpurun_out/call3.log; : > $log; : > gpurun_out/call3_bench.jsonl
run() { echo "== $*" | tee -a $log; "$@" 2>&1 | tail -${TAIL:-3} | cut -c1-600 | tee -a $log; }
bench() {
    echo "== bench $*" | tee -a $log
    env "$@" timeout 120 python bench.py --no-cpu --steps 20 2>>gpurun_out/call3.err | tail -1 > gpurun_out/_bench_line.json
    echo "{\"options\": \"$*\", \"line\": $(cat gpurun_out/_bench_line.json || echo null)}" >> gpurun_out/call3_bench.jsonl
    python -c "import json,sys; d=json.load(open('gpurun_out/_bench_line.json')); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'], '| launches/step', d['config'].get('launches_per_step'))" 2>&1 | tail -1 | tee -a $log
}
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/call3_pytest.log 2>&1; tail -25 gpurun_out/call3_pytest.log | cut -c1-300 | tee -a $log
bench QUASI_NOTHING=1
bench QUASI_CUDA_MIRROR=0
bench QUASI_CUDA_BATCH=1
bench QUASI_CUDA_BATCH=8
bench QUASI_CUDA_BATCH=16
bench QUASI_CUDA_BATCH=32
bench QUASI_CUDA_SCAN=generic
bench QUASI_CUDA_SCAN=boxes
bench QUASI_NOTHING=2
TAIL=1 run timeout 200 python bench.py --workload mesh --no-cpu --steps 5
timeout 200 python tools/post_bench.py --out gpurun_out/post_bench_c.json > gpurun_out/post_bench_c.log 2>&1
grep -E '"kernel"|"GB/s"|"size"' gpurun_out/post_bench_c.json | paste - - - | cut -c1-200 | tee -a $log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pathtrace_accumulate -s 3 -c 1 \
    -o gpurun_out/r2_final_batch_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_final_batch.log 2>&1
QUASI_CUDA_BATCH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:pathtrace_accumulate -s 70 -c 1 \
    -o gpurun_out/r2_final_single_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_final_single.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'accumulate_kernel|tonemap_kernel|combine_shards_kernel' -c 30 \
    -o gpurun_out/r2_post_final_full -f python tools/post_bench.py --ncu-friendly > gpurun_out/ncu_post_final.log 2>&1
BIN=quasi_b200/bin/quasi_headless; LIB=quasi_b200/lib/libquasi_cuda.so
timeout 120 $BIN --render 64 --repeat 8 --readback-each --width 1920 --height 1080 --json --out /tmp/host.exr $LIB > gpurun_out/host_repeat.log 2>&1
grep '^{' gpurun_out/host_repeat.log | tail -1 | cut -c1-400 | tee -a $log
ls -la gpurun_out | tail -14
