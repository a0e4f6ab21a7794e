# 2 GPUs: the fused exchange with arrival counters that the completing launch zeroes again: multi_gpu_check.py, the C++
# host on two GPUs, one bench line.
mkdir -p gpurun_out; log=gpurun_out/fused_reset_2.log; : > $log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29533 tests/multi_gpu_check.py > gpurun_out/fused_reset_check.log 2>&1; grep -E "FAIL|MULTI_GPU_CHECK|Error|error" gpurun_out/fused_reset_check.log | head -10 | cut -c1-300 | tee -a $log
timeout 300 python -m pytest tests/test_host_cpp.py -m gpu -q -k "several" 2>&1 | tail -1 | tee -a $log
timeout 200 $TR --master-port 29541 bench.py --gpus 2 --no-cpu > gpurun_out/bench_fused_reset_2.json 2> gpurun_out/bench_fused_reset_2.err
tail -1 gpurun_out/bench_fused_reset_2.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'], '| parity', d.get('parity'))" 2>&1 | tail -1 | cut -c1-400 | tee -a $log
