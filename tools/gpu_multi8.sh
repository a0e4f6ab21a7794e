# Under `gpurun --gpus 8`: parity on 8 ranks, then the headline frame strong-scaled over 8 GPUs: the push exchange with
# the slices going to the host by copy engine (default) and by the reduce kernel's own stores, the NCCL baseline
# (strong and weak), BASELINE config 4 (4K x 1024 spp) and 4 GPUs for the curve.
mkdir -p gpurun_out; log=gpurun_out/multi_8.log; : > $log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
echo "== multi_gpu_check.py on 8 ranks" | tee -a $log
timeout 600 $TR --nproc-per-node 8 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/multi_check_8.log 2>&1; grep -E "FAIL|MULTI_GPU_CHECK|Error|error" gpurun_out/multi_check_8.log | head -20 | cut -c1-300 | tee -a $log
i=0
run() {  # N, env assignments..., then bench options after --
  n=$1; shift; envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  i=$((i+1))
  echo "== ${envs[*]} bench.py --gpus $n $*" | tee -a $log
  env "${envs[@]}" timeout 400 $TR --nproc-per-node $n --master-port $((29540+i)) bench.py --gpus $n --no-cpu "$@" > gpurun_out/bench_multi_8_$i.json 2> gpurun_out/bench_multi_8_$i.err
  tail -1 gpurun_out/bench_multi_8_$i.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'], '|', d['scaling'], '| parity', d.get('parity'))" 2>&1 | tail -1 | cut -c1-400 | tee -a $log
  grep -E "Error|error|Traceback" gpurun_out/bench_multi_8_$i.err | head -5 | cut -c1-300 | tee -a $log
}
run 8 X=1 -- --reduce push --steps 40
run 8 QUASI_EXCHANGE_HOST_PATH=store -- --reduce push --steps 40 --no-parity
run 8 X=1 -- --reduce nccl --steps 20 --no-parity
run 8 NCCL_ALGO=Ring -- --reduce nccl --steps 20 --no-parity
run 8 X=1 -- --reduce nccl --scaling weak --steps 10 --no-parity
run 4 X=1 -- --reduce push --steps 40
run 8 X=1 -- --reduce push --workload demo4k --steps 5 --no-parity
