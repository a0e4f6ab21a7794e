# N GPUs (4): multi_gpu_check.py and bench.py --gpus N with the fused (default) and the push exchange, final binaries.
N=${1:-4}
mkdir -p gpurun_out; log=gpurun_out/final_$N.log; : > $log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multi_gpu_check.py on $N ranks" | tee -a $log
timeout 300 $TR --master-port 29533 tests/multi_gpu_check.py > gpurun_out/final_check_$N.log 2>&1; grep -E "FAIL|MULTI_GPU_CHECK|Error|error" gpurun_out/final_check_$N.log | head -20 | cut -c1-300 | tee -a $log
i=0
for opts in "" "--reduce push" "" "--reduce push"; do
  i=$((i+1))
  echo "== bench.py --gpus $N $opts" | tee -a $log
  timeout 200 $TR --master-port $((29540+i)) bench.py --gpus $N --no-cpu $opts > gpurun_out/bench_final_${N}_$i.json 2> gpurun_out/bench_final_${N}_$i.err
  tail -1 gpurun_out/bench_final_${N}_$i.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'], '|', d['scaling'], '| parity', d.get('parity'), '| issue', (d.get('roofline') or {}).get('issue'))" 2>&1 | tail -1 | cut -c1-500 | tee -a $log
  grep -E "Error|error|Traceback" gpurun_out/bench_final_${N}_$i.err | head -5 | cut -c1-300 | tee -a $log
done
