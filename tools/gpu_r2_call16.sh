# 8 (or N) GPUs: multi_gpu_check.py, then bench.py --gpus N: push against fused on the 1080p x 64 spp frame (strong
# scaling), fused at 4K x 1024 spp, fused with 64 spp per GPU (weak).
N=${1:-8}
mkdir -p gpurun_out; log=gpurun_out/fused8_$N.log; : > $log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
echo "== multi_gpu_check.py on $N ranks" | tee -a $log
timeout 300 $TR --master-port 29533 tests/multi_gpu_check.py > gpurun_out/fused8_check_$N.log 2>&1; grep -E "FAIL|MULTI_GPU_CHECK|Error|error" gpurun_out/fused8_check_$N.log | head -20 | cut -c1-300 | tee -a $log
i=0
for opts in "--reduce push" "--reduce fused" "--reduce push" "--reduce fused" "--reduce fused --workload demo4k --steps 3" "--reduce fused --scaling weak --steps 5"; do
  i=$((i+1))
  echo "== bench.py --gpus $N $opts" | tee -a $log
  timeout 200 $TR --master-port $((29540+i)) bench.py --gpus $N --no-cpu $opts > gpurun_out/bench_fused8_${N}_$i.json 2> gpurun_out/bench_fused8_${N}_$i.err
  tail -1 gpurun_out/bench_fused8_${N}_$i.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'], '|', d['scaling'], '| parity', d.get('parity'))" 2>&1 | tail -1 | cut -c1-400 | tee -a $log
  grep -E "Error|error|Traceback" gpurun_out/bench_fused8_${N}_$i.err | head -5 | cut -c1-300 | tee -a $log
done
