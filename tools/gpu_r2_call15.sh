# 2 GPUs: the fused exchange with the copies of a unit loaded one after the other (default build: no spills) against
# four at a time (variant v4: 30 bytes of spills in the path loop), and push for reference.
N=${1:-2}
mkdir -p gpurun_out; log=gpurun_out/fused_seq_$N.log; : > $log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
i=0
for cfg in "default fused" "v4 fused" "default push" "default fused" "v4 fused"; do
  set -- $cfg; i=$((i+1))
  echo "== $1 --reduce $2" | tee -a $log
  if [ $1 = default ]; then unset QUASI_PLUGIN_OVERRIDE; else export QUASI_PLUGIN_OVERRIDE=$PWD/quasi_b200/lib/variants/libquasi_cuda_$1.so; fi
  timeout 600 $TR --master-port $((29540+i)) bench.py --gpus $N --no-cpu --reduce $2 > gpurun_out/bench_seq_${N}_$i.json 2> gpurun_out/bench_seq_${N}_$i.err
  tail -1 gpurun_out/bench_seq_${N}_$i.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'], '| parity', d.get('parity'))" 2>&1 | tail -1 | cut -c1-300 | tee -a $log
  grep -E "Error|error|Traceback" gpurun_out/bench_seq_${N}_$i.err | head -5 | cut -c1-300 | tee -a $log
done
