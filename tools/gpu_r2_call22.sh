# 1 GPU: the quad kernel with tiles handed out centre-out (variant) against row by row (default): parity of the variant,
# then the headline bench alternately.
mkdir -p gpurun_out; log=gpurun_out/tile_order.log; : > $log
V=$PWD/quasi_b200/lib/variants/libquasi_cuda_centre_out.so
QUASI_PLUGIN_OVERRIDE=$V timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -q -x 2>&1 | tail -2 | tee -a $log
for i in 1 2; do
  for which in default variant; do
    if [ $which = default ]; then unset QUASI_PLUGIN_OVERRIDE; else export QUASI_PLUGIN_OVERRIDE=$V; fi
    echo "== $which" | tee -a $log
    timeout 300 python bench.py --no-cpu --no-extra --no-parity 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value', d['value'], 'ms', d['ms_per_step'], '| e2e', d['e2e']['value'], 'ms', d['e2e']['ms_per_step'])" | tee -a $log
  done
done
unset QUASI_PLUGIN_OVERRIDE
echo "== 4K default / variant" | tee -a $log
timeout 300 python bench.py --workload demo4k --steps 2 --no-cpu --no-extra --no-parity 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value', d['value'], 'ms', d['ms_per_step'])" | tee -a $log
QUASI_PLUGIN_OVERRIDE=$V timeout 300 python bench.py --workload demo4k --steps 2 --no-cpu --no-extra --no-parity 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('   value', d['value'], 'ms', d['ms_per_step'])" | tee -a $log
