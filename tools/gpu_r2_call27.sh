# 1 GPU: the tonemap pass with a narrower fallback band (variants tm1: 1e-3, tm07: 7e-4; default 2e-3): HBM fraction
# from tools/post_bench.py, and how many 8-bit values differ from the literal form (oracle) on a rendered image.
mkdir -p gpurun_out; log=gpurun_out/tonemap_band.log; : > $log
for v in default tm1 tm07; do
  if [ $v = default ]; then unset QUASI_PLUGIN_OVERRIDE; else export QUASI_PLUGIN_OVERRIDE=$PWD/quasi_b200/lib/variants/libquasi_cuda_$v.so; fi
  echo "== $v" | tee -a $log
  timeout 200 python tools/post_bench.py --out gpurun_out/post_$v.json > /dev/null 2>&1
  python -c "
import json
for p in json.load(open('gpurun_out/post_$v.json'))['passes']:
    if 'tonemap' in p['kernel']: print('  ', p['kernel'], p['size'], p['ms']*1e3, 'us', p['GB/s'], 'GB/s', p['frac_of_hbm_peak'])
" | tee -a $log
  timeout 200 python -m pytest tests/test_cuda_parity.py -m gpu -q -k "tonemap or standalone" 2>&1 | tail -1 | tee -a $log
done
