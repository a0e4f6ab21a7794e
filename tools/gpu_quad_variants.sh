# Parity tests + headline bench of every library under quasi_b200/lib/variants/ (tools/build_variant.sh).
mkdir -p gpurun_out; rm -f gpurun_out/quad_variants.log
for v in quasi_b200/lib/variants/*.so; do echo $v; QUASI_PLUGIN_OVERRIDE=$v timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -x -q 2>&1 | grep -E "passed|failed|Error|assert|differ" | head -8; QUASI_PLUGIN_OVERRIDE=$v timeout 120 python bench.py --no-cpu --steps 20 2>&1 | tail -1 | cut -c1-200 | tee -a gpurun_out/quad_variants.log; done
