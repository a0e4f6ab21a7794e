mkdir -p gpurun_out; rm -f gpurun_out/quad_variants.log
for v in quasi_b200/lib/variants/*.so; do echo $v; QUASI_PLUGIN_OVERRIDE=$v timeout 120 python bench.py --no-cpu --steps 20 2>&1 | tail -1 | cut -c1-200 | tee -a gpurun_out/quad_variants.log; done
