# 1 GPU: ncu of the final quad kernel (tiles handed out centre-out): launch list of a short bench run and one --set full
# capture of the batched launch.
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-parity > gpurun_out/ncu_launches.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pathtrace_accumulate -s 3 -c 1 -o gpurun_out/r2c_quad_batch_full -f python bench.py --steps 1 --warmup 3 --no-cpu --no-extra --no-parity > gpurun_out/ncu_quad.log 2>&1; tail -1 gpurun_out/ncu_quad.log
grep -c pathtrace gpurun_out/launches.csv
