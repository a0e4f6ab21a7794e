# 1 GPU: wavefront mesh form after the streaming-kernel changes (raygen without atomics, CTA-level compaction in shade,
# 8-byte hits): parity, config 3, the launch list of one frame, and ncu --set full of the second trace launch (the first
# wave of scattered rays) and of the first shade launch.
mkdir -p gpurun_out; log=gpurun_out/mesh_wave2.log; : > $log
timeout 600 python -m pytest tests/test_cuda_mesh.py -m gpu -q -x 2>&1 | tail -3 | tee -a $log
run() { echo "== $*" | tee -a $log; timeout 150 python tools/mesh_bench.py --steps 5 --warmup 2 "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  ', d['ms_per_frame_median'], 'ms', d['mrays_per_s'], 'Mrays/s nodes/ray', d['nodes_per_ray'], 'tris/ray', d['triangles_per_ray'])" 2>&1 | tail -1 | cut -c1-300 | tee -a $log; }
run
run --builder lbvh
run --env QUASI_CUDA_MESH_QUANT=0
ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file gpurun_out/mesh_wave2_launches.csv python tools/mesh_bench.py --steps 1 --warmup 0 > /dev/null 2>&1
python - <<'PY' | tee -a $log
import csv
rows=[r for r in csv.reader(open('gpurun_out/mesh_wave2_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
from collections import defaultdict
t=defaultdict(list)
for r in rows[1:]:
    try: t[r[ik][:60]].append(float(r[iv].replace(',','')))
    except: pass
for k,v in t.items(): print(f"{k:60s} n={len(v):3d} sum={sum(v)/1e6:8.3f} ms  each {[round(x/1e3) for x in v[:12]]} us")
PY
timeout 300 ncu --set full --import-source on --clock-control none -k regex:mesh_wave_trace --launch-skip 1 --launch-count 1 -o gpurun_out/r2_wave_trace_full -f python tools/mesh_bench.py --steps 1 --warmup 0 > gpurun_out/ncu_wave_trace.log 2>&1; tail -2 gpurun_out/ncu_wave_trace.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:mesh_wave_shade --launch-skip 0 --launch-count 1 -o gpurun_out/r2_wave_shade_full -f python tools/mesh_bench.py --steps 1 --warmup 0 > gpurun_out/ncu_wave_shade.log 2>&1; tail -2 gpurun_out/ncu_wave_shade.log
