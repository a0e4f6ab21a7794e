# 1 GPU: the wavefront form with the quantized tree: parity (test_cuda_mesh.py), config 3 with the float tree, the
# quantized tree and the refill thresholds, then the launch list of one frame.
mkdir -p gpurun_out; log=gpurun_out/mesh_quant.log; : > $log
timeout 600 python -m pytest tests/test_cuda_mesh.py -m gpu -q -x 2>&1 | tail -5 | tee -a $log
run() { echo "== $*" | tee -a $log; timeout 150 python tools/mesh_bench.py --steps 5 --warmup 2 "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  ', d['ms_per_frame_median'], 'ms', d['mrays_per_s'], 'Mrays/s nodes/ray', d['nodes_per_ray'], 'tris/ray', d['triangles_per_ray'])" 2>&1 | tail -1 | cut -c1-300 | tee -a $log; }
run --env QUASI_CUDA_MESH_WAVEFRONT=0
run --env QUASI_CUDA_MESH_QUANT=0
run
run --builder lbvh
for v in quasi_b200/lib/variants/libquasi_cuda_w*.so; do [ -f $v ] && run --plugin $v; done
run --env QUASI_CUDA_MESH_L2_KEEP=100
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/mesh_quant_launches.csv python tools/mesh_bench.py --steps 1 --warmup 1 > /dev/null 2>&1
python - <<'PY' | tee -a $log
import csv
rows=[r for r in csv.reader(open('gpurun_out/mesh_quant_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
from collections import defaultdict
t=defaultdict(list)
for r in rows[1:]:
    try: t[r[ik][:60]].append(float(r[iv].replace(',','')))
    except: pass
for k,v in t.items(): print(f"{k:60s} n={len(v):3d} sum={sum(v)/1e6:8.3f} ms  each {[round(x/1e3) for x in v[:12]]} us")
PY
