# 1 GPU: the recorded hit carries the material (no tri_v0 load in shading): mesh parity on both builders, config 3 timing,
# launch list of one frame.
mkdir -p gpurun_out; log=gpurun_out/mesh_hit_material.log; : > $log
timeout 600 python -m pytest tests/test_cuda_mesh.py -m gpu -q 2>&1 | tail -2 | tee -a $log
run() { echo "== $*" | tee -a $log; timeout 150 python tools/mesh_bench.py --steps 5 --warmup 2 "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  ', d['ms_per_frame_median'], 'ms', d['mrays_per_s'], 'Mrays/s')" 2>&1 | tail -1 | tee -a $log; }
run
run --builder lbvh
run --env QUASI_CUDA_MESH_WAVEFRONT=0
run
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/mesh_hit_material_launches.csv python tools/mesh_bench.py --steps 1 --warmup 0 > /dev/null 2>&1
python - <<'PY' | tee -a $log
import csv
rows=[r for r in csv.reader(open('gpurun_out/mesh_hit_material_launches.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
from collections import defaultdict
t=defaultdict(list)
for r in rows[1:]:
    try: t[r[ik][:60]].append(float(r[iv].replace(',','')))
    except: pass
for k,v in t.items(): print(f"{k:60s} n={len(v):3d} sum={sum(v)/1e6:8.3f} ms  each {[round(x/1e3) for x in v[:12]]} us")
PY
