mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_mesh.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python bench.py --workload mesh --steps 5 --no-cpu > gpurun_out/bench_mesh_b.json 2>> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench_mesh_b.json
timeout 200 python bench.py --workload mesh --mesh-builder lbvh --steps 5 --no-cpu > gpurun_out/bench_mesh_lbvh.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_mesh_lbvh.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['config']['plugin_messages'])
d=json.load(open('gpurun_out/bench_mesh_b.json')); print(d['value'], d['ms_per_step'], d['e2e'])"
