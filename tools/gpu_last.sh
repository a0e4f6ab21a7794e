mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_mesh.py -m gpu -x -q 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 100 python - <<'PY'
import time, sys
sys.path.insert(0, '.')
from quasi_b200.headless import headless_host
t = time.perf_counter()
h = headless_host(64, 36, env={"QUASI_CUDA_SCENE": "mesh", "QUASI_CUDA_MESH_GRID": 707, "QUASI_CUDA_MESH_SEED": 1, "QUASI_CUDA_MAX_BOUNCES": 4})
print("mesh plugin create (host SAH build on all cores): %.2f s" % (time.perf_counter() - t), h.messages[:1])
h.close()
PY
