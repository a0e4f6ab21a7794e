# compute-sanitizer passes over small invocations of every kernel. Output: gpurun_out/sanitize_*.log
mkdir -p gpurun_out
cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from quasi_b200.headless import headless_host
import numpy as np
h = headless_host(97, 61, present=True); h.render(3); a = h.readback_aov(); print("quad, one launch per call + tonemap", h.stats().rays); h.close()
h = headless_host(97, 61); h.render(7); a = h.readback_aov(); b = h.readback_aov(); print("quad, batched launch with the tile export", h.stats().rays, h.stats().kernel_launches); h.close()
h = headless_host(61, 35, env={"QUASI_CUDA_BATCH": 3, "QUASI_CUDA_MIRROR": 0, "QUASI_CUDA_SCAN": "generic"}); h.render(7); a = h.readback_aov(); print("quad, generic scan, batches of 3", h.stats().rays); h.close()
for builder in ("sah", "lbvh"):
    h = headless_host(80, 45, env={"QUASI_CUDA_SCENE": "mesh", "QUASI_CUDA_MESH_GRID": 40, "QUASI_CUDA_MESH_SEED": 1,
                                   "QUASI_CUDA_MAX_BOUNCES": 4, "QUASI_CUDA_MESH_BUILDER": builder})
    h.render(2); a = h.readback_aov(); print("mesh, wavefront form (quantized tree)", builder, h.stats().rays); h.close()
h = headless_host(77, 45, env={"QUASI_CUDA_SCENE": "mesh", "QUASI_CUDA_MESH_GRID": 40, "QUASI_CUDA_MESH_SEED": 1, "QUASI_CUDA_MAX_BOUNCES": 4,
                               "QUASI_CUDA_MESH_BATCH": 2, "QUASI_CUDA_MESH_QUANT": 0})
h.render(5); a = h.readback_aov(); print("mesh, wavefront form, float tree, ragged image, batches of 2", h.stats().rays); h.close()
h = headless_host(80, 45, env={"QUASI_CUDA_SCENE": "mesh", "QUASI_CUDA_MESH_GRID": 40, "QUASI_CUDA_MESH_SEED": 1, "QUASI_CUDA_MAX_BOUNCES": 4,
                               "QUASI_CUDA_MESH_WAVEFRONT": 0})
h.render(2); a = h.readback_aov(); print("mesh, one-kernel form", h.stats().rays); h.close()
h = headless_host(64, 36, keep_accumulation=True); h.render(2); h.reload(); h.render(2); print("persist", h.stats().frame_count); h.close()
PY
for tool in memcheck racecheck initcheck synccheck; do
  timeout 400 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_$tool.log | tail -1)"
done
