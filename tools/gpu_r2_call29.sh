# 2 GPUs: compute-sanitizer memcheck over tests/multi_gpu_check.py (push, fused, nccl exchanges; peer stores, system-scope
# atomics on peer counters, the device-side barrier), both ranks under the tool.
mkdir -p gpurun_out
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 --no-python \
  compute-sanitizer --tool memcheck --print-limit 20 --log-file gpurun_out/sanitize_multi_%p.log python tests/multi_gpu_check.py > gpurun_out/sanitize_multi.out 2>&1
grep -E "MULTI_GPU_CHECK|FAIL" gpurun_out/sanitize_multi.out | head -5
for f in gpurun_out/sanitize_multi_*.log; do echo "$f: $(grep -E 'ERROR SUMMARY' $f | tail -1)"; done
