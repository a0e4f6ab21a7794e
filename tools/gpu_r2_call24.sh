# 2 GPUs: the C++ host's fused exchange (and its pull fallback) against the oracle's combine, the pytest multi-GPU
# test, and the time of a 1080p x 64 spp run of the C++ host on 2 GPUs either way.
mkdir -p gpurun_out; log=gpurun_out/host_multi.log; : > $log
timeout 600 python -m pytest tests/test_host_cpp.py tests/test_cuda_multi_gpu.py -m gpu -q -k "several or two_gpu" 2>&1 | tail -8 | tee -a $log
for mode in fused pull; do
  echo "== quasi_headless --render 32 --gpus 2, QUASI_HOST_EXCHANGE=$mode" | tee -a $log
  QUASI_HOST_EXCHANGE=$mode timeout 120 quasi_b200/bin/quasi_headless --render 32 --gpus 2 --width 1920 --height 1080 --out gpurun_out/host_multi_$mode.exr --json quasi_b200/lib/libquasi_cuda.so 2>&1 | grep -E "Host\]|frames" | cut -c1-260 | tee -a $log
done
cmp gpurun_out/host_multi_fused.exr gpurun_out/host_multi_pull.exr && echo "EXR identical (fused == pull)" | tee -a $log
rm -f gpurun_out/host_multi_*.exr
