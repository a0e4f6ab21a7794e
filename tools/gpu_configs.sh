# BASELINE configs 4 and 5 through the C++ headless host. usage: bash tools/gpu_configs.sh N
N=${1:-1}
mkdir -p gpurun_out
BIN=quasi_b200/bin/quasi_headless
LIB=quasi_b200/lib/libquasi_cuda.so
# config 4: 3840x2160, 1024 spp in total, sample-sharded over G GPUs (G = 1 for the reference line)
for G in 1 $N; do
  [ "$G" = "1" ] && [ "$N" = "1" ] && [ -n "$seen" ] && continue
  seen=1
  timeout 300 $BIN --render $((1024 / G)) --width 3840 --height 2160 --gpus $G --json --out /tmp/config4_g$G.exr $LIB > gpurun_out/config4_g$G.log 2>&1
  grep '^{' gpurun_out/config4_g$G.log | tail -1 | tee gpurun_out/config4_g$G.json
  ls -la /tmp/config4_g$G.exr | awk '{print "exr bytes", $5}'
done
# config 5: progressive 1080p, 1 spp + tonemap per frame, library replaced mid-run
cp $LIB /tmp/libquasi_cuda_live.so
( sleep 3; cp $LIB /tmp/libquasi_cuda_live.so.new; mv /tmp/libquasi_cuda_live.so.new /tmp/libquasi_cuda_live.so ) &
timeout 120 $BIN --progressive 6 --width 1920 --height 1080 --watch --json /tmp/libquasi_cuda_live.so > gpurun_out/config5.log 2>&1
wait
grep '^{' gpurun_out/config5.log | tail -1 | tee gpurun_out/config5.json
# progressive without reload, for the steady-state latency
timeout 120 $BIN --progressive 4 --width 1920 --height 1080 --json $LIB > gpurun_out/config5_steady.log 2>&1
grep '^{' gpurun_out/config5_steady.log | tail -1 | tee gpurun_out/config5_steady.json
