# Round 2, first GPU call: everything written at the end of round 1 that has never run (tools/gpu_candidates.sh),
# then the evidence the round-1 review asked for: the HBM-bound passes against measured HBM bandwidth (timed, then
# one ncu --set full capture of each), a fresh full capture of the default quad kernel with source, and the C++
# host's wall time for the headline frame, cold (64 frames from process start) and warm (10 x 64 frames).
#   gpurun --timeout 1500 -- 'bash tools/gpu_r2_call1.sh'       (run tools/build_candidates.sh before, here)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
bash tools/gpu_candidates.sh > /dev/null 2>&1
tail -c 6000 gpurun_out/candidates.log | grep -E "^==|passed|failed|error|\"value\"" | cut -c1-260

timeout 200 python tools/post_bench.py --out gpurun_out/post_bench.json > gpurun_out/post_bench.log 2>&1
grep -E '"kernel"|"GB/s"|"size"' gpurun_out/post_bench.json | paste - - - | cut -c1-200
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'accumulate_kernel|tonemap_kernel|combine_shards_kernel' \
    -c 30 -o gpurun_out/r2_post_full -f python tools/post_bench.py --ncu-friendly > gpurun_out/ncu_post.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pathtrace_accumulate -s 70 -c 1 \
    -o gpurun_out/r2_quad_full -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_quad.log 2>&1

BIN=quasi_b200/bin/quasi_headless
LIB=quasi_b200/lib/libquasi_cuda.so
for n in 64 640; do
  timeout 120 $BIN --render $n --width 1920 --height 1080 --json --out /tmp/host_$n.exr $LIB > gpurun_out/host_render_$n.log 2>&1
  grep '^{' gpurun_out/host_render_$n.log | tail -1 | cut -c1-200
done
ls -la gpurun_out | tail -20
