# 1 GPU, the last pass of the round on the final binaries: the GPU suite, smoke(), the default bench and the reference
# arm exactly as the driver runs them, config 4 on one GPU, config 5 (progressive, with and without a reload).
mkdir -p gpurun_out; : > gpurun_out/bench.err
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 500 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -1 gpurun_out/bench.json | cut -c1-300; tail -2 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_reference.json | cut -c1-200
timeout 300 python bench.py --workload demo4k --steps 2 --no-cpu --no-extra > gpurun_out/bench_4k.json 2>> gpurun_out/bench.err; tail -1 gpurun_out/bench_4k.json | cut -c1-200
# config 5 only (config 4 through the C++ host takes a minute per line)
BIN=quasi_b200/bin/quasi_headless; LIB=quasi_b200/lib/libquasi_cuda.so
cp $LIB /tmp/libquasi_cuda_live.so
( sleep 3; cp $LIB /tmp/libquasi_cuda_live.so.new; mv /tmp/libquasi_cuda_live.so.new /tmp/libquasi_cuda_live.so ) &
timeout 120 $BIN --progressive 6 --width 1920 --height 1080 --watch --json /tmp/libquasi_cuda_live.so > gpurun_out/config5.log 2>&1
wait
grep '^{' gpurun_out/config5.log | tail -1 | tee gpurun_out/config5.json | cut -c1-330
timeout 120 $BIN --progressive 4 --width 1920 --height 1080 --json $LIB > gpurun_out/config5_steady.log 2>&1
grep '^{' gpurun_out/config5_steady.log | tail -1 | tee gpurun_out/config5_steady.json | cut -c1-330
