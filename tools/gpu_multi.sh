N=${1:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py 2>&1 | grep MULTI_GPU_CHECK | tee gpurun_out/multi_check_$N.log
for mode in nccl p2p; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --reduce $mode 2> gpurun_out/bench_multi_$mode.err | tail -1 | tee gpurun_out/bench_gpus${N}_$mode.json | cut -c1-330
done
timeout 120 quasi_b200/bin/quasi_headless --render $((1024 / N)) --width 3840 --height 2160 --gpus $N --json --out /tmp/config4.exr quasi_b200/lib/libquasi_cuda.so > gpurun_out/config4_g$N.log 2>&1
grep '^{' gpurun_out/config4_g$N.log | tail -1 | tee gpurun_out/config4_g$N.json
