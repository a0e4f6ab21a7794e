N=${1:-8}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py 2>&1 | grep MULTI_GPU_CHECK
for mode in p2p nccl; do
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 --reduce $mode 2> gpurun_out/bench_multi_$mode.err | tail -1 | tee gpurun_out/bench_gpus${N}_$mode.json | cut -c1-330
done
