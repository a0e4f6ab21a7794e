# usage: bash tools/gpu_multi.sh N   (under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_cuda_multi_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/pytest_multi.log
for mode in p2p p2p_host nccl; do
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --reduce $mode 2> gpurun_out/bench_multi_$mode.err | tail -1 | tee gpurun_out/bench_gpus${N}_$mode.json | cut -c1-400
done
