mkdir -p gpurun_out
for v in quasi_b200/lib/variants/*.so; do echo $v; timeout 90 python tools/mesh_bench.py --steps 5 --warmup 2 --plugin $v 2>&1 | tail -1 | cut -c1-260; done
