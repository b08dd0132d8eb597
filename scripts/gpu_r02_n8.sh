# round 2: torchrun N = 8 of the default bench (our arm)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29537 bench.py --gpus 8 --steps 20 --warmup 5 > $O/r02_bench_n8.json 2> $O/r02_bench_n8.err; echo "n8 rc=$?"; tail -c 300 $O/r02_bench_n8.err; tail -1 $O/r02_bench_n8.json | cut -c1-220
