set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_n2.log 2>&1; echo "n2 rc=$?"; tail -2 gpurun_out/bench_n2.log | cut -c1-700
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_ref_n2.log | cut -c1-400
timeout 300 python bench.py --steps 30 --warmup 5 --cpu-seconds 1 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-300
