# round 2, visit V (2 GPUs): encoder tests after the thin-conv guard, torchrun N = 2 of both arms
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_encoder.py -m gpu -q > $O/r02v_pytest_enc.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02v_pytest_enc.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02v_bench_n2.json 2> $O/r02v_bench_n2.err; echo "n2 rc=$?"; tail -c 300 $O/r02v_bench_n2.err; cut -c1-260 $O/r02v_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $O/r02v_bench_n2_ref.json 2> $O/r02v_bench_n2_ref.err; echo "n2 ref rc=$?"; cut -c1-260 $O/r02v_bench_n2_ref.json
