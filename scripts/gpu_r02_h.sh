# round 2, visit H: weight-tile multicast across clusters — numerics first (every cluster size), then timing
set -x
mkdir -p gpurun_out
O=gpurun_out
for cl in 1 2 4; do
  APS_B200_TC_CL=$cl timeout 600 python -m pytest tests/test_encoder.py -m gpu -q -x -k "tensor_core or split_k or c4_conformer or golden" > $O/r02h_pytest_cl$cl.log 2>&1; echo "pytest CL=$cl rc=$?"; tail -4 $O/r02h_pytest_cl$cl.log
done
timeout 900 python scripts/dev_tc_mode3.py > $O/r02h_tc_mode3.txt 2>&1; echo "microbench rc=$?"; tail -22 $O/r02h_tc_mode3.txt
for cl in 1 2 4; do
  APS_B200_TC_CL=$cl timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02h_bench_cl$cl.json 2> $O/r02h_bench_cl$cl.err; echo "bench CL=$cl rc=$?"; tail -c 400 $O/r02h_bench_cl$cl.err; cut -c1-200 $O/r02h_bench_cl$cl.json
done
timeout 1500 python -m pytest tests -m gpu -q > $O/r02h_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -12 $O/r02h_pytest_all.log
