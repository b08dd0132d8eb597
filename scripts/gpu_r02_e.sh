# round 2, visit E: full GPU suite after the LSTM / PDL fix, bench with the graph-replay kernel leg, secondary workloads
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r02e_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -5 $O/r02e_pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02e_bench_asr_encoder.json 2> $O/r02e_bench_asr_encoder.err; echo "bench rc=$?"; tail -c 800 $O/r02e_bench_asr_encoder.err; cut -c1-260 $O/r02e_bench_asr_encoder.json
timeout 300 python bench.py --workload encoder --steps 20 --warmup 5 --cpu-seconds 2 > $O/r02e_bench_encoder.json 2>&1; cut -c1-260 $O/r02e_bench_encoder.json
timeout 300 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 2 > $O/r02e_bench_dccrn.json 2>&1; cut -c1-260 $O/r02e_bench_dccrn.json
timeout 300 python bench.py --workload mvdr_tcn --steps 20 --warmup 5 --cpu-seconds 2 > $O/r02e_bench_mvdr_tcn.json 2>&1; cut -c1-260 $O/r02e_bench_mvdr_tcn.json
timeout 300 python __graft_entry__.py --smoke > $O/r02e_smoke.log 2>&1; tail -2 $O/r02e_smoke.log
