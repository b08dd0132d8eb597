# round 2, visit D: register-tiled attention + programmatic dependent launch
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_encoder.py -m gpu -q -x > $O/r02d_pytest_enc.log 2>&1; echo "pytest enc rc=$?"; tail -6 $O/r02d_pytest_enc.log
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02d_bench_nopdl.json 2> $O/r02d_bench_nopdl.err; echo "bench rc=$?"; tail -c 800 $O/r02d_bench_nopdl.err; cut -c1-260 $O/r02d_bench_nopdl.json
APS_B200_PDL=1 timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02d_bench_pdl.json 2> $O/r02d_bench_pdl.err; echo "bench pdl rc=$?"; tail -c 800 $O/r02d_bench_pdl.err; cut -c1-260 $O/r02d_bench_pdl.json
APS_B200_PDL=1 timeout 900 python -m pytest tests/test_encoder.py tests/test_tcn.py tests/test_dccrn.py tests/test_lstm.py -m gpu -q -x > $O/r02d_pytest_pdl.log 2>&1; echo "pytest pdl rc=$?"; tail -6 $O/r02d_pytest_pdl.log
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02d_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -5 $O/r02d_pytest_all.log
APS_B200_PDL=1 timeout 300 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 1 > $O/r02d_bench_dccrn_pdl.json 2>&1; cut -c1-260 $O/r02d_bench_dccrn_pdl.json
APS_B200_PDL=1 timeout 300 python bench.py --workload mvdr_tcn --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02d_bench_mvdr_tcn_pdl.json 2>&1; cut -c1-260 $O/r02d_bench_mvdr_tcn_pdl.json
