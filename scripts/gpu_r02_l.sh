# round 2, visit K: epilogue read-back fix (shared-space asm, batched) — tests, trace, microbench, benches
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02l_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02l_pytest_all.log
APS_B200_LIB=aps_b200/libaps_b200_trace.so timeout 300 python scripts/dev_tc_trace5.py > $O/r02l_tc_trace.txt 2>&1; grep -E "===|tile " $O/r02l_tc_trace.txt
timeout 900 python scripts/dev_tc_mode3.py > $O/r02l_tc_mode3.txt 2>&1; grep -E "tma|splitk|cluster|CL=|conv" $O/r02l_tc_mode3.txt | head -60
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02l_bench_asr_encoder.json 2> $O/r02l_bench_asr_encoder.err; echo "bench rc=$?"; tail -c 400 $O/r02l_bench_asr_encoder.err; cut -c1-200 $O/r02l_bench_asr_encoder.json
timeout 300 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 1 > $O/r02l_bench_dccrn.json 2>&1; cut -c1-200 $O/r02l_bench_dccrn.json
timeout 300 python bench.py --workload mvdr_tcn --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02l_bench_mvdr_tcn.json 2>&1; cut -c1-200 $O/r02l_bench_mvdr_tcn.json
