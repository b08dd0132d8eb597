# round 2, visit M: mode-dependent epilogue smem, re-fitted cost model
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02m_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02m_pytest_all.log
timeout 900 python scripts/dev_tc_mode3.py > $O/r02m_tc_mode3.txt 2>&1; sed -n '/cluster size sweep/,$p' $O/r02m_tc_mode3.txt
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02m_bench_asr_encoder.json 2> $O/r02m_bench_asr_encoder.err; echo "bench rc=$?"; tail -c 400 $O/r02m_bench_asr_encoder.err; cut -c1-200 $O/r02m_bench_asr_encoder.json
timeout 300 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 1 > $O/r02m_bench_dccrn.json 2>&1; cut -c1-200 $O/r02m_bench_dccrn.json
timeout 300 python bench.py --workload mvdr_tcn --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02m_bench_mvdr_tcn.json 2>&1; cut -c1-200 $O/r02m_bench_mvdr_tcn.json
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
    --csv --log-file $O/r02m_launches_asr_encoder.csv python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > $O/r02m_ncu_launches.log 2>&1
