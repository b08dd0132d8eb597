# round 2, visit C: coalesced tensor-core epilogue, iSTFT / F2 restructure — tests, graph-timed microbenchmark, bench, launch list
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > $O/r02c_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -8 $O/r02c_pytest_all.log
timeout 600 python scripts/dev_tc_mode3.py > $O/r02c_tc_mode3.txt 2>&1; echo "microbench rc=$?"; cat $O/r02c_tc_mode3.txt
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02c_bench_asr_encoder.json 2> $O/r02c_bench_asr_encoder.err; echo "bench rc=$?"; tail -c 1500 $O/r02c_bench_asr_encoder.err; cut -c1-400 $O/r02c_bench_asr_encoder.json
timeout 300 python bench.py --workload stft_istft --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02c_bench_stft_istft.json 2>&1; cut -c1-300 $O/r02c_bench_stft_istft.json
timeout 300 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 1 > $O/r02c_bench_dccrn.json 2>&1; cut -c1-300 $O/r02c_bench_dccrn.json
timeout 300 python bench.py --workload mvdr_tcn --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02c_bench_mvdr_tcn.json 2>&1; cut -c1-300 $O/r02c_bench_mvdr_tcn.json
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
    --csv --log-file $O/r02c_launches_asr_encoder.csv python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > $O/r02c_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel<128, 3>|tc_gemm_kernel<256, 3>|layernorm2_kernel' -s 40 -c 4 -f -o $O/r02c_prof_tc_mode3 \
    python bench.py --workload encoder --steps 3 --warmup 3 --cpu-seconds 0.1 > $O/r02c_ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'istft_kernel|frontend_kernel' -s 6 -c 2 -f -o $O/r02c_prof_stft_istft \
    python bench.py --workload stft_istft --steps 3 --warmup 3 --cpu-seconds 0.1 > $O/r02c_ncu_stft.log 2>&1
ls -la $O | tail -12
