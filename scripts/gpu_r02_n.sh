# round 2, visit N: state after the epilogue / codegen work — full suite, every bench line, launch list, ncu captures for profiles/
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02n_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02n_pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02n_bench_asr_encoder.json 2> $O/r02n_bench_asr_encoder.err; echo "bench rc=$?"; tail -c 400 $O/r02n_bench_asr_encoder.err; cut -c1-200 $O/r02n_bench_asr_encoder.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02n_bench_asr_encoder_ref.json 2>/dev/null; cut -c1-200 $O/r02n_bench_asr_encoder_ref.json
for w in encoder fbank stft_istft mvdr_tcn dccrn; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 --cpu-seconds 4 > $O/r02n_bench_$w.json 2> $O/r02n_bench_$w.err; cut -c1-200 $O/r02n_bench_$w.json
done
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
    --csv --log-file $O/r02n_launches_asr_encoder.csv python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > $O/r02n_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel' -s 8 -c 8 -f -o $O/r02n_prof_tc \
    python bench.py --workload encoder --steps 1 --warmup 3 --cpu-seconds 0.1 > $O/r02n_ncu_tc.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mhsa_tiled_kernel|layernorm2_kernel|conv2d_thin3x3_kernel|dwconv1d_kernel' -s 4 -c 6 -f -o $O/r02n_prof_enc_small \
    python bench.py --workload encoder --steps 1 --warmup 3 --cpu-seconds 0.1 > $O/r02n_ncu_small.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'frontend_kernel' -s 2 -c 1 -f -o $O/r02n_prof_f1 \
    python bench.py --workload fbank --steps 3 --warmup 3 --cpu-seconds 0.1 > $O/r02n_ncu_f1.log 2>&1
ls -la $O | grep r02n
