# round 2, visit A: tests + the new default bench + launch list with tensor-pipe % + full captures of kernels without one
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02a_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02a_bench_asr_encoder.json 2> $O/r02a_bench_asr_encoder.err; echo "bench rc=$?"; tail -c 600 $O/r02a_bench_asr_encoder.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02a_bench_asr_encoder_ref.json 2>&1
timeout 300 python bench.py --workload fbank --steps 50 --warmup 5 > $O/r02a_bench_fbank.json 2> $O/r02a_bench_fbank.err
timeout 300 python bench.py --workload stft_istft --steps 20 --warmup 5 --cpu-seconds 3 > $O/r02a_bench_stft_istft.json 2> $O/r02a_bench_stft_istft.err
timeout 300 python bench.py --workload mvdr_tcn --steps 20 --warmup 5 --cpu-seconds 3 > $O/r02a_bench_mvdr_tcn.json 2> $O/r02a_bench_mvdr_tcn.err
timeout 300 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 3 > $O/r02a_bench_dccrn.json 2> $O/r02a_bench_dccrn.err
# launch list of the default step (graph nodes are profiled one by one), with the tensor-pipe share of every launch
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
    --csv --log-file $O/r02a_launches_asr_encoder.csv python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > $O/r02a_ncu_launches.log 2>&1
# full captures (one launch each) of kernels that have none yet
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'istft_kernel|frontend_kernel' -s 6 -c 2 -f -o $O/r02a_prof_stft_istft \
    python bench.py --workload stft_istft --steps 3 --warmup 3 --cpu-seconds 0.1 > $O/r02a_ncu_stft.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'covar_kernel|beamform_kernel|dwconv1d_kernel' -s 20 -c 3 -f -o $O/r02a_prof_mvdr \
    python bench.py --workload mvdr_tcn --steps 3 --warmup 3 --cpu-seconds 0.1 > $O/r02a_ncu_mvdr.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lstm_step_kernel' -s 600 -c 1 -f -o $O/r02a_prof_lstm \
    python bench.py --workload dccrn --steps 3 --warmup 3 --cpu-seconds 0.1 > $O/r02a_ncu_lstm.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'mhsa_kernel|conv2d_narrow_kernel|layernorm_kernel' -s 30 -c 3 -f -o $O/r02a_prof_enc_small \
    python bench.py --workload encoder --steps 3 --warmup 3 --cpu-seconds 0.1 > $O/r02a_ncu_enc_small.log 2>&1
ls -la $O | tail -20
