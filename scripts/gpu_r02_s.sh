# round 2, visit S: evidence for profiles/ after the TMA-fed convolution — every bench line, launch list, ncu captures
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02s_bench_asr_encoder.json 2> $O/r02s_bench_asr_encoder.err; echo "bench rc=$?"; cut -c1-200 $O/r02s_bench_asr_encoder.json
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02s_bench_asr_encoder_ref.json 2>/dev/null; cut -c1-200 $O/r02s_bench_asr_encoder_ref.json
for w in encoder fbank stft_istft mvdr_tcn dccrn; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 --cpu-seconds 4 > $O/r02s_bench_$w.json 2> $O/r02s_bench_$w.err; cut -c1-200 $O/r02s_bench_$w.json
done
timeout 900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none \
    --csv --log-file $O/r02s_launches_asr_encoder.csv python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > $O/r02s_ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_gemm_kernel|conv2d_thin3x3_kernel' -s 0 -c 3 -f -o $O/r02s_prof_front \
    python bench.py --workload encoder --steps 1 --warmup 0 --cpu-seconds 0.1 > $O/r02s_ncu_front.log 2>&1
ls -la $O | grep r02s
