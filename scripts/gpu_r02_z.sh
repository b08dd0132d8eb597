# round 2, final validation (third run: the state at the end of the round): full GPU suite, smoke, every bench line
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02_final3_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -4 $O/r02_final3_pytest_gpu.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_final3_smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_final3_smoke.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_final3_bench_asr_encoder_ref.json 2>/dev/null; cut -c1-200 $O/r02_final3_bench_asr_encoder_ref.json
timeout 600 python bench.py > $O/r02_final3_bench_asr_encoder.json 2> $O/r02_final3_bench.err; echo "bench rc=$?"; tail -c 300 $O/r02_final3_bench.err; cut -c1-200 $O/r02_final3_bench_asr_encoder.json
for w in encoder fbank stft_istft mvdr_tcn dccrn; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 --cpu-seconds 4 > $O/r02_final3_bench_$w.json 2>/dev/null; cut -c1-200 $O/r02_final3_bench_$w.json
done
