# round 2, visit G: completeness kernels (featops, ragged decoding, padded LSTM, fused SpecAugment), full suite + bench sanity
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > $O/r02g_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -25 $O/r02g_pytest_all.log
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02g_bench_asr_encoder.json 2> $O/r02g_bench_asr_encoder.err; echo "bench rc=$?"; tail -c 600 $O/r02g_bench_asr_encoder.err; cut -c1-240 $O/r02g_bench_asr_encoder.json
timeout 300 python bench.py --workload fbank --steps 50 --warmup 5 --cpu-seconds 1 > $O/r02g_bench_fbank.json 2>&1; cut -c1-240 $O/r02g_bench_fbank.json
timeout 300 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 1 > $O/r02g_bench_dccrn.json 2>&1; cut -c1-240 $O/r02g_bench_dccrn.json
