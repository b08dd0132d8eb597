# round 2, visit B: MODE 3 / split-K / companions — tests first, then the microbenchmark and the default bench
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_encoder.py -m gpu -q -x > $O/r02b_pytest_enc.log 2>&1; echo "pytest enc rc=$?"; tail -15 $O/r02b_pytest_enc.log
timeout 600 python scripts/dev_tc_mode3.py > $O/r02b_tc_mode3.txt 2>&1; echo "microbench rc=$?"; cat $O/r02b_tc_mode3.txt
timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02b_bench_asr_encoder.json 2> $O/r02b_bench_asr_encoder.err; echo "bench rc=$?"; tail -c 1500 $O/r02b_bench_asr_encoder.err; cut -c1-1200 $O/r02b_bench_asr_encoder.json
APS_B200_ENC_PAIRS=0 timeout 600 python bench.py --steps 20 --warmup 5 --cpu-seconds 1 > $O/r02b_bench_asr_encoder_nopairs.json 2>&1; cut -c1-300 $O/r02b_bench_asr_encoder_nopairs.json
timeout 900 python -m pytest tests -m gpu -q -x > $O/r02b_pytest_all.log 2>&1; echo "pytest all rc=$?"; tail -5 $O/r02b_pytest_all.log
