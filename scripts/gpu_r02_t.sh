# round 2, visit T: raw split-K store path; plain epilogue variant (A/B)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_encoder.py tests/test_dropin.py -m gpu -q > $O/r02t_pytest_enc.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02t_pytest_enc.log
APS_B200_LIB=$PWD/aps_b200/libaps_b200_plain.so timeout 600 python -m pytest tests/test_encoder.py -m gpu -q > $O/r02t_pytest_plain.log 2>&1; echo "pytest plain rc=$?"; tail -3 $O/r02t_pytest_plain.log
for rep in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02t_bench_def_$rep.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02t_bench_def_$rep.json'));print('default', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
  APS_B200_LIB=$PWD/aps_b200/libaps_b200_plain.so timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02t_bench_plain_$rep.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02t_bench_plain_$rep.json'));print('plain', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
timeout 400 python scripts/dev_tc_mode3.py 2>&1 | head -68 > $O/r02t_tc_mode3.txt; grep -n "ffn_a swish\|ffn_b k5\|front k5\|qkv" $O/r02t_tc_mode3.txt
