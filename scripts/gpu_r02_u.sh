# round 2, visit U: rolled epilogue loop for the TMA-fed linear kernel (A/B)
set -x
mkdir -p gpurun_out
O=gpurun_out
V=$PWD/aps_b200/libaps_b200_rolled.so
APS_B200_LIB=$V timeout 600 python -m pytest tests/test_encoder.py tests/test_dropin.py -m gpu -q > $O/r02u_pytest_rolled.log 2>&1; echo "pytest rolled rc=$?"; tail -3 $O/r02u_pytest_rolled.log
for rep in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02u_bench_def_$rep.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02u_bench_def_$rep.json'));print('default', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
  APS_B200_LIB=$V timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02u_bench_rolled_$rep.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02u_bench_rolled_$rep.json'));print('rolled', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
APS_B200_LIB=$V timeout 400 python scripts/dev_tc_mode3.py 2>&1 | head -68 > $O/r02u_tc_mode3_rolled.txt; grep -n "tma\|splitk" $O/r02u_tc_mode3_rolled.txt
