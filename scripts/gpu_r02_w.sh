# round 2, visit W: sliding depthwise conv; GRP = 8 epilogue variant (A/B)
set -x
mkdir -p gpurun_out
O=gpurun_out
V=$PWD/aps_b200/libaps_b200_grp8.so
timeout 600 python -m pytest tests/test_encoder.py tests/test_tcn.py tests/test_dropin.py -m gpu -q > $O/r02w_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02w_pytest.log
timeout 300 python scripts/dev_r02o.py 2>&1 | tail -3
for rep in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02w_bench_def_$rep.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02w_bench_def_$rep.json'));print('default', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
  APS_B200_LIB=$V timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02w_bench_grp8_$rep.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02w_bench_grp8_$rep.json'));print('grp8', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
APS_B200_LIB=$V timeout 300 python -m pytest tests/test_encoder.py -m gpu -q 2>&1 | tail -2
timeout 300 python bench.py --workload mvdr_tcn --steps 20 --warmup 5 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('mvdr_tcn', d['ms_per_step'])"
