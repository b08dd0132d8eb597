# round 2, visit AD: pair schedule (TMA-fed 1x1 convolutions) in the TCN repeat stack
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_tcn.py tests/test_mvdr.py tests/test_dropin.py -m gpu -q -x > $O/r02ad_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r02ad_pytest.log
for rep in 1 2; do
timeout 300 python bench.py --workload mvdr_tcn --steps 20 --warmup 5 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('mvdr_tcn pairs', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])"
APS_B200_TCN_PAIRS=0 timeout 300 python bench.py --workload mvdr_tcn --steps 20 --warmup 5 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('mvdr_tcn gather', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
