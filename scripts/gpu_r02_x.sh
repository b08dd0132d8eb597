# round 2, visit X: depthwise convolutions with pipelined loads
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_encoder.py tests/test_tcn.py tests/test_dropin.py tests/test_mvdr.py -m gpu -q > $O/r02x_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02x_pytest.log
timeout 300 python scripts/dev_r02o.py 2>&1 | tail -3
for w in asr_encoder mvdr_tcn; do
timeout 300 python bench.py --workload $w --steps 20 --warmup 5 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('$w', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
