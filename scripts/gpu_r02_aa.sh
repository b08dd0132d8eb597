# round 2, visit AA: transposed convolutions on the TMA-fed kernel (MODE 4)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_encoder.py tests/test_dccrn.py tests/test_lstm.py -m gpu -q -x > $O/r02aa_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02aa_pytest.log
timeout 400 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('dccrn tma', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])"
APS_B200_NO_CONV_TMA=1 timeout 400 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('dccrn gather', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
