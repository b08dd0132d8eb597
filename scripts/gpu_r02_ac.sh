# round 2, visit AC: LSTM recurrence on the tensor-core engine (grouped MODE 5 GEMM + cell kernel)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_lstm.py tests/test_dccrn.py -m gpu -q -x > $O/r02ac_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/r02ac_pytest.log
timeout 400 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('dccrn tc-lstm', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])"
APS_B200_LSTM=simt timeout 400 python bench.py --workload dccrn --steps 10 --warmup 3 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('dccrn simt-lstm', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('asr_encoder', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
