# round 2, visit Q: TMA-fed stride-2 convolution (MODE 4)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_encoder.py -m gpu -q -x -k "tma_fed_conv" > $O/r02q_pytest_conv.log 2>&1; echo "pytest rc=$?"; tail -25 $O/r02q_pytest_conv.log
timeout 200 python scripts/dev_conv_ab.py; APS_B200_NO_CONV_TMA=1 timeout 200 python scripts/dev_conv_ab.py
timeout 600 python -m pytest tests/test_encoder.py tests/test_dropin.py tests/test_dccrn.py -m gpu -q > $O/r02q_pytest_enc.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02q_pytest_enc.log
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02q_bench_asr_encoder.json 2>$O/r02q_bench.err; tail -3 $O/r02q_bench.err; python -c "import json;d=json.load(open('$O/r02q_bench_asr_encoder.json'));print('asr_encoder', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
