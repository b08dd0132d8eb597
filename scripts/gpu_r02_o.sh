# round 2, visit O: 48-tile attention, split-K LayerNorm preload, fast sigmoid in the GEMM epilogue (variant build), write roofline of the thin conv
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_encoder.py tests/test_dropin.py -m gpu -q -x > $O/r02o_pytest_enc.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02o_pytest_enc.log
APS_B200_LIB=$PWD/aps_b200/libaps_b200_ftz.so timeout 600 python -m pytest tests/test_encoder.py tests/test_tcn.py tests/test_dccrn.py -m gpu -q -x > $O/r02o_pytest_ftz.log 2>&1; echo "pytest ftz rc=$?"; tail -3 $O/r02o_pytest_ftz.log
timeout 300 python scripts/dev_r02o.py > $O/r02o_small_microbench.txt 2>&1; cat $O/r02o_small_microbench.txt
for rep in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02o_bench_def_$rep.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02o_bench_def_$rep.json'));print('default', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
  APS_B200_LIB=$PWD/aps_b200/libaps_b200_ftz.so timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02o_bench_ftz_$rep.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02o_bench_ftz_$rep.json'));print('ftz', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
timeout 200 python scripts/dev_conv_ab.py; APS_B200_LIB=$PWD/aps_b200/libaps_b200_ftz.so timeout 200 python scripts/dev_conv_ab.py
timeout 300 python scripts/dev_dccrn_ab.py aps_b200/libaps_b200.so aps_b200/libaps_b200_ftz.so
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'conv2d_thin3x3_kernel' -s 2 -c 1 -f -o $O/r02o_prof_thin \
    python bench.py --steps 1 --warmup 3 --cpu-seconds 0.1 > $O/r02o_ncu_thin.log 2>&1
ls -la $O | grep r02o
