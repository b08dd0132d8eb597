# round 2, visit R: TMA-fed convolution (MODE 4) — full suite, all workloads
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02r_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -12 $O/r02r_pytest_all.log
for w in asr_encoder encoder mvdr_tcn dccrn; do
  timeout 400 python bench.py --workload $w --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02r_bench_$w.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02r_bench_$w.json'));print('$w', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['roofline']['frac'])"
done
APS_B200_NO_CONV_TMA=1 timeout 400 python bench.py --workload dccrn --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02r_bench_dccrn_gather.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02r_bench_dccrn_gather.json'));print('dccrn gather', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
