# round 2, visit P: scalar-epilogue fix in the second epilogue body, NaN-poisoned test outputs, staged thin conv
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r02p_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -15 $O/r02p_pytest_all.log
timeout 300 python scripts/dev_r02o.py > $O/r02p_small_microbench.txt 2>&1; cat $O/r02p_small_microbench.txt
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 > $O/r02p_bench_asr_encoder.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02p_bench_asr_encoder.json'));print('asr_encoder', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
for w in mvdr_tcn dccrn; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --cpu-seconds 0.2 > $O/r02p_bench_$w.json 2>/dev/null; python -c "import json;d=json.load(open('$O/r02p_bench_$w.json'));print('$w', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
timeout 200 python scripts/dev_conv_ab.py
ls -la $O | grep r02p
