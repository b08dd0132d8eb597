set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
for w in encoder mvdr_tcn; do
timeout 300 python bench.py --workload $w --steps 10 --warmup 4 > gpurun_out/bench_$w.log 2>&1; python -c "
import json
d=json.loads(open('gpurun_out/bench_$w.log').read().strip().splitlines()[-1]); print('WL $w', round(d['ms_per_step'],3),'ms', round(d['roofline']['achieved'],1), d['roofline']['unit'], round(d['value']/1e6,2),'M frames/s')"
done
