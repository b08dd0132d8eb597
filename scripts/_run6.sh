set -x
mkdir -p gpurun_out
true
timeout 900 python -m pytest tests/test_encoder.py -m gpu -q -x > gpurun_out/pytest_enc.log 2>&1; tail -15 gpurun_out/pytest_enc.log
for eng in simt tc; do
APS_B200_GEMM=$eng timeout 300 python bench.py --workload encoder --steps 10 --warmup 4 > gpurun_out/bench_encoder_$eng.log 2>&1; python -c "
import json
d=json.loads(open('gpurun_out/bench_encoder_$eng.log').read().strip().splitlines()[-1]); print('ENC $eng', round(d['ms_per_step'],3),'ms', round(d['roofline']['achieved'],1),'TF/s', round(d['value']/1e6,2),'M frames/s')"
done
APS_B200_GEMM=tc timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 420 -c 280 --csv --log-file gpurun_out/launches_encoder_tc.csv python bench.py --workload encoder --steps 2 --warmup 2 > gpurun_out/ncu_enc_tc.log 2>&1
