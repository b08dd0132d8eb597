# round 2, visit Y: programmatic dependent launch re-measured on the final kernels
set -x
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2; do
  timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('default', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
  APS_B200_PDL=1 timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('pdl', d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"
done
APS_B200_PDL=1 timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
