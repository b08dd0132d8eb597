# round 2, visit AG: cheaper staleness check on the exposed host path
set -x
for rep in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 5 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('asr_encoder', d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'])"
done
timeout 300 python bench.py --workload encoder --steps 20 --warmup 5 --cpu-seconds 0.2 2>/dev/null | python -c "import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('encoder', d['ms_per_step'])"
timeout 300 python -m pytest tests/test_host_logic.py tests/test_encoder.py tests/test_dropin.py -m gpu -q 2>&1 | tail -2
