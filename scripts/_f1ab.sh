set -x
mkdir -p gpurun_out
for v in "" _old; do
  APS_B200_LIB=$PWD/aps_b200/libaps_b200$v.so timeout 200 python scripts/dev_f1.py 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_transform_gpu.py -m gpu -q -x 2>&1 | tail -3
