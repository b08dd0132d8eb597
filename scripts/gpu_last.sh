set +e
cd $GRAFT_REPO_ROOT
timeout 200 python -m pytest tests/test_encoder.py tests/test_dccrn.py -m gpu -q -x 2>&1 | tail -2
timeout 100 python scripts/dev_dccrn_layers.py 2>&1 | tail -2 | cut -c1-30,100-200
APS_B200_TCONV_NARROW_1PX=1 timeout 100 python scripts/dev_dccrn_layers.py 2>&1 | tail -2 | cut -c1-30,100-200
for i in 1 2; do
timeout 100 python bench.py --workload encoder --cpu-seconds 0.3 2>/dev/null | cut -c90-150
APS_B200_LIB=$PWD/aps_b200/libaps_b200_pwl8.so timeout 100 python bench.py --workload encoder --cpu-seconds 0.3 2>/dev/null | cut -c90-150
done
APS_B200_LIB=$PWD/aps_b200/libaps_b200_pwl8.so timeout 100 python bench.py --workload mvdr_tcn --cpu-seconds 0.3 2>/dev/null | cut -c90-150
timeout 100 python bench.py --workload dccrn --cpu-seconds 0.3 2>/dev/null | cut -c90-150
