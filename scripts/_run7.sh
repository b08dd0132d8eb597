set -x
mkdir -p gpurun_out
APS_B200_GEMM=tc APS_B200_GRAPHS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gemm_kernel -s 12 -c 4 -o gpurun_out/prof_tc_r1 -f python bench.py --workload encoder --steps 1 --warmup 1 > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
