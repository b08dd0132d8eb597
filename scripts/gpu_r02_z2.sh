# round 2: ncu --set full of the TMA-fed convolution kernel (evidence for profiles/)
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_gemm_kernel<\(int\)256, \(int\)4' -s 6 -c 2 -f -o $O/r02z_prof_conv_tma \
    python bench.py --workload encoder --steps 1 --warmup 3 --cpu-seconds 0.1 > $O/r02z_ncu_conv_tma.log 2>&1
tail -c 300 $O/r02z_ncu_conv_tma.log; ls -la $O | grep r02z
