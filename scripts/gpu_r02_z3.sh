# round 2: ncu --set full of the grouped recurrent GEMM (MODE 5), the LSTM cell kernel and a TMA-fed transposed convolution
set -x
mkdir -p gpurun_out
O=gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:tc_gemm_kernel<\(int\)128, \(int\)5|lstm_cell_kernel|tc_gemm_kernel<\(int\)64, \(int\)4' -s 40 -c 5 -f -o $O/r02z_prof_dccrn \
    python scripts/dev_prof_step.py dccrn 32 > $O/r02z_ncu_dccrn.log 2>&1
tail -c 200 $O/r02z_ncu_dccrn.log; ls -la $O | grep r02z_prof_dccrn
