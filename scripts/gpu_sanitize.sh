# compute-sanitizer memcheck over one small invocation of every kernel family
set -x
mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20"
timeout 900 $S python __graft_entry__.py --smoke > gpurun_out/san_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/san_smoke.log
timeout 1500 $S python -m pytest tests -m gpu -q -x -k "enc_0 or mvdr_0 or tcn_0 or dccrn_0 or asr_grid_0 or stft_0 or istft_0 or enh_1 or tensor_core_gemm or tensor_core_conv" > gpurun_out/san_tests.log 2>&1; echo "tests rc=$?"; tail -8 gpurun_out/san_tests.log
grep -c "ERROR SUMMARY" gpurun_out/san_*.log; grep -h "ERROR SUMMARY" gpurun_out/san_*.log | sort | uniq -c
