# round 2: compute-sanitizer memcheck over the kernels added / rewritten this round
set -x
mkdir -p gpurun_out
S="compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20"
timeout 1500 $S python -m pytest tests -m gpu -q -x -k "tma_fed_conv or tensor_core_conv or thin or lstm_tensor_core or lstm_matches_torch or ops_vs_torch or dccrn_0 or enc_0 or tcn_0 or register_tiled or split_k or featops or batch_decoding" > gpurun_out/r02_san_tests.log 2>&1; echo "tests rc=$?"; tail -6 gpurun_out/r02_san_tests.log
grep -h "ERROR SUMMARY" gpurun_out/r02_san_tests.log | sort | uniq -c
