set +e
cd $GRAFT_REPO_ROOT
timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/final_pytest.txt
timeout 200 python bench.py > gpurun_out/final_bench_fbank.json 2> gpurun_out/final_bench_fbank.err; tail -c 600 gpurun_out/final_bench_fbank.json
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3 | tee gpurun_out/final_smoke.txt
timeout 200 python bench.py --workload dccrn --cpu-seconds 2 > gpurun_out/final_bench_dccrn.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_dccrn.json
timeout 200 python bench.py --workload encoder --cpu-seconds 2 > gpurun_out/final_bench_encoder.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_encoder.json
timeout 200 python bench.py --workload mvdr_tcn --cpu-seconds 2 > gpurun_out/final_bench_mvdr_tcn.json 2>/dev/null; cut -c1-200 gpurun_out/final_bench_mvdr_tcn.json
timeout 150 compute-sanitizer --tool memcheck python -m pytest tests/test_lstm.py tests/test_encoder.py -m gpu -q -x -k "lstm_matches_torch or lstm_multi or narrow_output" > gpurun_out/final_sanitizer_new_kernels.log 2>&1; tail -3 gpurun_out/final_sanitizer_new_kernels.log
