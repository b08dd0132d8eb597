# usage: bash scripts/gpu_profile.sh <tag> [kernel-regex]
# launch list + one full ncu capture of the dominant kernel of bench.py (1 GPU; never multi-rank)
set -x
TAG=${1:-r1}
KRE=${2:-frontend_kernel}
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 4 -c 40 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 8 --warmup 3 --cpu-seconds 0.2 > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KRE} -s 3 -c 2 \
    -o gpurun_out/prof_${TAG} -f python bench.py --steps 5 --warmup 3 --cpu-seconds 0.2 >> gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ls -la gpurun_out
