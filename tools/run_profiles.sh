# dev: round-end evidence run (tools/run_profiles.sh <tag>): GPU suite, bench line, reference arm, ncu captures, eager launch list
set -x
TAG=${1:-r02}
python -m pytest tests -m gpu -q -s > gpurun_out/t_$TAG.log 2>&1; tail -3 gpurun_out/t_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_$TAG.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2> gpurun_out/bench_ref_$TAG.err; tail -c 300 gpurun_out/bench_ref_$TAG.log
python tools/gemm_probe.py && ncu --set full --clock-control none --import-source on -k regex:"conv3x3|wgrad|bn_" -s 11 -c 11 -o gpurun_out/prof_gemm_$TAG -f python tools/gemm_probe.py > gpurun_out/ncu_gemm_$TAG.log 2>&1; tail -3 gpurun_out/ncu_gemm_$TAG.log
if [ "$2" = "launches" ]; then
DM_BENCH_GRAPH=0 DM_BENCH_FAST=1 python bench.py --steps 1 --warmup 3 > /dev/null 2>&1 && DM_BENCH_GRAPH=0 DM_BENCH_FAST=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 5200 -c 3000 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_$TAG.log 2>&1; tail -2 gpurun_out/ncu_$TAG.log; wc -l gpurun_out/launches_$TAG.csv
fi
