# dev: GPU test suite + default bench line (tools/run_check.sh <tag>)
set -x
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/t_$TAG.log 2>&1; tail -3 gpurun_out/t_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; tail -c 600 gpurun_out/bench_$TAG.log
