set -x
TAG=${1:-x}
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_$TAG.log; tail -3 gpurun_out/bench_$TAG.err
