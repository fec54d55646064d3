set -x
TAG=${1:-x}
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "upcat" > gpurun_out/t_$TAG.log 2>&1; tail -3 gpurun_out/t_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_$TAG.log; tail -3 gpurun_out/bench_$TAG.err
