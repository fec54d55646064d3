# dev: GPU tests + upsample probe timings + bench line
set -x
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/t_$TAG.log 2>&1; tail -3 gpurun_out/t_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"upcat" --csv --log-file gpurun_out/up_$TAG.csv python tools/new_kernels_probe.py > /dev/null 2>&1
grep upcat gpurun_out/up_$TAG.csv | awk -F'","' '{print $5, $9, $15}' | cut -c1-140
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.log 2> gpurun_out/bench_$TAG.err; tail -c 300 gpurun_out/bench_$TAG.log
