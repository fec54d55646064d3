set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench29.log 2> gpurun_out/bench29.err; tail -c 1500 gpurun_out/bench29.log
python tools/new_kernels_probe.py && ncu --set full --clock-control none --import-source on -k regex:"upcat|skinny|linear_act" -c 12 -o gpurun_out/prof_new_r1 -f python tools/new_kernels_probe.py > gpurun_out/ncu_new.log 2>&1; tail -3 gpurun_out/ncu_new.log
DM_BENCH_GRAPH=0 DM_BENCH_FAST=1 python bench.py --steps 1 --warmup 3 > /dev/null 2>&1 && DM_BENCH_GRAPH=0 DM_BENCH_FAST=1 ncu --metrics gpu__time_duration.sum --clock-control none -s 5200 -c 3000 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 1 --warmup 3 > gpurun_out/ncu_r1f.log 2>&1; tail -2 gpurun_out/ncu_r1f.log; wc -l gpurun_out/launches_r1f.csv
