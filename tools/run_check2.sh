# dev: GPU tests, then per-launch times (ncu, duration only) of the MLP / CoordAttn probes
set -x
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/t_$TAG.log 2>&1; tail -3 gpurun_out/t_$TAG.log
python tools/ca_probe.py 1536 16 4; python tools/ca_probe.py 768 32 4; python tools/ca_probe.py 192 128 4
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/ca_$TAG.csv python tools/ca_probe.py 1536 16 4 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"linear_act|sum_parts|skinny" --csv --log-file gpurun_out/mlp_$TAG.csv python tools/new_kernels_probe.py > /dev/null 2>&1
