set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_${1:-x}.log 2>&1; tail -15 gpurun_out/t_${1:-x}.log
