# dev: GPU tests + sampling per-entry profile + CoordAttn probe timings
set -x
TAG=${1:-x}
python -m pytest tests -m gpu -x -q > gpurun_out/t_$TAG.log 2>&1; tail -3 gpurun_out/t_$TAG.log
python tools/sample_profile.py > gpurun_out/sample_profile_$TAG.log 2>&1; head -40 gpurun_out/sample_profile_$TAG.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"ca_wgrad" --csv --log-file gpurun_out/ca_$TAG.csv python tools/ca_probe.py 1536 16 4 > /dev/null 2>&1; tail -4 gpurun_out/ca_$TAG.csv | awk -F'","' '{print $5, $9, $15}' | cut -c1-120
