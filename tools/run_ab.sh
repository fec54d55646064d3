# dev: A/B of two library builds on one box (sampling per-entry profile)
set -x
python tools/sample_profile.py > gpurun_out/sample_new.log 2>&1; head -8 gpurun_out/sample_new.log
DM_B200_LIB=$PWD/tools/ab/libdm_head.so python tools/sample_profile.py > gpurun_out/sample_head.log 2>&1; head -8 gpurun_out/sample_head.log
python tools/sample_profile.py > gpurun_out/sample_new2.log 2>&1; head -8 gpurun_out/sample_new2.log
