set -x
python tools/single_step.py > gpurun_out/plain_step.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_launches_single_step.csv python tools/single_step.py > gpurun_out/ncu_step.log 2>&1
tail -2 gpurun_out/plain_step.log gpurun_out/ncu_step.log
wc -l gpurun_out/r2_launches_single_step.csv
