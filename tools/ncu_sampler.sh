set -x
python tools/profile_sampler.py 256 7.5 persistent > gpurun_out/plain_sampler.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:den_persist_kernel -s 3 -c 1 -o gpurun_out/r2_den_persist python tools/profile_sampler.py 256 7.5 persistent > gpurun_out/ncu_sampler.log 2>&1
tail -3 gpurun_out/ncu_sampler.log
ls -la gpurun_out/*.ncu-rep
