# usage: tools/ncu_sampler.sh <backend> <kernel regex> <out name>
set -x
BE=${1:-persistent}; RX=${2:-den_persist_kernel}; OUT=${3:-r2_den_persist}
python tools/profile_sampler.py 256 7.5 $BE > gpurun_out/plain_sampler.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:$RX -s 3 -c 1 -o gpurun_out/$OUT python tools/profile_sampler.py 256 7.5 $BE > gpurun_out/ncu_sampler.log 2>&1
ls -la gpurun_out/$OUT.ncu-rep
