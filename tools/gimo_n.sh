N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 500)) bench.py --gpus $N --config gimo --steps 16 --warmup 3 > gpurun_out/r2b_gimo_n$N.json 2> gpurun_out/r2b_gimo_n$N.err
tail -n 1 gpurun_out/r2b_gimo_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])"
