set -x
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --no-hbm --no-cpu 2>gpurun_out/bench_n4.err | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['e2e']['value'], d['e2e']['predict']['value'])"
tail -2 gpurun_out/bench_n4.err | cut -c1-200
