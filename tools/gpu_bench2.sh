set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -3 gpurun_out/bench_n2.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2.json'))
for k in ("value","ms_per_step","n_gpus","gpu_launches","clocks","roofline","e2e"):
    print(k, json.dumps(d.get(k))[:700])
print("predict", json.dumps(d["predict"])[:600])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>/dev/null | cut -c1-600
