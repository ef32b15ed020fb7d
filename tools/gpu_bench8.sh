set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --predict-images 8 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
tail -2 gpurun_out/r2_bench_n8.err | cut -c1-300
python bench.py --gpus 1 --quick --steps 20 --warmup 5 --no-predict --no-cpu 2>/dev/null | cut -c1-260
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
for k in ("value","ms_per_step","n_gpus","clocks"):
    print(k, json.dumps(d.get(k))[:400])
print("e2e", json.dumps(d["e2e"])[:1200])
PY
