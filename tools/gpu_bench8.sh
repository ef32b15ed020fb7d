set -x
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 20 --warmup 5 --predict-images 8 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
tail -2 gpurun_out/bench_n8.err | cut -c1-300
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n8.json'))
for k in ("value","ms_per_step","n_gpus","clocks","e2e"):
    print(k, json.dumps(d.get(k))[:900])
print("predict", json.dumps(d["predict"])[:900])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29515 bench.py --gpus 8 --large-tiles --steps 3 --warmup 3 > gpurun_out/r2_large_tiles_n8.json 2> gpurun_out/r2_large_tiles_n8.err
tail -2 gpurun_out/r2_large_tiles_n8.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/r2_large_tiles_n8.json'))
for r in d['large_tiles']['train']: print(r)
for r in d['large_tiles']['predict']: print(r)
"
