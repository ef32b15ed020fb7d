set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --dump-layers gpurun_out/layers.json > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
tail -3 gpurun_out/bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
for k in ("value","ms_per_step","gpu_launches","clocks","roofline","roofline_wgrad","roofline_network","other_kernels_ms_per_step","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:900])
print("e2e", json.dumps(d["e2e"])[:1200])
print("predict", json.dumps(d["predict"])[:1500])
for k,v in d["hbm_kernels"]["kernels"].items(): print("  %-40s %s"%(k,v))
print(json.dumps(d.get("launches_per_step")))
PY
