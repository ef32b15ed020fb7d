set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.txt; cat gpurun_out/r2_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee -a gpurun_out/r2_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 --dump-layers gpurun_out/r2_layers.json > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -2 gpurun_out/r2_bench_n1.err | cut -c1-300
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null; cut -c1-400 gpurun_out/r2_bench_ref.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
for k in ("value","ms_per_step","gpu_launches","clocks","roofline","roofline_wgrad","roofline_network","other_kernels_ms_per_step"):
    print(k, json.dumps(d.get(k))[:1200])
print("cpu", json.dumps(d["cpu_baseline"])[:1500])
print("e2e", json.dumps(d["e2e"])[:1200])
print("predict", json.dumps(d["predict"])[:1500])
for k,v in d["hbm_kernels"]["kernels"].items(): print("  %-40s %s"%(k,v))
PY
