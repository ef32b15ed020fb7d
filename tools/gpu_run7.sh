set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_images_gpu.py -m gpu -q -x 2>&1 | tail -5 | cut -c1-300
python tools/bench_hbm.py --json gpurun_out/hbm.json 2>&1 | tail -20
