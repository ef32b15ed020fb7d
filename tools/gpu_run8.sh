set -x
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x 2>&1 | tail -3
python tools/bench_upconv.py 2>&1 | tail -5 | cut -c1-250
python bench.py --quick --steps 10 --warmup 3 --no-predict --no-cpu 2>/dev/null | cut -c1-300
