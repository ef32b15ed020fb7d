set -x
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 | cut -c1-300
python bench.py --quick --steps 10 --warmup 3 --no-predict --no-cpu 2>/dev/null | cut -c1-300
