set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "halo_cta_pair" 2>&1 | tail -15 | cut -c1-300
