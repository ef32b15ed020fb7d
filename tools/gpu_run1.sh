set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
RSU_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "wgrad_cta_pair" 2>&1 | tail -15 > gpurun_out/r2_wgrad_pair_test.txt
cat gpurun_out/r2_wgrad_pair_test.txt
timeout 1500 python -m pytest tests -m gpu -q -s --durations=15 2>&1 | tail -150 > gpurun_out/r2_pytest_gpu_1.txt
tail -60 gpurun_out/r2_pytest_gpu_1.txt
if grep -q passed gpurun_out/r2_wgrad_pair_test.txt && ! grep -q failed gpurun_out/r2_wgrad_pair_test.txt; then
  timeout 300 python tools/bench_wgrad_pair.py > gpurun_out/r2_wgrad_pair_ab.txt 2>&1
  cat gpurun_out/r2_wgrad_pair_ab.txt
fi
