set -x
mkdir -p gpurun_out
L="conv_0/conv2,conv_10/conv1,conv_10/conv2,conv_dilut_0/atrous_conv2,conv_1/conv1,conv_1/conv2,conv_9/conv1,conv_9/conv2"
for V in "base" "RSU_HALO_EPI=0" "RSU_HALO_MT=1" "RSU_HALO_EPI=0 RSU_HALO_MT=1"; do
  echo "=== $V"
  if [ "$V" = "base" ]; then python tools/bench_layers.py --only "$L" --out gpurun_out/l_base.json
  else env $V python tools/bench_layers.py --only "$L" --out gpurun_out/l_x.json; fi
done > gpurun_out/r2_halo_knobs.txt 2>&1
cut -c1-260 gpurun_out/r2_halo_knobs.txt
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_kernels_gpu.py -m gpu -q -s -k "batch4 or wgrad" > gpurun_out/r2_pytest_gpu_3.txt 2>&1
grep -v "^$" gpurun_out/r2_pytest_gpu_3.txt | tail -80 | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
