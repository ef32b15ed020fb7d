set -x
mkdir -p gpurun_out
timeout 120 tools/ws_probe > gpurun_out/r2_ws_probe.txt 2>&1
cat gpurun_out/r2_ws_probe.txt
RSU_HALO_WS=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "conv3x3" 2>&1 | tail -5
L="conv_0/conv2,conv_10/conv1,conv_10/conv2,conv_dilut_0/atrous_conv2,conv_1/conv1,conv_1/conv2,conv_9/conv1,conv_9/conv2"
for V in "RSU_HALO_WS=0" "RSU_HALO_WS=1"; do
  echo "=== $V"
  env $V python tools/bench_layers.py --only "$L" --out gpurun_out/l_x.json
done > gpurun_out/r2_halo_ws.txt 2>&1
cut -c1-260 gpurun_out/r2_halo_ws.txt
