set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm2 --launch-skip 2 -c 1 -o gpurun_out/up4_fwd -f python tools/bench_upconv.py > gpurun_out/up4_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_gemm --launch-skip 8 -c 1 -o gpurun_out/up4_dgrad -f python tools/bench_upconv.py > gpurun_out/up4_dgrad.log 2>&1
tail -3 gpurun_out/up4_dgrad.log
ls -la gpurun_out/*.ncu-rep
