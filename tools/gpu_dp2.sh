set -x
mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -12
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x -k "dp_momentum or conv3x3_fwd or head" 2>&1 | tail -3
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/test_dp.py --flagship --steps 10 > gpurun_out/r2_dp_n2.txt 2>&1
tail -40 gpurun_out/r2_dp_n2.txt
