set -x
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/test_dp.py --flagship --steps 10 --reps 1 > gpurun_out/r2_dp_n8.txt 2>&1
grep -v "^$" gpurun_out/r2_dp_n8.txt | grep -v "W1017\|^\*\*\*\|Setting OMP\|FutureWarning\|symm.enable\|UserWarning\|return func" | tail -25 | cut -c1-300
