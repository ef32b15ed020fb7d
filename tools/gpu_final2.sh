set -x
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --quick --steps 1 --warmup 3 --no-predict --no-cpu > gpurun_out/launches.log 2>&1
bash tools/gpu_final.sh
