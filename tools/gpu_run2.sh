set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_pipeline_gpu.py -m gpu -q -s -k "teacher_forced or batch4 or main_train or sharded_train or eval_every" > gpurun_out/r2_pytest_gpu_2.txt 2>&1
grep -v "^$" gpurun_out/r2_pytest_gpu_2.txt | grep -i "worst\|bottom\|passed\|failed\|error\|flagship\|agreement" | cut -c1-1500
