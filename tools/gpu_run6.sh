set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x -s -k "large_tile" 2>&1 | tail -5 | cut -c1-400
timeout 900 python bench.py --large-tiles --steps 3 --warmup 3 > gpurun_out/r2_large_tiles_n1.json 2> gpurun_out/r2_large_tiles_n1.err
tail -5 gpurun_out/r2_large_tiles_n1.err | cut -c1-300
python -c "
import json
d=json.load(open('gpurun_out/r2_large_tiles_n1.json'))
for r in d['large_tiles']['train']: print(r)
for r in d['large_tiles']['predict']: print(r)
"
