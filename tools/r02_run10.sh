# round 2, call 10: weight-gradient GEMM with the bias gradient fused in — checks, gradient parity, training bench
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag_train.py wgrad > gpurun_out/r02j_diag.txt 2>&1
tail -20 gpurun_out/r02j_diag.txt
timeout 900 python -m pytest tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/r02j_pytest.txt
timeout 600 python bench.py --workload train --steps 10 --warmup 3 > gpurun_out/r02j_train.json 2> gpurun_out/r02j_train.err
tail -3 gpurun_out/r02j_train.err
python -c "
import json;d=json.loads(open('gpurun_out/r02j_train.json').read().strip().splitlines()[-1])
print('train', d['value'], d['ms_per_step'])
for k,v in list(d['breakdown'].items())[:14]: print(k, v)
"
