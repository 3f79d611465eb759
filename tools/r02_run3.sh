# round 2, call 3: new spatial attention kernel — correctness checks, isolated A/B timing vs the round-1 kernel, bench
set -x
mkdir -p gpurun_out
python tools/gpu_diag.py attn_spatial > gpurun_out/r02c_diag.txt 2>&1
tail -25 gpurun_out/r02c_diag.txt
for i in 1 2; do
TCOW_B200_LIB=$PWD/build/libtcow_spatial_r01.so python tools/prof_op.py spatial 5
python tools/prof_op.py spatial 5
done 2>&1 | tee gpurun_out/r02c_ab.txt
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_train_gpu.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r02c_pytest.txt
python bench.py --no-cpu-baseline --no-eager-baseline --no-train > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err
tail -3 gpurun_out/r02c_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r02c_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])
for k,v in d['breakdown'].items(): print(k, v)
"
