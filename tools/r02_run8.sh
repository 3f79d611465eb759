# round 2, call 8: spatial attention v3 (double-buffered score blocks, single TMEM read, lazy rescale)
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py attn_spatial > gpurun_out/r02h_diag.txt 2>&1
tail -22 gpurun_out/r02h_diag.txt
for i in 1 2; do
TCOW_SPATIAL_IMPL=r timeout 120 python tools/prof_op.py spatial 20 100
timeout 120 python tools/prof_op.py spatial 20 100
done 2>&1 | grep "avg us" | tee gpurun_out/r02h_ab.txt
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_train_gpu.py tests/test_train_kernels_gpu.py -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/r02h_pytest.txt
timeout 600 python bench.py --no-cpu-baseline --no-eager-baseline --no-train > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err
tail -3 gpurun_out/r02h_bench.err
python -c "
import json;d=json.loads(open('gpurun_out/r02h_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])
for k,v in d['breakdown'].items(): print(k, v)
"
ncu --set full --clock-control none --import-source on -k regex:attn_spatial_pp -s 3 -c 1 -o gpurun_out/r02h_pp python tools/prof_op.py spatial 5 > gpurun_out/r02h_ncu.log 2>&1
