# round 2, call 9: fused patch-embed kernel — checks, parity, bench A/B (fused vs three-kernel form)
set -x
mkdir -p gpurun_out
timeout 300 python tools/gpu_diag.py patch_embed embed_init patch_gather > gpurun_out/r02i_diag.txt 2>&1
tail -12 gpurun_out/r02i_diag.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/r02i_pytest.txt
for v in 1 0; do
TCOW_FUSE_PATCH=$v timeout 600 python bench.py --no-cpu-baseline --no-eager-baseline --no-train > gpurun_out/r02i_bench_$v.json 2> gpurun_out/r02i_bench_$v.err
tail -3 gpurun_out/r02i_bench_$v.err
python -c "
import json;d=json.loads(open('gpurun_out/r02i_bench_$v.json').read().strip().splitlines()[-1])
print('FUSE_PATCH=$v', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'])
for k,v in d['breakdown'].items():
    if k in ('patch_embed','patch_gather','gemm_patch','embed_init','ln','attn_spatial'): print(k, v)
"
done
