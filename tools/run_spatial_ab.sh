# A/B of the spatial-attention kernels on ONE box: kernel checks, then isolated timings (TCOW_SPATIAL_IMPL: p = ping-pong two
# threads per row, unset = one thread per row with register-resident scores).
set -x
mkdir -p gpurun_out
T=${1:-r02q}
for v in ${VARIANTS:-p r}; do
  TCOW_SPATIAL_IMPL=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attn_spatial" 2>&1 | tail -3
  for i in 1 2; do TCOW_SPATIAL_IMPL=$v python tools/prof_op.py spatial 20 100; done
done 2>&1 | tee gpurun_out/${T}_ab.txt
