# A/B of the spatial-attention softmax variants on ONE box: kernel checks, then isolated timings (TCOW_SPATIAL_IMPL = 0 two-pass,
# 1 five chunks resident, 2 four resident + one re-read).
set -x
mkdir -p gpurun_out
T=${1:-r02q}
for v in ${VARIANTS:-0 1 2}; do
  TCOW_SPATIAL_IMPL=$v timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attn_spatial" 2>&1 | tail -3
  for i in 1 2; do TCOW_SPATIAL_IMPL=$v python tools/prof_op.py spatial 20 100; done
done 2>&1 | tee gpurun_out/${T}_ab.txt
