# Spatial attention on ONE box: kernel checks, then isolated timings (B=8, T=30, N=300) — run it for two builds
# (TCOW_B200_LIB, see the Makefile's BUILD / LIB / EXTRA) to compare kernels.
set -x
mkdir -p gpurun_out
T=${1:-spatial}
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attn_spatial" 2>&1 | tail -3 | tee gpurun_out/${T}_ab.txt
for i in 1 2; do python tools/prof_op.py spatial 20 100; done 2>&1 | tee -a gpurun_out/${T}_ab.txt
