set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:attn_spatial_pp -s 3 -c 1 -o gpurun_out/r02f_pp python tools/prof_op.py spatial 5 > gpurun_out/r02f_ncu.log 2>&1
ls -la gpurun_out/r02f_pp.ncu-rep
