set -x
mkdir -p gpurun_out
for v in default build/libtcow_spatial_r01.so build/libtcow_sp_exp1.so build/libtcow_sp_exp2.so build/libtcow_sp_exp3.so build/libtcow_sp_exp4.so build/libtcow_sp_exp6.so; do
  if [ $v = default ]; then python tools/prof_op.py spatial 20 100; else TCOW_B200_LIB=$PWD/$v python tools/prof_op.py spatial 20 100; fi
done 2>&1 | grep "avg us" | tee gpurun_out/r02d_exp.txt
ncu --set full --clock-control none --import-source on -k regex:attn_spatial_tc -s 3 -c 1 -o gpurun_out/r02d_sp python tools/prof_op.py spatial 5 > gpurun_out/r02d_ncu.log 2>&1
ls -la gpurun_out/r02d_sp.ncu-rep
