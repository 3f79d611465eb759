set -x
mkdir -p gpurun_out
for i in 1 2; do
for v in default build/libtcow_pp_exp1.so build/libtcow_pp_exp2.so; do
  if [ $v = default ]; then timeout 120 python tools/prof_op.py spatial 20 100; else TCOW_B200_LIB=$PWD/$v timeout 120 python tools/prof_op.py spatial 20 100; fi
done; done 2>&1 | grep "avg us" | tee gpurun_out/r02g_exp.txt
