for L in base spin; do
  if [ $L = base ]; then unset TCOW_B200_LIB; else export TCOW_B200_LIB=$PWD/tcow_b200/libtcow_b200_spin.so; fi
  echo "== $L"; timeout 600 python tools/train_bench.py --profile 2>&1 | grep -E "train step|attn_spatial" | cut -c1-100
done
