for grp in wgrad gemm_ ln_ colsum embed mask_head attn_temporal attn_spatial; do
  timeout 300 python tools/gpu_diag_train.py $grp 2>&1 | tail -40
done > gpurun_out/diag_train1.txt 2>&1
tail -5 gpurun_out/diag_train1.txt
