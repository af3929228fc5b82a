for v in "ITN_GEMM_RZ_COMP=0.5" "ITN_GEMM_RZ_COMP=0.7" "ITN_GEMM_SPLITACC=1 ITN_GEMM_RZ_COMP=0.5" "ITN_GEMM_SPLITACC=1 ITN_GEMM_RZ_COMP=0.7" "ITN_GEMM_SPLITACC=1 ITN_GEMM_RZ_COMP=0.3"; do
  echo "=== $v"
  env $v timeout 300 python tools/meta_parity.py 2>&1 | grep -A5 "group  " | grep -v "worst\|group"
done
