for mb in 4 5 6; do
  cp gpurun_tmp_mb$mb.so interactive_zkp_study_b200/libzkp_b200.so
  echo "== min blocks $mb"
  ZKP_B200_TRACE=1 python tools/msm_once.py 20 3 17 2>&1 | grep -E "accumulate|msm 2" | tail -2
  python tools/msm_once.py 20 3 0 | tail -1
done
