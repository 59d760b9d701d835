for c in 16 17; do echo "== precompute c=$c"; python tools/msm_once.py 20 3 $c | tail -2; done
ZKP_B200_TRACE=1 python tools/msm_once.py 20 2 17 2>&1 | tail -26
for lg in 16 18 22; do echo "== 2^$lg plain"; python tools/msm_once.py $lg 3 0 | tail -1; done
echo "== 2^18 c=15"; python tools/msm_once.py 18 3 15 | tail -1
echo "== 2^22 c=19"; python tools/msm_once.py 22 3 19 | tail -1
echo "== 2^22 c=18"; python tools/msm_once.py 22 3 18 | tail -1
