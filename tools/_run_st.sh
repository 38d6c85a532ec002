for v in "" "CGAN3D_BNRED_U8=1" "CGAN3D_RED_BLOCKS=2" "CGAN3D_RED_BLOCKS=2 CGAN3D_BNRED_U8=1" "CGAN3D_RED_BLOCKS=8" "CGAN3D_RED_BLOCKS=6 CGAN3D_BNRED_U8=1"; do
echo "== $v"; env $v timeout 200 python tools/bench_ew.py 2>&1 | grep "bn_bwd_reduce\|bn_stats" | cut -c1-120
done
