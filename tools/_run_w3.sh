B="timeout 120 python tools/bench_conv.py --cases first_c3 --ops wgrad --impls tc --iters 3"
for d in 16 23 17 18; do CGAN3D_W2_DEBUG=$d $B 2>&1 | grep -v "^{" | tail -8 > gpurun_out/w3_prof$d.txt; done
