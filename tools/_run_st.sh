for w in 0.02 0.05 0.1 0.2 0.4; do
CGAN3D_THIN_HALO_W=$w timeout 120 python tools/bench_conv.py --cases first_c3 --ops gather --impls tc --iters 20 2>&1 | grep '"ms"' | cut -c1-90 | sed "s/^/w=$w /"
CGAN3D_THIN_HALO_W=$w timeout 120 python tools/bench_conv.py --cases last_c3 --ops scatter --impls tc --iters 20 2>&1 | grep '"ms"' | cut -c1-90 | sed "s/^/w=$w /"
done
