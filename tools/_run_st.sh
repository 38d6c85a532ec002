timeout 200 python tools/bench_pad.py
for l in 1 2 8; do CGAN3D_PADBWD_LPB=$l timeout 200 python tools/bench_pad.py | tail -1 | sed "s/^/lpb=$l /"; done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "pad or reflect or train or generator" 2>&1 | tail -3
