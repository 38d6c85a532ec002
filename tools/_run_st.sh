timeout 600 python -m pytest tests/test_gpu_c3_shapes.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/st_test.log 2>&1; echo test_rc=$?
tail -15 gpurun_out/st_test.log
B="timeout 120 python tools/bench_conv.py --cases res,res_b2,c32_64 --ops gather,scatter --impls tc --iters 20"
$B > gpurun_out/st_on.jsonl 2>&1
CGAN3D_NO_STACK=1 $B > gpurun_out/st_off.jsonl 2>&1
grep -h '"ms"' gpurun_out/st_on.jsonl gpurun_out/st_off.jsonl | cut -c1-150
