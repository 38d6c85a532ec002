timeout 600 python -m pytest tests/test_gpu_c3_shapes.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/st_test.log 2>&1; echo test_rc=$?
tail -3 gpurun_out/st_test.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --breakdown gpurun_out/r2_break16.txt > gpurun_out/r2_b16.json 2> gpurun_out/r2_b16.err; echo bench_rc=$?
