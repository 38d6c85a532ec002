timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_t18.log 2>&1; echo test_rc=$?
tail -3 gpurun_out/r2_t18.log
timeout 600 python bench.py --steps 20 --warmup 5 --breakdown gpurun_out/r2_break18.txt > gpurun_out/r2_b18.json 2> gpurun_out/r2_b18.err; echo bench_rc=$?
timeout 900 python tools/bench_conv.py --sweep --impls tc,cudnn --ops gather,scatter,wgrad --iters 10 > gpurun_out/r2_c5_sweep2.jsonl 2> gpurun_out/r2_c5_2.err; echo sweep_rc=$?
timeout 300 python tools/bench_conv.py --cases res,down0_c3,down1,first_c3,last_c3,d_first,d_mid0 --impls tc,cudnn --ops gather,scatter,wgrad --iters 10 > gpurun_out/r2_vs_cudnn.jsonl 2> gpurun_out/r2_vs_cudnn.err; echo vs_rc=$?
