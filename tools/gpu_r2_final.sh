mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
ZB_BENCH_COMPRESS_SEGMENTS=1024 ZB_BENCH_DICT_RECORDS=32768 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/bench_under_ncu.log 2>&1
tail -n 3 gpurun_out/r2_bench_launches.csv
