mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -n 2 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -n 2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -n 2 gpurun_out/bench.err; cut -c1-600 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench_ref.json
MB=128 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"zb_chase|zb_scan_blocks_big|zb_scan_frames_big" -c 8 -o gpurun_out/r2_chase -f python tools/gpu_prof_c5.py > gpurun_out/ncu_chase.log 2>&1; tail -n 2 gpurun_out/ncu_chase.log
MB=128 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_large_frame_launches.csv python tools/gpu_prof_c5.py > /dev/null 2>&1; tail -n 2 gpurun_out/r2_large_frame_launches.csv | cut -c1-300
