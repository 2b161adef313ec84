python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "dp" 2>&1 | tail -8 > gpurun_out/r02_gputests_j.log
python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/j_c3.json 2> gpurun_out/j_c3.err
BMM_DP_LOGFORM=1 python bench.py --workload c3 --steps 3 --warmup 3 --no-cpu > gpurun_out/j_c3_log.json 2>/dev/null
tail -4 gpurun_out/r02_gputests_j.log
