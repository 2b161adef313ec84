python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or grid or rerun" 2>&1 | tail -5 > gpurun_out/r02_gputests_n5.log
BMM_SWEEP_EVENTS=0 python tools/jobs/ws_trace.py 1250000 > gpurun_out/n5_trace.txt 2>&1
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/n5_c4.json 2> gpurun_out/n5_c4.err
python bench.py --workload c4 --n 1250000 --steps 3 --warmup 3 --no-cpu > gpurun_out/n5_c4_n125.json 2>/dev/null
tail -2 gpurun_out/r02_gputests_n5.log
