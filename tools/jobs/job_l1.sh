python tools/jobs/ws_trace.py 1250000 > gpurun_out/l_trace.txt 2>&1
python tools/jobs/ws_trace.py 10000000 >> gpurun_out/l_trace.txt 2>&1
BMM_SWEEP_EVENTS=1 python tools/jobs/ws_trace.py 1250000 >> gpurun_out/l_trace.txt 2>&1
python bench.py --workload c4 --n 1250000 --steps 3 --warmup 3 --no-cpu > gpurun_out/l_c4_n125.json 2>/dev/null
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/l_c2_widen.json 2>/dev/null
BMM_FETCH_WIDEN=0 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/l_c2_nowiden.json 2>/dev/null
cat gpurun_out/l_trace.txt
