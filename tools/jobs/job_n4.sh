BMM_SWEEP_EVENTS=0 python tools/jobs/ws_trace.py 1250000 > gpurun_out/n4_trace.txt 2>&1
