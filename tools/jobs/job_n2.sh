BMM_SWEEP_EVENTS=0 BMM_PDL=0 python tools/jobs/ws_trace.py 1250000 > gpurun_out/n2_trace.txt 2>&1
BMM_SWEEP_EVENTS=0 BMM_PDL=0 python tools/jobs/ws_trace.py 10000000 >> gpurun_out/n2_trace.txt 2>&1
grep -v "^N=" gpurun_out/n2_trace.txt
