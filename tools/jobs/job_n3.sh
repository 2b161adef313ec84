BMM_SWEEP_EVENTS=0 python tools/jobs/ws_trace.py 1250000 > gpurun_out/n3_trace.txt 2>&1
BMM_SWEEP_EVENTS=0 BMM_PDL=0 python tools/jobs/ws_trace.py 1250000 >> gpurun_out/n3_trace.txt 2>&1
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/n3_c4.json 2> gpurun_out/n3_c4.err
python bench.py --workload c4 --n 1250000 --steps 3 --warmup 3 --no-cpu > gpurun_out/n3_c4_n125.json 2>/dev/null
