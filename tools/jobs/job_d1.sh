python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_gputests_d.log
B="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu"
$B > gpurun_out/d_c4_1gpu.json 2> gpurun_out/d_c4_1gpu.err
BMM_PDL=0 $B > gpurun_out/d_c4_1gpu_nopdl.json 2>/dev/null
BMM_SWEEP_EVENTS=0 $B > gpurun_out/d_c4_1gpu_noev.json 2>/dev/null
tail -5 gpurun_out/r02_gputests_d.log
