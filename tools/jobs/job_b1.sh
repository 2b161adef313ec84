# scratch GPU job (1 GPU): GPU tests + C4 bench variants
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_gputests_b.log
B="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu"
$B > gpurun_out/b_c4_1gpu.json 2> gpurun_out/b_c4_1gpu.err
BMM_GRAPH=0 $B > gpurun_out/b_c4_1gpu_nograph.json 2>/dev/null
BMM_SWEEP_EVENTS=0 $B > gpurun_out/b_c4_1gpu_noev.json 2>/dev/null
BMM_SWEEP_EVENTS=1 $B > gpurun_out/b_c4_1gpu_ev1.json 2>/dev/null
tail -5 gpurun_out/r02_gputests_b.log
