python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_gputests_e.log
B="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu"
$B > gpurun_out/e_c4_1gpu.json 2> gpurun_out/e_c4_1gpu.err
python bench.py --workload c4relabel --steps 3 --warmup 3 --no-cpu > gpurun_out/e_c4rel_1gpu.json 2> gpurun_out/e_c4rel_1gpu.err
tail -5 gpurun_out/r02_gputests_e.log
