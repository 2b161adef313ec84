timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_pass or relabel or rerun or summaries or thinning" 2>&1 | tail -30 > gpurun_out/r02_gputests_r4.log
tail -5 gpurun_out/r02_gputests_r4.log
B="timeout 300 python bench.py --workload c4relabel --steps 3 --warmup 3 --no-cpu"
$B > gpurun_out/r4_c4relabel.json 2> gpurun_out/r4_c4relabel.err
BMM_ASSIGN_FORK=0 $B > gpurun_out/r4_c4relabel_nofork.json 2> gpurun_out/r4_c4relabel_nofork.err
BMM_GRAPH=0 $B > gpurun_out/r4_c4relabel_nograph.json 2> gpurun_out/r4_c4relabel_nograph.err
python tools/showbench.py gpurun_out/r4_c4relabel.json gpurun_out/r4_c4relabel_nofork.json gpurun_out/r4_c4relabel_nograph.json; tail -3 gpurun_out/r4_c4relabel.err
A="--workload c4relabel --nsamples 12 --steps 1 --warmup 3 --no-cpu"
BMM_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:big_relabel_ws -s 12 -c 1 -o gpurun_out/r02_wsr_v3 -f python bench.py $A > gpurun_out/ncu_wsr.log 2>&1
