timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_pass or many_tiles or relabel or rerun" 2>&1 | tail -30 > gpurun_out/r02_gputests_r5.log
tail -5 gpurun_out/r02_gputests_r5.log
grep -q passed gpurun_out/r02_gputests_r5.log || exit 1
grep -q failed gpurun_out/r02_gputests_r5.log && exit 1
B="timeout 60 python bench.py --workload c4relabel --steps 3 --warmup 3 --no-cpu"
$B > gpurun_out/r5_c4relabel.json 2> gpurun_out/r5_c4relabel.err || { tail -n 3 gpurun_out/r5_c4relabel.err; exit 1; }
BMM_ASSIGN_FORK=0 $B > gpurun_out/r5_c4relabel_nofork.json 2> gpurun_out/r5_c4relabel_nofork.err
BMM_LIB=$PWD/bmm_mcmc_b200/libbmm_b200_nepi3.so $B > gpurun_out/r5_c4relabel_nepi3.json 2> gpurun_out/r5_c4relabel_nepi3.err
python tools/showbench.py gpurun_out/r5_c4relabel.json gpurun_out/r5_c4relabel_nofork.json gpurun_out/r5_c4relabel_nepi3.json
A="--workload c4relabel --nsamples 12 --steps 1 --warmup 3 --no-cpu"
BMM_GRAPH=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:big_relabel_ws -s 12 -c 1 -o gpurun_out/r02_wsr_v3 -f python bench.py $A > gpurun_out/ncu_wsr.log 2>&1
