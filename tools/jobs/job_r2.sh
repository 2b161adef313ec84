timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_pass or relabel or rerun or summaries or assign or grid" 2>&1 | tail -30 > gpurun_out/r02_gputests_r2.log
tail -5 gpurun_out/r02_gputests_r2.log
timeout 300 python bench.py --workload c4relabel --steps 3 --warmup 3 --no-cpu > gpurun_out/r2_c4relabel.json 2> gpurun_out/r2_c4relabel.err
timeout 300 python bench.py --workload c5 --steps 2 --warmup 2 --no-cpu > gpurun_out/r2_c5.json 2> gpurun_out/r2_c5.err
python tools/showbench.py gpurun_out/r2_c4relabel.json gpurun_out/r2_c5.json; tail -3 gpurun_out/r2_c4relabel.err
A="--workload c4relabel --nsamples 12 --steps 1 --warmup 3 --no-cpu"
BMM_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c4relabel_v2.csv python bench.py $A > gpurun_out/ncu_wsr2.log 2>&1
