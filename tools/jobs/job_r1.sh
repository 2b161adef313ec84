timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "single_pass or relabel or rerun or summaries" 2>&1 | tail -30 > gpurun_out/r02_gputests_r1.log
tail -30 gpurun_out/r02_gputests_r1.log
timeout 300 python bench.py --workload c4relabel --steps 3 --warmup 3 --no-cpu > gpurun_out/r1_c4relabel.json 2> gpurun_out/r1_c4relabel.err
BMM_RELABEL_FUSED=0 timeout 300 python bench.py --workload c4relabel --steps 3 --warmup 3 --no-cpu > gpurun_out/r1_c4relabel_unfused.json 2> gpurun_out/r1_c4relabel_unfused.err
python tools/showbench.py gpurun_out/r1_c4relabel.json gpurun_out/r1_c4relabel_unfused.json; tail -3 gpurun_out/r1_c4relabel.err
