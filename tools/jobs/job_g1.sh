python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or grid" 2>&1 | tail -5 > gpurun_out/r02_gputests_g.log
python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/g_c4.json 2> gpurun_out/g_c4.err
tail -3 gpurun_out/r02_gputests_g.log
