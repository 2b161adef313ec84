python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fetch or summaries or thinning or rerun" 2>&1 | tail -8 > gpurun_out/r02_gputests_q.log
BMM_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/q_c2.json 2> gpurun_out/q_c2.err
BMM_FETCH_DERIVE=0 BMM_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/q_c2_noderive.json 2> gpurun_out/q_c2_noderive.err
for t in 8 16; do BMM_FETCH_THREADS=$t BMM_TRACE=1 python bench.py --steps 3 --warmup 3 --no-cpu --no-extra > gpurun_out/q_c2_t$t.json 2> gpurun_out/q_c2_t$t.err; done
tail -4 gpurun_out/r02_gputests_q.log; tail -3 gpurun_out/q_c2.err; tail -2 gpurun_out/q_c2_noderive.err; tail -2 gpurun_out/q_c2_t8.err; tail -2 gpurun_out/q_c2_t16.err
