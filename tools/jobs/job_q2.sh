python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r02_gputests_q2.log
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-extra"
BMM_TRACE=1 $B > gpurun_out/q2_c2.json 2> gpurun_out/q2_c2.err
BMM_FETCH_SEGMENTS=1 BMM_FETCH_DMA_FRAC=0 BMM_TRACE=1 $B > gpurun_out/q2_c2_s1f0.json 2> gpurun_out/q2_c2_s1f0.err
BMM_FETCH_SEGMENTS=1 BMM_TRACE=1 $B > gpurun_out/q2_c2_s1.json 2> gpurun_out/q2_c2_s1.err
BMM_FETCH_DMA_FRAC=0 BMM_TRACE=1 $B > gpurun_out/q2_c2_f0.json 2> gpurun_out/q2_c2_f0.err
BMM_FETCH_DMA_FRAC=0.3 BMM_TRACE=1 $B > gpurun_out/q2_c2_f30.json 2> gpurun_out/q2_c2_f30.err
BMM_FETCH_DMA_FRAC=0.1 BMM_TRACE=1 $B > gpurun_out/q2_c2_f10.json 2> gpurun_out/q2_c2_f10.err
BMM_FETCH_DERIVE=0 BMM_TRACE=1 $B > gpurun_out/q2_c2_nd.json 2> gpurun_out/q2_c2_nd.err
BMM_FETCH_SEGMENTS=8 BMM_TRACE=1 $B > gpurun_out/q2_c2_s8.json 2> gpurun_out/q2_c2_s8.err
tail -4 gpurun_out/r02_gputests_q2.log
for f in q2_c2 q2_c2_s1f0 q2_c2_s1 q2_c2_f0 q2_c2_f30 q2_c2_f10 q2_c2_nd q2_c2_s8; do echo $f; grep "bmm trace" gpurun_out/$f.err | tail -3; done
