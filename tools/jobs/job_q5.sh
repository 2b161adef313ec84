python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fetch or summaries or thinning or rerun" 2>&1 | tail -4 > gpurun_out/r02_gputests_q5.log
python tools/probes/widen_probe.py > gpurun_out/q5_widen_probe.txt 2>&1
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-extra"
for rep in 1 2; do for s in 1 2 3; do BMM_FETCH_SEGMENTS=$s $B > gpurun_out/q5_c2_s${s}_r$rep.json 2> gpurun_out/q5_c2_s${s}_r$rep.err; done; done
BMM_FETCH_SEGMENTS=2 BMM_FETCH_THREADS=16 $B > gpurun_out/q5_c2_s2_t16.json 2>/dev/null
BMM_FETCH_SEGMENTS=2 BMM_WIDEN_ISA=avx2 $B > gpurun_out/q5_c2_s2_avx2.json 2>/dev/null
tail -2 gpurun_out/r02_gputests_q5.log
python tools/showbench.py gpurun_out/q5_c2_*.json | grep e2e
grep "derive\|plain" gpurun_out/q5_widen_probe.txt | grep "T=12"
