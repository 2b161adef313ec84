lscpu | head -30 > gpurun_out/q3_lscpu.txt
B="python bench.py --steps 4 --warmup 3 --no-cpu --no-extra"
for s in 1 2 3 4; do BMM_FETCH_DMA_FRAC=0 BMM_FETCH_SEGMENTS=$s $B > gpurun_out/q3_c2_s$s.json 2> gpurun_out/q3_c2_s$s.err; done
BMM_FETCH_DMA_FRAC=0 BMM_FETCH_SEGMENTS=1 BMM_FETCH_THREADS=16 $B > gpurun_out/q3_c2_s1_t16.json 2>/dev/null
BMM_FETCH_DMA_FRAC=0 BMM_FETCH_SEGMENTS=2 BMM_FETCH_THREADS=16 $B > gpurun_out/q3_c2_s2_t16.json 2>/dev/null
BMM_FETCH_DMA_FRAC=0.06 BMM_FETCH_SEGMENTS=2 $B > gpurun_out/q3_c2_s2_f6.json 2>/dev/null
python tools/showbench.py gpurun_out/q3_c2_*.json | grep e2e
grep -E "Model name|^CPU\(s\)|Flags" gpurun_out/q3_lscpu.txt | cut -c1-400
