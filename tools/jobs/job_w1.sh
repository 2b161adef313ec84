timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fetch or summaries or rerun" 2>&1 | tail -3
B="timeout 120 python bench.py --steps 4 --warmup 3 --no-cpu --no-extra"
for rep in 1 2 3; do for t in 12 14 16; do BMM_FETCH_THREADS=$t $B > gpurun_out/w1_t${t}_r$rep.json 2>/dev/null; done; done
python tools/showbench.py gpurun_out/w1_t*.json | grep e2e | sed 's/| launches.*//'
