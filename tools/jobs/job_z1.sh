B="timeout 120 python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu"
$B > gpurun_out/z1_c4.json 2>/dev/null
BMM_LIB=$PWD/bmm_mcmc_b200/libbmm_b200_nepi4.so $B > gpurun_out/z1_c4_nepi4.json 2>/dev/null
BMM_LIB=$PWD/bmm_mcmc_b200/libbmm_b200_nepi4.so timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_tiles or c4_shape or tensor_path_small" 2>&1 | tail -2
python tools/showbench.py gpurun_out/z1_c4.json gpurun_out/z1_c4_nepi4.json | grep value
A="--workload c4 --nsamples 14 --steps 1 --warmup 3 --no-cpu"
BMM_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:big_sweep_ws -s 45 -c 1 -o gpurun_out/r02_ws_v5 -f python bench.py $A > gpurun_out/z1_ncu1.log 2>&1
