B="python bench.py --workload c4 --steps 3 --warmup 3 --no-cpu"
for v in nb8 nb6na6 nepi2; do
  BMM_LIB=$PWD/bmm_mcmc_b200/libbmm_b200_$v.so $B > gpurun_out/f_c4_$v.json 2> gpurun_out/f_c4_$v.err
done
