echo "== qfence cap 20"; BMM_LIB=$PWD/bmm_mcmc_b200/libbmm_b200_qfence.so timeout 120 python tools/probes/wsr_debug.py 20
echo "== qfence cap 3"; BMM_LIB=$PWD/bmm_mcmc_b200/libbmm_b200_qfence.so timeout 120 python tools/probes/wsr_debug.py 3
