python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "collapsed or rerun or chain_split" 2>&1 | tail -8 > gpurun_out/r02_gputests_i.log
for w in collapsed c1; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/i_$w.json 2> gpurun_out/i_$w.err
  BMM_COLLAPSED_KERNEL=log python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/i_${w}_log.json 2>/dev/null
done
tail -4 gpurun_out/r02_gputests_i.log
