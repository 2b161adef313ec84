# ncu capture of the single-pass relabelling kernel (C4 shape, N = 1e7)
A="--workload c4relabel --nsamples 12 --steps 1 --warmup 3 --no-cpu"
python bench.py $A > gpurun_out/ncu_wsr_plain.json 2> gpurun_out/ncu_wsr_plain.err || exit 1
BMM_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:big_relabel_ws -s 12 -c 1 -o gpurun_out/r02_wsr_v1 -f \
  python bench.py $A > gpurun_out/ncu_wsr.log 2>&1
tail -3 gpurun_out/ncu_wsr.log
BMM_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c4relabel.csv python bench.py $A > gpurun_out/ncu_wsr2.log 2>&1
tail -2 gpurun_out/ncu_wsr2.log
