# ncu capture of the tensor sweep kernel (C4 shape, N = 4e6, a few sweeps)
python bench.py --workload c4 --n 4000000 --nsamples 14 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_ws_plain.json 2> gpurun_out/ncu_ws_plain.err || exit 1
BMM_GRAPH=0 ncu --set full --clock-control none --import-source on -k regex:big_sweep_ws -s 45 -c 1 -o gpurun_out/r02_ws_v3 -f \
  python bench.py --workload c4 --n 4000000 --nsamples 14 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_ws.log 2>&1
tail -3 gpurun_out/ncu_ws.log
