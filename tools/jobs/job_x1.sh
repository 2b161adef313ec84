# final ncu captures: tensor sweep kernel (C4) and single-pass relabelling kernel (C4 + relabel) at N = 1e7, launch list of C4
A="--workload c4 --nsamples 14 --steps 1 --warmup 3 --no-cpu"
timeout 200 python bench.py $A > gpurun_out/x1_plain.json 2>/dev/null || exit 1
BMM_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:big_sweep_ws -s 45 -c 1 -o gpurun_out/r02_ws_v4 -f python bench.py $A > gpurun_out/x1_ncu1.log 2>&1
BMM_GRAPH=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r02_launches_c4.csv python bench.py $A > gpurun_out/x1_ncu2.log 2>&1
R="--workload c4relabel --nsamples 12 --steps 1 --warmup 3 --no-cpu"
BMM_GRAPH=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:big_relabel_ws -s 12 -c 1 -o gpurun_out/r02_wsr_v4 -f python bench.py $R > gpurun_out/x1_ncu3.log 2>&1
python tools/launch_agg.py gpurun_out/r02_launches_c4.csv
