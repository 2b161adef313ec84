A="--workload c4relabel --nsamples 22 --steps 1 --warmup 2 --no-cpu"
BMM_GRAPH=0 BMM_ASSIGN_FORK=0 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c4relabel_v3.csv python bench.py $A > gpurun_out/ncu_wsr2.log 2>&1
python tools/launch_agg.py gpurun_out/r02_launches_c4relabel_v3.csv
grep grid_assign gpurun_out/r02_launches_c4relabel_v3.csv | awk -F'","' '{print $NF}' | tr -d '"' | head -70 | tr '\n' ' '
