timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_gputests_1gpu.log
tail -3 gpurun_out/r02_gputests_1gpu.log
grep -q failed gpurun_out/r02_gputests_1gpu.log && exit 1
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/v_default.json 2> gpurun_out/v_default.err
timeout 300 python bench.py --workload c4relabel --steps 3 --warmup 3 --no-cpu > gpurun_out/v_c4relabel.json 2> gpurun_out/v_c4relabel.err
python tools/showbench.py gpurun_out/v_default.json gpurun_out/v_c4relabel.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_default.json').read().strip().splitlines()[-1])
print(json.dumps(d['extra'].get('cpu_optimised'))[:600])
sh=d['extra']['sharded']
for k in ('relabel_off','relabel_on'):
    v=sh[k]; print(k, "value %.3e ms/sweep %.4f sweep_us %.1f exch_us %.1f relabel_us %.1f hbm_frac %.3f" % (v['value'], v['ms_per_sweep'], v['sweep_kernel_us'], v['exchange_us'], v['relabel_kernel_us'], v['hbm_frac']))
PY
