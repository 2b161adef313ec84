# state snapshot (1 GPU): GPU tests, default bench line + reference arm, every workload, launch list of the default line
python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_gputests_p.log
( time python bench.py ) > gpurun_out/p_default.json 2> gpurun_out/p_default.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/p_reference.json 2> gpurun_out/p_reference.err
for w in c4 c4relabel c5 collapsed c1 c3; do
  python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/p_$w.json 2> gpurun_out/p_$w.err
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/p_launches_c2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extra > gpurun_out/p_ncu.log 2>&1
tail -4 gpurun_out/r02_gputests_p.log; tail -4 gpurun_out/p_default.err; nproc; free -g | head -2
