# round-2 snapshot (1 GPU): all GPU tests, default line + reference arm, per-workload lines, launch list + ncu of the C2 kernels
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_gputests_1gpu.log
tail -3 gpurun_out/r02_gputests_1gpu.log
grep -q failed gpurun_out/r02_gputests_1gpu.log && exit 1
s=$(date +%s); timeout 600 python bench.py > gpurun_out/t_default.json 2> gpurun_out/t_default.err; e=$(date +%s); echo "default wall $((e-s)) s" > gpurun_out/t_wall.txt
s=$(date +%s); timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/t_reference.json 2> gpurun_out/t_reference.err; e=$(date +%s); echo "reference wall $((e-s)) s" >> gpurun_out/t_wall.txt
for w in c4 c4relabel c5 collapsed c1 c3; do
  timeout 300 python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/t_$w.json 2> gpurun_out/t_$w.err
done
python tools/showbench.py gpurun_out/t_default.json gpurun_out/t_c4.json gpurun_out/t_c4relabel.json gpurun_out/t_c5.json gpurun_out/t_collapsed.json gpurun_out/t_c1.json gpurun_out/t_c3.json
cat gpurun_out/t_wall.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_c2_default.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/t_ncu1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:full_chain_kernel -s 12 -c 1 -o gpurun_out/r02_full_chain -f python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/t_ncu2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:collapsed -s 3 -c 1 -o gpurun_out/r02_collapsed -f python bench.py --workload collapsed --steps 1 --warmup 3 --no-cpu > gpurun_out/t_ncu3.log 2>&1
