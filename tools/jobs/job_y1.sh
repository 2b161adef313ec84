timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tensor or grid or many_tiles or single_pass or large_p" 2>&1 | tail -6 > gpurun_out/r02_gputests_y1.log
tail -3 gpurun_out/r02_gputests_y1.log
grep -q failed gpurun_out/r02_gputests_y1.log && exit 1
for w in c4 c4relabel c5; do timeout 200 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu > gpurun_out/y1_$w.json 2> gpurun_out/y1_$w.err; done
python tools/showbench.py gpurun_out/y1_c4.json gpurun_out/y1_c4relabel.json gpurun_out/y1_c5.json
