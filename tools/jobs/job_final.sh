timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r02_gputests_1gpu.log
tail -2 gpurun_out/r02_gputests_1gpu.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 300 python bench.py > gpurun_out/f_default.json 2> gpurun_out/f_default.err
python tools/showbench.py gpurun_out/f_default.json | head -1
