python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r02_gputests_o.log
tail -6 gpurun_out/r02_gputests_o.log
