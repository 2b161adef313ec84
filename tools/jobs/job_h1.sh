( time python bench.py ) > gpurun_out/h_default.json 2> gpurun_out/h_default.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/h_reference.json 2> gpurun_out/h_reference.err
tail -4 gpurun_out/h_default.err; tail -4 gpurun_out/h_reference.err; free -g | head -2
