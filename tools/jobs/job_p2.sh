# default bench line + reference arm (1 GPU), wall-clocked
s=$(date +%s.%N); python bench.py > gpurun_out/p_default.json 2> gpurun_out/p_default.err; e=$(date +%s.%N); echo "default wall $(echo "$e - $s" | bc) s" > gpurun_out/p_wall.txt
s=$(date +%s.%N); python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/p_reference.json 2> gpurun_out/p_reference.err; e=$(date +%s.%N); echo "reference wall $(echo "$e - $s" | bc) s" >> gpurun_out/p_wall.txt
BMM_TRACE=1 python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/p_c2_trace.json 2> gpurun_out/p_c2_trace.err
cat gpurun_out/p_wall.txt; tail -3 gpurun_out/p_default.err
