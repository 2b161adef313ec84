# 8 GPUs: the driver's scaling command (default line: C2 chains split + sharded C4 legs + collapsed leg), then the sharded workloads alone
nproc > gpurun_out/s8_box.txt; free -g | head -2 >> gpurun_out/s8_box.txt
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8"
$T --steps 3 --warmup 3 > gpurun_out/s8_default.json 2> gpurun_out/s8_default.err
$T --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/s8_c4.json 2> gpurun_out/s8_c4.err
$T --workload c4relabel --steps 3 --warmup 3 --no-cpu > gpurun_out/s8_c4relabel.json 2> gpurun_out/s8_c4relabel.err
python tools/showbench.py gpurun_out/s8_default.json gpurun_out/s8_c4.json gpurun_out/s8_c4relabel.json
tail -n 2 gpurun_out/s8_default.err; cat gpurun_out/s8_box.txt
