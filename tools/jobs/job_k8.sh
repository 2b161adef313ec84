free -g | head -2 > gpurun_out/k8_box.txt; nproc >> gpurun_out/k8_box.txt
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8"
( time $T --steps 3 --warmup 3 ) > gpurun_out/k8_default.json 2> gpurun_out/k8_default.err
$T --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/k8_c4.json 2> gpurun_out/k8_c4.err
BMM_P2P=0 $T --workload c4 --steps 3 --warmup 3 --no-cpu > gpurun_out/k8_c4_nccl.json 2> gpurun_out/k8_c4_nccl.err
tail -3 gpurun_out/k8_default.err; cat gpurun_out/k8_box.txt
