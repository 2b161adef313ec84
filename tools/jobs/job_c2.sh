# scratch GPU job (2 GPUs): multi-GPU tests + C4 bench sharded
python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r02_gputests_2gpu.log
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload c4 --steps 3 --warmup 3 --no-cpu"
$T > gpurun_out/c_c4_2gpu.json 2> gpurun_out/c_c4_2gpu.err
BMM_P2P=0 $T > gpurun_out/c_c4_2gpu_nccl.json 2> gpurun_out/c_c4_2gpu_nccl.err
tail -5 gpurun_out/r02_gputests_2gpu.log
