# 2 GPUs: multi-GPU parity tests, then the default bench line (chains split + sharded C4 legs) and the c4relabel workload
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_gputests_2gpu.log
tail -4 gpurun_out/r02_gputests_2gpu.log
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2"
$T --steps 3 --warmup 3 > gpurun_out/s2_default.json 2> gpurun_out/s2_default.err
$T --workload c4relabel --steps 3 --warmup 3 --no-cpu > gpurun_out/s2_c4relabel.json 2> gpurun_out/s2_c4relabel.err
python tools/showbench.py gpurun_out/s2_default.json gpurun_out/s2_c4relabel.json
tail -n 3 gpurun_out/s2_default.err
