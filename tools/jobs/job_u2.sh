timeout 900 python -m pytest tests/test_multigpu.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r02_gputests_2gpu.log
tail -4 gpurun_out/r02_gputests_2gpu.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "many_tiles or fetch or single_pass" 2>&1 | tail -5
T="timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus 2"
$T --steps 3 --warmup 3 > gpurun_out/u2_default.json 2> gpurun_out/u2_default.err
python tools/showbench.py gpurun_out/u2_default.json
