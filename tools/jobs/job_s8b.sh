# 8 GPUs: C2 end to end with host widening forced on (default for > 2 ranks is device widening + int32 DMA)
T="timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu --no-extra"
BMM_FETCH_WIDEN=1 $T > gpurun_out/s8b_widen.json 2> gpurun_out/s8b_widen.err
BMM_FETCH_WIDEN=1 BMM_FETCH_THREADS=4 $T > gpurun_out/s8b_widen_t4.json 2> gpurun_out/s8b_widen_t4.err
python tools/showbench.py gpurun_out/s8b_widen.json gpurun_out/s8b_widen_t4.json | grep e2e
