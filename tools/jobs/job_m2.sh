python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_p_tensor_path_vs_oracle" 2>&1 | tail -40 > gpurun_out/r02_gputests_m2.log
