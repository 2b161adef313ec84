#pragma once
#include <cstddef>
#include <cuda_runtime.h>
namespace bmm {
int dist_rank();
int dist_world();
const char *dist_error();
int dist_allreduce_i32(int *buf, size_t n, cudaStream_t st);
int dist_allreduce_f64(double *buf, size_t n, cudaStream_t st);
// one-shot all-reduce over IPC-mapped peer memory (bmm_dist_p2p_local / bmm_dist_p2p_attach)
bool dist_p2p_ready(size_t n);
int dist_p2p_allreduce_i32(int *buf, size_t n, int *status, cudaStream_t st);
}
