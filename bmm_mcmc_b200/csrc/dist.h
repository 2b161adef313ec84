#pragma once
#include <cstddef>
#include <cuda_runtime.h>
namespace bmm {
int dist_rank();
int dist_world();
const char *dist_error();
int dist_allreduce_i32(int *buf, size_t n, cudaStream_t st);
int dist_allreduce_f64(double *buf, size_t n, cudaStream_t st);
}
