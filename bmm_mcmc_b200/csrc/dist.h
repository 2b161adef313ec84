#pragma once
#include <cstddef>
#include <cuda_runtime.h>
namespace bmm {
int dist_rank();
int dist_world();
const char *dist_error();
// capi.cu: what bmm_last_error() returns for this thread
void set_last_error(const char *msg);
int dist_allreduce_i32(int *buf, size_t n, cudaStream_t st);
int dist_allreduce_f64(double *buf, size_t n, cudaStream_t st);
// count exchange over IPC-mapped peer memory (bmm_dist_p2p_local / bmm_dist_p2p_attach); see dist.cu
struct P2PView {
    int2 *const *peer = nullptr;  // device array [world]: every rank's inbox block as mapped into this process
    int2 *local = nullptr;        // this rank's inbox block: [2][world][cap] words of (count, exchange number)
    const int *seq = nullptr;     // device: seq[0] + j = exchange number of sweep j of the current run
    size_t cap = 0;
    int world = 1, rank = 0;
};
bool dist_p2p_ready(size_t n);
P2PView dist_p2p_view();
int dist_p2p_begin_run(int n_sweeps, cudaStream_t st);
int dist_p2p_publish(const int *counts, size_t n, int j, cudaStream_t st);
}
