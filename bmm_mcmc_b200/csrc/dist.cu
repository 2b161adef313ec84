// Multi-GPU plumbing: one process per GPU.  Independent chains need no collective (chains are split
// by chain_offset).  The N-sharded uncollapsed samplers all-reduce their integer count tensors once
// per sweep; NCCL is loaded lazily with dlopen so the library itself has no link-time dependency
// (inside a torch process the already-loaded libnccl.so.2 is reused).
#include <dlfcn.h>
#include <cstring>
#include <string>
#include "../../include/bmm_capi.h"
#include "dist.h"

namespace bmm {

namespace {
typedef struct { char internal[128]; } ncclUniqueId_t;
typedef void *ncclComm_t_;
typedef int (*fn_getid)(ncclUniqueId_t *);
typedef int (*fn_init)(ncclComm_t_ *, int, ncclUniqueId_t, int);
typedef int (*fn_destroy)(ncclComm_t_);
typedef int (*fn_allreduce)(const void *, void *, size_t, int, int, ncclComm_t_, cudaStream_t);
typedef const char *(*fn_errstr)(int);

void *g_h = nullptr;
fn_getid p_getid = nullptr;
fn_init p_init = nullptr;
fn_destroy p_destroy = nullptr;
fn_allreduce p_allreduce = nullptr;
fn_errstr p_errstr = nullptr;
ncclComm_t_ g_comm = nullptr;
int g_rank = 0, g_world = 1;
std::string g_derr;

bool load_nccl() {
    if (g_h) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_h) break;
    }
    if (!g_h) { g_derr = "cannot dlopen libnccl.so.2"; return false; }
    p_getid = (fn_getid)dlsym(g_h, "ncclGetUniqueId");
    p_init = (fn_init)dlsym(g_h, "ncclCommInitRank");
    p_destroy = (fn_destroy)dlsym(g_h, "ncclCommDestroy");
    p_allreduce = (fn_allreduce)dlsym(g_h, "ncclAllReduce");
    p_errstr = (fn_errstr)dlsym(g_h, "ncclGetErrorString");
    if (!p_getid || !p_init || !p_destroy || !p_allreduce) { g_derr = "libnccl lacks expected symbols"; return false; }
    return true;
}
}  // namespace

int dist_rank() { return g_rank; }
int dist_world() { return g_world; }
const char *dist_error() { return g_derr.c_str(); }

// sum-all-reduce of int32 (dtype 2 = ncclInt32, op 0 = ncclSum) in place on `st`
int dist_allreduce_i32(int *buf, size_t n, cudaStream_t st) {
    if (g_world == 1) return 0;
    if (!g_comm) { g_derr = "bmm_dist_init has not been called"; return -1; }
    int rc = p_allreduce(buf, buf, n, 2, 0, g_comm, st);
    if (rc) { g_derr = std::string("ncclAllReduce: ") + (p_errstr ? p_errstr(rc) : "error"); return -1; }
    return 0;
}
// double (dtype 8 = ncclFloat64)
int dist_allreduce_f64(double *buf, size_t n, cudaStream_t st) {
    if (g_world == 1) return 0;
    if (!g_comm) { g_derr = "bmm_dist_init has not been called"; return -1; }
    int rc = p_allreduce(buf, buf, n, 8, 0, g_comm, st);
    if (rc) { g_derr = std::string("ncclAllReduce: ") + (p_errstr ? p_errstr(rc) : "error"); return -1; }
    return 0;
}


// ---- one-shot all-reduce over NVLink peer memory ------------------------------------------------------
// The per-sweep exchange is a few KB of int32 counts: latency-bound, and an NCCL call costs ~25 us of
// the ~70 us sweep at 8 GPUs.  Every rank owns an inbox [2 parities][world][cap] ints plus flags
// [2][world] in one cudaMalloc'ed block that the peers map through CUDA IPC.  After its sweep a rank
// PUSHES its counts into slot `rank` of every peer's inbox (remote stores), fences, and writes the sweep
// number into the peer's flag; the gather kernel of each rank spins on its LOCAL flags and sums its
// local inbox.  Parity double-buffering is enough: a rank cannot publish sweep j+2 before it has
// received every peer's sweep j+1, which those peers only send after they consumed sweep j.
namespace {
struct P2P {
    int *local = nullptr;            // this rank's block
    int *peer[64] = {nullptr};       // peer[r] = rank r's block mapped here (peer[rank] = local)
    int **peer_dev = nullptr;        // device copy of peer[]
    size_t cap = 0;                  // ints per (parity, source) slot
    bool attached = false;
} g_p2p;

__host__ __device__ inline size_t p2p_block_ints(size_t cap, int world) { return 2 * (size_t)world * cap + 2 * (size_t)world; }

__global__ void p2p_publish_kernel(int **peer, const int *counts, size_t n, size_t cap, int world, int rank, int parity,
                                   int sweep) {
    int *dst_block = peer[blockIdx.x];
    int *dst = dst_block + ((size_t)parity * world + rank) * cap;
    for (size_t e = threadIdx.x; e < n; e += blockDim.x) dst[e] = counts[e];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile int *flag = dst_block + 2 * (size_t)world * cap + (size_t)parity * world + rank;
        *flag = sweep;
        __threadfence_system();
    }
}

__global__ void p2p_gather_kernel(const int *local, int *counts, size_t n, size_t cap, int world, int parity, int sweep,
                                  int *status) {
    __shared__ int ok_sh;
    if (threadIdx.x == 0) ok_sh = 1;
    __syncthreads();
    if (threadIdx.x < world) {
        const volatile int *flag = local + 2 * (size_t)world * cap + (size_t)parity * world + threadIdx.x;
        long long spins = 0;
        while (*flag != sweep) {
            if (++spins > (1ll << 26)) { ok_sh = 0; break; }   // a peer never published: do not hang the GPU
        }
    }
    __syncthreads();
    if (!ok_sh) { if (threadIdx.x == 0) *status = -7; return; }   // BMM_ERR_NCCL: exchange failed
    __threadfence_system();
    const int *in = local + (size_t)parity * world * cap;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        int acc = 0;
        for (int r = 0; r < world; ++r) acc += ((const volatile int *)in)[(size_t)r * cap + e];
        counts[e] = acc;
    }
}
}  // namespace

bool dist_p2p_ready(size_t n) { return g_p2p.attached && n <= g_p2p.cap && g_world > 1 && g_world <= 64; }

// sum-all-reduce of int32 counts in place.  Every rank makes the same sequence of calls, so a per-process
// call counter is a consistent exchange number; consecutive exchanges alternate the inbox parity.
int dist_p2p_allreduce_i32(int *buf, size_t n, int *status, cudaStream_t st) {
    static int seq = 0;
    const int sweep = ++seq, parity = sweep & 1;
    p2p_publish_kernel<<<g_world, 256, 0, st>>>(g_p2p.peer_dev, buf, n, g_p2p.cap, g_world, g_rank, parity, sweep);
    const int blocks = (int)((n + 255) / 256 < 64 ? (n + 255) / 256 : 64);
    p2p_gather_kernel<<<blocks, 256, 0, st>>>(g_p2p.local, buf, n, g_p2p.cap, g_world, parity, sweep, status);
    if (cudaGetLastError() != cudaSuccess) { g_derr = "p2p all-reduce launch failed"; return -1; }
    return 0;
}

}  // namespace bmm

#pragma GCC visibility push(default)
extern "C" {

int bmm_dist_unique_id(uint8_t id_out[128]) {
    if (!id_out) return BMM_ERR_INVALID;
    if (!bmm::load_nccl()) return BMM_ERR_NCCL;
    bmm::ncclUniqueId_t id;
    if (bmm::p_getid(&id)) { bmm::g_derr = "ncclGetUniqueId failed"; return BMM_ERR_NCCL; }
    memcpy(id_out, id.internal, 128);
    return BMM_OK;
}

int bmm_dist_init(int32_t rank, int32_t world, const uint8_t id[128], int32_t device) {
    if (world < 1 || rank < 0 || rank >= world) return BMM_ERR_INVALID;
    bmm::g_rank = rank; bmm::g_world = world;
    if (world == 1) return BMM_OK;
    if (!id) return BMM_ERR_INVALID;
    if (!bmm::load_nccl()) return BMM_ERR_NCCL;
    if (cudaSetDevice(device) != cudaSuccess) return BMM_ERR_CUDA;
    bmm::ncclUniqueId_t uid;
    memcpy(uid.internal, id, 128);
    int rc = bmm::p_init(&bmm::g_comm, world, uid, rank);
    if (rc) { bmm::g_derr = std::string("ncclCommInitRank: ") + (bmm::p_errstr ? bmm::p_errstr(rc) : "error"); return BMM_ERR_NCCL; }
    return BMM_OK;
}

int bmm_dist_p2p_local(uint64_t cap_ints, uint8_t handle_out[64]) {
    using namespace bmm;
    if (!handle_out || cap_ints == 0 || g_world < 2 || g_world > 64) return BMM_ERR_INVALID;
    if (g_p2p.local) return BMM_ERR_INVALID;
    const size_t ints = p2p_block_ints(cap_ints, g_world);
    if (cudaMalloc((void **)&g_p2p.local, ints * sizeof(int)) != cudaSuccess) { g_derr = "p2p: cudaMalloc failed"; return BMM_ERR_CUDA; }
    cudaMemset(g_p2p.local, 0xFF, ints * sizeof(int));     // flags = -1: no sweep published yet
    cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, g_p2p.local) != cudaSuccess) { g_derr = "p2p: cudaIpcGetMemHandle failed"; cudaGetLastError(); return BMM_ERR_CUDA; }
    static_assert(sizeof(h) == 64, "CUDA IPC handle size");
    memcpy(handle_out, &h, 64);
    g_p2p.cap = cap_ints;
    return BMM_OK;
}

int bmm_dist_p2p_attach(const uint8_t *handles) {
    using namespace bmm;
    if (!handles || !g_p2p.local) return BMM_ERR_INVALID;
    for (int r = 0; r < g_world; ++r) {
        if (r == g_rank) { g_p2p.peer[r] = g_p2p.local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void *ptr = nullptr;
        if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            g_derr = "p2p: cudaIpcOpenMemHandle failed (no peer access between the devices?)";
            cudaGetLastError();
            return BMM_ERR_CUDA;
        }
        g_p2p.peer[r] = (int *)ptr;
    }
    if (cudaMalloc((void **)&g_p2p.peer_dev, 64 * sizeof(int *)) != cudaSuccess) return BMM_ERR_CUDA;
    cudaMemcpy(g_p2p.peer_dev, g_p2p.peer, 64 * sizeof(int *), cudaMemcpyHostToDevice);
    g_p2p.attached = true;
    return BMM_OK;
}

int bmm_dist_p2p_detach(void) {   // fall back to NCCL (all ranks must agree)
    bmm::g_p2p.attached = false;
    return BMM_OK;
}

int bmm_dist_finalize(void) {
    if (bmm::g_p2p.local) {
        cudaDeviceSynchronize();
        for (int r = 0; r < bmm::g_world; ++r)
            if (r != bmm::g_rank && bmm::g_p2p.peer[r]) cudaIpcCloseMemHandle(bmm::g_p2p.peer[r]);
        if (bmm::g_p2p.peer_dev) cudaFree(bmm::g_p2p.peer_dev);
        cudaFree(bmm::g_p2p.local);
        bmm::g_p2p = bmm::P2P{};
    }
    if (bmm::g_comm) { bmm::p_destroy(bmm::g_comm); bmm::g_comm = nullptr; }
    bmm::g_rank = 0; bmm::g_world = 1;
    return BMM_OK;
}

}  // extern "C"
#pragma GCC visibility pop
