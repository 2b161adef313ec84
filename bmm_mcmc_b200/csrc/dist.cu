// Multi-GPU plumbing: one process per GPU.  Independent chains need no collective (chains are split
// by chain_offset).  The N-sharded uncollapsed samplers all-reduce their integer count tensors once
// per sweep; NCCL is loaded lazily with dlopen so the library itself has no link-time dependency
// (inside a torch process the already-loaded libnccl.so.2 is reused).
#include <dlfcn.h>
#include <cstring>
#include <string>
#include "../../include/bmm_capi.h"
#include "dist.h"
#include "kernels.h"

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
void derr(const std::string &m) { g_derr = m; set_last_error(m.c_str()); }

bool load_nccl() {
    if (g_h) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *n : names) {
        g_h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_h) break;
    }
    if (!g_h) { derr("cannot dlopen libnccl.so.2"); return false; }
    p_getid = (fn_getid)dlsym(g_h, "ncclGetUniqueId");
    p_init = (fn_init)dlsym(g_h, "ncclCommInitRank");
    p_destroy = (fn_destroy)dlsym(g_h, "ncclCommDestroy");
    p_allreduce = (fn_allreduce)dlsym(g_h, "ncclAllReduce");
    p_errstr = (fn_errstr)dlsym(g_h, "ncclGetErrorString");
    if (!p_getid || !p_init || !p_destroy || !p_allreduce) { derr("libnccl lacks expected symbols"); return false; }
    return true;
}
}  // namespace

int dist_rank() { return g_rank; }
int dist_world() { return g_world; }
const char *dist_error() { return g_derr.c_str(); }

// sum-all-reduce of int32 (dtype 2 = ncclInt32, op 0 = ncclSum) in place on `st`
int dist_allreduce_i32(int *buf, size_t n, cudaStream_t st) {
    if (g_world == 1) return 0;
    if (!g_comm) { derr("bmm_dist_init has not been called"); return -1; }
    int rc = p_allreduce(buf, buf, n, 2, 0, g_comm, st);
    if (rc) { derr(std::string("ncclAllReduce: ") + (p_errstr ? p_errstr(rc) : "error")); return -1; }
    return 0;
}
// double (dtype 8 = ncclFloat64)
int dist_allreduce_f64(double *buf, size_t n, cudaStream_t st) {
    if (g_world == 1) return 0;
    if (!g_comm) { derr("bmm_dist_init has not been called"); return -1; }
    int rc = p_allreduce(buf, buf, n, 8, 0, g_comm, st);
    if (rc) { derr(std::string("ncclAllReduce: ") + (p_errstr ? p_errstr(rc) : "error")); return -1; }
    return 0;
}


// ---- count exchange over NVLink peer memory ----------------------------------------------------------
// The per-sweep exchange is a few KB of int32 counts: latency-bound.  Every rank owns an inbox
// [2 parities][world][cap] of 8-byte words (count, exchange number) in one cudaMalloc'ed block that the peers
// map through CUDA IPC.  After its sweep a rank PUSHES its counts into slot `rank` of every rank's inbox
// (remote 8-byte stores over NVLink, its own slot included); the consumer (big_update_kernel) re-reads each
// word it needs until the word carries the number of the exchange it waits for, and sums over the ranks.
// No fence, no flag, no gather launch (system-scope fences after remote stores cost ~10 us per sweep on this
// box, which is what made the first, flag-based version of this exchange slower than NCCL), and on the
// warp-specialised tensor path no publish launch either: the last CTA of the sweep kernel to flush its
// counts does the push (kern_big_ws.cu).
// Parity double-buffering is enough: a rank cannot publish exchange s+2 before it has received every
// peer's s+1, which those peers only send after they consumed s.
// Exchange numbers come from a device-side counter so that a captured CUDA graph replays with fresh
// numbers: seq[0] = number of sweep 0 of the current run, seq[1] = next free number.
namespace {
struct P2P {
    int2 *local = nullptr;           // this rank's block
    int2 *peer[64] = {nullptr};      // peer[r] = rank r's block mapped here (peer[rank] = local)
    int2 **peer_dev = nullptr;       // device copy of peer[]
    int *seq = nullptr;              // device: [0] base of the current run, [1] next free exchange number
    size_t cap = 0;                  // ints per (parity, source) slot
    bool attached = false;
} g_p2p;

__host__ __device__ inline size_t p2p_block_words(size_t cap, int world) { return 2 * (size_t)world * cap; }

// generic publish (sweep kernels without a fused push): block r pushes to rank r
__global__ void p2p_publish_kernel(int2 *const *peer, const int *counts, size_t n, size_t cap, int world, int rank,
                                   const int *seq, int j) {
    const int s = seq[0] + j;
    int2 *dst = peer[blockIdx.x] + x_slot_off(s, world, rank, cap);
    for (size_t e = threadIdx.x; e < n; e += blockDim.x) x_store(dst + e, counts[e], s);
}
}  // namespace

bool dist_p2p_ready(size_t n) { return g_p2p.attached && n <= g_p2p.cap && g_world > 1 && g_world <= 64; }

P2PView dist_p2p_view() {
    P2PView v;
    v.peer = g_p2p.peer_dev; v.local = g_p2p.local; v.seq = g_p2p.seq; v.cap = g_p2p.cap;
    v.world = g_world; v.rank = g_rank;
    return v;
}

// reserve the exchange numbers of a run of n sweeps (stream-ordered; every rank makes the same calls)
int dist_p2p_begin_run(int n, cudaStream_t st) {
    if (launch_x_begin_run(g_p2p.seq, n, st) != cudaSuccess) { derr("p2p begin_run launch failed"); return -1; }
    return 0;
}

int dist_p2p_publish(const int *counts, size_t n, int j, cudaStream_t st) {
    p2p_publish_kernel<<<g_world, 256, 0, st>>>(g_p2p.peer_dev, counts, n, g_p2p.cap, g_world, g_rank, g_p2p.seq, j);
    if (cudaGetLastError() != cudaSuccess) { derr("p2p publish launch failed"); return -1; }
    return 0;
}

}  // namespace bmm

#pragma GCC visibility push(default)
extern "C" {

int bmm_dist_unique_id(uint8_t id_out[128]) {
    if (!id_out) return BMM_ERR_INVALID;
    if (!bmm::load_nccl()) return BMM_ERR_NCCL;
    bmm::ncclUniqueId_t id;
    if (bmm::p_getid(&id)) { bmm::derr("ncclGetUniqueId failed"); return BMM_ERR_NCCL; }
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
    if (rc) { bmm::derr(std::string("ncclCommInitRank: ") + (bmm::p_errstr ? bmm::p_errstr(rc) : "error")); return BMM_ERR_NCCL; }
    return BMM_OK;
}

int bmm_dist_p2p_local(uint64_t cap_ints, uint8_t handle_out[64]) {
    using namespace bmm;
    if (!handle_out || cap_ints == 0 || g_world < 2 || g_world > 64) return BMM_ERR_INVALID;
    if (g_p2p.local) return BMM_ERR_INVALID;
    const size_t words = p2p_block_words(cap_ints, g_world);
    if (cudaMalloc((void **)&g_p2p.local, words * sizeof(int2)) != cudaSuccess) { derr("p2p: cudaMalloc failed"); return BMM_ERR_CUDA; }
    cudaMemset(g_p2p.local, 0xFF, words * sizeof(int2));     // tags = -1: nothing published yet
    if (cudaMalloc((void **)&g_p2p.seq, 2 * sizeof(int)) != cudaSuccess) { derr("p2p: cudaMalloc failed"); return BMM_ERR_CUDA; }
    const int seq0[2] = {1, 1};
    cudaMemcpy(g_p2p.seq, seq0, sizeof seq0, cudaMemcpyHostToDevice);
    cudaDeviceSynchronize();
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, g_p2p.local) != cudaSuccess) { derr("p2p: cudaIpcGetMemHandle failed"); cudaGetLastError(); return BMM_ERR_CUDA; }
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
            derr("p2p: cudaIpcOpenMemHandle failed (no peer access between the devices?)");
            cudaGetLastError();
            return BMM_ERR_CUDA;
        }
        g_p2p.peer[r] = (int2 *)ptr;
    }
    if (cudaMalloc((void **)&g_p2p.peer_dev, 64 * sizeof(int2 *)) != cudaSuccess) return BMM_ERR_CUDA;
    cudaMemcpy(g_p2p.peer_dev, g_p2p.peer, 64 * sizeof(int2 *), cudaMemcpyHostToDevice);
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
        if (bmm::g_p2p.seq) cudaFree(bmm::g_p2p.seq);
        cudaFree(bmm::g_p2p.local);
        bmm::g_p2p = bmm::P2P{};
    }
    if (bmm::g_comm) { bmm::p_destroy(bmm::g_comm); bmm::g_comm = nullptr; }
    bmm::g_rank = 0; bmm::g_world = 1;
    return BMM_OK;
}

}  // extern "C"
#pragma GCC visibility pop
