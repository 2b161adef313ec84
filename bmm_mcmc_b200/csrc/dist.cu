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

int bmm_dist_finalize(void) {
    if (bmm::g_comm) { bmm::p_destroy(bmm::g_comm); bmm::g_comm = nullptr; }
    bmm::g_rank = 0; bmm::g_world = 1;
    return BMM_OK;
}

}  // extern "C"
#pragma GCC visibility pop
