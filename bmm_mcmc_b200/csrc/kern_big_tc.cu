// Tensor-core z-sweep of the grid path (BMM_FP32): the N x K Bernoulli log-likelihood of
// /root/reference/src/full_gibbs.cpp:92-106 (stickbreaking.cpp:75-92) as a tcgen05 contraction, and the
// sufficient statistics of full_gibbs.cpp:182-200 (stickbreaking.cpp:164-186) as a second one, fused
// in one persistent kernel so every bit-packed row is read from HBM exactly once per sweep.
//
//   loglh_k(x_i) = sum_d x_id (log th_kd - log(1-th_kd)) + sum_d log(1-th_kd)  =  (X D^T)_ik + b_k
//   GEMM1  [128 obs x Pd] (bf16 0/1, K-major)  x  [3*KC x Pd]^T (D split hi|mid|lo into three bf16
//          terms, fp32 accumulate in TMEM)  ->  128 x 3*KC accumulator; thread i owns TMEM lane i.
//   epilogue (one observation per thread): logits = hi+mid+lo + b_k + log pi_k, max, exp2, sum,
//          inverse-CDF draw with the observation's Philox uniform, 1-byte allocation to HBM,
//          one-hot row (bf16) to shared memory.
//   GEMM2  [X | 1]^T (128 x 128 obs, the same shared-memory tile read MN-major)  x  one-hot
//          [128 obs x KC] (MN-major)  ->  V_kd^T (rows d < P) and c_k (the all-ones row), exact integer
//          counts in fp32, accumulated in TMEM over all tiles of the CTA, flushed once with atomics.
//
// Shared-memory operand layout (no swizzle): 16-byte chunk c of row r at  base + c*2048 + r*16, so a
// UMMA core matrix (8 rows x 16 B) is 128 contiguous bytes, SBO (next 8 rows) = 128 B and LBO (next
// chunk) = 2048 B for the K-major reading; the MN-major reading of the same bytes has LBO = 128 B,
// SBO = 2048 B.  One warpgroup (128 threads) per tile; NWG warpgroups per CTA work on different tiles
// concurrently, each with its own operand buffers, TMEM accumulators and mbarriers; one CTA per SM.
#include <cuda_bf16.h>

#include "common.cuh"
#include "kernels.h"

namespace bmm {
namespace {

constexpr int TC_TILE = 128;
constexpr int TC_CHUNK = TC_TILE * 16;  // bytes of one 16-byte-chunk column over 128 rows
constexpr int TC_ACHUNKS = 16;          // A region: 128 "d" positions = UMMA M of GEMM2
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// UMMA shared-memory descriptor, SWIZZLE_NONE, version 1 (sm_100): start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
           ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}

// kind::f16 instruction descriptor: D = f32 (bit 4), A = B = bf16 (bits 7, 10), majors, N >> 3, M >> 4
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void wg_barrier(int wg) { asm volatile("bar.sync %0, %1;" :: "r"(1 + wg), "r"(128) : "memory"); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
// Bounded wait (a lost arrival must not hang the GPU): false on time-out.
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (int spin = 0; spin < (1 << 22); ++spin) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// 2 bits -> two bf16 (0.0 / 1.0) packed in a u32
__device__ __forceinline__ uint32_t bits2_bf16x2(uint32_t t) {
    return (((t & 3u) * 0x8001u) & 0x00010001u) * 0x3F80u;
}

template <int KC, int NWG>
struct TcLayout {
    static constexpr int A_BYTES = TC_ACHUNKS * TC_CHUNK;          // 32 KB per warpgroup
    static constexpr int B2_BYTES = (KC / 8) * TC_CHUNK;           // one-hot, MN-major
    static constexpr int WG_BYTES = A_BYTES + B2_BYTES;
    static constexpr int B1_ROW = 3 * KC * 16;                     // bytes of one d-chunk of the split table
    static constexpr int B1_OFF = NWG * WG_BYTES;
    static constexpr int TMEM_COLS = 512;
    static constexpr int ACC_COLS = 4 * KC;                        // per warpgroup: 3*KC (GEMM1) + KC (GEMM2)
    static_assert(NWG * ACC_COLS <= TMEM_COLS, "TMEM budget");
    __host__ __device__ static constexpr int bias_off(int nch) { return B1_OFF + nch * B1_ROW; }
    __host__ __device__ static constexpr int lpi_off(int nch) { return bias_off(nch) + KC * 4; }
    __host__ __device__ static constexpr int bar_off(int nch) { return lpi_off(nch) + KC * 4; }
    __host__ __device__ static constexpr int total(int nch) { return bar_off(nch) + 2 * NWG * 8 + 16; }
};

template <int KC, int NWG>
__global__ void __launch_bounds__(NWG * 128, 1) big_sweep_tc_kernel(const BigParams p, const int j, const int nch) {
    using L = TcLayout<KC, NWG>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, wq = (tid >> 5) & 3;
    const int K = p.K, P = p.P, W = p.W;
    const int ONES = nch * 8;                      // d index of the all-ones column of [X | 1]
    unsigned char *A = smem + wg * L::WG_BYTES;
    unsigned char *B2 = A + L::A_BYTES;
    unsigned char *B1 = smem + L::B1_OFF;
    float *bias = (float *)(smem + L::bias_off(nch));
    float *lpis = (float *)(smem + L::lpi_off(nch));
    uint64_t *bars = (uint64_t *)(smem + L::bar_off(nch));
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * NWG);
    const uint32_t bar1 = smem_u32(&bars[2 * wg]), bar2 = smem_u32(&bars[2 * wg + 1]);

    // ---- prologue: TMEM, barriers, split weight table, constant part of [X | 1] ----------------
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(L::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 32) {
        for (int b = 0; b < 2 * NWG; ++b) mbar_init(smem_u32(&bars[b]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // D_kd = log th_kd - log(1 - th_kd) split into three bf16 terms: row (part*KC + k), column d
    for (int e = tid; e < KC * nch * 8; e += NWG * 128) {
        const int k = e % KC, d = e / KC;
        double D = 0.0;
        if (k < K && d < P) D = p.w1[k + K * d] - p.w0[k + K * d];
        const __nv_bfloat16 hi = __double2bfloat16(D);
        const double r1 = D - (double)__bfloat162float(hi);
        const __nv_bfloat16 mid = __double2bfloat16(r1);
        const __nv_bfloat16 lo = __double2bfloat16(r1 - (double)__bfloat162float(mid));
        unsigned char *col = B1 + (d >> 3) * L::B1_ROW + (d & 7) * 2;
        *(__nv_bfloat16 *)(col + (0 * KC + k) * 16) = hi;
        *(__nv_bfloat16 *)(col + (1 * KC + k) * 16) = mid;
        *(__nv_bfloat16 *)(col + (2 * KC + k) * 16) = lo;
    }
    for (int k = tid; k < KC; k += NWG * 128) {
        float b = -INFINITY, lp = 0.f;
        if (k < K) {
            double s0 = 0.0;
            for (int d = 0; d < P; ++d) s0 += p.w0[k + K * d];
            b = (float)(p.lpi[k] + s0);
            lp = (float)p.lpi[k];
        }
        bias[k] = b;
        lpis[k] = lp;
    }
    // chunks nch..15 of this warpgroup's A region never change: zero, except the ones column
    for (int c = nch; c < TC_ACHUNKS; ++c) {
        uint4 v = make_uint4(c == nch ? 0x3F80u : 0u, 0u, 0u, 0u);
        *(uint4 *)(A + c * TC_CHUNK + t * 16) = v;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc1 = tmem_base + (uint32_t)(wg * L::ACC_COLS);
    const uint32_t acc2 = acc1 + 3 * KC;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;

    constexpr uint32_t IDESC1 = umma_idesc(128, 3 * KC, 0, 0);
    constexpr uint32_t IDESC2 = umma_idesc(128, KC, 1, 1);
    const uint2 key = make_uint2((uint32_t)p.seed, (uint32_t)p.chain_offset);
    const uint32_t sid = ST_Z ^ ((uint32_t)(p.seed >> 32) << 8);
    uint8_t *zrow = p.zhist ? p.zhist + (size_t)(p.keep_history ? j : 0) * p.N_local : nullptr;
    const long long ntiles = ((long long)p.N_local + TC_TILE - 1) / TC_TILE;
    bool ok = true;
    int it = 0;

    for (long long tile = (long long)blockIdx.x * NWG + wg; tile < ntiles && ok; tile += (long long)gridDim.x * NWG, ++it) {
        const long long i = tile * TC_TILE + t;
        const bool valid = i < p.N_local;
        uint32_t xw[4] = {0u, 0u, 0u, 0u};
        if (valid) {
            const uint32_t *xb = p.xbits + (size_t)i * W;
#pragma unroll
            for (int w = 0; w < 4; ++w) if (w < W) xw[w] = xb[w];
        }
        const unsigned long long gi = (unsigned long long)p.row_offset + (unsigned long long)i;
        // the previous tile's GEMM2 must have finished reading A / B2 before they are overwritten
        if (it > 0) ok = mbar_wait(bar2, (uint32_t)((it - 1) & 1));
        if (!ok) break;
#pragma unroll
        for (int c = 0; c < TC_ACHUNKS - 1; ++c) {
            if (c < nch) {
                const uint32_t byte = (xw[c >> 2] >> ((c & 3) * 8)) & 0xFFu;
                uint4 v = make_uint4(bits2_bf16x2(byte), bits2_bf16x2(byte >> 2), bits2_bf16x2(byte >> 4), bits2_bf16x2(byte >> 6));
                *(uint4 *)(A + c * TC_CHUNK + t * 16) = v;
            }
        }
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(A), b0 = smem_u32(B1);
            for (int kk = 0; kk < nch / 2; ++kk)
                umma_bf16(acc1, umma_desc(a0 + kk * 2 * TC_CHUNK, TC_CHUNK, 128),
                          umma_desc(b0 + kk * 2 * L::B1_ROW, L::B1_ROW, 128), IDESC1, kk > 0);
            umma_commit(bar1);
        }
        // this observation's uniform (same counters as the other uncollapsed kernels)
        const uint4 rnd = philox4x32_10(make_uint4((uint32_t)(gi >> 1), (uint32_t)(gi >> 33), sid, (uint32_t)j), key);
        const float u = (float)((gi & 1) ? u53(rnd.z, rnd.w) : u53(rnd.x, rnd.y));
        ok = mbar_wait(bar1, (uint32_t)(it & 1));
        if (!ok) break;
        tc_fence_after();
        float l[KC];
        {
            uint32_t v[32];
#pragma unroll
            for (int part = 0; part < 3; ++part) {
#pragma unroll
                for (int c0 = 0; c0 < KC; c0 += 32) {
                    tmem_ld32(acc1 + lane_sel + (uint32_t)(part * KC + c0), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const float f = __uint_as_float(v[q]);
                        l[c0 + q] = part == 0 ? f : l[c0 + q] + f;
                    }
                }
            }
        }
        float mx = -INFINITY;
#pragma unroll
        for (int k = 0; k < KC; ++k) { l[k] += bias[k]; mx = fmaxf(mx, l[k]); }
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < KC; ++k) { l[k] = exp2f((l[k] - mx) * LOG2E); sum += l[k]; }
        if (!(sum > 0.f) || !isfinite(sum)) *p.status = -9;  // BMM_ERR_PROB
        const float target = u * sum;
        float c = 0.f;
        int z = 0;
#pragma unroll
        for (int k = 0; k < KC; ++k) { c += l[k]; z += (k < K - 1 && c <= target) ? 1 : 0; }
        if (valid) {
            if (zrow) zrow[i] = (uint8_t)(z + 1);
            if (p.probs_out) {
                const float inv = 1.f / sum;
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k < K) p.probs_out[(size_t)j * p.N_local * K + i + (size_t)p.N_local * k] = (double)(l[k] * inv);
            }
        }
        // one-hot row of the allocation, MN-major: chunk z/8 holds 1.0 at element z%8
#pragma unroll
        for (int cc = 0; cc < KC / 8; ++cc) {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (valid && (z >> 3) == cc) {
                const uint32_t h = (z & 1) ? 0x3F800000u : 0x3F80u;
                const int wsel = (z & 7) >> 1;
                v.x = wsel == 0 ? h : 0u; v.y = wsel == 1 ? h : 0u; v.z = wsel == 2 ? h : 0u; v.w = wsel == 3 ? h : 0u;
            }
            *(uint4 *)(B2 + cc * TC_CHUNK + t * 16) = v;
        }
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(A), b0 = smem_u32(B2);
            for (int kk = 0; kk < TC_TILE / 16; ++kk)
                umma_bf16(acc2, umma_desc(a0 + kk * 256, 128, TC_CHUNK), umma_desc(b0 + kk * 256, 128, TC_CHUNK),
                          IDESC2, (it > 0 || kk > 0) ? 1u : 0u);
            umma_commit(bar2);
        }
    }
    // ---- flush this warpgroup's counts: TMEM lane d holds V_kd (d < P) or c_k (d == ONES) -------
    if (ok && it > 0) ok = mbar_wait(bar2, (uint32_t)((it - 1) & 1));
    if (ok && it > 0) {
        tc_fence_after();
        int *gcnt = p.counts + (size_t)(j & 1) * (K + K * P);
#pragma unroll
        for (int c0 = 0; c0 < KC; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(acc2 + lane_sel + (uint32_t)c0, v);
            tmem_ld_wait();
            if (t < P || t == ONES) {
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const int k = c0 + q;
                    const int n = (int)(__uint_as_float(v[q]) + 0.5f);
                    if (k < K && n) atomicAdd(t == ONES ? &gcnt[k] : &gcnt[K + k + K * t], n);
                }
            }
        }
    }
    if (!ok) *p.status = -10;  // BMM_ERR_TIMEOUT: an mbarrier never completed
    tc_fence_before();
    __syncthreads();
    if (tid < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(L::TMEM_COLS) : "memory");
    }
}

}  // namespace

// The tensor-core sweep covers the float path for K <= 32 clusters and P <= 112 variables.
bool big_tc_supported(const BigParams &p) {
    return p.precision == 1 && p.K <= 32 && p.P <= 112 && p.W <= 4 && p.ru == nullptr && p.loglik_out == nullptr;
}

cudaError_t launch_big_sweep_tc(const BigParams &p, int j, int sm_count, cudaStream_t st) {
    constexpr int KC = 32, NWG = 4;
    using L = TcLayout<KC, NWG>;
    const int nch = 2 * ((p.P + 15) / 16);
    const size_t smem = (size_t)L::total(nch);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(big_sweep_tc_kernel<KC, NWG>, cudaFuncAttributeMaxDynamicSharedMemorySize, L::total(14));
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const long long ntiles = ((long long)p.N_local + TC_TILE - 1) / TC_TILE;
    long long ctas = (ntiles + NWG - 1) / NWG;
    if (ctas > sm_count) ctas = sm_count;
    if (ctas < 1) ctas = 1;
    big_sweep_tc_kernel<KC, NWG><<<(int)ctas, NWG * 128, smem, st>>>(p, j, nch);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace bmm
