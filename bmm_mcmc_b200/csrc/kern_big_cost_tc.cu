// Stephens cost matrix of the grid path on tcgen05 for 8 <= K <= 128, K % 8 == 0:
//   G(k,l) = sum_i log q_ik * p_il,   s_l = sum_i p_il^2  (online)  or  sum_i p_il log p_il  (batch)
// (/root/reference/src/stephens.cpp:45-53,76-84), a K x K contraction over the N observations of two
// row-major fp32 matrices.  The CUDA-core version in kern_big_relabel.cu runs at a quarter of the FMA
// peak (1.97 ms at N = 1e6, K = 128); here the contraction is a 128 x 128 x N GEMM whose operands are
// produced on the fly: 8 warps stream 32 (wide) observations of P and Q per stage through a cp.async ring, take the logarithm, split
// every value into fp16 hi + lo (22 significant bits) and store the MN-major operand image; one thread
// issues hi*hi + hi*lo + lo*hi (6 MMAs M128 N128 K16 per stage) into one fp32 accumulator in TMEM.
// The kernel is then bound by reading P and Q once (2 x N x K x 4 bytes).  Per-CTA partial sums go to
// the fp64 cost buffer with atomics, like the CUDA-core kernel.
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"

namespace bmm {
namespace {

constexpr int CT_ROWS = 32;                      // (wide) observations per stage
constexpr int CT_MAT = 16 * CT_ROWS * 16;        // [16 column groups of 8][32 obs][16 B] = 8 KB
constexpr int CT_STAGE = 4 * CT_MAT;             // LQ hi | LQ lo | P hi | P lo
constexpr int CT_NS = 3;                         // operand stages
constexpr int CT_NR = 3;                         // raw fp32 stages (cp.async ring)
constexpr int CT_SL = 2;                         // work items per producer thread and stage
constexpr int CT_RROW = 512 + 16;                // raw row pitch: 128 floats + 16 B so that 8 rows hit 32 distinct banks
constexpr int CT_RAW = 2 * CT_ROWS * CT_RROW;    // [P | Q][32 rows][528 B] = 33 KB
constexpr int CT_RAW_OFF = CT_NS * CT_STAGE;
constexpr int CT_BAR_OFF = CT_RAW_OFF + CT_NR * CT_RAW;
constexpr int CT_THREADS = 288;                  // warps 0-7 producers, warp 8 MMA issuer
constexpr int CT_SMEM = CT_BAR_OFF + 2 * CT_NS * 8 + 8 + 16;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(src) : "memory");
}

// 8 floats -> 8 fp16 hi (one 16-byte chunk) and 8 fp16 lo
__device__ __forceinline__ void split8(const float (&v)[8], uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const __half2 hh = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * q] - back.x, v[2 * q + 1] - back.y);
        h[q] = *(const uint32_t *)&hh;
        l[q] = *(const uint32_t *)&ll;
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__global__ void __launch_bounds__(CT_THREADS, 1) grid_cost_tc_kernel(long long N, int K, const float *__restrict__ P,
                                                                     const float *__restrict__ Q, int use_logp,
                                                                     double *out, int *status) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint64_t *bars = (uint64_t *)(smem + CT_BAR_OFF);             // full[NS], empty[NS], done
    uint32_t *tmem_slot = (uint32_t *)(bars + 2 * CT_NS + 1);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[CT_NS]), done = smem_u32(&bars[2 * CT_NS]);
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(128) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // column groups beyond K are never written: they stay zero (rows k >= K of the accumulator are unused)
    for (int e = tid; e < CT_NS * CT_STAGE / 16; e += CT_THREADS)   // operand stages only ((uint4 *)smem)[e] = make_uint4(0u, 0u, 0u, 0u);
    fence_async_smem();
    if (tid == 0) {
        for (int s = 0; s < CT_NS; ++s) { mbar_init(full0 + 8 * s, 256); mbar_init(empty0 + 8 * s, 1); }
        mbar_init(done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t acc = *tmem_slot;
    // K dividing 128: `fold` consecutive observations are read as one wide row of KW = 128 columns (the matrices
    // are row-major and contiguous), which keeps every column group of the operand images busy; the wide
    // contraction holds G in its `fold` diagonal K x K blocks, summed by the epilogue.
    const int fold = (128 % K) == 0 ? 128 / K : 1, KW = K * fold;
    const long long total = N * (long long)K, NWIDE = (N + fold - 1) / fold;
    const long long nst = (NWIDE + CT_ROWS - 1) / CT_ROWS;
    // contiguous range of stages per CTA
    const long long per = (nst + gridDim.x - 1) / gridDim.x;
    const long long st0 = (long long)blockIdx.x * per, st1 = min(st0 + per, nst);
    bool ok = true;

    const int ncg = KW >> 3;                                 // column groups of 8 (K % 8 == 0)
    if (warp < 8) {
        // ---- producers: work item = (row octet, column group); the 8 lanes of a quarter-warp phase take the 8
        //      rows of one item, so every 16-byte store phase covers 128 contiguous bytes of one column group,
        //      and consecutive items are consecutive column groups of the same rows (coalesced reads).
        //      Slot m of a thread is item (tid >> 3) + 32 m in every stage, hence a fixed column group. ----
        const int rl = lane & 7;
        int cgm[CT_SL], octm[CT_SL];
        bool act[CT_SL];
#pragma unroll
        for (int m = 0; m < CT_SL; ++m) {
            const int pi = (tid >> 3) + 32 * m;
            act[m] = pi < (CT_ROWS / 8) * ncg;
            cgm[m] = pi % ncg; octm[m] = pi / ncg;
        }
        float sacc[CT_SL][8];
#pragma unroll
        for (int m = 0; m < CT_SL; ++m)
#pragma unroll
            for (int q = 0; q < 8; ++q) sacc[m][q] = 0.f;
        // The rows travel global -> shared with cp.async, a warp instruction per 512-byte row (whole 32-byte
        // sectors, once: 16-byte pieces of a sector fetched by different instructions doubled the L2 traffic
        // and bound the first version), two stages ahead of their conversion; registers cannot hold that much
        // data in flight (a register double buffer spilled and was slower).
        const uint32_t raw0 = smem_u32(smem + CT_RAW_OFF);
        auto issue = [&](long long st, int r) {
            if (st < st1 && lane * 4 < KW) {
#pragma unroll
                for (int rr = 0; rr < CT_ROWS / 8; ++rr) {
                    const int row = warp + 8 * rr;
                    const long long e = (st * CT_ROWS + row) * KW + lane * 4;
                    if (e < total) {
                        const uint32_t d = raw0 + r * CT_RAW + row * CT_RROW + lane * 16;
                        cp_async16(d, P + e);
                        cp_async16(d + CT_ROWS * CT_RROW, Q + e);
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
#pragma unroll
        for (int r = 0; r < CT_NR - 1; ++r) issue(st0 + r, r);
        long long g = 0;
        for (long long st = st0; st < st1; ++st, ++g) {   // no early exit: every producer reaches every bar.sync
            // own copies of stage g have landed; after the barrier everybody's have, and everybody has read
            // stage g - 1 out of the slot that the next issue overwrites
            asm volatile("cp.async.wait_group %0;" :: "n"(CT_NR - 2) : "memory");
            asm volatile("bar.sync 1, 256;" ::: "memory");
            issue(st + CT_NR - 1, (int)((g + CT_NR - 1) % CT_NR));
            float4 pv[CT_SL][2], qv[CT_SL][2];
            {
                const unsigned char *rb = smem + CT_RAW_OFF + (int)(g % CT_NR) * CT_RAW;
#pragma unroll
                for (int m = 0; m < CT_SL; ++m) {
                    const int row = octm[m] * 8 + rl;
                    const long long e = (st * CT_ROWS + row) * KW + cgm[m] * 8;
                    if (act[m] && e < total) {
                        const unsigned char *src = rb + row * CT_RROW + cgm[m] * 32;
                        pv[m][0] = *(const float4 *)(src);
                        pv[m][1] = *(const float4 *)(src + 16);
                        qv[m][0] = *(const float4 *)(src + CT_ROWS * CT_RROW);
                        qv[m][1] = *(const float4 *)(src + CT_ROWS * CT_RROW + 16);
                    } else {
                        pv[m][0] = pv[m][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        qv[m][0] = qv[m][1] = make_float4(1.f, 1.f, 1.f, 1.f);
                    }
                }
            }
            const int s = (int)(g % CT_NS);
            const long long n = g / CT_NS;
            if (ok && n > 0) ok = mbar_wait(empty0 + 8 * s, (uint32_t)((n - 1) & 1));   // after a timeout: stop waiting
            unsigned char *stage = smem + s * CT_STAGE;
#pragma unroll
            for (int m = 0; m < CT_SL; ++m) {
                if (!act[m]) continue;
                const float p8[8] = {pv[m][0].x, pv[m][0].y, pv[m][0].z, pv[m][0].w, pv[m][1].x, pv[m][1].y, pv[m][1].z, pv[m][1].w};
                const float q8[8] = {qv[m][0].x, qv[m][0].y, qv[m][0].z, qv[m][0].w, qv[m][1].x, qv[m][1].y, qv[m][1].z, qv[m][1].w};
                float l8[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    l8[q] = __logf(q8[q]);
                    sacc[m][q] += use_logp ? (p8[q] > 0.f ? p8[q] * __logf(p8[q]) : 0.f) : p8[q] * p8[q];
                }
                uint4 hi, lo;
                const int off = cgm[m] * (CT_ROWS * 16) + (octm[m] * 8 + rl) * 16;
                split8(l8, hi, lo);
                *(uint4 *)(stage + 0 * CT_MAT + off) = hi;
                *(uint4 *)(stage + 1 * CT_MAT + off) = lo;
                split8(p8, hi, lo);
                *(uint4 *)(stage + 2 * CT_MAT + off) = hi;
                *(uint4 *)(stage + 3 * CT_MAT + off) = lo;
            }
            fence_async_smem();
            mbar_arrive(full0 + 8 * s);
        }
        // column sums: the 8 lanes of an item hold the same columns
#pragma unroll
        for (int m = 0; m < CT_SL; ++m)
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float v = sacc[m][q];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                if (rl == 0 && act[m]) atomicAdd(&out[(size_t)K * K + (cgm[m] * 8 + q) % K], (double)v);
            }
    } else if (tid == 256) {
        // ---- MMA issuer ----
        const uint32_t IDESC = umma_idesc_f16(128, (KW + 15) & ~15, 1, 1);   // N = width rounded up to the MMA granule
        long long g = 0;
        for (long long st = st0; st < st1 && ok; ++st, ++g) {
            const int s = (int)(g % CT_NS);
            ok = mbar_wait(full0 + 8 * s, (uint32_t)((g / CT_NS) & 1));
            if (!ok) break;
            tc_fence_after();
            const uint32_t b = smem_u32(smem + s * CT_STAGE);
#pragma unroll
            for (int kk = 0; kk < CT_ROWS / 16; ++kk) {
                const uint64_t ah = umma_desc(b + 0 * CT_MAT + kk * 256, 128, CT_ROWS * 16);
                const uint64_t al = umma_desc(b + 1 * CT_MAT + kk * 256, 128, CT_ROWS * 16);
                const uint64_t bh = umma_desc(b + 2 * CT_MAT + kk * 256, 128, CT_ROWS * 16);
                const uint64_t bl = umma_desc(b + 3 * CT_MAT + kk * 256, 128, CT_ROWS * 16);
                umma_f16(acc, ah, bh, IDESC, (g > 0 || kk > 0) ? 1u : 0u);
                umma_f16(acc, ah, bl, IDESC, 1u);
                umma_f16(acc, al, bh, IDESC, 1u);
            }
            umma_commit(empty0 + 8 * s);
        }
        umma_commit(done);
    }
    __syncwarp();
    // ---- epilogue: warps 0-3, TMEM lane = k, column = l ----
    if (warp < 4 && st0 < st1) {
        if (ok) ok = mbar_wait(done, 0u);
        if (ok) {
            tc_fence_after();
            const int kw = warp * 32 + lane;                 // wide row of G = TMEM lane
            for (int c0 = 0; c0 < KW; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(acc + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
                tmem_ld_wait();
                if (kw < KW) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        const int cw = c0 + q;
                        if (cw < KW && cw / K == kw / K)     // diagonal blocks only
                            atomicAdd(&out[(kw % K) + (size_t)K * (cw % K)], (double)__uint_as_float(v[q]));
                    }
                }
            }
        }
    }
    if (!ok) *status = -10;  // BMM_ERR_TIMEOUT
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(acc), "r"(128) : "memory");
    }
}

}  // namespace

bool grid_cost_tc_supported(long long N, int K) { return K >= 8 && K <= 128 && (K % 8) == 0 && N >= 1; }

// acc must be zero on entry
cudaError_t launch_grid_cost_tc(long long N, int K, const float *P, const float *Q, int use_logp, double *acc, int *status,
                                int sm_count, cudaStream_t st) {
    static FuncAttrCache attr;
    if (cudaError_t e = attr.ensure_smem(grid_cost_tc_kernel, CT_SMEM)) return e;
    const int fold = (128 % K) == 0 ? 128 / K : 1;
    const long long nst = ((N + fold - 1) / fold + CT_ROWS - 1) / CT_ROWS;
    const int grid = (int)(nst < sm_count ? nst : sm_count);
    grid_cost_tc_kernel<<<grid, CT_THREADS, CT_SMEM, st>>>(N, K, P, Q, use_logp, acc, status);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace bmm
