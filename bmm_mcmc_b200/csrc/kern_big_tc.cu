// Tensor-core z-sweep of the grid path (BMM_FP32): the N x K Bernoulli log-likelihood of
// /root/reference/src/full_gibbs.cpp:92-106 (stickbreaking.cpp:75-92) as a tcgen05 contraction, and the
// sufficient statistics of full_gibbs.cpp:182-200 (stickbreaking.cpp:164-186) as a second one, fused
// in one persistent kernel so every bit-packed row is read from HBM exactly once per sweep.
//
//   loglh_k(x_i) = sum_d x_id (log th_kd - log(1-th_kd)) + sum_d log(1-th_kd)  =  (X D^T)_ik + b_k
//   GEMM1  [128 obs x Pd] (bf16 0/1, K-major)  x  [3*KC x Pd]^T: D (in log2 units) split into hi, mid,
//          lo bf16 terms side by side, fp32 accumulate in TMEM -> 128 x 3*KC; thread i owns TMEM lane i.
//   epilogue (one observation per thread): logits = acc + b_k + log pi_k, max, exp2, running sum,
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
#include "umma.cuh"

namespace bmm {
namespace {

constexpr int TC_TILE = 128;
constexpr int TC_CHUNK = TC_TILE * 16;  // bytes of one 16-byte-chunk column over 128 rows
constexpr int TC_ACHUNKS = 16;          // A region: 128 "d" positions = UMMA M of GEMM2

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
    // split table: [d-chunk][row = part*KC + k][16 B], parts = hi, mid, lo
    __host__ __device__ static constexpr int bias_off(int nch) { return B1_OFF + nch * B1_ROW; }
    __host__ __device__ static constexpr int bar_off(int nch) { return bias_off(nch) + KC * 4; }
    __host__ __device__ static constexpr int total(int nch) { return bar_off(nch) + 2 * NWG * 8 + 16; }
};

template <int KC, int NWG, int NCH>
__global__ void __launch_bounds__(NWG * 128, 1) big_sweep_tc_kernel(const BigParams p, const int j) {
    using L = TcLayout<KC, NWG>;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, wg = tid >> 7, t = tid & 127, wq = (tid >> 5) & 3;
    const int K = p.K, P = p.P, W = p.W;
    constexpr int ONES = NCH * 8;                  // d index of the all-ones column of [X | 1]
    constexpr int NW = (NCH + 3) / 4;              // 32-bit words of a packed row that carry data
    unsigned char *A = smem + wg * L::WG_BYTES;
    unsigned char *B2 = A + L::A_BYTES;
    unsigned char *B1 = smem + L::B1_OFF;
    float *bias = (float *)(smem + L::bias_off(NCH));
    uint64_t *bars = (uint64_t *)(smem + L::bar_off(NCH));
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
    // log2(e) * (log th_kd - log(1 - th_kd)) split into three bf16 terms (24 mantissa bits); the
    // accumulator is then already in log2 units and the epilogue needs no multiply before exp2.
    for (int e = tid; e < KC * NCH * 8; e += NWG * 128) {
        const int k = e % KC, d = e / KC;
        double D = 0.0;
        if (k < K && d < P) D = (p.w1[k + K * d] - p.w0[k + K * d]) * 1.4426950408889634;
        D = fmin(fmax(D, -1.0e4), 1.0e4);     // theta exactly 0 / 1: log 0 = -inf would make 0 * inf = NaN in the contraction
        if (D != D) D = 0.0;
        const __nv_bfloat16 hi = __double2bfloat16(D);
        const double r1 = D - (double)__bfloat162float(hi);
        const __nv_bfloat16 mid = __double2bfloat16(r1);
        const __nv_bfloat16 lo = __double2bfloat16(r1 - (double)__bfloat162float(mid));
        unsigned char *cell = B1 + (d >> 3) * L::B1_ROW + k * 16 + (d & 7) * 2;
        *(__nv_bfloat16 *)(cell + 0 * KC * 16) = hi;
        *(__nv_bfloat16 *)(cell + 1 * KC * 16) = mid;
        *(__nv_bfloat16 *)(cell + 2 * KC * 16) = lo;
    }
    for (int k = tid; k < KC; k += NWG * 128) {
        float b = -INFINITY;
        if (k < K) {
            double s0 = 0.0;
            for (int d = 0; d < P; ++d) s0 += p.w0[k + K * d];
            const double bb = (p.lpi[k] + s0) * 1.4426950408889634;
            b = bb == bb ? (float)fmax(bb, -3.0e38) : -INFINITY;
        }
        bias[k] = b;
    }
    // chunks NCH..15 of this warpgroup's A region never change: zero, except the ones column
#pragma unroll
    for (int c = NCH; c < TC_ACHUNKS; ++c)
        *(uint4 *)(A + c * TC_CHUNK + t * 16) = make_uint4(c == NCH ? 0x3F80u : 0u, 0u, 0u, 0u);
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
    const long long tstride = (long long)gridDim.x * NWG;
    bool ok = true;
    int it = 0;

    auto load_row = [&](long long tile, uint32_t (&xw)[NW]) {
        const long long i = tile * TC_TILE + t;
#pragma unroll
        for (int w = 0; w < NW; ++w) xw[w] = 0u;
        if (tile < ntiles && i < p.N_local) {
            const uint32_t *xb = p.xbits + (size_t)i * W;
            if (NW == 2 && W == 2) {
                const uint2 v = *(const uint2 *)xb;
                xw[0] = v.x; xw[NW - 1] = v.y;
            } else {
#pragma unroll
                for (int w = 0; w < NW; ++w) if (w < W) xw[w] = xb[w];
            }
        }
    };

    long long tile = (long long)blockIdx.x * NWG + wg;
    uint32_t xw[NW];
    load_row(tile, xw);
    for (; tile < ntiles && ok; tile += tstride, ++it) {
        const long long i = tile * TC_TILE + t;
        const bool valid = i < p.N_local;
        const unsigned long long gi = (unsigned long long)p.row_offset + (unsigned long long)i;
        // bits -> bf16 in registers (overlaps the previous tile's GEMM2), then wait until that GEMM2 has
        // finished reading A / B2 before overwriting them
        uint4 ex[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const uint32_t byte = xw[c >> 2] >> ((c & 3) * 8);
            ex[c] = make_uint4(bits2_bf16x2(byte), bits2_bf16x2(byte >> 2), bits2_bf16x2(byte >> 4), bits2_bf16x2(byte >> 6));
        }
        if (it > 0) ok = mbar_wait(bar2, (uint32_t)((it - 1) & 1));
        if (!ok) break;
#pragma unroll
        for (int c = 0; c < NCH; ++c) *(uint4 *)(A + c * TC_CHUNK + t * 16) = ex[c];
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(A), b0 = smem_u32(B1);
            // one pass: the hi | mid | lo terms land in three groups of KC accumulator columns (a small-N
            // tcgen05.mma costs ~50 cycles whatever N is, so 3 x fewer instructions beat 3 x narrower ones)
#pragma unroll
            for (int kk = 0; kk < NCH / 2; ++kk)
                umma_f16(acc1, umma_desc(a0 + kk * 2 * TC_CHUNK, TC_CHUNK, 128),
                          umma_desc(b0 + kk * 2 * L::B1_ROW, L::B1_ROW, 128), IDESC1, kk ? 1u : 0u);
            umma_commit(bar1);
        }
        load_row(tile + tstride, xw);   // prefetch the next tile's row; used at the top of the next iteration
        // this observation's uniform (same counters as the other uncollapsed kernels)
        const uint4 rnd = philox4x32_10(make_uint4((uint32_t)(gi >> 2), (uint32_t)(gi >> 34), sid, (uint32_t)j), key);
        // top 24 bits of the 32-bit word the fp64 kernels use for this observation
        const float u = u32_unit_f(philox_word(rnd, (int)(gi & 3)));
        ok = mbar_wait(bar1, (uint32_t)(it & 1));
        if (!ok) break;
        tc_fence_after();
        float l[KC];
#pragma unroll
        for (int c0 = 0; c0 < KC; c0 += 32) {
#pragma unroll
            for (int part = 0; part < 3; ++part) {
                uint32_t v[32];
                tmem_ld32(acc1 + lane_sel + (uint32_t)(part * KC + c0), v);
                tmem_ld_wait();
#pragma unroll
                for (int q = 0; q < 32; ++q)
                    l[c0 + q] = part == 0 ? __uint_as_float(v[q]) + bias[c0 + q] : l[c0 + q] + __uint_as_float(v[q]);
            }
        }
        float mx = l[0];
#pragma unroll
        for (int k = 1; k < KC; ++k) mx = fmaxf(mx, l[k]);
        // l[k] becomes the running sum of exp2(logit - max): the inverse CDF up to a factor
        float run = 0.f;
        if (p.probs_out == nullptr && p.probs_f32 == nullptr) {
#pragma unroll
            for (int k = 0; k < KC; ++k) { run += ex2_ftz(l[k] - mx); l[k] = run; }
        } else {             // relabelling / probe: conditional probabilities of this sweep
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < KC; ++k) { l[k] = ex2_ftz(l[k] - mx); sum += l[k]; }
            const float inv = 1.f / sum;
            if (valid && p.probs_out) {
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k < K) p.probs_out[(size_t)j * p.N_local * K + i + (size_t)p.N_local * k] = (double)(l[k] * inv);
            }
            if (valid && p.probs_f32) {
#pragma unroll
                for (int k = 0; k < KC; ++k)
                    if (k < K) p.probs_f32[(size_t)i * K + k] = l[k] * inv;
            }
#pragma unroll
            for (int k = 0; k < KC; ++k) { run += l[k]; l[k] = run; }
        }
        if (!(run > 0.f) || !isfinite(run)) *p.status = -9;  // BMM_ERR_PROB
        const float target = u * run;
        int z = 0;
#pragma unroll
        for (int k = 0; k < KC; ++k) z += (l[k] <= target) ? 1 : 0;
        z = min(z, K - 1);
        if (valid && zrow) zrow[i] = (uint8_t)(z + 1);
        // one-hot row of the allocation, MN-major: chunk z/8 holds 1.0 at element z%8
        {
            const uint32_t h = valid ? ((z & 1) ? 0x3F800000u : 0x3F80u) : 0u;
            const int wsel = (z & 7) >> 1, csel = z >> 3;
            const uint4 hot = make_uint4(wsel == 0 ? h : 0u, wsel == 1 ? h : 0u, wsel == 2 ? h : 0u, wsel == 3 ? h : 0u);
            const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int cc = 0; cc < KC / 8; ++cc) *(uint4 *)(B2 + cc * TC_CHUNK + t * 16) = (csel == cc) ? hot : zero;
        }
        fence_async_smem();
        tc_fence_before();
        wg_barrier(wg);
        if (t == 0) {
            tc_fence_after();
            const uint32_t a0 = smem_u32(A), b0 = smem_u32(B2);
#pragma unroll
            for (int kk = 0; kk < TC_TILE / 16; ++kk)
                umma_f16(acc2, umma_desc(a0 + kk * 256, 128, TC_CHUNK), umma_desc(b0 + kk * 256, 128, TC_CHUNK),
                          IDESC2, (it > 0 || kk > 0) ? 1u : 0u);
            umma_commit(bar2);
        }
    }
    // ---- flush the counts: TMEM lane d of each warpgroup holds V_kd (d < P) or c_k (d == ONES);
    //      the warpgroups are summed in shared memory first, then one global atomic per entry and CTA
    if (ok && it > 0) ok = mbar_wait(bar2, (uint32_t)((it - 1) & 1));
    if (!ok) *p.status = -10;  // BMM_ERR_TIMEOUT: an mbarrier never completed
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    int *scratch = (int *)smem;                    // A region of warpgroup 0, free now
    const int ncnt = K + K * P;
    for (int e = tid; e < ncnt; e += NWG * 128) scratch[e] = 0;
    __syncthreads();
    if (ok && it > 0) {
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
                    if (k < K && n) atomicAdd(t == ONES ? &scratch[k] : &scratch[K + k + K * t], n);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    int *gcnt = p.counts + (size_t)(j & 1) * ncnt;
    for (int e = tid; e < ncnt; e += NWG * 128) {
        const int n = scratch[e];
        if (n) atomicAdd(&gcnt[e], n);
    }
    if (tid < 32) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(L::TMEM_COLS) : "memory");
    }
}

template <int KC, int NWG, int NCH>
cudaError_t launch_tc_nch(const BigParams &p, int j, int sm_count, cudaStream_t st) {
    using L = TcLayout<KC, NWG>;
    const size_t smem = (size_t)L::total(NCH);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(big_sweep_tc_kernel<KC, NWG, NCH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const long long ntiles = ((long long)p.N_local + TC_TILE - 1) / TC_TILE;
    long long ctas = (ntiles + NWG - 1) / NWG;
    if (ctas > sm_count) ctas = sm_count;
    if (ctas < 1) ctas = 1;
    big_sweep_tc_kernel<KC, NWG, NCH><<<(int)ctas, NWG * 128, smem, st>>>(p, j);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace

// The tensor-core sweep covers the float path for K <= 32 clusters and P <= 112 variables.
bool big_tc_supported(const BigParams &p) {
    return p.precision == 1 && p.K <= 32 && p.P <= 112 && p.W <= 4 && p.ru == nullptr && p.loglik_out == nullptr;
}

cudaError_t launch_big_sweep_tc(const BigParams &p, int j, int sm_count, cudaStream_t st) {
    constexpr int KC = 32, NWG = 4;
    switch ((p.P + 15) / 16) {
        case 1: return launch_tc_nch<KC, NWG, 2>(p, j, sm_count, st);
        case 2: return launch_tc_nch<KC, NWG, 4>(p, j, sm_count, st);
        case 3: return launch_tc_nch<KC, NWG, 6>(p, j, sm_count, st);
        case 4: return launch_tc_nch<KC, NWG, 8>(p, j, sm_count, st);
        case 5: return launch_tc_nch<KC, NWG, 10>(p, j, sm_count, st);
        case 6: return launch_tc_nch<KC, NWG, 12>(p, j, sm_count, st);
        default: return launch_tc_nch<KC, NWG, 14>(p, j, sm_count, st);
    }
}

}  // namespace bmm
