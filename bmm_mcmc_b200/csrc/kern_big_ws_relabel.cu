// Online Stephens relabelling of the grid path in ONE pass over Q per sweep, for the tensor-path shapes
// (K <= 32, P <= 112, BASELINE config C4 with relabelling).
//
// The reference's step (/root/reference/src/stephens.cpp:66-94) per sweep j >= burnin is
//     C_j(k,l) = sum_i p_il (p_il - log q_ik),  perm_j = assignment(C_j),  Q_j = j (Q_{j-1} + p_j[:, perm_j]) / (j + 1)
// with p_j the N x K conditional probabilities of sweep j (full_gibbs.cpp:146-156).  Stored and streamed literally
// (kern_big_relabel.cu) that is three passes over N x K floats per sweep: P written by the sweep kernel, P and Q read
// by the cost kernel, Q and P read and Q written by the update -- 7.7 GB per sweep at N = 1e7, K = 32 against 2.65 GB
// for "Q once in, once out" (SURVEY 8d).  Here nothing but Q and the packed rows touches HBM:
//   * P is never stored.  p_j is a function of the packed row and of sweep j's parameters, so it is RECOMPUTED from
//     the row with the same tcgen05 contraction as the sweep kernel (kern_big_ws.cu): X D^T with the fp16 hi | lo
//     operand image of the log-odds table the update kernel writes.
//   * The Q update is DEFERRED into the next sweep's pass: perm_j only exists after the whole cost matrix has been
//     reduced, so pass j applies update j - 1 -- it recomputes p_{j-1} as well (one contraction with N = 128:
//     [cur hi | cur lo | prev hi | prev lo] x 32 clusters), with the rows of the previous table taken in perm_{j-1}
//     order so that column k of the accumulator IS p_{j-1}[:, perm(k)], forms Q_{j-1} in registers, writes it back
//     and uses its logarithm for C_j.  A last pass after the final sweep applies the pending update.
//   * The K x K contraction G(k,l) = sum_i log q_ik p_il runs on tcgen05 as well: four observations form one 128-wide
//     row (warp w of a tile supplies column block w), so the 128 x 128 accumulator holds G in its four diagonal
//     32 x 32 blocks; operands are fp16 hi + lo, three products hi*hi + hi*lo + lo*hi.
//   * The column constant s_l = sum_i p_il^2 (p log p in the fixed mode) of C(k,l) = s_l - G(k,l) is left out: it adds
//     the same amount to every assignment, so the optimal permutation is the one of -G.
// Q lives in a tiled layout owned by this kernel, [tile of 128 rows][8 chunks of 4 clusters][128 rows][16 B]: thread t
// of a tile reads and writes chunk c at c * 2048 + t * 16, i.e. coalesced 16-byte accesses with the row in
// registers (TMEM lane = observation).  The first pass reads the row-major Q of the batch step, the last one writes it.
//
//   warps 0-3   producers: packed row -> fp16 A stage
//   warp  4     GEMM1 issuer (one thread): [128 x P] . [P x 128] -> accumulator k % NA
//   warp  5     cost-MMA issuer (one thread): 6 x (M128 N128 K16) per tile from the operand stage k % NB
//   warps 8-    epilogue warpgroups (tile k -> warpgroup k % NEPI)
#include <cstdlib>
#include <cuda_fp16.h>

#include "common.cuh"
#include "kernels.h"
#include "umma.cuh"
#include "ws_table.cuh"

namespace bmm {
namespace {

// Epilogue warpgroups for P <= 64.  Measured at C4 (N = 1e7, P = 64, K = 32), kernel alone under ncu / 40-sweep step:
// three warpgroups (96 registers, three cost stages, three Q tiles in flight) 573 us / 42.1 ms; two warpgroups (110
// registers, two cost stages, six Q tiles) 667 us / 42.8 ms.
#ifndef WSR_NEPI_CFG
#define WSR_NEPI_CFG 3
#endif
constexpr int WSR_NA = 6;                 // GEMM1 accumulators, 64 TMEM columns each (cur | prev); the cost accumulator takes the last 128
constexpr int WSR_QTILE = 128 * WS_KC * 4;
constexpr int WSR_CHUNK = 2048;           // [128 rows][16 B]
constexpr int WSR_MAT = 16 * 32 * 16;     // one cost operand image: [16 column groups][32 wide rows][16 B] = 8 KB
constexpr int WSR_STAGE = 4 * WSR_MAT;    // LQ hi | LQ lo | P hi | P lo
constexpr int WSR_B_ROW = 2 * WS_KC * 16; // bytes per feature chunk of one weight image: [cur | prev] x 32 clusters x 16 B

// Shared memory: [A ring][cost operand stages][Q ring][weight images][bias][barriers][tmem slot].  What is left after the
// operand rings goes to the Q ring: a tile of Q is 16 KB and the SM needs ~30 KB per microsecond of HBM latency in flight to
// keep its share of the bandwidth (the first version, three tiles deep, had its epilogue warps waiting for Q).
template <int NCH>
struct WsrLayout {
    static constexpr int A_STAGE = NCH * WSR_CHUNK;
    // Ring depths are multiples of the number of epilogue warpgroups: tile k - depth then belongs to the warpgroup of tile k,
    // which has finished it, so a parity wait can never be two phases ahead of its barrier (a warpgroup that runs ahead of a
    // slot another warpgroup still owes would pass the wait on the aliased parity -- seen as time-outs at 528 tiles per CTA
    // with two cost stages under three warpgroups).
    static constexpr int NEPI = NCH > 8 ? 2 : WSR_NEPI_CFG;   // epilogue warpgroups
    static constexpr int THREADS = 256 + 128 * NEPI;
    static constexpr int NS = NCH > 8 ? 2 : 3;     // A stages: held only until the contraction of their tile has run
    static constexpr int NB = NEPI;                // cost operand stages (32 KB each)
    static constexpr int FIXED = NS * A_STAGE + NB * WSR_STAGE + 2 * NCH * WSR_B_ROW + 2 * WS_KC * 4 + 64 * 8 + 16;
    static constexpr int NQ_FIT = (226 * 1024 - FIXED) / WSR_QTILE / NEPI * NEPI;
    static constexpr int NQ = NQ_FIT > 6 ? 6 : NQ_FIT;    // Q tiles in flight (one bulk copy per tile)
    static constexpr int ST_OFF = NS * A_STAGE;
    static constexpr int QR_OFF = ST_OFF + NB * WSR_STAGE;
    static constexpr int B1_OFF = QR_OFF + NQ * WSR_QTILE;
    static constexpr int BIAS_OFF = B1_OFF + 2 * NCH * WSR_B_ROW;
    static constexpr int BAR_OFF = BIAS_OFF + 2 * WS_KC * 4;
    static constexpr int NBAR = 2 * NS + 2 * WSR_NA + 2 * NB + 2 * NQ + 1;
    static constexpr int TOTAL = BAR_OFF + NBAR * 8 + 16;
    static_assert(NQ >= 2 && NQ % NEPI == 0 && NB % NEPI == 0 && WSR_NA % NEPI == 0, "ring depths must be multiples of NEPI");
    static_assert(NBAR <= 64, "barrier block");
};

// 8 floats (four packed pairs) -> 8 fp16 hi (one 16-byte cell) and 8 fp16 lo
__device__ __forceinline__ void split8(const unsigned long long *v2, uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float a, b;
        f2_unpack(v2[q], a, b);
        const __half2 hh = __floats2half2_rn(a, b);
        const float2 back = __half22float2(hh);
        float ra, rb;
        f2_unpack(f2_add(v2[q], f2_pack(-back.x, -back.y)), ra, rb);
        const __half2 ll = __floats2half2_rn(ra, rb);
        h[q] = *(const uint32_t *)&hh;
        l[q] = *(const uint32_t *)&ll;
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// logits of one part (32 clusters) of an accumulator (hi and lo terms were summed by the contraction) + bias, then the
// softmax numerators e_k = 2^(l_k - max) as packed pairs in e2[] and the inverse of their sum; two fp32 lanes per
// instruction throughout (FADD2)
__device__ __forceinline__ float softmax32(uint32_t taddr, const float *bias, unsigned long long (&e2)[WS_KC / 2]) {
    float l[WS_KC];
    {
        uint32_t v[32];
        tmem_ld32(taddr, v);
        tmem_ld_wait();
        const float2 *b2 = (const float2 *)bias;
#pragma unroll
        for (int q = 0; q < WS_KC / 2; ++q) {
            const float2 bq = b2[q];
            e2[q] = f2_add(f2_pack(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])), f2_pack(bq.x, bq.y));
            f2_unpack(e2[q], l[2 * q], l[2 * q + 1]);
        }
    }
    float mx;
    {
        float m[10];
#pragma unroll
        for (int q = 0; q < 10; ++q) m[q] = fmax3(l[3 * q], l[3 * q + 1], l[3 * q + 2]);
        const float m0 = fmax3(m[0], m[1], m[2]), m1 = fmax3(m[3], m[4], m[5]), m2 = fmax3(m[6], m[7], m[8]);
        mx = fmaxf(fmax3(m0, m1, m2), fmax3(m[9], l[30], l[31]));
    }
    const unsigned long long nmx = f2_pack(-mx, -mx);
    unsigned long long s0 = f2_pack(0.f, 0.f), s1 = s0;
#pragma unroll
    for (int q = 0; q < WS_KC / 2; q += 2) {
        float a, b, c, d;
        f2_unpack(f2_add(e2[q], nmx), a, b);
        f2_unpack(f2_add(e2[q + 1], nmx), c, d);
        e2[q] = f2_pack(ex2_ftz(a), ex2_ftz(b));
        e2[q + 1] = f2_pack(ex2_ftz(c), ex2_ftz(d));
        s0 = f2_add(s0, e2[q]);
        s1 = f2_add(s1, e2[q + 1]);
    }
    float sa, sb;
    f2_unpack(f2_add(s0, s1), sa, sb);
    return 1.f / (sa + sb);
}

template <int NCH>
__global__ void __launch_bounds__(WsrLayout<NCH>::THREADS, 1) big_relabel_ws_kernel(const WsRelabelParams p) {
    using L = WsrLayout<NCH>;
    constexpr int NS = L::NS, WSR_NB = L::NB, WSR_NQ = L::NQ, WSR_NEPI = L::NEPI, WSR_THREADS = L::THREADS;
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int K = p.K, W = p.W;
    constexpr int NW = (NCH + 3) / 4;
    unsigned char *B1 = smem + L::B1_OFF;
    float *bias = (float *)(smem + L::BIAS_OFF);          // [0, 32) current sweep, [32, 64) pending update (permuted)
    uint64_t *bars = (uint64_t *)(smem + L::BAR_OFF);
    uint32_t *tmem_slot = (uint32_t *)(bars + L::NBAR);
    const uint32_t full_a = smem_u32(bars), free_a = full_a + 8 * NS;
    const uint32_t acc_full = free_a + 8 * NS, acc_free = acc_full + 8 * WSR_NA;
    const uint32_t st_full = acc_free + 8 * WSR_NA, st_free = st_full + 8 * WSR_NB;
    const uint32_t q_full = st_free + 8 * WSR_NB, q_free = q_full + 8 * WSR_NQ;
    const uint32_t all_done = q_free + 8 * WSR_NQ;
    const bool do_cost = p.do_cost != 0, upd = p.upd != 0;

    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 192) {
        for (int s = 0; s < NS; ++s) { mbar_init(full_a + 8 * s, 128); mbar_init(free_a + 8 * s, 1); }
        for (int a = 0; a < WSR_NA; ++a) { mbar_init(acc_full + 8 * a, 1); mbar_init(acc_free + 8 * a, 128); }
        for (int b = 0; b < WSR_NB; ++b) { mbar_init(st_full + 8 * b, 128); mbar_init(st_free + 8 * b, 1); }
        for (int s = 0; s < WSR_NQ; ++s) { mbar_init(q_full + 8 * s, 1); mbar_init(q_free + 8 * s, 128); }
        mbar_init(all_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- weight images, 64 rows per feature chunk: [cur | prev] x 32 clusters, the hi terms in one image and the lo terms
    //      in a second one behind it (both are contracted into the same accumulator); the rows of the previous table are
    //      taken in the order of the pending permutation ----
    const unsigned char *prev_img = upd ? p.b1_hist + (size_t)p.hist_prev * NCH * WS_B1_ROW : p.b1_cur;
    const float *prev_bias = p.bias_hist + 32 * p.hist_prev;
    for (int e = tid; e < NCH * 128; e += WSR_THREADS) {
        const int c = e >> 7, r = e & 127, lo_part = r >> 6, prev = (r >> 5) & 1, k = r & 31;
        const int ks = (prev && upd && k < K) ? p.perm[k] : k;
        const unsigned char *src = (prev ? prev_img : p.b1_cur) + (size_t)c * WS_B1_ROW + lo_part * (WS_KC * 16) + ks * 16;
        *(uint4 *)(B1 + lo_part * (NCH * WSR_B_ROW) + c * WSR_B_ROW + (r & 63) * 16) = __ldcg((const uint4 *)src);
    }
    for (int k = tid; k < 2 * WS_KC; k += WSR_THREADS) {
        float b = -INFINITY;
        const int kk = k & 31;
        if (kk < K) {
            if (k < WS_KC || !upd) {
                const double bb = (p.lpi[kk] + p.s0[kk]) * 1.4426950408889634;
                b = bb == bb ? (float)fmax(bb, -3.0e38) : -INFINITY;
            } else b = prev_bias[p.perm[kk]];
        }
        bias[k] = b;
    }
    // this sweep's table and bias become the "previous" ones of the next pass (CTA 0 keeps the copy; the other parity is
    // the one everybody reads in this launch)
    if (blockIdx.x == 0 && p.hist_next >= 0) {
        for (int e = tid; e < NCH * WS_B1_ROW / 16; e += WSR_THREADS)
            ((uint4 *)(p.b1_hist + (size_t)p.hist_next * NCH * WS_B1_ROW))[e] = __ldcg((const uint4 *)p.b1_cur + e);
        for (int k = tid; k < WS_KC; k += WSR_THREADS) {
            float b = -INFINITY;
            if (k < K) {
                const double bb = (p.lpi[k] + p.s0[k]) * 1.4426950408889634;
                b = bb == bb ? (float)fmax(bb, -3.0e38) : -INFINITY;
            }
            p.bias_hist[32 * p.hist_next + k] = b;
        }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc_cost = tmem_base + WSR_NA * 2 * WS_KC;
    const long long ntiles = ((long long)p.N_local + 127) / 128;
    const int T = blockIdx.x < ntiles ? (int)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    bool ok = true;

    if (warp < 4) {
        // ================= producers (as in kern_big_ws.cu) =================
        const int t = tid;
        auto load_row = [&](long long k, uint32_t (&xw)[NW]) {
            const long long i = (blockIdx.x + k * gridDim.x) * 128 + t;
#pragma unroll
            for (int w = 0; w < NW; ++w) xw[w] = 0u;
            if (k < T && i < p.N_local) {
                const uint32_t *xb = p.xbits + (size_t)i * W;
                if (NW == 2 && W == 2) { const uint2 v = *(const uint2 *)xb; xw[0] = v.x; xw[NW - 1] = v.y; }
                else {
#pragma unroll
                    for (int w = 0; w < NW; ++w) if (w < W) xw[w] = xb[w];
                }
            }
        };
        constexpr int PF = 4;
        uint32_t ring[PF][NW];
#pragma unroll
        for (int d = 0; d < PF; ++d) load_row(d, ring[d]);
        int s = 0;
        uint32_t ph = 0;
        bool first_pass = true;
        for (int k0 = 0; k0 < T && ok; k0 += PF) {
#pragma unroll
            for (int d = 0; d < PF; ++d) {
                const int k = k0 + d;
                if (k >= T || !ok) break;
                uint4 ex[NCH];
#pragma unroll
                for (int c = 0; c < NCH; ++c) {
                    const uint32_t y = __byte_perm(ring[d][c >> 2], 0u, 0x4440u + (c & 3)) * 0x8001u;
                    ex[c] = make_uint4((y & 0x10001u) * 0x3C00u, ((y >> 2) & 0x10001u) * 0x3C00u,
                                       ((y >> 4) & 0x10001u) * 0x3C00u, ((y >> 6) & 0x10001u) * 0x3C00u);
                }
                load_row(k + PF, ring[d]);
                if (!first_pass) ok = mbar_wait(free_a + 8 * s, ph ^ 1u);
                if (!ok) break;
                unsigned char *A = smem + s * L::A_STAGE;
#pragma unroll
                for (int c = 0; c < NCH; ++c) *(uint4 *)(A + c * WSR_CHUNK + t * 16) = ex[c];
                fence_async_smem();
                mbar_arrive(full_a + 8 * s);
                if (++s == NS) { s = 0; ph ^= 1u; first_pass = false; }
            }
        }
    } else if (warp == 4) {
        // ================= GEMM1 issuer =================
        // all 32 lanes walk the loop, one elected lane issues (umma.cuh::elect_one)
        {
            constexpr uint32_t IDESC1 = umma_idesc_f16(128, 2 * WS_KC, 0, 0);
            const uint64_t da0 = umma_desc(smem_u32(smem), WSR_CHUNK, 128), db0 = umma_desc(smem_u32(B1), WSR_B_ROW, 128);
            int s = 0, a = 0;
            uint32_t ph_s = 0, ph_a = 0;
            for (int k = 0; k < T; ++k) {
                ok = mbar_wait(full_a + 8 * s, ph_s);
                if (ok && k >= WSR_NA) ok = mbar_wait(acc_free + 8 * a, ph_a ^ 1u);
                ok = __all_sync(0xffffffffu, ok);
                if (!ok) break;
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t da = da0 + (uint64_t)((s * L::A_STAGE) >> 4);
                    // X . hi^T + X . lo^T into one accumulator: the A stage is read twice, the two terms never meet in registers
#pragma unroll
                    for (int kk = 0; kk < NCH; ++kk) {
                        const int ka = kk % (NCH / 2);
                        umma_f16(tmem_base + (uint32_t)(a * 2 * WS_KC), da + (uint64_t)((ka * 2 * WSR_CHUNK) >> 4),
                                  db0 + (uint64_t)(((kk / (NCH / 2)) * NCH * WSR_B_ROW + ka * 2 * WSR_B_ROW) >> 4), IDESC1, kk ? 1u : 0u);
                    }
                    umma_commit(acc_full + 8 * a);
                    umma_commit(free_a + 8 * s);        // the A stage is only read by this contraction
                }
                __syncwarp();
                if (++s == NS) { s = 0; ph_s ^= 1u; }
                if (++a == WSR_NA) { a = 0; ph_a ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp == 5) {
        // ================= cost-MMA issuer =================
        if (do_cost) {
            constexpr uint32_t IDESC2 = umma_idesc_f16(128, 128, 1, 1);
            int b = 0;
            uint32_t ph_b = 0;
            for (int q = 0; q < T; ++q) {
                ok = __all_sync(0xffffffffu, mbar_wait(st_full + 8 * b, ph_b));
                if (!ok) break;
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t sb = smem_u32(smem + L::ST_OFF + b * WSR_STAGE);
#pragma unroll
                    for (int kk = 0; kk < 2; ++kk) {
                        const uint64_t ah = umma_desc(sb + 0 * WSR_MAT + kk * 256, 128, 512);
                        const uint64_t al = umma_desc(sb + 1 * WSR_MAT + kk * 256, 128, 512);
                        const uint64_t bh = umma_desc(sb + 2 * WSR_MAT + kk * 256, 128, 512);
                        const uint64_t bl = umma_desc(sb + 3 * WSR_MAT + kk * 256, 128, 512);
                        umma_f16(acc_cost, ah, bh, IDESC2, (q > 0 || kk > 0) ? 1u : 0u);
                        umma_f16(acc_cost, ah, bl, IDESC2, 1u);
                        umma_f16(acc_cost, al, bh, IDESC2, 1u);
                    }
                    umma_commit(st_free + 8 * b);
                }
                __syncwarp();
                if (++b == WSR_NB) { b = 0; ph_b ^= 1u; }
            }
            if (elect_one()) umma_commit(all_done);
        }
        __syncwarp();
    } else if (warp == 6) {
        // ================= Q loader: one 16 KB bulk copy per tile (the tiled layout makes a tile contiguous) =================
        if (tid == 192 && !p.q_in_rowmajor && !p.q_direct) {
            int s = 0;
            uint32_t ph = 0;
            for (int k = 0; k < T && ok; ++k) {
                if (k >= WSR_NQ) ok = mbar_wait(q_free + 8 * s, ph ^ 1u);
                if (!ok) break;
                const long long tile = (long long)blockIdx.x + (long long)k * gridDim.x;
                mbar_arrive_expect_tx(q_full + 8 * s, WSR_QTILE);
                bulk_g2s(smem_u32(smem + L::QR_OFF + s * WSR_QTILE), p.Q_tiled + (size_t)tile * (128 * WS_KC), WSR_QTILE, q_full + 8 * s);
                if (++s == WSR_NQ) { s = 0; ph ^= 1u; }
            }
        }
        __syncwarp();
    } else if (warp >= 8) {
        // ================= epilogue warpgroups =================
        const int e = (warp - 8) >> 2, t = (tid - 256) & 127, wq = warp & 3, lane = tid & 31;
        const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
        const float cq = p.cq, cp = p.cp;
        for (int k = e; k < T && ok; k += WSR_NEPI) {
            // tile k uses accumulator k % NA, cost stage k % NB, Q stage k % NQ; the n-th use of a ring slot waits for phase n & 1
            const int a = k % WSR_NA, b = k % WSR_NB, sq = k % WSR_NQ;
            const uint32_t ph_a = (uint32_t)((k / WSR_NA) & 1), ph_b = (uint32_t)((k / WSR_NB) & 1), ph_q = (uint32_t)((k / WSR_NQ) & 1);
            const long long tile = (long long)blockIdx.x + (long long)k * gridDim.x;
            const long long i = tile * 128 + t;
            const bool valid = i < p.N_local;
            ok = mbar_wait(acc_full + 8 * a, ph_a);
            if (!ok) break;
            tc_fence_after();
            const uint32_t tacc = tmem_base + lane_sel + (uint32_t)(a * 2 * WS_KC);
            unsigned long long e2[WS_KC / 2];
            float cpi = 0.f;
            // p_{j-1}[:, perm]: softmax of the permuted previous table's logits
            if (upd) cpi = cp * softmax32(tacc + WS_KC, bias + WS_KC, e2);
            // this observation's row of Q (reference probabilities) from the tile the loader thread brought in
            unsigned long long q2[WS_KC / 2];
            if (p.q_in_rowmajor) {
#pragma unroll
                for (int c = 0; c < WS_KC / 2; ++c)
                    q2[c] = f2_pack((valid && 2 * c < K) ? p.Q_rm[(size_t)i * K + 2 * c] : 1.f,
                                    (valid && 2 * c + 1 < K) ? p.Q_rm[(size_t)i * K + 2 * c + 1] : 1.f);
            } else if (p.q_direct) {       // A/B switch: straight from global memory, no ring
                const float4 *src = (const float4 *)(p.Q_tiled + (size_t)tile * (128 * WS_KC)) + t;
#pragma unroll
                for (int c = 0; c < WS_KC / 4; ++c) {
                    const float4 v = __ldcs(src + c * 128);
                    q2[2 * c] = f2_pack(v.x, v.y); q2[2 * c + 1] = f2_pack(v.z, v.w);
                }
            } else {
                ok = mbar_wait(q_full + 8 * sq, ph_q);
                if (!ok) break;
                const float4 *src = (const float4 *)(smem + L::QR_OFF + sq * WSR_QTILE) + t;
#pragma unroll
                for (int c = 0; c < WS_KC / 4; ++c) {
                    const float4 v = src[c * 128];
                    q2[2 * c] = f2_pack(v.x, v.y); q2[2 * c + 1] = f2_pack(v.z, v.w);
                }
                // The slot is refilled by the async proxy (bulk copy): without the cross-proxy fence the refill overtook these
                // generic-proxy reads -- the second and later uses of a slot came back with the next tile's rows (caught by
                // test_grid_tensor_kernels_many_tiles_per_cta; one or two tiles per CTA never reuse a slot).
                fence_async_smem();
                mbar_arrive(q_free + 8 * sq);
            }
            if (upd) {     // Q_{j-1} = cq * Q_{j-2} + cp * p_{j-1}[:, perm]  (stephens.cpp:87-92; running mean in the fixed mode)
                const unsigned long long cq2 = f2_pack(cq, cq), cpi2 = f2_pack(cpi, cpi);
#pragma unroll
                for (int c = 0; c < WS_KC / 2; ++c) q2[c] = f2_fma(cq2, q2[c], f2_mul(cpi2, e2[c]));
            }
            if (!valid) {
#pragma unroll
                for (int c = 0; c < WS_KC / 2; ++c) q2[c] = f2_pack(1.f, 1.f);
            }
            if (valid) {
                if (p.q_out_rowmajor) {
#pragma unroll
                    for (int c = 0; c < WS_KC / 2; ++c) {
                        float x, y;
                        f2_unpack(q2[c], x, y);
                        if (2 * c < K) p.Q_rm[(size_t)i * K + 2 * c] = x;
                        if (2 * c + 1 < K) p.Q_rm[(size_t)i * K + 2 * c + 1] = y;
                    }
                } else if (upd || p.q_in_rowmajor) {
                    float4 *dst = (float4 *)(p.Q_tiled + (size_t)tile * (128 * WS_KC)) + t;
#pragma unroll
                    for (int c = 0; c < WS_KC / 4; ++c) {
                        float4 v;
                        f2_unpack(q2[2 * c], v.x, v.y); f2_unpack(q2[2 * c + 1], v.z, v.w);
                        __stcs(dst + c * 128, v);
                    }
                }
            }
            if (!do_cost) {
                tc_fence_before();
                mbar_arrive(acc_free + 8 * a);
                continue;
            }
            // log2 q (the factor ln 2 is applied once, when the cost matrix leaves the CTA)
#pragma unroll
            for (int c = 0; c < WS_KC / 2; ++c) {
                float x, y;
                f2_unpack(q2[c], x, y);
                q2[c] = f2_pack(lg2_approx(x), lg2_approx(y));
            }
            if (k >= WSR_NB) ok = mbar_wait(st_free + 8 * b, ph_b ^ 1u);
            if (!ok) break;
            unsigned char *stage = smem + L::ST_OFF + b * WSR_STAGE + (wq * 4) * 512 + lane * 16;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 hi, lo;
                split8(q2 + 4 * c, hi, lo);
                *(uint4 *)(stage + 0 * WSR_MAT + c * 512) = hi;
                *(uint4 *)(stage + 1 * WSR_MAT + c * 512) = lo;
            }
            // p_j: this sweep's conditional probabilities (full_gibbs.cpp:97-122), recomputed
            const float inv = softmax32(tacc, bias, e2);
            tc_fence_before();
            mbar_arrive(acc_free + 8 * a);
            const float pin = valid ? inv : 0.f;
            const unsigned long long pin2 = f2_pack(pin, pin);
#pragma unroll
            for (int c = 0; c < WS_KC / 2; ++c) e2[c] = f2_mul(e2[c], pin2);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint4 hi, lo;
                split8(e2 + 4 * c, hi, lo);
                *(uint4 *)(stage + 2 * WSR_MAT + c * 512) = hi;
                *(uint4 *)(stage + 3 * WSR_MAT + c * 512) = lo;
            }
            fence_async_smem();
            mbar_arrive(st_full + 8 * b);
        }
    }
    // ---- the CTA's share of G: the four diagonal 32 x 32 blocks of the accumulator, summed in shared memory ----
    float *gsum = (float *)(smem + L::ST_OFF);   // [4][32][33] floats; the cost operand stages are idle once all_done has fired
    const bool flush = do_cost && T > 0;
    if (warp >= 8 && warp < 12 && flush) {
        if (ok) ok = mbar_wait(all_done, 0u);
        if (ok) {
            tc_fence_after();
            const int wq = warp & 3, lane = tid & 31;
            uint32_t v[32];
            tmem_ld32(acc_cost + ((uint32_t)(wq * 32) << 16) + (uint32_t)(wq * 32), v);   // lane = k, columns = l of block wq
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) gsum[(wq * 32 + lane) * 33 + c] = __uint_as_float(v[c]);
        }
    }
    if (!ok) *p.status = -10;  // BMM_ERR_TIMEOUT
    tc_fence_before();
    __syncthreads();
    if (flush && ok) {
        for (int e2 = tid; e2 < K * K; e2 += WSR_THREADS) {
            const int k = e2 % K, l2 = e2 / K;
            const float g = (gsum[(0 * 32 + k) * 33 + l2] + gsum[(1 * 32 + k) * 33 + l2]) +
                            (gsum[(2 * 32 + k) * 33 + l2] + gsum[(3 * 32 + k) * 33 + l2]);
            atomicAdd(&p.cost_out[k + (size_t)K * l2], (double)g * 0.6931471805599453);
        }
    }
    if (warp == 4) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(512) : "memory");
    }
}

template <int NCH>
cudaError_t launch_wsr_nch(const WsRelabelParams &p, int sm_count, cudaStream_t st) {
    using L = WsrLayout<NCH>;
    static FuncAttrCache attr;
    if (cudaError_t e = attr.ensure_smem(big_relabel_ws_kernel<NCH>, (int)L::TOTAL)) return e;
    const long long ntiles = ((long long)p.N_local + 127) / 128;
    long long ctas = ntiles < sm_count ? ntiles : sm_count;
    if (const char *e = getenv("BMM_GRID_MAX_CTAS")) { const int cap = atoi(e); if (cap > 0 && ctas > cap) ctas = cap; }   // tests: many tiles per CTA
    if (ctas < 1) ctas = 1;
    big_relabel_ws_kernel<NCH><<<(unsigned)ctas, L::THREADS, L::TOTAL, st>>>(p);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace

size_t wsr_q_tiled_bytes(long long N) { return (size_t)((N + 127) / 128) * 128 * WS_KC * sizeof(float); }

cudaError_t launch_big_relabel_ws(const WsRelabelParams &p, int sm_count, cudaStream_t st) {
    switch (ws_nch(p.P)) {
        case 2: return launch_wsr_nch<2>(p, sm_count, st);
        case 4: return launch_wsr_nch<4>(p, sm_count, st);
        case 6: return launch_wsr_nch<6>(p, sm_count, st);
        case 8: return launch_wsr_nch<8>(p, sm_count, st);
        case 10: return launch_wsr_nch<10>(p, sm_count, st);
        case 12: return launch_wsr_nch<12>(p, sm_count, st);
        default: return launch_wsr_nch<14>(p, sm_count, st);
    }
}

}  // namespace bmm
