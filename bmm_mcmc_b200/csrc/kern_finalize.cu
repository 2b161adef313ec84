// History layout conversion.  The sampler kernels append allocations as one byte per draw in
// [chain][sweep][observation] order (sequential, coalesced).  The reference returns z as an
// S x N IntegerMatrix, column-major, i.e. sweep fastest (full_gibbs.cpp:56,240-245); this kernel
// transposes 32x32 tiles through shared memory so both the read and the write are coalesced, and
// applies the per-sweep relabelling z_rel = perm[z-1]+1 (full_gibbs.cpp:171-174) on the way.
#include "kernels.h"

namespace bmm {
namespace {

template <typename OutT>
__global__ void finalize_z_kernel(int N, int nsamples, int burnin, int thin, int K, int s_lo, int s_hi,
                                  const uint8_t *__restrict__ zhist, const int *__restrict__ perm_out, OutT *__restrict__ z_orig,
                                  OutT *__restrict__ z_rel) {
    __shared__ uint8_t tile[32][33];
    const int c = blockIdx.z, S = hist_count(nsamples, burnin, thin), L = s_hi - s_lo;
    const int i0 = blockIdx.x * 32, s0 = s_lo + blockIdx.y * 32;
    const uint8_t *src = zhist + ((size_t)c * nsamples + burnin) * N;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int s = s0 + r, i = i0 + threadIdx.x;
        tile[r][threadIdx.x] = (s < s_hi && i < N) ? src[(size_t)s * thin * N + i] : 0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        const int i = i0 + r, s = s0 + threadIdx.x;
        if (i < N && s < s_hi) {
            const int z = tile[threadIdx.x][r];
            const size_t o = ((size_t)c * N + i) * L + (s - s_lo);
            if (z_orig) z_orig[o] = (OutT)z;
            if (z_rel) z_rel[o] = (OutT)((z >= 1 && z <= K) ? perm_out[(size_t)c * S * K + s + (size_t)S * (z - 1)] + 1 : 0);
        }
    }
}

// Byte-wide output with N and S multiples of 4: 64 x 64 tiles, 32-bit loads along the observations and
// 32-bit stores along the sweeps (the 1-byte-per-thread version above reached ~1 TB/s of the 6.4).
__global__ void __launch_bounds__(256) finalize_z_u8x4_kernel(int N, int nsamples, int burnin, int thin, int K, int s_lo, int s_hi,
                                                              const uint8_t *__restrict__ zhist, const int *__restrict__ perm_out,
                                                              uint8_t *__restrict__ z_orig, uint8_t *__restrict__ z_rel) {
    __shared__ uint8_t tile[64][68];
    const int c = blockIdx.z, S = hist_count(nsamples, burnin, thin), t = threadIdx.x, L = s_hi - s_lo;
    const int i0 = blockIdx.x * 64, s0 = s_lo + blockIdx.y * 64;
    const uint8_t *src = zhist + ((size_t)c * nsamples + burnin) * N;
    {
        const int col4 = t & 15;
        for (int r = t >> 4; r < 64; r += 16) {
            const int s = s0 + r, i = i0 + 4 * col4;
            uint32_t v = 0u;
            if (s < s_hi && i < N) v = *(const uint32_t *)(src + (size_t)s * thin * N + i);   // N % 4 == 0: i + 3 < N
            *(uint32_t *)&tile[r][4 * col4] = v;
        }
    }
    __syncthreads();
    const int w = t & 15;
    const int sw = s0 + 4 * w;
    if (sw < s_hi) {                                  // L % 4 == 0: sw + 3 < s_hi
        for (int r = t >> 4; r < 64; r += 16) {
            const int i = i0 + r;
            if (i >= N) break;
            uint32_t o = 0u, rl = 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t z = tile[4 * w + q][r];
                o |= z << (8 * q);
                if (z_rel) rl |= (uint32_t)((z >= 1 && (int)z <= K) ? perm_out[(size_t)c * S * K + (sw + q) + (size_t)S * (z - 1)] + 1 : 0) << (8 * q);
            }
            const size_t off = ((size_t)c * N + i) * L + (sw - s_lo);
            if (z_orig) *(uint32_t *)(z_orig + off) = o;
            if (z_rel) *(uint32_t *)(z_rel + off) = rl;
        }
    }
}

__global__ void expand_rows_kernel(int N, int U, int K, const int *__restrict__ rowid, const double *__restrict__ src,
                                   double *__restrict__ dst) {
    const int c = blockIdx.y;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < (size_t)N * K; e += (size_t)gridDim.x * blockDim.x)
        dst[(size_t)c * N * K + e] = src[(size_t)c * U * K + rowid[e % N] + (size_t)U * (e / N)];
}

}  // namespace

// Kept-history slots [s_lo, s_hi) of every chain (s_hi < 0: all of them) into [chain][observation][s_hi - s_lo]: with the
// whole range that is the reference's S x N column-major matrix per chain, with a sub-range one segment of it.
cudaError_t launch_finalize_z(int n_chains, int N, int nsamples, int burnin, int thin, int K, const uint8_t *zhist,
                              const int *perm_out, void *z_orig, void *z_rel, int elem_bytes, cudaStream_t st, int s_lo, int s_hi) {
    if (thin < 1) thin = 1;
    const int S = hist_count(nsamples, burnin, thin);
    if (s_hi < 0) { s_lo = 0; s_hi = S; }
    const int L = s_hi - s_lo;
    if (S <= 0 || N <= 0 || L <= 0 || n_chains <= 0) return cudaSuccess;
    dim3 block(32, 8);
    for (int c0 = 0; c0 < n_chains; c0 += 65535) {
        const int nc = n_chains - c0 < 65535 ? n_chains - c0 : 65535;
        dim3 grid((N + 31) / 32, (L + 31) / 32, nc);
        const uint8_t *zh = zhist + (size_t)c0 * nsamples * N;
        const int *pm = perm_out ? perm_out + (size_t)c0 * S * K : nullptr;
        const size_t off = (size_t)c0 * L * N;
        if (elem_bytes == 4)
            finalize_z_kernel<int32_t><<<grid, block, 0, st>>>(N, nsamples, burnin, thin, K, s_lo, s_hi, zh, pm,
                z_orig ? (int32_t *)z_orig + off : nullptr, z_rel ? (int32_t *)z_rel + off : nullptr);
        else if (N % 4 == 0 && L % 4 == 0)
            finalize_z_u8x4_kernel<<<dim3((N + 63) / 64, (L + 63) / 64, nc), 256, 0, st>>>(N, nsamples, burnin, thin, K, s_lo, s_hi, zh, pm,
                z_orig ? (uint8_t *)z_orig + off : nullptr, z_rel ? (uint8_t *)z_rel + off : nullptr);
        else
            finalize_z_kernel<uint8_t><<<grid, block, 0, st>>>(N, nsamples, burnin, thin, K, s_lo, s_hi, zh, pm,
                z_orig ? (uint8_t *)z_orig + off : nullptr, z_rel ? (uint8_t *)z_rel + off : nullptr);
        g_launches++;
    }
    return cudaGetLastError();
}

// chain-parallel posterior summary: one thread per (chain, observation) walks the kept sweeps of the raw history
// [chain][sweep][observation] (coalesced over the observations).  perm != NULL: sweep s counts towards perm(s, z - 1),
// the relabelled allocation (full_gibbs.cpp:171-174).
__global__ void chain_zfreq_kernel(int N, int nsamples, int burnin, int thin, int K, const uint8_t *__restrict__ zhist,
                                   const int *__restrict__ perm, unsigned *__restrict__ zfreq) {
    const int c = blockIdx.y, S = hist_count(nsamples, burnin, thin);
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const uint8_t *col = zhist + ((size_t)c * nsamples + burnin) * N + i;
    const int *pc = perm ? perm + (size_t)c * S * K : nullptr;
    for (int k0 = 0; k0 < K; k0 += 8) {
        unsigned cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int s = 0; s < S; ++s) {
            int zl = (int)col[(size_t)s * thin * N] - 1;
            if (pc) zl = (zl >= 0 && zl < K) ? pc[s + (size_t)S * zl] : -1;
            const int r = zl - k0;
#pragma unroll
            for (int q = 0; q < 8; ++q) cnt[q] += (r == q) ? 1u : 0u;
        }
#pragma unroll
        for (int q = 0; q < 8; ++q)
            if (k0 + q < K) zfreq[(size_t)c * N * K + i + (size_t)N * (k0 + q)] = cnt[q];
    }
}

cudaError_t launch_chain_zfreq(int n_chains, int N, int nsamples, int burnin, int thin, int K, const uint8_t *zhist, const int *perm,
                               unsigned *zfreq, cudaStream_t st) {
    if (thin < 1) thin = 1;
    const int S = hist_count(nsamples, burnin, thin);
    if (n_chains < 1 || N < 1 || S < 1) return cudaSuccess;
    for (int c0 = 0; c0 < n_chains; c0 += 65535) {
        const int nc = n_chains - c0 < 65535 ? n_chains - c0 : 65535;
        dim3 grid((N + 127) / 128, nc);
        chain_zfreq_kernel<<<grid, 128, 0, st>>>(N, nsamples, burnin, thin, K, zhist + (size_t)c0 * nsamples * N,
                                                 perm ? perm + (size_t)c0 * S * K : nullptr, zfreq + (size_t)c0 * N * K);
        g_launches++;
    }
    return cudaGetLastError();
}

cudaError_t launch_expand_rows(int n_chains, int N, int U, int K, const int *rowid, const double *src, double *dst,
                               cudaStream_t st) {
    dim3 grid((unsigned)(((size_t)N * K + 255) / 256), n_chains);
    if (grid.x > 1024) grid.x = 1024;
    expand_rows_kernel<<<grid, 256, 0, st>>>(N, U, K, rowid, src, dst);
    g_launches++;
    return cudaGetLastError();
}

}  // namespace bmm
