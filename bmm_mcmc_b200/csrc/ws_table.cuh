// Operand image of the tensor-path weight table (kern_big_ws.cu), shared by the kernel that reads it and the
// kernels that write it (kern_big.cu).  D_kd = (log theta_kd - log(1 - theta_kd)) * log2(e) is split into two
// fp16 terms hi + lo (22 significant bits; the 0/1 rows are exact in fp16) laid side by side as 2 x 32
// accumulator columns.  No-swizzle core-matrix layout: 16-byte cells of 8 consecutive d, row stride 16 bytes,
// chunk (d / 8) stride WS_B1_ROW; the hi block first, the lo block WS_KC * 16 bytes further.
#pragma once
#include <cuda_fp16.h>

namespace bmm {

constexpr int WS_KC = 32;
constexpr int WS_PARTS = 2;
constexpr int WS_B1_ROW = WS_PARTS * WS_KC * 16;

__host__ __device__ inline int ws_nch(int P) { const int c = (P + 15) / 16; return 2 * (c < 1 ? 1 : (c > 7 ? 7 : c)); }

__device__ __forceinline__ void ws_store_cell(unsigned char *B1, int k, int d, double l1, double l0) {
    double D = (l1 - l0) * 1.4426950408889634;
    D = fmin(fmax(D, -1.0e4), 1.0e4);     // theta exactly 0 / 1: log 0 = -inf would make 0 * inf = NaN in the contraction
    if (D != D) D = 0.0;
    const __half hi = __double2half(D);
    const __half lo = __double2half(D - (double)__half2float(hi));
    unsigned char *cell = B1 + (d >> 3) * WS_B1_ROW + k * 16 + (d & 7) * 2;
    *(__half *)(cell + 0 * WS_KC * 16) = hi;
    *(__half *)(cell + 1 * WS_KC * 16) = lo;
}

}  // namespace bmm
