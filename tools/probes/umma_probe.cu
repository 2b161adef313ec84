// Micro-benchmark (not product code): cycles per tcgen05.mma for the operand layouts kern_big_ws.cu uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && ./umma_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__host__ __device__ constexpr uint32_t umma_idesc(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// 8-bit operands (e4m3 x e4m3 -> f32), K = 32 per instruction
__device__ __forceinline__ void umma_f8(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
                 :: "r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc_f8(int M, int N, int a_mn, int b_mn) {
    return (1u << 4) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    for (int s = 0; s < (1 << 22); ++s) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// mode 0: K-major A (LBO 2048, SBO 128) x K-major B;  mode 1: MN-major A and B (LBO 128, SBO 2048)
// nacc: number of distinct accumulators the MMAs rotate over (1 = fully dependent chain)
template <int N, int MODE, int NMMA, int NACC, int M = 128, int F8 = 0>
__global__ void probe(int reps, long long *out) {
    constexpr int mode = MODE;
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) ((uint32_t *)smem)[i] = 0x3F803F80u;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(&slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&bar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 48 * 1024;
        const uint32_t idesc = F8 ? umma_idesc_f8(M, N, mode, mode) : umma_idesc(M, N, mode, mode);
        long long best = 1ll << 60;
        for (int r = 0; r < reps; ++r) {
            long long t0 = clock64();
#pragma unroll
            for (int m = 0; m < NMMA; ++m) {
                const uint32_t off = (m & 3) * (mode ? 256 : 4096);
                const uint64_t da = mode ? umma_desc(a0 + off, 128, 2048) : umma_desc(a0 + off, 2048, 128);
                const uint64_t db = mode ? umma_desc(b0 + off, 128, 2048) : umma_desc(b0 + (m & 3) * 2 * (N * 16), N * 16, 128);
                if (F8) umma_f8(tm + (uint32_t)((m % NACC) * N), da, db, idesc, m >= NACC);
                else umma(tm + (uint32_t)((m % NACC) * N), da, db, idesc, m >= NACC);
            }
            commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), r & 1);
            long long t1 = clock64();
            if (t1 - t0 < best) best = t1 - t0;
        }
        out[0] = best;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tm), "r"(512) : "memory");
    }
}


template <int N, int MODE, int NMMA, int NACC, int M = 128, int F8 = 0>
void run(long long *d) {
    long long h;
    cudaFuncSetAttribute(probe<N, MODE, NMMA, NACC, M, F8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    probe<N, MODE, NMMA, NACC, M, F8><<<1, 128, 96 * 1024>>>(20, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("M %d f8 %d mode %d N %d: error %s\n", M, F8, MODE, N, cudaGetErrorString(e)); cudaGetLastError(); return; }
    cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
    printf("%d %3d %d %2d %6lld %7.1f   M=%d %s\n", MODE, N, NACC, NMMA, h, (double)h / NMMA, M, F8 ? "e4m3 K=32" : "bf16 K=16");
}
template <int N, int MODE>
void runN(long long *d) {
    run<N, MODE, 1, 1>(d); run<N, MODE, 8, 1>(d); run<N, MODE, 32, 1>(d); run<N, MODE, 32, 2>(d);
}
int main() {
    long long *d;
    cudaMalloc(&d, 8);
    printf("mode N nacc nmma cycles cycles/mma\n");
    runN<16, 0>(d); runN<32, 0>(d); runN<64, 0>(d); runN<96, 0>(d); runN<128, 0>(d);
    runN<32, 1>(d); runN<64, 1>(d); runN<128, 1>(d);
    // M = 64 and 8-bit operands (round 2: candidates for the count contraction)
    run<32, 1, 32, 1, 64, 0>(d); run<32, 0, 32, 1, 64, 0>(d); run<64, 0, 32, 1, 64, 0>(d);
    run<32, 0, 32, 1, 128, 1>(d); run<32, 0, 32, 1, 64, 1>(d); run<64, 0, 32, 1, 128, 1>(d);
    run<32, 1, 32, 1, 128, 1>(d); run<32, 1, 32, 1, 64, 1>(d);
    return 0;
}
