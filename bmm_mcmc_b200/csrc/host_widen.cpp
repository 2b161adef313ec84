// Host side of the result download: the device keeps allocation histories as one byte per draw; the
// reference's returned `z` is an IntegerMatrix (int32, full_gibbs.cpp:56,240-245).  Moving bytes over
// PCIe and widening them on the host cores (AVX2, non-temporal stores) is cheaper than moving 4 B per
// allocation: the S x N matrices are >90 % of the bytes a run returns.
#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

// Full-line (64-byte) streaming stores where the CPU has AVX-512: a cache line leaves the core in one piece instead of
// being merged from two halves in a write-combining buffer.  BMM_WIDEN_ISA=avx2 keeps the 32-byte stores (A/B).
bool use_avx512() {
#if defined(__x86_64__)
    static const bool on = [] {
        const char *e = getenv("BMM_WIDEN_ISA");
        if (e && e[0] == 'a' && e[3] == '2') return false;
        return __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw");
    }();
    return on;
#else
    return false;
#endif
}

#if defined(__x86_64__)
__attribute__((target("avx512f,avx512bw"))) void widen_range_512(const uint8_t *src, int32_t *dst, size_t n) {
    size_t i = 0;
    while (i < n && ((uintptr_t)(dst + i) & 63u)) { dst[i] = src[i]; ++i; }
    for (; i + 64 <= n; i += 64) {
        const __m128i b0 = _mm_loadu_si128((const __m128i *)(src + i)), b1 = _mm_loadu_si128((const __m128i *)(src + i + 16));
        const __m128i b2 = _mm_loadu_si128((const __m128i *)(src + i + 32)), b3 = _mm_loadu_si128((const __m128i *)(src + i + 48));
        _mm512_stream_si512((__m512i *)(dst + i), _mm512_cvtepu8_epi32(b0));
        _mm512_stream_si512((__m512i *)(dst + i + 16), _mm512_cvtepu8_epi32(b1));
        _mm512_stream_si512((__m512i *)(dst + i + 32), _mm512_cvtepu8_epi32(b2));
        _mm512_stream_si512((__m512i *)(dst + i + 48), _mm512_cvtepu8_epi32(b3));
    }
    for (; i + 16 <= n; i += 16)
        _mm512_stream_si512((__m512i *)(dst + i), _mm512_cvtepu8_epi32(_mm_loadu_si128((const __m128i *)(src + i))));
    _mm_sfence();
    for (; i < n; ++i) dst[i] = src[i];
}

// One (chain, observation) run of the relabel-deriving widening from element q on (outputs 16-byte aligned at q, both
// with the same phase): 16- and 32-byte streaming stores up to the first cache-line boundary, whole lines, then the
// same downwards -- a run of a sweep segment starts and ends anywhere, and scalar stores at its ends were a third of the
// time of 450-sweep runs.  Returns the elements done (the caller finishes < 4 of them).
__attribute__((target("avx512f,avx512bw,avx2"))) size_t derive_run_512(const uint8_t *zs, const uint8_t *tb, size_t pitch, int K,
                                                                        size_t q, size_t len, int32_t *oz, int32_t *oo) {
    auto relabel16 = [&](const __m128i z, size_t at) {
        __m128i rl = _mm_setzero_si128();
        for (int k = 0; k < K; ++k) {
            const __m128i m = _mm_cmpeq_epi8(z, _mm_set1_epi8((char)(k + 1)));
            rl = _mm_blendv_epi8(rl, _mm_loadu_si128((const __m128i *)(tb + (size_t)k * pitch + at)), m);
        }
        return rl;
    };
    auto piece = [&](int cnt) {      // cnt = 4 or 8 elements at q; the source is read exactly, the table rows are padded
        __m128i z;
        if (cnt == 8) z = _mm_loadl_epi64((const __m128i *)(zs + q));
        else { int w; memcpy(&w, zs + q, 4); z = _mm_cvtsi32_si128(w); }
        if (oz) {
            __m128i rl = _mm_setzero_si128();
            for (int k = 0; k < K; ++k) {
                const __m128i m = _mm_cmpeq_epi8(z, _mm_set1_epi8((char)(k + 1)));
                rl = _mm_blendv_epi8(rl, _mm_loadl_epi64((const __m128i *)(tb + (size_t)k * pitch + q)), m);
            }
            if (cnt == 8) _mm256_stream_si256((__m256i *)(oz + q), _mm256_cvtepu8_epi32(rl));
            else _mm_stream_si128((__m128i *)(oz + q), _mm_cvtepu8_epi32(rl));
        }
        if (oo) {
            if (cnt == 8) _mm256_stream_si256((__m256i *)(oo + q), _mm256_cvtepu8_epi32(z));
            else _mm_stream_si128((__m128i *)(oo + q), _mm_cvtepu8_epi32(z));
        }
        q += cnt;
    };
    const int32_t *lead = oo ? oo : oz;
    if (q + 4 <= len && ((uintptr_t)(lead + q) & 31u)) piece(4);
    if (q + 8 <= len && ((uintptr_t)(lead + q) & 63u)) piece(8);
    if (((uintptr_t)(lead + q) & 63u) == 0) {
        for (; q + 16 <= len; q += 16) {
            const __m128i z = _mm_loadu_si128((const __m128i *)(zs + q));
            if (oz) _mm512_stream_si512((__m512i *)(oz + q), _mm512_cvtepu8_epi32(relabel16(z, q)));
            if (oo) _mm512_stream_si512((__m512i *)(oo + q), _mm512_cvtepu8_epi32(z));
        }
        if (q + 8 <= len) piece(8);
        if (q + 4 <= len) piece(4);
    }
    return q;
}
#endif

void widen_range(const uint8_t *src, int32_t *dst, size_t n) {
#if defined(__x86_64__)
    if (use_avx512()) { widen_range_512(src, dst, n); return; }
#endif
    size_t i = 0;
#if defined(__x86_64__) && defined(__AVX2__)
    while (i < n && ((uintptr_t)(dst + i) & 31u)) { dst[i] = src[i]; ++i; }
    for (; i + 16 <= n; i += 16) {
        const __m128i b = _mm_loadu_si128((const __m128i *)(src + i));
        _mm256_stream_si256((__m256i *)(dst + i), _mm256_cvtepu8_epi32(b));
        _mm256_stream_si256((__m256i *)(dst + i + 8), _mm256_cvtepu8_epi32(_mm_srli_si128(b, 8)));
    }
    _mm_sfence();
#endif
    for (; i < n; ++i) dst[i] = src[i];
}

// Persistent workers: creating threads per 32 MB chunk cost more than the widening itself.
class Pool {
public:
    explicit Pool(int n) : n_(n) {
        for (int t = 0; t < n_; ++t) workers_.emplace_back([this, t] { loop(t); });
    }
    ~Pool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; ++gen_; }
        cv_.notify_all();
        for (auto &w : workers_) w.join();
    }
    // fn(lo, hi) over [0, n) in 256K-element pieces handed out dynamically: on the shared 16-vCPU boxes a static split
    // waits for whichever worker was descheduled
    void run(size_t n, const std::function<void(size_t, size_t)> &fn) {
        std::unique_lock<std::mutex> g(m_);
        fn_ = &fn; count_ = n; pending_ = n_; next_.store(0, std::memory_order_relaxed); ++gen_;
        cv_.notify_all();
        done_.wait(g, [this] { return pending_ == 0; });
    }
    int size() const { return n_; }

private:
    void loop(int t) {
        unsigned long seen = 0;
        for (;;) {
            const std::function<void(size_t, size_t)> *fn; size_t n;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                fn = fn_; n = count_;
            }
            (void)t;
            constexpr size_t PIECE = (size_t)256 << 10;
            for (;;) {
                const size_t lo = next_.fetch_add(PIECE, std::memory_order_relaxed);
                if (lo >= n) break;
                (*fn)(lo, lo + PIECE <= n ? lo + PIECE : n);
            }
            {
                std::lock_guard<std::mutex> g(m_);
                if (--pending_ == 0) done_.notify_one();
            }
        }
    }
    int n_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_;
    const std::function<void(size_t, size_t)> *fn_ = nullptr; size_t count_ = 0;
    std::atomic<size_t> next_{0};
    int pending_ = 0; unsigned long gen_ = 0; bool stop_ = false;
};

}  // namespace

static std::mutex pool_m;
static Pool *pool = nullptr;              // leaked on purpose: workers must outlive static destruction order

extern "C" __attribute__((visibility("default"))) void bmm_widen_u8_i32(const uint8_t *src, int32_t *dst, size_t n, int threads) {
    if (threads < 1) threads = 1;
    if (threads == 1 || n < (1u << 20)) { widen_range(src, dst, n); return; }
    std::lock_guard<std::mutex> g(pool_m);
    if (!pool || pool->size() != threads) { delete pool; pool = new Pool(threads); }
    pool->run(n, [&](size_t lo, size_t hi) { widen_range(src + lo, dst + lo, hi - lo); });
}

// Widening of a byte stream of allocations into the int32 matrices the reference returns, optionally deriving the
// relabelled matrix on the way, for a source that holds one SEGMENT of the kept sweeps.
//
// Source: elements [x0, x0 + n) of a buffer laid out [chain][observation][L sweeps of the segment] (labels 1..K).
// Output: the full S x N column-major matrices per chain, element (chain c, observation i, sweep s_off + sl) at
// ((c N + i) S + s_off + sl).  L == S, s_off == 0 is the unsegmented case (source and output indices coincide).
// perm == NULL: dst_zo receives the widened stream (dst_z unused).  perm != NULL (the returned `permutations`,
// [chain][S x K column-major]): the stream is z_original, dst_zo receives it and dst_z = perm(s, z_original - 1) + 1
// (full_gibbs.cpp:171-174), so only one of the two matrices has to cross PCIe; needs K <= 16.  Either output may be NULL.
extern "C" __attribute__((visibility("default"))) void bmm_widen_runs_u8_i32(
    const uint8_t *src, size_t x0, size_t n, int L, int S, int s_off, int N, int K, const int32_t *perm, int32_t *dst_z,
    int32_t *dst_zo, int threads) {
    if (threads < 1) threads = 1;
    if (!perm) dst_z = nullptr;
    if (!dst_z && !dst_zo) return;
    auto work = [&](size_t lo, size_t hi) {
        const size_t pitch = (size_t)L + 32;
        std::vector<uint8_t> tab(perm ? (size_t)K * pitch : 0);   // tab[k][sl] = perm(s_off + sl, k) + 1 of the current chain
        long long tab_chain = -1;
        size_t x = x0 + lo;
        const size_t xend = x0 + hi;
        while (x < xend) {
            const size_t r = x / (size_t)L, sl0 = x - r * (size_t)L;     // r = chain * N + observation
            size_t len = (size_t)L - sl0;                                // rest of this (chain, observation) run
            if (len > xend - x) len = xend - x;
            const size_t c = r / (size_t)N;
            if (perm && (long long)c != tab_chain) {
                const int32_t *pc = perm + c * (size_t)S * (size_t)K + (size_t)s_off;
                for (int k = 0; k < K; ++k)
                    for (int sl = 0; sl < L; ++sl) tab[(size_t)k * pitch + sl] = (uint8_t)(pc[(size_t)sl + (size_t)S * k] + 1);
                tab_chain = (long long)c;
            }
            const uint8_t *zs = src + (x - x0);
            const uint8_t *tb = perm ? tab.data() + sl0 : nullptr;
            const size_t e = r * (size_t)S + (size_t)s_off + sl0;        // output element
            int32_t *oz = dst_z ? dst_z + e : nullptr, *oo = dst_zo ? dst_zo + e : nullptr;
            size_t q = 0;
#if defined(__x86_64__) && defined(__AVX2__)
            const int32_t *lead = oo ? oo : oz;
            const bool wide = use_avx512();
            const uintptr_t amask = wide ? 15u : 31u;      // the wide path works its own way up to a cache line
            const bool same_phase = !oz || !oo || ((((uintptr_t)oz) ^ ((uintptr_t)oo)) & 63u) == 0;
            if (same_phase) {
                while (q < len && ((uintptr_t)(lead + q) & amask)) {
                    const uint8_t z = zs[q];
                    if (oo) oo[q] = z;
                    if (oz) oz[q] = (z >= 1 && z <= K) ? tb[(size_t)(z - 1) * pitch + q] : 0;
                    ++q;
                }
                if (wide) q = derive_run_512(zs, tb, pitch, K, q, len, oz, oo);
                for (; !wide && q + 16 <= len; q += 16) {
                    const __m128i z = _mm_loadu_si128((const __m128i *)(zs + q));
                    if (oz) {
                        __m128i rl = _mm_setzero_si128();
                        for (int k = 0; k < K; ++k) {
                            const __m128i m = _mm_cmpeq_epi8(z, _mm_set1_epi8((char)(k + 1)));
                            rl = _mm_blendv_epi8(rl, _mm_loadu_si128((const __m128i *)(tb + (size_t)k * pitch + q)), m);
                        }
                        _mm256_stream_si256((__m256i *)(oz + q), _mm256_cvtepu8_epi32(rl));
                        _mm256_stream_si256((__m256i *)(oz + q + 8), _mm256_cvtepu8_epi32(_mm_srli_si128(rl, 8)));
                    }
                    if (oo) {
                        _mm256_stream_si256((__m256i *)(oo + q), _mm256_cvtepu8_epi32(z));
                        _mm256_stream_si256((__m256i *)(oo + q + 8), _mm256_cvtepu8_epi32(_mm_srli_si128(z, 8)));
                    }
                }
            }
#endif
            for (; q < len; ++q) {
                const uint8_t z = zs[q];
                if (oo) oo[q] = z;
                if (oz) oz[q] = (z >= 1 && z <= K) ? tb[(size_t)(z - 1) * pitch + q] : 0;
            }
            x += len;
        }
#if defined(__x86_64__) && defined(__AVX2__)
        _mm_sfence();
#endif
    };
    if (threads == 1 || n < (1u << 20)) { work(0, n); return; }
    std::lock_guard<std::mutex> g(pool_m);
    if (!pool || pool->size() != threads) { delete pool; pool = new Pool(threads); }
    pool->run(n, work);
}
