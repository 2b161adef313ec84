// Host side of the result download: the device keeps allocation histories as one byte per draw; the
// reference's returned `z` is an IntegerMatrix (int32, full_gibbs.cpp:56,240-245).  Moving bytes over
// PCIe and widening them on the host cores (AVX2, non-temporal stores) is cheaper than moving 4 B per
// allocation: the S x N matrices are >90 % of the bytes a run returns.
#include <cstddef>
#include <cstdint>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace {

void widen_range(const uint8_t *src, int32_t *dst, size_t n) {
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
    void run(const uint8_t *src, int32_t *dst, size_t n) {
        std::unique_lock<std::mutex> g(m_);
        src_ = src; dst_ = dst; count_ = n; pending_ = n_; ++gen_;
        cv_.notify_all();
        done_.wait(g, [this] { return pending_ == 0; });
    }
    int size() const { return n_; }

private:
    void loop(int t) {
        unsigned long seen = 0;
        for (;;) {
            const uint8_t *src; int32_t *dst; size_t n;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                src = src_; dst = dst_; n = count_;
            }
            const size_t per = ((n + n_ - 1) / n_ + 63) & ~(size_t)63, lo = (size_t)t * per;
            if (lo < n) widen_range(src + lo, dst + lo, lo + per <= n ? per : n - lo);
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
    const uint8_t *src_ = nullptr; int32_t *dst_ = nullptr; size_t count_ = 0;
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
    pool->run(src, dst, n);
}
