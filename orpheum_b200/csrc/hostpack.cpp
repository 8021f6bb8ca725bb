// Host-side ingest helpers: ASCII -> 2-bit packing with SSSE3 / AVX2 and a small persistent thread pool.
//
// A 150-nt read is 150 bytes as ASCII but 37.5 bytes as 2-bit codes, and the host-buffer scoring call is bound
// by PCIe, so orph_score_reads packs part of every batch on the host cores while the
// copy engine ships the rest as ASCII (see orpheum_b200.cu).  Layout of the packed stream: include/orpheum_b200.h
// (ORPH_READS_PACKED2): base g at bits [2*(g%32), +2) of little-endian u64 word g/32, code = (ascii >> 1) & 3.
#include "hostpack.h"

#include <immintrin.h>

#include <algorithm>
#include <cstring>

namespace orph_host {

// ------------------------------------------------------------------------------------------------
// thread pool
// ------------------------------------------------------------------------------------------------
Pool::Pool(int n_threads)
{
    if (n_threads < 1) n_threads = 1;
    n_workers_ = n_threads - 1;  // the caller of run() is the last worker
    for (int i = 0; i < n_workers_; i++) threads_.emplace_back([this] { worker(); });
}

Pool::~Pool()
{
    {
        std::lock_guard<std::mutex> lk(mu_);
        stop_ = true;
        generation_++;
    }
    cv_start_.notify_all();
    for (auto &t : threads_) t.join();
}

void Pool::drain()
{
    for (;;) {
        const uint64_t i = next_.fetch_add(1, std::memory_order_relaxed);
        if (i >= n_items_) break;
        (*fn_)(i);
    }
}

void Pool::worker()
{
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(mu_);
            cv_start_.wait(lk, [&] { return generation_ != seen; });
            seen = generation_;
            if (stop_) return;
        }
        drain();
        {
            std::lock_guard<std::mutex> lk(mu_);
            if (--active_ == 0) cv_done_.notify_one();
        }
    }
}

void Pool::run(uint64_t n_items, const std::function<void(uint64_t)> &fn)
{
    if (n_items == 0) return;
    if (n_workers_ == 0 || n_items == 1) {
        for (uint64_t i = 0; i < n_items; i++) fn(i);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(mu_);
        fn_ = &fn;
        n_items_ = n_items;
        next_.store(0, std::memory_order_relaxed);
        active_ = n_workers_;
        generation_++;
    }
    cv_start_.notify_all();
    drain();
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return active_ == 0; });
    fn_ = nullptr;
}

// ------------------------------------------------------------------------------------------------
// packing kernels: each packs whole 32-base words; `bad` gets bit i set when byte i is not upper-case ACGT
// ------------------------------------------------------------------------------------------------
static inline uint64_t pack_word_scalar(const uint8_t *p, uint32_t count, uint32_t *bad_out)
{
    uint64_t bits = 0;
    uint32_t bad = 0;
    for (uint32_t i = 0; i < count; i++) {
        const uint8_t b = p[i];
        const uint32_t c = (b >> 1) & 3u;
        bits |= (uint64_t)c << (2 * i);
        bad |= (uint32_t)(b != (uint8_t)"ACTG"[c]) << i;
    }
    *bad_out = bad;
    return bits;
}

__attribute__((target("ssse3"))) static inline uint32_t pack16_ssse3(const uint8_t *p, uint32_t *bad16)
{
    const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(p));
    const __m128i code = _mm_and_si128(_mm_srli_epi16(v, 1), _mm_set1_epi8(3));
    const __m128i expect = _mm_shuffle_epi8(_mm_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0), code);
    *bad16 = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(expect, v)) ^ 0xFFFFu;
    const __m128i t = _mm_maddubs_epi16(code, _mm_set1_epi16(0x0401));             // c0 + 4 c1 per 16-bit lane
    const __m128i u = _mm_madd_epi16(t, _mm_set1_epi32(0x00100001));               // + 16 (c2 + 4 c3) per 32-bit lane
    const __m128i g = _mm_shuffle_epi8(u, _mm_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1));
    return (uint32_t)_mm_cvtsi128_si32(g);
}

__attribute__((target("ssse3"))) static void pack_words_ssse3(const uint8_t *src, uint64_t n_full_words, uint64_t *dst, uint32_t *bad_words,
                                                              uint64_t *n_bad)
{
    for (uint64_t w = 0; w < n_full_words; w++) {
        uint32_t b0, b1;
        const uint32_t lo = pack16_ssse3(src + 32 * w, &b0), hi = pack16_ssse3(src + 32 * w + 16, &b1);
        dst[w] = (uint64_t)lo | ((uint64_t)hi << 32);
        const uint32_t bad = b0 | (b1 << 16);
        if (bad) {
            bad_words[2 * *n_bad] = (uint32_t)w;
            bad_words[2 * *n_bad + 1] = bad;
            ++*n_bad;
        }
    }
}

__attribute__((target("avx2"))) static void pack_words_avx2(const uint8_t *src, uint64_t n_full_words, uint64_t *dst, uint32_t *bad_words,
                                                            uint64_t *n_bad)
{
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i lut = _mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i w1 = _mm256_set1_epi16(0x0401), w2 = _mm256_set1_epi32(0x00100001);
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1,
                                            -1, -1, -1, -1);
    for (uint64_t w = 0; w < n_full_words; w++) {
        const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(src + 32 * w));
        const __m256i code = _mm256_and_si256(_mm256_srli_epi16(v, 1), three);
        const __m256i expect = _mm256_shuffle_epi8(lut, code);
        const uint32_t bad = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(expect, v));
        const __m256i g = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(code, w1), w2), gather);
        // streaming store: the words go to pinned staging that only the DMA engine reads -- no read-for-ownership of the
        // destination line, no cache pollution (the packer is bound by the host's memory system, not by its arithmetic)
        _mm_stream_si64(reinterpret_cast<long long *>(dst + w),
                        (long long)((uint64_t)(uint32_t)_mm256_extract_epi32(g, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(g, 4) << 32)));
        if (bad) {
            bad_words[2 * *n_bad] = (uint32_t)w;
            bad_words[2 * *n_bad + 1] = bad;
            ++*n_bad;
        }
    }
}

static void pack_words_scalar(const uint8_t *src, uint64_t n_full_words, uint64_t *dst, uint32_t *bad_words, uint64_t *n_bad)
{
    for (uint64_t w = 0; w < n_full_words; w++) {
        uint32_t bad;
        dst[w] = pack_word_scalar(src + 32 * w, 32, &bad);
        if (bad) {
            bad_words[2 * *n_bad] = (uint32_t)w;
            bad_words[2 * *n_bad + 1] = bad;
            ++*n_bad;
        }
    }
}

typedef void (*pack_fn)(const uint8_t *, uint64_t, uint64_t *, uint32_t *, uint64_t *);

static pack_fn pick_pack_fn()
{
    __builtin_cpu_init();
    if (__builtin_cpu_supports("avx2")) return pack_words_avx2;
    if (__builtin_cpu_supports("ssse3")) return pack_words_ssse3;
    return pack_words_scalar;
}

const char *simd_level()
{
    __builtin_cpu_init();
    return __builtin_cpu_supports("avx2") ? "avx2" : __builtin_cpu_supports("ssse3") ? "ssse3" : "scalar";
}

// Packs bases [0, total) of `src` (absolute base index abase0 at src[0]) into dst[0 .. ceil(total/32)).  Reads that
// hold a byte other than upper-case ACGT get needs_ascii[r] = 1 (r relative to `offsets`, which has n_reads + 1
// absolute entries starting at abase0).  Work is cut into blocks of kBlockWords words spread over the pool.
void pack_stream(Pool *pool, const uint8_t *src, uint64_t abase0, uint64_t total, const uint64_t *offsets, uint64_t n_reads,
                 uint64_t *dst, uint8_t *needs_ascii)
{
    static const pack_fn fn = pick_pack_fn();
    const uint64_t n_words = (total + 31) / 32, n_full = total / 32;
    constexpr uint64_t kBlockWords = 1 << 13;  // 256 KiB of ASCII per work item
    const uint64_t n_blocks = (n_words + kBlockWords - 1) / kBlockWords;
    auto mark = [&](uint64_t word, uint32_t bad) {
        while (bad) {
            const int i = __builtin_ctz(bad);
            bad &= bad - 1;
            const uint64_t g = abase0 + 32 * word + (uint64_t)i;
            const uint64_t *it = std::upper_bound(offsets, offsets + n_reads + 1, g);
            uint64_t r = (uint64_t)(it - offsets) - 1;
            if (r >= n_reads) r = n_reads - 1;
            __atomic_store_n(&needs_ascii[r], (uint8_t)1, __ATOMIC_RELAXED);
        }
    };
    auto work = [&](uint64_t blk) {
        const uint64_t w0 = blk * kBlockWords, w1 = std::min(n_words, w0 + kBlockWords);
        const uint64_t full1 = std::min(n_full, w1);
        uint32_t bad_words[2 * 64];
        uint64_t w = w0;
        while (w < full1) {  // sub-blocks of 64 words so that the bad-word list stays on the stack
            const uint64_t m = std::min<uint64_t>(64, full1 - w);
            uint64_t n_bad = 0;
            fn(src + 32 * w, m, dst + w, bad_words, &n_bad);
            if (needs_ascii)
                for (uint64_t b = 0; b < n_bad; b++) mark(w + bad_words[2 * b], bad_words[2 * b + 1]);
            w += m;
        }
        if (w1 > full1) {  // the partial last word of the stream
            uint32_t bad;
            dst[full1] = pack_word_scalar(src + 32 * full1, (uint32_t)(total - 32 * full1), &bad);
            if (bad && needs_ascii) mark(full1, bad);
        }
        _mm_sfence();  // streaming stores are weakly ordered: visible before the item is handed back
    };
    if (pool) pool->run(n_blocks, work);
    else
        for (uint64_t b = 0; b < n_blocks; b++) work(b);
    _mm_sfence();  // (streaming stores of the calling thread; the pool's hand-back already orders the workers')
}

}  // namespace orph_host
