// Host-side ingest helpers of the C-ABI library: persistent thread pool + SIMD ASCII -> 2-bit packer.
#pragma once
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace orph_host {

// Fixed set of worker threads; run() executes fn(0 .. n_items-1) with dynamic item hand-out and returns when every
// item is done.  The calling thread works too.  One run() at a time.
class Pool {
public:
    explicit Pool(int n_threads);
    ~Pool();
    void run(uint64_t n_items, const std::function<void(uint64_t)> &fn);
    int size() const { return n_workers_ + 1; }

private:
    void worker();
    void drain();
    std::vector<std::thread> threads_;
    int n_workers_ = 0;
    std::mutex mu_;
    std::condition_variable cv_start_, cv_done_;
    const std::function<void(uint64_t)> *fn_ = nullptr;
    uint64_t n_items_ = 0;
    std::atomic<uint64_t> next_{0};
    int active_ = 0;
    uint64_t generation_ = 0;
    bool stop_ = false;
};

const char *simd_level();

void pack_stream(Pool *pool, const uint8_t *src, uint64_t abase0, uint64_t total, const uint64_t *offsets, uint64_t n_reads,
                 uint64_t *dst, uint8_t *needs_ascii);

}  // namespace orph_host
