// C ABI of the B200-native orpheum translate/index hot path (see include/orpheum_b200.h).
// Host side: handles, device memory, launch plumbing.  Kernels live in score_kernels.cuh and
// bloom_kernels.cuh.  Built with: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <sched.h>
#include <thread>
#include <vector>

#include "bloom_kernels.cuh"
#include "hostpack.h"
#include "long_kernels.cuh"
#include "score_kernels.cuh"
#include "synth_kernels.cuh"
#include "task_kernels.cuh"

using namespace orph;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// shared with fastx.cpp
int orph_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return fail(e__ == cudaErrorMemoryAllocation ? ORPH_E_NOMEM : ORPH_E_CUDA, "%s: %s (%s:%d)", #expr, \
                        cudaGetErrorString(e__), __FILE__, __LINE__);                               \
    } while (0)

extern "C" const char *orph_last_error(void) { return g_err; }
extern "C" int orph_abi_version(void) { return ORPH_ABI_VERSION; }

// ------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------
// Test aids (environment variables, read once):
//   ORPHEUM_B200_POISON=1  every scratch allocation is pre-filled with 0xCD, so that a kernel that reads memory nothing
//                          has written shows up as a parity failure instead of working by luck on zeroed pages;
//   ORPHEUM_B200_GUARD=1   the slack behind the requested size of every scratch buffer is filled with 0xA5 and checked
//                          after each scoring / index call: a kernel writing past its buffer fails the call
//                          (compute-sanitizer is not available on the GPU pool).
static bool env_flag(const char *name)
{
    const char *v = getenv(name);
    return v && *v && *v != '0';
}
static const bool g_poison = env_flag("ORPHEUM_B200_POISON");
static const bool g_guard = env_flag("ORPHEUM_B200_GUARD");

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    size_t used = 0;  // bytes asked for by the last ensure(); [used, cap) is slack
    int ensure(size_t need)
    {
        if (need > cap) {
            if (p) cudaFree(p);  // cudaFree waits for the device: work still using the old buffer finishes first
            p = nullptr;
            cap = 0;
            size_t want = need + need / 8 + 256;
            CUDA_TRY(cudaMalloc(&p, want));
            cap = want;
            if (g_poison) CUDA_TRY(cudaMemset(p, 0xCD, want));
        }
        used = need;
        if (g_guard) {  // (synchronises: a test aid; earlier launches may still be using the part that becomes slack)
            CUDA_TRY(cudaDeviceSynchronize());
            CUDA_TRY(cudaMemset((uint8_t *)p + used, 0xA5, cap - used));
        }
        return ORPH_OK;
    }
    // guard mode: true if nothing wrote into the slack
    bool slack_intact() const
    {
        if (!g_guard || !p || cap <= used) return true;
        std::vector<uint8_t> h(cap - used);
        if (cudaMemcpy(h.data(), (const uint8_t *)p + used, h.size(), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
        for (uint8_t b : h)
            if (b != 0xA5) return false;
        return true;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        used = 0;
    }
};

// Process-wide cache of pinned host buffers.  cudaHostAlloc pins page by page (tens of milliseconds per 100 MB), which
// would dominate a short scoring stream or the first call of every context; buffers handed back are kept (up to kCap
// bytes) and given to the next request they fit.
class PinnedCache {
public:
    void *get(size_t need, size_t *got)
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            size_t best = free_.size();
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i].second >= need && free_[i].second <= 2 * need + (1u << 20) && (best == free_.size() || free_[i].second < free_[best].second)) best = i;
            if (best < free_.size()) {
                void *p = free_[best].first;
                *got = free_[best].second;
                cached_ -= free_[best].second;
                free_.erase(free_.begin() + (long)best);
                live_.push_back({p, *got});
                return p;
            }
        }
        void *p = nullptr;
        if (cudaHostAlloc(&p, need, cudaHostAllocDefault) != cudaSuccess) {
            cudaGetLastError();
            trim(0);  // give the cached buffers back and try once more
            if (cudaHostAlloc(&p, need, cudaHostAllocDefault) != cudaSuccess) {
                cudaGetLastError();
                return nullptr;
            }
        }
        *got = need;
        std::lock_guard<std::mutex> lk(mu_);
        live_.push_back({p, need});
        return p;
    }
    void put(void *p)
    {
        if (!p) return;
        size_t size = 0;
        {
            std::lock_guard<std::mutex> lk(mu_);
            for (size_t i = 0; i < live_.size(); i++)
                if (live_[i].first == p) {
                    size = live_[i].second;
                    live_.erase(live_.begin() + (long)i);
                    break;
                }
            if (size && cached_ + size <= kCap) {
                free_.push_back({p, size});
                cached_ += size;
                return;
            }
        }
        cudaFreeHost(p);
    }
    void trim(size_t keep)
    {
        std::vector<void *> drop;
        {
            std::lock_guard<std::mutex> lk(mu_);
            while (cached_ > keep && !free_.empty()) {
                drop.push_back(free_.back().first);
                cached_ -= free_.back().second;
                free_.pop_back();
            }
        }
        for (void *p : drop) cudaFreeHost(p);
    }

private:
    static constexpr size_t kCap = 4ull << 30;
    std::mutex mu_;
    std::vector<std::pair<void *, size_t>> free_, live_;
    size_t cached_ = 0;
};
static PinnedCache g_pinned;

// pinned host buffer
struct PinnedBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t need)
    {
        if (need <= cap) return ORPH_OK;
        g_pinned.put(p);
        p = nullptr;
        cap = 0;
        size_t got = 0;
        p = g_pinned.get(need + need / 8 + 256, &got);
        if (!p) return fail(ORPH_E_NOMEM, "cudaHostAlloc of %zu bytes failed", need);
        cap = got;
        return ORPH_OK;
    }
    void release()
    {
        g_pinned.put(p);
        p = nullptr;
        cap = 0;
    }
};

constexpr int kLaneSlots = 6;   // chunks in flight on the host-packing lane of the host-buffer pipeline (ingest_pipeline.inl):
                                // its small copies wait behind the other lane's large ones
constexpr int kFrontSlots = 4;  // chunks in flight on the copy-engine lane

// staging of one chunk in flight: device buffers, the pinned buffer a host-packed chunk is built in, and the events
// that order its three stages (H2D -> kernels -> D2H)
struct PipeSlot {
    DevBuf in_reads, in_offsets, out_results;
    PinnedBuf h_stage;
    cudaEvent_t ev_in = nullptr, ev_kernels = nullptr, ev_out = nullptr;  // ev_out: blocking-sync (the host waits on it to reuse the slot)
    cudaEvent_t ev_h2d_begin = nullptr, ev_h2d_end = nullptr;             // timed: the H2D rate reported by orph_ctx_ingest_stats
    uint64_t h2d_bytes = 0;
    bool used = false;
};

struct orph_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 0;
    int max_smem_optin = 0;
    size_t l2_persist_max = 0, l2_window_max = 0;
    const void *l2_window_base = nullptr;  // filter buffer the access-policy window currently covers
    uint64_t l2_window_bytes = 0;
    bool force_exact = false;
    bool disable_task_kernel = false;  // test hook: route every read through the warp-per-read kernel
    uint64_t launches = 0;
    uint32_t sticky_status = 0;  // device status bits folded in by the host-buffer entry points
    DevBuf counters;             // ScoreCounters (one per in-flight call; calls on a ctx are serialised on its stream)
    DevBuf queue, bytes_queue;   // uint32 read indices
    DevBuf task_queue;           // uint32 (read << 3 | frame), up to 6 per read
    DevBuf task_queue_sorted;    // the same entries ordered by peptide length (mixed-length batches)
    DevBuf needs_bytes;          // ASCII input: uint8 per read "holds a byte other than ACGTN", then uint8 per read "holds an N"
    DevBuf nmask;                // ASCII input: the N plane of the 2-bit stream (one bit per base)
    DevBuf task_queue_n;         // tasks of the reads that hold N
    cudaStream_t side = nullptr; // the N reads' two kernels run here, next to the main kernels of the batch (fork / join by events)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    DevBuf packed;               // 2-bit stream built from ASCII input
    DevBuf in_reads, in_offsets, out_results;  // staging for the index build
    PipeSlot lane_slots[2][kLaneSlots];                  // [0] the copy-engine lane, [1] the host-packing lane of orph_score_reads
    cudaStream_t s_in[2] = {nullptr, nullptr}, s_out[2] = {nullptr, nullptr};  // H2D and D2H stream per lane
    orph_ctx *lane_ctx = nullptr;  // private context (kernel stream + scratch) of the host-packing lane, so that neither
                                   // lane's kernels queue behind the other lane's copies
    bool pipe_ready = false;
    int host_threads = -1;                               // host packing threads: -1 auto, 0 = never pack on the host
    orph_host::Pool *pool = nullptr;
    double pack_s_per_byte = 0;                          // measured host packing cost (running average)
    double h2d_bytes_per_s = 50e9;                       // planning estimate of the H2D rate
    uint64_t chunks_packed_on_host = 0, chunks_sent_ascii = 0;
    uint32_t pack_lane_skips = 0;                        // calls that went without the packing lane because it was the slower route
    DevBuf misc;                 // small transfers (hash batches, ...)
    bool timing = false;
    struct Interval { cudaEvent_t a, b; int id; };
    std::vector<Interval> intervals;
    DevBuf seen;                 // hash set of k-mer hashes seen during one index-build chunk
    DevBuf long_scratch;         // per-block peptide + de-duplication table of score_long_kernel
};

// guard mode: every scratch buffer of the context, by name
static int verify_guards(orph_ctx *ctx)
{
    if (!g_guard) return ORPH_OK;
    cudaDeviceSynchronize();
    struct Named { const char *name; const DevBuf *b; };
    std::vector<Named> all = {{"counters", &ctx->counters}, {"queue", &ctx->queue}, {"bytes_queue", &ctx->bytes_queue}, {"task_queue", &ctx->task_queue},
                              {"task_queue_sorted", &ctx->task_queue_sorted}, {"needs_bytes", &ctx->needs_bytes}, {"packed", &ctx->packed},
                              {"nmask", &ctx->nmask}, {"task_queue_n", &ctx->task_queue_n},
                              {"in_reads", &ctx->in_reads}, {"in_offsets", &ctx->in_offsets}, {"out_results", &ctx->out_results}, {"misc", &ctx->misc},
                              {"seen", &ctx->seen}, {"long_scratch", &ctx->long_scratch}};
    for (auto &lane : ctx->lane_slots)
        for (PipeSlot &q : lane) {
            all.push_back({"slot.in_reads", &q.in_reads});
            all.push_back({"slot.in_offsets", &q.in_offsets});
            all.push_back({"slot.out_results", &q.out_results});
        }
    for (const Named &n : all)
        if (!n.b->slack_intact()) return fail(ORPH_E_CUDA, "ORPHEUM_B200_GUARD: a kernel wrote past the end of scratch buffer '%s'", n.name);
    if (ctx->lane_ctx) return verify_guards(ctx->lane_ctx);
    return ORPH_OK;
}

struct orph_filter {
    int device = 0;
    uint32_t ksize = 0, n_tables = 0;
    uint64_t sizes[ORPH_MAX_TABLES] = {0};
    uint64_t offsets[ORPH_MAX_TABLES] = {0};  // byte offset of each table in buf
    uint64_t padded[ORPH_MAX_TABLES] = {0};   // bytes reserved per table (multiple of 256)
    uint8_t *buf = nullptr;
    uint64_t n_bytes = 0;
    unsigned long long *d_scalars = nullptr;  // [0] n_unique, [1] popcount scratch, [2] next record
    uint64_t n_unique = 0;
    orph_ctx *ctx = nullptr;  // context the filter was created on (its stream orders builds)
};

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev)
    {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static FilterDev filter_dev(const orph_filter *f)
{
    FilterDev F;
    memset(&F, 0, sizeof(F));
    F.ksize = f->ksize;
    F.n_tables = f->n_tables;
    bool small = true, tiny = true;
    for (uint32_t i = 0; i < f->n_tables; i++) {
        F.size[i] = f->sizes[i];
        F.magic[i] = ~0ULL / f->sizes[i];
        F.table[i] = f->buf + f->offsets[i];
        if (f->sizes[i] >= (1ULL << 31)) small = false;
        if (f->sizes[i] < 2 || f->sizes[i] > kTinyPrimeMax) tiny = false;
    }
    F.small = small ? 1u : 0u;
    F.tiny = tiny ? 1u : 0u;
    return F;
}

static int grid_for(uint64_t n_items, int block, int sm_count, int waves = 8)
{
    uint64_t g = (n_items + block - 1) / block;
    uint64_t cap = (uint64_t)sm_count * waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
extern "C" int orph_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();  // not sticky: a machine without a GPU is an answer, not a failure
        return 0;
    }
    return n;
}

extern "C" int orph_ctx_create(int device, void *stream, orph_ctx **out)
{
    if (!out) return fail(ORPH_E_INVALID, "orph_ctx_create: out is NULL");
    *out = nullptr;
    int n_dev = 0;
    // the very first CUDA call of a process can fail transiently while a previous process on the same
    // GPU is still tearing down; retry briefly before giving up
    for (int attempt = 0;; attempt++) {
        cudaError_t e = cudaGetDeviceCount(&n_dev);
        if (e == cudaSuccess) break;
        cudaGetLastError();
        if (attempt >= 4) return fail(ORPH_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
        std::this_thread::sleep_for(std::chrono::milliseconds(250 * (attempt + 1)));
    }
    if (device < 0 || device >= n_dev) return fail(ORPH_E_INVALID, "orph_ctx_create: device %d of %d", device, n_dev);
    DeviceGuard guard(device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return fail(ORPH_E_INVALID, "orph_ctx_create: device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    orph_ctx *ctx = new (std::nothrow) orph_ctx();
    if (!ctx) return fail(ORPH_E_NOMEM, "orph_ctx_create: out of host memory");
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    ctx->l2_persist_max = (size_t)prop.persistingL2CacheMaxSize;
    ctx->l2_window_max = (size_t)prop.accessPolicyMaxWindowSize;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) {
            delete ctx;
            return fail(ORPH_E_CUDA, "cudaStreamCreate: %s", cudaGetErrorString(e));
        }
        ctx->own_stream = true;
    }
    int rc = ctx->counters.ensure(sizeof(ScoreCounters));
    if (rc == ORPH_OK) {
        // cudaMalloc does not clear: the sticky status word must start at zero (recycled device memory would otherwise
        // show up as a spurious ORPH_E_EMPTY_READ / ORPH_E_TOO_LONG on the context's first call)
        cudaError_t e = cudaMemsetAsync(ctx->counters.p, 0, sizeof(ScoreCounters), ctx->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) rc = fail(ORPH_E_CUDA, "orph_ctx_create: %s", cudaGetErrorString(e));
    }
    if (rc != ORPH_OK) {
        orph_ctx_destroy(ctx);
        return rc;
    }
    *out = ctx;
    return ORPH_OK;
}

extern "C" void orph_ctx_destroy(orph_ctx *ctx)
{
    if (!ctx) return;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (DevBuf *b : {&ctx->counters, &ctx->queue, &ctx->bytes_queue, &ctx->task_queue, &ctx->task_queue_sorted, &ctx->needs_bytes, &ctx->packed, &ctx->nmask, &ctx->task_queue_n, &ctx->in_reads,
                      &ctx->in_offsets, &ctx->out_results, &ctx->misc, &ctx->seen, &ctx->long_scratch})
        b->release();
    for (auto &lane : ctx->lane_slots)
        for (PipeSlot &q : lane) {
            q.in_reads.release();
            q.in_offsets.release();
            q.out_results.release();
            q.h_stage.release();
            for (cudaEvent_t e : {q.ev_in, q.ev_kernels, q.ev_out, q.ev_h2d_begin, q.ev_h2d_end})
                if (e) cudaEventDestroy(e);
        }
    delete ctx->pool;
    if (ctx->side) cudaStreamDestroy(ctx->side);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->lane_ctx) orph_ctx_destroy(ctx->lane_ctx);
    for (cudaStream_t st : {ctx->s_in[0], ctx->s_out[0], ctx->s_out[1]})  // s_in[1] is s_in[0]
        if (st) cudaStreamDestroy(st);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" int orph_ctx_synchronize(orph_ctx *ctx)
{
    if (!ctx) return fail(ORPH_E_INVALID, "orph_ctx_synchronize: ctx is NULL");
    DeviceGuard guard(ctx->device);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ORPH_OK;
}

static int status_to_code(uint32_t bits)
{
    if (bits & kStatusEmptyRead) return fail(ORPH_E_EMPTY_READ, "a read of length 0 was passed (the reference raises ZeroDivisionError, translate.py:83)");
    if (bits & kStatusTooLong) return fail(ORPH_E_TOO_LONG, "a read is longer than the declared bound / ORPH_MAX_READ_LEN (%u)", ORPH_MAX_READ_LEN);
    return ORPH_OK;
}

extern "C" int orph_ctx_check(orph_ctx *ctx)
{
    if (!ctx) return fail(ORPH_E_INVALID, "orph_ctx_check: ctx is NULL");
    DeviceGuard guard(ctx->device);
    ScoreCounters c;
    CUDA_TRY(cudaMemcpyAsync(&c, ctx->counters.p, sizeof(c), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    uint32_t bits = c.status | ctx->sticky_status;
    ctx->sticky_status = 0;
    if (g_guard) {
        const int rg = verify_guards(ctx);
        if (rg != ORPH_OK) return rg;
    }
    CUDA_TRY(cudaMemsetAsync((uint8_t *)ctx->counters.p + offsetof(ScoreCounters, status), 0, sizeof(uint32_t), ctx->stream));
    return status_to_code(bits);
}

extern "C" int orph_ctx_set_force_exact_dedup(orph_ctx *ctx, int on)
{
    if (!ctx) return fail(ORPH_E_INVALID, "ctx is NULL");
    ctx->force_exact = on != 0;
    return ORPH_OK;
}

// RAII helper: brackets one kernel launch with events when timing is on
struct Timed {
    orph_ctx *ctx;
    cudaEvent_t a = nullptr, b = nullptr;
    int id;
    cudaStream_t on;
    Timed(orph_ctx *c, int kernel_id, cudaStream_t stream = nullptr) : ctx(c), id(kernel_id), on(stream ? stream : c->stream)
    {
        if (!ctx->timing) return;
        if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) { a = b = nullptr; cudaGetLastError(); return; }
        cudaEventRecord(a, on);
    }
    ~Timed()
    {
        if (!a) return;
        cudaEventRecord(b, on);
        ctx->intervals.push_back({a, b, id});
    }
};

extern "C" int orph_ctx_set_kernel_timing(orph_ctx *ctx, int on)
{
    if (!ctx) return fail(ORPH_E_INVALID, "ctx is NULL");
    ctx->timing = on != 0;
    return ORPH_OK;
}

extern "C" int orph_ctx_kernel_times(orph_ctx *ctx, double *ms, uint64_t *launches)
{
    if (!ctx) return fail(ORPH_E_INVALID, "ctx is NULL");
    DeviceGuard guard(ctx->device);
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < ORPH_N_TIMED_KERNELS; i++) {
        if (ms) ms[i] = 0;
        if (launches) launches[i] = 0;
    }
    for (auto &iv : ctx->intervals) {
        float t = 0;
        if (cudaEventElapsedTime(&t, iv.a, iv.b) == cudaSuccess && iv.id >= 0 && iv.id < ORPH_N_TIMED_KERNELS) {
            if (ms) ms[iv.id] += t;
            if (launches) launches[iv.id]++;
        }
        cudaEventDestroy(iv.a);
        cudaEventDestroy(iv.b);
    }
    ctx->intervals.clear();
    return ORPH_OK;
}

extern "C" int orph_ctx_set_thread_path(orph_ctx *ctx, int on)
{
    if (!ctx) return fail(ORPH_E_INVALID, "ctx is NULL");
    ctx->disable_task_kernel = on == 0;
    return ORPH_OK;
}

extern "C" uint64_t orph_ctx_launch_count(const orph_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ------------------------------------------------------------------------------------------------
// filter
// ------------------------------------------------------------------------------------------------
static bool is_prime_u64(uint64_t n)
{
    if (n < 2) return false;
    if (n < 4) return true;
    if (!(n & 1)) return false;
    for (uint64_t d = 3; d * d <= n; d += 2)
        if (n % d == 0) return false;
    return true;
}

extern "C" int orph_table_sizes(uint64_t tablesize, uint32_t n_tables, uint64_t *out_sizes)
{
    if (!out_sizes || n_tables == 0 || n_tables > ORPH_MAX_TABLES || tablesize < 3)
        return fail(ORPH_E_INVALID, "orph_table_sizes: bad arguments (tablesize %llu, n_tables %u)", (unsigned long long)tablesize, n_tables);
    uint64_t i = tablesize - 1;
    if (!(i & 1)) i--;
    uint32_t found = 0;
    while (found < n_tables && i > 0) {
        if (is_prime_u64(i)) out_sizes[found++] = i;
        if (i < 2) break;
        i -= 2;
    }
    if (found != n_tables) return fail(ORPH_E_INVALID, "orph_table_sizes: fewer than %u primes below %llu", n_tables, (unsigned long long)tablesize);
    return ORPH_OK;
}

extern "C" void orph_filter_destroy(orph_filter *f)
{
    if (!f) return;
    DeviceGuard guard(f->device);
    if (f->ctx && f->ctx->l2_window_base == f->buf) f->ctx->l2_window_base = nullptr;
    if (f->buf) cudaFree(f->buf);
    if (f->d_scalars) cudaFree(f->d_scalars);
    delete f;
}

extern "C" int orph_filter_create(orph_ctx *ctx, uint32_t ksize, uint32_t n_tables, const uint64_t *table_sizes, orph_filter **out)
{
    if (!ctx || !out || !table_sizes) return fail(ORPH_E_INVALID, "orph_filter_create: NULL argument");
    *out = nullptr;
    if (ksize == 0 || ksize > (uint32_t)kMaxKsize) return fail(ORPH_E_INVALID, "orph_filter_create: ksize %u not in 1..%d", ksize, kMaxKsize);
    if (n_tables == 0 || n_tables > ORPH_MAX_TABLES) return fail(ORPH_E_INVALID, "orph_filter_create: n_tables %u not in 1..%d", n_tables, ORPH_MAX_TABLES);
    DeviceGuard guard(ctx->device);
    orph_filter *f = new (std::nothrow) orph_filter();
    if (!f) return fail(ORPH_E_NOMEM, "out of host memory");
    f->device = ctx->device;
    f->ctx = ctx;
    f->ksize = ksize;
    f->n_tables = n_tables;
    uint64_t total = 0;
    for (uint32_t i = 0; i < n_tables; i++) {
        if (table_sizes[i] < 2 || table_sizes[i] >= (1ULL << 48)) {
            delete f;
            return fail(ORPH_E_INVALID, "orph_filter_create: table size %llu out of range", (unsigned long long)table_sizes[i]);
        }
        f->sizes[i] = table_sizes[i];
        f->offsets[i] = total;
        f->padded[i] = (table_sizes[i] / 8 + 1 + 255) & ~255ULL;
        total += f->padded[i];
    }
    f->n_bytes = total;
    cudaError_t e = cudaMalloc((void **)&f->buf, total);
    if (e == cudaSuccess) e = cudaMalloc((void **)&f->d_scalars, 4 * sizeof(unsigned long long));
    if (e == cudaSuccess) e = cudaMemsetAsync(f->buf, 0, total, ctx->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(f->d_scalars, 0, 4 * sizeof(unsigned long long), ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        int code = fail(e == cudaErrorMemoryAllocation ? ORPH_E_NOMEM : ORPH_E_CUDA, "orph_filter_create (%llu bytes): %s", (unsigned long long)total, cudaGetErrorString(e));
        orph_filter_destroy(f);
        return code;
    }
    *out = f;
    return ORPH_OK;
}

extern "C" int orph_filter_from_host(orph_ctx *ctx, uint32_t ksize, uint32_t n_tables, const uint64_t *table_sizes,
                                     const uint8_t *const *tables, orph_filter **out)
{
    if (!tables) return fail(ORPH_E_INVALID, "orph_filter_from_host: tables is NULL");
    int rc = orph_filter_create(ctx, ksize, n_tables, table_sizes, out);
    if (rc != ORPH_OK) return rc;
    orph_filter *f = *out;
    DeviceGuard guard(ctx->device);
    for (uint32_t i = 0; i < n_tables; i++) {
        cudaError_t e = cudaMemcpyAsync(f->buf + f->offsets[i], tables[i], f->sizes[i] / 8 + 1, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
            orph_filter_destroy(f);
            *out = nullptr;
            return fail(ORPH_E_CUDA, "orph_filter_from_host: %s", cudaGetErrorString(e));
        }
    }
    CUDA_TRY(cudaStreamSynchronize(ctx->stream));
    return ORPH_OK;
}

extern "C" int orph_filter_popcount(const orph_filter *f, uint32_t table, uint64_t *out)
{
    if (!f || !out || table >= f->n_tables) return fail(ORPH_E_INVALID, "orph_filter_popcount: bad arguments");
    DeviceGuard guard(f->device);
    cudaStream_t s = f->ctx->stream;
    CUDA_TRY(cudaMemsetAsync(f->d_scalars + 1, 0, sizeof(unsigned long long), s));
    // only the bytes that exist in the file take part; the padding is zero anyway
    const uint64_t n_words = f->padded[table] / 4;
    popcount_kernel<<<grid_for(n_words, 256, f->ctx->sm_count), 256, 0, s>>>(reinterpret_cast<const uint32_t *>(f->buf + f->offsets[table]), n_words, f->d_scalars + 1);
    f->ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    unsigned long long v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, f->d_scalars + 1, sizeof(v), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    *out = v;
    return ORPH_OK;
}

extern "C" int orph_filter_to_host(const orph_filter *f, uint8_t *const *tables, uint64_t *occupied_bins_table0)
{
    if (!f) return fail(ORPH_E_INVALID, "orph_filter_to_host: filter is NULL");
    DeviceGuard guard(f->device);
    cudaStream_t s = f->ctx->stream;
    if (tables) {
        for (uint32_t i = 0; i < f->n_tables; i++)
            CUDA_TRY(cudaMemcpyAsync(tables[i], f->buf + f->offsets[i], f->sizes[i] / 8 + 1, cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
    }
    if (occupied_bins_table0) return orph_filter_popcount(f, 0, occupied_bins_table0);
    return ORPH_OK;
}

extern "C" int orph_filter_info(const orph_filter *f, uint32_t *ksize, uint32_t *n_tables, uint64_t *table_sizes)
{
    if (!f) return fail(ORPH_E_INVALID, "orph_filter_info: filter is NULL");
    if (ksize) *ksize = f->ksize;
    if (n_tables) *n_tables = f->n_tables;
    if (table_sizes) memcpy(table_sizes, f->sizes, f->n_tables * sizeof(uint64_t));
    return ORPH_OK;
}

extern "C" uint64_t orph_filter_n_unique_kmers(const orph_filter *f) { return f ? f->n_unique : 0; }

extern "C" int orph_filter_device_buffer(const orph_filter *f, void **device_ptr, uint64_t *n_bytes, uint64_t *table_offsets)
{
    if (!f) return fail(ORPH_E_INVALID, "orph_filter_device_buffer: filter is NULL");
    if (device_ptr) *device_ptr = f->buf;
    if (n_bytes) *n_bytes = f->n_bytes;
    if (table_offsets) memcpy(table_offsets, f->offsets, f->n_tables * sizeof(uint64_t));
    return ORPH_OK;
}

extern "C" int orph_filter_replicate(const orph_filter *src, orph_ctx *dst_ctx, orph_filter **out)
{
    if (!src || !dst_ctx || !out) return fail(ORPH_E_INVALID, "orph_filter_replicate: NULL argument");
    int rc = orph_filter_create(dst_ctx, src->ksize, src->n_tables, src->sizes, out);
    if (rc != ORPH_OK) return rc;
    orph_filter *dst = *out;
    {
        DeviceGuard g(src->device);
        cudaStreamSynchronize(src->ctx->stream);  // the source must be fully built
    }
    DeviceGuard guard(dst_ctx->device);
    if (src->device != dst_ctx->device) {
        int can = 0;
        cudaDeviceCanAccessPeer(&can, dst_ctx->device, src->device);
        if (can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(src->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
            else cudaGetLastError();
        }
    }
    cudaError_t e = cudaMemcpyPeerAsync(dst->buf, dst_ctx->device, src->buf, src->device, src->n_bytes, dst_ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(dst_ctx->stream);
    if (e != cudaSuccess) {
        orph_filter_destroy(dst);
        *out = nullptr;
        return fail(ORPH_E_CUDA, "orph_filter_replicate: %s", cudaGetErrorString(e));
    }
    dst->n_unique = src->n_unique;
    return ORPH_OK;
}

static int fetch_n_unique(orph_filter *f, uint64_t *inc)
{
    unsigned long long v = 0;
    CUDA_TRY(cudaMemcpyAsync(&v, f->d_scalars, sizeof(v), cudaMemcpyDeviceToHost, f->ctx->stream));
    CUDA_TRY(cudaStreamSynchronize(f->ctx->stream));
    if (inc) *inc = v - f->n_unique;
    f->n_unique = v;
    return ORPH_OK;
}

extern "C" int orph_filter_add_peptides(orph_filter *f, const uint8_t *residues, const uint64_t *offsets, uint64_t n_records,
                                        const uint8_t lut[256], uint64_t *n_unique_kmers_inc)
{
    if (!f || !offsets || !lut || (!residues && n_records)) return fail(ORPH_E_INVALID, "orph_filter_add_peptides: NULL argument");
    if (n_unique_kmers_inc) *n_unique_kmers_inc = 0;
    if (n_records == 0) return ORPH_OK;
    orph_ctx *ctx = f->ctx;
    DeviceGuard guard(f->device);
    cudaStream_t s = ctx->stream;
    BuildParams P;
    P.F = filter_dev(f);
    memcpy(P.lut, lut, 256);
    const uint64_t before = f->n_unique;
    // chunks of records bounded by residues, so that the per-call hash set of seen k-mers stays <= 1 GiB
    const uint64_t kMaxChunkResidues = 1ULL << 26;
    uint64_t lo = 0;
    while (lo < n_records) {
        uint64_t hi = lo + 1;
        while (hi < n_records && offsets[hi + 1] - offsets[lo] <= kMaxChunkResidues) hi++;
        const uint64_t m = hi - lo;
        if (offsets[hi] < offsets[lo]) return fail(ORPH_E_INVALID, "orph_filter_add_peptides: offsets are not monotonic");
        const uint64_t base = offsets[lo], total = offsets[hi] - base;
        int rc;
        if ((rc = ctx->in_reads.ensure(total + 16)) != ORPH_OK) return rc;
        if ((rc = ctx->in_offsets.ensure((m + 1) * sizeof(uint64_t))) != ORPH_OK) return rc;
        // the hash set of seen k-mers only serves n_unique_kmers: skipped when the caller does not ask for the count
        const bool count_unique = n_unique_kmers_inc != nullptr;
        uint64_t slots = 1024;
        if (count_unique) {
            while (slots < 2 * total) slots <<= 1;
            if ((rc = ctx->seen.ensure(slots * 8)) != ORPH_OK) return rc;
            CUDA_TRY(cudaMemsetAsync(ctx->seen.p, 0xFF, slots * 8, s));
        }
        if (total) CUDA_TRY(cudaMemcpyAsync(ctx->in_reads.p, residues + base, total, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(ctx->in_offsets.p, offsets + lo, (m + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemsetAsync(f->d_scalars + 2, 0, sizeof(unsigned long long), s));
        const int grid = ctx->sm_count * 8;
        // the residues pointer is shifted so that the absolute offsets index it directly
        bloom_add_peptides_kernel<<<grid, 256, 0, s>>>(P, (const uint8_t *)ctx->in_reads.p - base, (const uint64_t *)ctx->in_offsets.p, m,
                                                      f->d_scalars + 2, f->d_scalars, count_unique ? (unsigned long long *)ctx->seen.p : nullptr,
                                                      slots - 1);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaStreamSynchronize(s));  // the staging buffers are reused by the next chunk
        lo = hi;
    }
    int rc = fetch_n_unique(f, nullptr);
    if (rc != ORPH_OK) return rc;
    if (n_unique_kmers_inc) *n_unique_kmers_inc = f->n_unique - before;
    if (g_guard && (rc = verify_guards(ctx)) != ORPH_OK) return rc;
    return ORPH_OK;
}

extern "C" int orph_filter_add_hashes(orph_filter *f, const uint64_t *hashes, uint64_t n, uint8_t *is_new)
{
    if (!f || (!hashes && n)) return fail(ORPH_E_INVALID, "orph_filter_add_hashes: NULL argument");
    if (n == 0) return ORPH_OK;
    orph_ctx *ctx = f->ctx;
    DeviceGuard guard(f->device);
    cudaStream_t s = ctx->stream;
    int rc;
    if ((rc = ctx->misc.ensure(n * 9 + 16)) != ORPH_OK) return rc;
    uint64_t *d_h = (uint64_t *)ctx->misc.p;
    uint8_t *d_new = (uint8_t *)ctx->misc.p + n * 8;
    CUDA_TRY(cudaMemcpyAsync(d_h, hashes, n * 8, cudaMemcpyHostToDevice, s));
    bloom_add_hashes_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, s>>>(filter_dev(f), d_h, n, is_new ? d_new : nullptr, f->d_scalars);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    if (is_new) CUDA_TRY(cudaMemcpyAsync(is_new, d_new, n, cudaMemcpyDeviceToHost, s));
    return fetch_n_unique(f, nullptr);
}

extern "C" int orph_filter_query_hashes(const orph_filter *f, const uint64_t *hashes, uint64_t n, uint8_t *out01)
{
    if (!f || ((!hashes || !out01) && n)) return fail(ORPH_E_INVALID, "orph_filter_query_hashes: NULL argument");
    if (n == 0) return ORPH_OK;
    orph_ctx *ctx = f->ctx;
    DeviceGuard guard(f->device);
    cudaStream_t s = ctx->stream;
    int rc;
    if ((rc = ctx->misc.ensure(n * 9 + 16)) != ORPH_OK) return rc;
    uint64_t *d_h = (uint64_t *)ctx->misc.p;
    uint8_t *d_o = (uint8_t *)ctx->misc.p + n * 8;
    CUDA_TRY(cudaMemcpyAsync(d_h, hashes, n * 8, cudaMemcpyHostToDevice, s));
    bloom_query_hashes_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, s>>>(filter_dev(f), d_h, n, d_o);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out01, d_o, n, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return ORPH_OK;
}

extern "C" int orph_hash_kmers(orph_ctx *ctx, const uint8_t *keys, uint32_t len, uint64_t n, uint64_t *out_hashes)
{
    if (!ctx || ((!keys && len) || !out_hashes) && n) return fail(ORPH_E_INVALID, "orph_hash_kmers: NULL argument");
    if (n == 0) return ORPH_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = ctx->stream;
    int rc;
    const uint64_t key_bytes = (uint64_t)len * n;
    const uint64_t key_pad = (key_bytes + 15) & ~15ULL;
    if ((rc = ctx->misc.ensure(key_pad + n * 8 + 16)) != ORPH_OK) return rc;
    uint8_t *d_k = (uint8_t *)ctx->misc.p;
    uint64_t *d_o = (uint64_t *)((uint8_t *)ctx->misc.p + key_pad);
    if (key_bytes) CUDA_TRY(cudaMemcpyAsync(d_k, keys, key_bytes, cudaMemcpyHostToDevice, s));
    hash_kmers_kernel<<<grid_for(n, 256, ctx->sm_count), 256, 0, s>>>(d_k, len, n, d_o);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(out_hashes, d_o, n * 8, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return ORPH_OK;
}

// ------------------------------------------------------------------------------------------------
// scoring
// ------------------------------------------------------------------------------------------------
static uint32_t next_pow2(uint32_t x)
{
    uint32_t p = 1;
    while (p < x) p <<= 1;
    return p;
}

// Pin the filter's bit tables in L2 when they fit (BASELINE north_star): a persisting access-policy
// window on the context's stream over the single contiguous filter allocation.
static void apply_l2_window(orph_ctx *ctx, const orph_filter *f)
{
    if (ctx->l2_window_base == f->buf && ctx->l2_window_bytes == f->n_bytes) return;
    ctx->l2_window_base = f->buf;
    ctx->l2_window_bytes = f->n_bytes;
    static const bool no_window = env_flag("ORPHEUM_B200_NO_L2_WINDOW");
    if (no_window) return;
    if (ctx->l2_persist_max == 0 || ctx->l2_window_max == 0) return;
    // The persisting carve-out is a DEVICE-wide limit shared by every context of the process: it only ever grows, to the
    // largest filter seen on the device (capped by the hardware), so two contexts with different filters do not shrink it
    // under each other.  Not simply the maximum: the carve-out is taken from the L2 the streaming kernels use (measured,
    // profiles/r02_kernel_ab_stage_split_carve.log: finalize_best 0.13 -> 0.24 ms per 10 M reads with the maximum carved
    // out for a 50 MB filter).
    size_t want = std::min<size_t>(f->n_bytes, ctx->l2_persist_max);
    {
        static std::mutex mu;
        static size_t carved[64] = {0};
        std::lock_guard<std::mutex> lk(mu);
        const int d = ctx->device < 64 ? ctx->device : 63;
        if (want > carved[d]) {
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) cudaGetLastError();
            carved[d] = want;
        }
    }
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = f->buf;
    attr.accessPolicyWindow.num_bytes = std::min<size_t>(f->n_bytes, ctx->l2_window_max);
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    if (cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
}

// Raises a kernel's dynamic shared-memory allowance to the device's opt-in maximum (never to the size of the launch at hand):
// the attribute belongs to the kernel, not to the launch, and contexts on several host threads launch the same kernels with
// different sizes -- a thread that set it to ITS size between another thread's set and launch made that launch fail.
template <class Kernel>
static cudaError_t allow_max_dynamic_smem(const orph_ctx *ctx, Kernel kern, size_t needed)
{
    cudaFuncAttributes attr;
    const cudaError_t e = cudaFuncGetAttributes(&attr, kern);
    if (e != cudaSuccess) return e;
    const size_t room = (size_t)ctx->max_smem_optin - std::min<size_t>((size_t)ctx->max_smem_optin, attr.sharedSizeBytes);  // static + dynamic share the limit
    if (needed > room) return cudaErrorInvalidValue;
    if ((size_t)attr.maxDynamicSharedSizeBytes >= room) return cudaSuccess;  // already raised
    return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)room);
}

// Resident blocks per SM of a kernel at a block size and dynamic shared-memory size: asked of the runtime once per
// (device, kernel, geometry) of the process, not on every chunk.
template <class Kernel>
static int resident_blocks(const orph_ctx *ctx, Kernel kern, int block, size_t smem, int *out)
{
    struct Key {
        int device;
        const void *kern;
        int block;
        size_t smem;
        bool operator<(const Key &o) const { return std::tie(device, kern, block, smem) < std::tie(o.device, o.kern, o.block, o.smem); }
    };
    static std::mutex mu;
    static std::map<Key, int> cache;
    const Key key{ctx->device, reinterpret_cast<const void *>(kern), block, smem};
    {
        std::lock_guard<std::mutex> lk(mu);
        auto it = cache.find(key);
        if (it != cache.end()) {
            *out = it->second;
            return ORPH_OK;
        }
    }
    int per_sm = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, block, smem));
    if (per_sm < 1) per_sm = 1;
    std::lock_guard<std::mutex> lk(mu);
    cache[key] = per_sm;
    *out = per_sm;
    return ORPH_OK;
}

template <int K>
static int launch_score_frames(orph_ctx *ctx, const ScoreParams &P, const uint32_t *packed32, const uint64_t *d_offsets,
                               uint64_t pbase0, const uint32_t *queue, ScoreCounters *counters, orph_read_result *d_out)
{
    const int warps = 8;
    const uint32_t per_warp = (P.slots * 4 + P.read_words * 4 + P.pep_bytes + 15) & ~15u;
    const size_t smem = (size_t)per_warp * warps;
    if (smem > (size_t)ctx->max_smem_optin) return fail(ORPH_E_INVALID, "score_frames: %zu bytes of shared memory needed", smem);
    auto kern = score_frames_kernel<K>;
    CUDA_TRY(allow_max_dynamic_smem(ctx, kern, smem));
    int per_sm = 0, rc_occ;
    if ((rc_occ = resident_blocks(ctx, kern, warps * 32, smem, &per_sm)) != ORPH_OK) return rc_occ;
    {
        Timed timed(ctx, 2);
        kern<<<ctx->sm_count * per_sm, warps * 32, smem, ctx->stream>>>(P, packed32, d_offsets, pbase0, queue, counters, d_out);
    }
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return ORPH_OK;
}

// score_tasks_kernel's per-thread column: the 2-bit strand of a read of up to tier_len bases (incl. the words a look-ahead
// window fetch may touch), later the frame's residues, four to a word, plus the two words the k-mer loop fetches ahead
static uint32_t task_stage_words(uint32_t tier_len) { return (2 * tier_len + 34) / 32 + 2; }
static uint32_t task_column_words(uint32_t tier_len)
{
    return std::max(task_stage_words(tier_len), (tier_len / 3 + 3) / 4 + 2);
}

#ifndef ORPH_TINY4
#define ORPH_TINY4 1  // 0: the generic table geometry code for every filter (A/B switch)
#endif

// four tables of at most kTinyPrimeMax bits: the geometry score_tasks_kernel<K, NM, true> is compiled for
static bool filter_is_tiny4(const FilterDev &F)
{
#if ORPH_TINY4
    return F.n_tables == 4 && F.tiny;
#else
    (void)F;
    return false;
#endif
}

template <int K, bool T4>
static int launch_score_tasks_geom(orph_ctx *ctx, const ScoreParams &P_in, uint32_t tier_len, int long_tier, const uint64_t *packed,
                                   const uint64_t *d_offsets, uint64_t pbase0, const uint32_t *tasks, const uint32_t *tasks_sorted,
                                   ScoreCounters *counters, orph_read_result *d_out)
{
    ScoreParams P = P_in;
    P.read_words = task_column_words(tier_len);
    P.stage_words = task_stage_words(tier_len);
    P.seen_words = long_tier ? kSeenWordsLong : kSeenWordsShort;
    const size_t smem = (size_t)kTaskBlock * (P.seen_words * 8 + P.read_words * 4);
    auto kern = score_tasks_kernel<K, false, T4>;
    // attribute + occupancy query once per (device, smem size): they cost tens of microseconds per call
    static thread_local int cached_dev = -1, cached_per_sm[2] = {0, 0};
    static thread_local size_t cached_smem[2] = {0, 0};
    if (cached_dev != ctx->device || cached_smem[long_tier] != smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)kTaskBlock * (kSeenWordsLong * 8 + task_column_words(kThreadMaxLenLong) * 4))));
        int q = 0;
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, kern, kTaskBlock, smem));
        if (cached_dev != ctx->device) cached_smem[0] = cached_smem[1] = 0;
        cached_dev = ctx->device;
        cached_smem[long_tier] = smem;
        cached_per_sm[long_tier] = q < 1 ? 1 : q;
    }
    const int per_sm = cached_per_sm[long_tier];
    {
        Timed timed(ctx, 4);
        kern<<<ctx->sm_count * per_sm, kTaskBlock, smem, ctx->stream>>>(P, reinterpret_cast<const uint32_t *>(packed), d_offsets,
                                                                          pbase0, tasks, tasks_sorted, long_tier, counters, d_out, nullptr);
    }
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return ORPH_OK;
}

template <int K>
static int launch_score_tasks(orph_ctx *ctx, const ScoreParams &P, uint32_t tier_len, int long_tier, const uint64_t *packed,
                              const uint64_t *d_offsets, uint64_t pbase0, const uint32_t *tasks, const uint32_t *tasks_sorted,
                              ScoreCounters *counters, orph_read_result *d_out)
{
    if (filter_is_tiny4(P.F))
        return launch_score_tasks_geom<K, true>(ctx, P, tier_len, long_tier, packed, d_offsets, pbase0, tasks, tasks_sorted, counters, d_out);
    return launch_score_tasks_geom<K, false>(ctx, P, tier_len, long_tier, packed, d_offsets, pbase0, tasks, tasks_sorted, counters, d_out);
}

// the instantiation with an N plane, over the queue of reads that hold N (ASCII input only; usually a few per cent of a batch)
template <int K>
static int launch_score_tasks_n(orph_ctx *ctx, const ScoreParams &P_in, uint32_t tier_len, const uint64_t *packed, const uint64_t *d_offsets,
                                uint64_t pbase0, const uint32_t *tasks_n, const uint32_t *nmask, ScoreCounters *counters, orph_read_result *d_out)
{
    ScoreParams P = P_in;
    P.read_words = task_column_words(tier_len);
    P.stage_words = task_stage_words(tier_len);
    P.seen_words = tier_len > kThreadMaxLenShort ? kSeenWordsLong : kSeenWordsShort;
    const uint32_t mask_words = (P.stage_words + 1) / 2 + 1;
    const size_t smem = (size_t)kTaskBlock * (P.seen_words * 8 + P.read_words * 4 + mask_words * 4);
    auto kern = score_tasks_kernel<K, true, false>;
    CUDA_TRY(allow_max_dynamic_smem(ctx, kern, smem));
    int per_sm = 0, rc_occ;
    if ((rc_occ = resident_blocks(ctx, kern, kTaskBlock, smem, &per_sm)) != ORPH_OK) return rc_occ;
    {
        // few tasks (a few per cent of the batch): one block per SM is plenty, and leaves the SMs to the main kernel
        Timed timed(ctx, 4, ctx->side);
        kern<<<ctx->sm_count * std::min(per_sm, 1), kTaskBlock, smem, ctx->side>>>(P, reinterpret_cast<const uint32_t *>(packed), d_offsets, pbase0, tasks_n, nullptr, 2,
                                                                                    counters, d_out, nmask);
    }
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    return ORPH_OK;
}

template <int K>
static int launch_score_task_tiers(orph_ctx *ctx, const ScoreParams &P, uint32_t thread_len, uint32_t short_len, const uint64_t *packed,
                                   const uint64_t *d_offsets, uint64_t pbase0, const uint32_t *tasks, const uint32_t *tasks_sorted,
                                   ScoreCounters *counters, orph_read_result *d_out)
{
    int rc = launch_score_tasks<K>(ctx, P, short_len, 0, packed, d_offsets, pbase0, tasks, tasks_sorted, counters, d_out);
    if (rc == ORPH_OK && thread_len > short_len)
        rc = launch_score_tasks<K>(ctx, P, thread_len, 1, packed, d_offsets, pbase0, tasks, tasks_sorted, counters, d_out);
    return rc;
}

// k-mer sizes with a compiled thread-per-frame kernel (the defaults of every alphabet, sequence_encodings.py:402-419)
static bool has_task_kernel(uint32_t k)
{
    switch (k) {
    case 7: case 8: case 9: case 10: case 11: case 12: case 16: case 21: case 31: return true;
    default: return false;
    }
}

// Core pipeline on device-resident inputs.  Exactly one of (d_ascii, d_packed_in) is the source:
//   ASCII : pack_ascii -> prefilter -> score_frames -> score_bytes<ASCII> on flagged reads
//   PACKED: prefilter -> score_frames -> score_bytes<PACKED> on over-long reads
//   PACKED + N plane (d_nmask_in / d_has_n_in): as PACKED, reads with N on the N-aware kernels
static int score_device_impl(orph_ctx *ctx, const orph_filter *f, const uint8_t *d_ascii, uint64_t abase0,
                             const uint64_t *d_packed_in, uint64_t pbase0_in, uint64_t total_bases_ascii,
                             const uint64_t *d_offsets, uint64_t n_reads, uint32_t max_read_len, const uint8_t lut[256],
                             double thr, orph_read_result *d_out, const uint32_t *d_nmask_in = nullptr, const uint8_t *d_has_n_in = nullptr)
{
    if (n_reads > (1ULL << 26)) return fail(ORPH_E_INVALID, "internal: more than 2^26 reads in one launch");
    cudaStream_t s = ctx->stream;
    int rc;
    const uint32_t bound = (max_read_len == 0 || max_read_len > ORPH_MAX_READ_LEN) ? ORPH_MAX_READ_LEN : max_read_len;
    const uint32_t fast_len = std::min(bound, kFastMaxLen);
    if ((rc = ctx->queue.ensure(n_reads * 4)) != ORPH_OK) return rc;
    if ((rc = ctx->bytes_queue.ensure(n_reads * 4)) != ORPH_OK) return rc;
    const uint32_t thread_len = has_task_kernel(f->ksize) && !ctx->disable_task_kernel ? std::min(bound, kThreadMaxLen) : 0u;
    const uint32_t short_len = std::min(thread_len, kThreadMaxLenShort);
    if (thread_len && (rc = ctx->task_queue.ensure(n_reads * 6 * 4)) != ORPH_OK) return rc;
    if (thread_len && (rc = ctx->task_queue_sorted.ensure(n_reads * 6 * 4)) != ORPH_OK) return rc;
    ScoreCounters *counters = (ScoreCounters *)ctx->counters.p;
    // status is sticky across calls until orph_ctx_check; everything else restarts
    CUDA_TRY(cudaMemsetAsync(counters, 0, offsetof(ScoreCounters, status), s));
    apply_l2_window(ctx, f);

    ScoreParams P;
    memset(&P, 0, sizeof(P));
    P.F = filter_dev(f);
    memcpy(P.lut, lut, 256);
    P.jaccard_threshold = thr;
    P.k = f->ksize;
    P.max_len = bound;
    P.fast_max_len = fast_len;
    P.thread_max_len = thread_len;
    P.thread_short_len = short_len;
    P.task_capacity = (uint32_t)(n_reads * 6);
    P.force_exact = ctx->force_exact ? 1u : 0u;
    // the warp kernel's de-duplication slots keep 11 position bits and reserve position 2047 for "empty": with k = 1 a
    // 2048-residue frame would reach it, so those reads take the block-per-read kernel
    const uint32_t bytes_max_len = f->ksize == 1 ? kBytesMaxLen - 3 : kBytesMaxLen;

    const uint64_t *d_packed = d_packed_in;
    uint64_t pbase0 = pbase0_in;
    const uint8_t *d_needs = nullptr, *d_has_n = nullptr;
    const uint32_t *d_nmask = nullptr;
    uint32_t *d_tasks_n = nullptr;
    if (d_ascii) {
        const uint64_t n_words = (total_bases_ascii + 31) / 32;
        if ((rc = ctx->packed.ensure((n_words + 2) * 8)) != ORPH_OK) return rc;
        if ((rc = ctx->needs_bytes.ensure(2 * n_reads)) != ORPH_OK) return rc;
        if ((rc = ctx->nmask.ensure((n_words + 2) * 4)) != ORPH_OK) return rc;
        if (thread_len && (rc = ctx->task_queue_n.ensure(n_reads * 7 * 4)) != ORPH_OK) return rc;
        CUDA_TRY(cudaMemsetAsync(ctx->needs_bytes.p, 0, 2 * n_reads, s));
        pbase0 = abase0;
        if (n_words) {
            Timed timed(ctx, 0);
            pack_ascii_kernel<<<grid_for(n_words, 256, ctx->sm_count, 16), 256, 0, s>>>(
                d_ascii, abase0, d_offsets, n_reads, pbase0, total_bases_ascii, (uint64_t *)ctx->packed.p, (uint8_t *)ctx->needs_bytes.p,
                (uint32_t *)ctx->nmask.p, (uint8_t *)ctx->needs_bytes.p + n_reads);
            ctx->launches++;
            CUDA_TRY(cudaGetLastError());
        }
        d_packed = (const uint64_t *)ctx->packed.p;
        d_needs = (const uint8_t *)ctx->needs_bytes.p;
        d_has_n = d_needs + n_reads;
        d_nmask = (const uint32_t *)ctx->nmask.p;
        if (thread_len) d_tasks_n = (uint32_t *)ctx->task_queue_n.p;
    } else if (d_nmask_in && d_has_n_in) {
        if (thread_len && (rc = ctx->task_queue_n.ensure(n_reads * 7 * 4)) != ORPH_OK) return rc;
        d_has_n = d_has_n_in;
        d_nmask = d_nmask_in;
        if (thread_len) d_tasks_n = (uint32_t *)ctx->task_queue_n.p;
    }

    {
        Timed timed(ctx, 1);
        prefilter_kernel<<<grid_for(n_reads, 256, ctx->sm_count, 16), 256, 0, s>>>(
            d_packed, d_offsets, pbase0, n_reads, d_needs, P.k, fast_len, thread_len, short_len, P.task_capacity, bound, bytes_max_len, d_out,
            (uint32_t *)ctx->queue.p,
            (uint32_t *)ctx->task_queue.p, (uint32_t *)ctx->bytes_queue.p, counters, d_nmask, d_has_n, d_tasks_n);
    }
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());

    // short reads: one thread per open frame
    if (thread_len) {
        const uint32_t *tq = (const uint32_t *)ctx->task_queue.p;
        const uint32_t *tq_sorted = (const uint32_t *)ctx->task_queue_sorted.p;
        if (d_tasks_n) {
            // reads that hold N: their own prefilter (warp per read), then the same kernel with an N plane -- on the side
            // stream, so that this small, latency-bound launch runs next to the main kernel instead of after it
            if (!ctx->side) {
                CUDA_TRY(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
                CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
                CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
            }
            CUDA_TRY(cudaEventRecord(ctx->ev_fork, s));
            CUDA_TRY(cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
            {
                Timed timed(ctx, 1, ctx->side);
                prefilter_n_kernel<<<grid_for(n_reads, 256, ctx->sm_count, 4), 256, 0, ctx->side>>>(d_packed, d_nmask, d_offsets, pbase0, P.k, P.task_capacity, d_tasks_n,
                                                                                                 counters, d_out);
            }
            ctx->launches++;
            CUDA_TRY(cudaGetLastError());
            switch (P.k) {
            case 7: rc = launch_score_tasks_n<7>(ctx, P, thread_len, d_packed, d_offsets, pbase0, d_tasks_n, d_nmask, counters, d_out); break;
            case 8: rc = launch_score_tasks_n<8>(ctx, P, thread_len, d_packed, d_offsets, pbase0, d_tasks_n, d_nmask, counters, d_out); break;
            case 9: rc = launch_score_tasks_n<9>(ctx, P, thread_len, d_packed, d_offsets, pbase0, d_tasks_n, d_nmask, counters, d_out); break;
            case 10: rc = launch_score_tasks_n<10>(ctx, P, thread_len, d_packed, d_offsets, pbase0, d_tasks_n, d_nmask, counters, d_out); break;
            case 11: rc = launch_score_tasks_n<11>(ctx, P, thread_len, d_packed, d_offsets, pbase0, d_tasks_n, d_nmask, counters, d_out); break;
            case 12: rc = launch_score_tasks_n<12>(ctx, P, thread_len, d_packed, d_offsets, pbase0, d_tasks_n, d_nmask, counters, d_out); break;
            case 16: rc = launch_score_tasks_n<16>(ctx, P, thread_len, d_packed, d_offsets, pbase0, d_tasks_n, d_nmask, counters, d_out); break;
            case 21: rc = launch_score_tasks_n<21>(ctx, P, thread_len, d_packed, d_offsets, pbase0, d_tasks_n, d_nmask, counters, d_out); break;
            case 31: rc = launch_score_tasks_n<31>(ctx, P, thread_len, d_packed, d_offsets, pbase0, d_tasks_n, d_nmask, counters, d_out); break;
            default: break;
            }
            if (rc != ORPH_OK) return rc;
            CUDA_TRY(cudaEventRecord(ctx->ev_join, ctx->side));
        }
        {
            Timed timed(ctx, 6);
            task_scan_kernel<<<1, 64, 0, s>>>(counters);
            task_scatter_kernel<<<grid_for(n_reads / 8 + 1, 256, ctx->sm_count, ORPH_SORT_LOCAL ? 2 : 4), kScatterBlock, 0, s>>>(
                d_offsets, tq, (uint32_t *)ctx->task_queue_sorted.p, P.task_capacity, counters);
        }
        ctx->launches += 2;
        CUDA_TRY(cudaGetLastError());
        switch (P.k) {
        case 7: rc = launch_score_task_tiers<7>(ctx, P, thread_len, short_len, d_packed, d_offsets, pbase0, tq, tq_sorted, counters, d_out); break;
        case 8: rc = launch_score_task_tiers<8>(ctx, P, thread_len, short_len, d_packed, d_offsets, pbase0, tq, tq_sorted, counters, d_out); break;
        case 9: rc = launch_score_task_tiers<9>(ctx, P, thread_len, short_len, d_packed, d_offsets, pbase0, tq, tq_sorted, counters, d_out); break;
        case 10: rc = launch_score_task_tiers<10>(ctx, P, thread_len, short_len, d_packed, d_offsets, pbase0, tq, tq_sorted, counters, d_out); break;
        case 11: rc = launch_score_task_tiers<11>(ctx, P, thread_len, short_len, d_packed, d_offsets, pbase0, tq, tq_sorted, counters, d_out); break;
        case 12: rc = launch_score_task_tiers<12>(ctx, P, thread_len, short_len, d_packed, d_offsets, pbase0, tq, tq_sorted, counters, d_out); break;
        case 16: rc = launch_score_task_tiers<16>(ctx, P, thread_len, short_len, d_packed, d_offsets, pbase0, tq, tq_sorted, counters, d_out); break;
        case 21: rc = launch_score_task_tiers<21>(ctx, P, thread_len, short_len, d_packed, d_offsets, pbase0, tq, tq_sorted, counters, d_out); break;
        case 31: rc = launch_score_task_tiers<31>(ctx, P, thread_len, short_len, d_packed, d_offsets, pbase0, tq, tq_sorted, counters, d_out); break;
        default: rc = fail(ORPH_E_INVALID, "internal: no task kernel for k=%u", P.k); break;
        }
        if (rc != ORPH_OK) return rc;
        if (d_tasks_n) CUDA_TRY(cudaStreamWaitEvent(s, ctx->ev_join, 0));  // the N reads' kernels have written their records
        {
            Timed timed(ctx, 5);
            finalize_best_kernel<<<grid_for(n_reads, 256, ctx->sm_count, 16), 256, 0, s>>>(d_out, n_reads);
        }
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
    }

    // medium reads (or k without a task kernel): one warp per read
    if (bound > thread_len) {
    P.slots = std::max(128u, next_pow2(4 * std::max(1u, fast_len / 3)));
    P.pep_bytes = (fast_len / 3 + kPepPad + 3) & ~3u;
    P.read_words = (2 * fast_len + 62) / 32 + 2;
    const uint32_t *packed32 = reinterpret_cast<const uint32_t *>(d_packed);
    switch (P.k) {
    case 7: rc = launch_score_frames<7>(ctx, P, packed32, d_offsets, pbase0, (const uint32_t *)ctx->queue.p, counters, d_out); break;
    case 12: rc = launch_score_frames<12>(ctx, P, packed32, d_offsets, pbase0, (const uint32_t *)ctx->queue.p, counters, d_out); break;
    case 31: rc = launch_score_frames<31>(ctx, P, packed32, d_offsets, pbase0, (const uint32_t *)ctx->queue.p, counters, d_out); break;
    default: rc = launch_score_frames<0>(ctx, P, packed32, d_offsets, pbase0, (const uint32_t *)ctx->queue.p, counters, d_out); break;
    }
    if (rc != ORPH_OK) return rc;
    }

    // byte-exact path for whatever prefilter routed away (non-ACGT bytes, reads longer than the fast path)
    if (d_ascii || d_nmask || bound > fast_len) {
        ScoreParams B = P;
        const uint32_t bbound = std::min(bound, bytes_max_len);
        B.max_len = bbound;
        B.slots = std::max(128u, next_pow2(2 * std::max(1u, bbound / 3)));
        B.pep_bytes = (bbound / 3 + kPepPad + 3) & ~3u;
        B.read_words = 0;
        const uint32_t per_warp = (B.slots * 4 + B.pep_bytes + ((bbound + 7) & ~3u) + 15) & ~15u;
        int warps = (int)std::min<size_t>(8, std::max<size_t>(1, (size_t)96 * 1024 / per_warp));
        const size_t smem = (size_t)per_warp * warps;
        Timed timed(ctx, 3);
        if (d_ascii) {
            auto kern = score_bytes_kernel<false>;
            CUDA_TRY(allow_max_dynamic_smem(ctx, kern, smem));
            int per_sm = 0;
            if ((rc = resident_blocks(ctx, kern, warps * 32, smem, &per_sm)) != ORPH_OK) return rc;
            kern<<<ctx->sm_count * std::max(1, per_sm), warps * 32, smem, s>>>(B, d_ascii, d_offsets, abase0, n_reads, (const uint32_t *)ctx->bytes_queue.p, counters, d_out,
                                                                               nullptr);
        } else {
            auto kern = score_bytes_kernel<true>;
            CUDA_TRY(allow_max_dynamic_smem(ctx, kern, smem));
            int per_sm = 0;
            if ((rc = resident_blocks(ctx, kern, warps * 32, smem, &per_sm)) != ORPH_OK) return rc;
            kern<<<ctx->sm_count * std::max(1, per_sm), warps * 32, smem, s>>>(B, d_packed, d_offsets, pbase0, n_reads, (const uint32_t *)ctx->bytes_queue.p, counters, d_out,
                                                                               d_nmask);
        }
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
    }

    // reads beyond the warp kernel's shared-memory staging: one block per read, scratch in global memory
    if (bound > bytes_max_len) {
        const int grid = ctx->sm_count;
        if ((rc = ctx->long_scratch.ensure(long_scratch_bytes(bound) * grid)) != ORPH_OK) return rc;
        Timed timed(ctx, 7);
        if (d_ascii)
            score_long_kernel<false><<<grid, kLongBlock, 0, s>>>(P, d_ascii, d_offsets, abase0, (const uint32_t *)ctx->bytes_queue.p, (uint32_t)n_reads,
                                                              counters, d_out, (uint8_t *)ctx->long_scratch.p, nullptr);
        else
            score_long_kernel<true><<<grid, kLongBlock, 0, s>>>(P, d_packed, d_offsets, pbase0, (const uint32_t *)ctx->bytes_queue.p, (uint32_t)n_reads,
                                                             counters, d_out, (uint8_t *)ctx->long_scratch.p, d_nmask);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return ORPH_OK;
}

static int check_score_args(orph_ctx *ctx, const orph_filter *f, const void *reads, int fmt, const uint64_t *offsets,
                            uint64_t n_reads, const uint8_t *lut, double thr, const void *out)
{
    if (!ctx || !f || !lut) return fail(ORPH_E_INVALID, "orph_score_reads: NULL ctx / filter / lut");
    if (n_reads && (!reads || !offsets || !out)) return fail(ORPH_E_INVALID, "orph_score_reads: NULL buffer");
    if (fmt != ORPH_READS_ASCII && fmt != ORPH_READS_PACKED2) return fail(ORPH_E_INVALID, "orph_score_reads: unknown reads_format %d", fmt);
    if (f->device != ctx->device) return fail(ORPH_E_INVALID, "orph_score_reads: filter lives on device %d, context on %d", f->device, ctx->device);
    if (std::isnan(thr)) return fail(ORPH_E_INVALID, "orph_score_reads: jaccard_threshold is NaN");
    return ORPH_OK;
}

extern "C" int orph_score_reads_device(orph_ctx *ctx, const orph_filter *f, const void *d_reads, int reads_format,
                                       const uint64_t *d_offsets, uint64_t n_reads, uint32_t max_read_len,
                                       const uint8_t lut[256], double jaccard_threshold, orph_read_result *d_out)
{
    int rc = check_score_args(ctx, f, d_reads, reads_format, d_offsets, n_reads, lut, jaccard_threshold, d_out);
    if (rc != ORPH_OK) return rc;
    if (n_reads == 0) return ORPH_OK;
    if (reinterpret_cast<uintptr_t>(d_reads) & 15) return fail(ORPH_E_INVALID, "orph_score_reads_device: reads must be 16-byte aligned");
    if (reinterpret_cast<uintptr_t>(d_out) & 15) return fail(ORPH_E_INVALID, "orph_score_reads_device: out must be 16-byte aligned");
    DeviceGuard guard(ctx->device);
    // one launch set per 2^26 reads (queue entries keep 6 bits for the open-frame mask)
    const uint64_t kMaxLaunchReads = 1ULL << 26;
    for (uint64_t lo = 0; lo < n_reads; lo += kMaxLaunchReads) {
        const uint64_t m = std::min(kMaxLaunchReads, n_reads - lo);
        if (reads_format == ORPH_READS_PACKED2) {
            rc = score_device_impl(ctx, f, nullptr, 0, (const uint64_t *)d_reads, 0, 0, d_offsets + lo, m, max_read_len, lut,
                                   jaccard_threshold, d_out + lo);
        } else {
            // ASCII: the size of the packed scratch depends on the first and last offset, which live on the device
            uint64_t ends[2];
            CUDA_TRY(cudaMemcpyAsync(&ends[0], d_offsets + lo, 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaMemcpyAsync(&ends[1], d_offsets + lo + m, 8, cudaMemcpyDeviceToHost, ctx->stream));
            CUDA_TRY(cudaStreamSynchronize(ctx->stream));
            if (ends[1] < ends[0]) return fail(ORPH_E_INVALID, "orph_score_reads_device: offsets are not monotonic");
            // d_reads[0] is absolute base 0; pack only [ends[0], ends[1])
            rc = score_device_impl(ctx, f, (const uint8_t *)d_reads + ends[0], ends[0], nullptr, 0, ends[1] - ends[0], d_offsets + lo, m,
                                   max_read_len, lut, jaccard_threshold, d_out + lo);
        }
        if (rc != ORPH_OK) return rc;
    }
    return ORPH_OK;
}

extern "C" int orph_pack_reads_device(orph_ctx *ctx, const uint8_t *d_ascii, const uint64_t *d_offsets, uint64_t n_reads, uint64_t total_bases,
                                      uint64_t *d_packed_out, uint32_t *d_nmask_out, uint8_t *d_has_n_out, uint8_t *d_needs_bytes_out)
{
    if (!ctx || (n_reads && (!d_ascii || !d_offsets || !d_packed_out || !d_nmask_out || !d_has_n_out || !d_needs_bytes_out)))
        return fail(ORPH_E_INVALID, "orph_pack_reads_device: NULL argument");
    if (n_reads == 0) return ORPH_OK;
    DeviceGuard guard(ctx->device);
    cudaStream_t s = ctx->stream;
    const uint64_t n_words = (total_bases + 31) / 32;
    CUDA_TRY(cudaMemsetAsync(d_has_n_out, 0, n_reads, s));
    CUDA_TRY(cudaMemsetAsync(d_needs_bytes_out, 0, n_reads, s));
    if (n_words) {
        pack_ascii_kernel<<<grid_for(n_words, 256, ctx->sm_count, 16), 256, 0, s>>>(d_ascii, 0, d_offsets, n_reads, 0, total_bases, d_packed_out, d_needs_bytes_out,
                                                                                     d_nmask_out, d_has_n_out);
        ctx->launches++;
        CUDA_TRY(cudaGetLastError());
    }
    return ORPH_OK;
}

extern "C" int orph_score_reads_device_n(orph_ctx *ctx, const orph_filter *f, const void *d_packed, const uint32_t *d_nmask, const uint8_t *d_has_n,
                                         const uint64_t *d_offsets, uint64_t n_reads, uint32_t max_read_len, const uint8_t lut[256],
                                         double jaccard_threshold, orph_read_result *d_out)
{
    int rc = check_score_args(ctx, f, d_packed, ORPH_READS_PACKED2, d_offsets, n_reads, lut, jaccard_threshold, d_out);
    if (rc != ORPH_OK) return rc;
    if (n_reads == 0) return ORPH_OK;
    if (!d_nmask || !d_has_n) return fail(ORPH_E_INVALID, "orph_score_reads_device_n: NULL N plane");
    if ((reinterpret_cast<uintptr_t>(d_packed) & 15) || (reinterpret_cast<uintptr_t>(d_out) & 15))
        return fail(ORPH_E_INVALID, "orph_score_reads_device_n: reads and out must be 16-byte aligned");
    DeviceGuard guard(ctx->device);
    const uint64_t kMaxLaunchReads = 1ULL << 26;
    for (uint64_t lo = 0; lo < n_reads; lo += kMaxLaunchReads) {
        const uint64_t m = std::min(kMaxLaunchReads, n_reads - lo);
        rc = score_device_impl(ctx, f, nullptr, 0, (const uint64_t *)d_packed, 0, 0, d_offsets + lo, m, max_read_len, lut, jaccard_threshold, d_out + lo, d_nmask,
                               d_has_n + lo);
        if (rc != ORPH_OK) return rc;
    }
    return ORPH_OK;
}

#include "ingest_pipeline.inl"
#include "stream.inl"

extern "C" int orph_translate_frames(orph_ctx *ctx, const uint8_t *reads, const uint64_t *offsets, uint64_t n_reads,
                                     const uint64_t *item_read, const int8_t *item_frame, uint64_t n_items,
                                     const uint64_t *out_offsets, uint8_t *out_peptides)
{
    if (!ctx) return fail(ORPH_E_INVALID, "orph_translate_frames: ctx is NULL");
    if (n_items == 0) return ORPH_OK;
    if (!reads || !offsets || !item_read || !item_frame || !out_offsets || !out_peptides || n_reads == 0)
        return fail(ORPH_E_INVALID, "orph_translate_frames: NULL argument");
    for (uint64_t i = 0; i < n_items; i++) {
        const int f = item_frame[i];
        if (item_read[i] >= n_reads || f == 0 || f < -3 || f > 3)
            return fail(ORPH_E_INVALID, "orph_translate_frames: item %llu is (read %llu, frame %d)", (unsigned long long)i,
                        (unsigned long long)item_read[i], f);
        const uint64_t L = offsets[item_read[i] + 1] - offsets[item_read[i]];
        const uint64_t phase = (uint64_t)(f > 0 ? f - 1 : -f - 1);
        const uint64_t plen = L >= phase ? (L - phase) / 3 : 0;
        if (out_offsets[i + 1] - out_offsets[i] != plen)
            return fail(ORPH_E_INVALID, "orph_translate_frames: out_offsets of item %llu holds %llu bytes, the frame has %llu residues",
                        (unsigned long long)i, (unsigned long long)(out_offsets[i + 1] - out_offsets[i]), (unsigned long long)plen);
    }
    DeviceGuard guard(ctx->device);
    cudaStream_t s = ctx->stream;
    const uint64_t b0 = offsets[0], n_bytes = offsets[n_reads] - b0, n_out = out_offsets[n_items] - out_offsets[0];
    int rc;
    if ((rc = ctx->in_reads.ensure(n_bytes + 64)) != ORPH_OK) return rc;
    if ((rc = ctx->in_offsets.ensure((n_reads + 1) * 8)) != ORPH_OK) return rc;
    // items: [read u64 | out offset u64 (n+1) | frame i8], then the output bytes
    const uint64_t items_bytes = n_items * 8 + (n_items + 1) * 8 + ((n_items + 15) & ~15ULL);
    if ((rc = ctx->misc.ensure(items_bytes + n_out + 64)) != ORPH_OK) return rc;
    uint8_t *base = (uint8_t *)ctx->misc.p;
    uint64_t *d_item_read = (uint64_t *)base;
    uint64_t *d_out_off = (uint64_t *)(base + n_items * 8);
    int8_t *d_item_frame = (int8_t *)(base + n_items * 8 + (n_items + 1) * 8);
    uint8_t *d_out = base + items_bytes;
    if (n_bytes) CUDA_TRY(cudaMemcpyAsync(ctx->in_reads.p, reads + b0, n_bytes, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(ctx->in_offsets.p, offsets, (n_reads + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_item_read, item_read, n_items * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_out_off, out_offsets, (n_items + 1) * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(d_item_frame, item_frame, n_items, cudaMemcpyHostToDevice, s));
    translate_frames_kernel<<<grid_for(n_items * 32, 256, ctx->sm_count, 16), 256, 0, s>>>(
        (const uint8_t *)ctx->in_reads.p, b0, (const uint64_t *)ctx->in_offsets.p, d_item_read, d_item_frame, n_items, d_out_off,
        d_out - out_offsets[0]);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    if (n_out) CUDA_TRY(cudaMemcpyAsync(out_peptides, d_out, n_out, cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    return ORPH_OK;
}

extern "C" double orph_jaccard(const orph_read_result *r, int frame_index)
{
    if (!r || frame_index < 0 || frame_index > 5) return NAN;
    const uint8_t c = r->category[frame_index];
    if (c != ORPH_CAT_CODING && c != ORPH_CAT_NON_CODING) return NAN;
    return (double)r->n_hits[frame_index] / (double)r->n_kmers[frame_index];
}

// ------------------------------------------------------------------------------------------------
// host-side 2-bit packer (for callers that ship packed reads over PCIe)
// ------------------------------------------------------------------------------------------------
extern "C" int orph_pack_reads(const uint8_t *reads, const uint64_t *offsets, uint64_t n_reads, uint64_t *out_packed,
                               uint8_t *needs_ascii, int n_threads)
{
    if (!offsets || (n_reads && (!reads || !out_packed))) return fail(ORPH_E_INVALID, "orph_pack_reads: NULL argument");
    if (n_reads == 0) return ORPH_OK;
    const uint64_t b0 = offsets[0], total = offsets[n_reads] - b0;
    const uint64_t n_words = (total + 31) / 32;
    if (needs_ascii) memset(needs_ascii, 0, n_reads);
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<uint64_t>(n_threads, std::max<uint64_t>(1, n_words / 8192));
    if (n_threads <= 1) {
        orph_host::pack_stream(nullptr, reads + b0, b0, total, offsets, n_reads, out_packed, needs_ascii);
    } else {
        orph_host::Pool pool(n_threads);
        orph_host::pack_stream(&pool, reads + b0, b0, total, offsets, n_reads, out_packed, needs_ascii);
    }
    return ORPH_OK;
}

// ------------------------------------------------------------------------------------------------
// measurement helpers
// ------------------------------------------------------------------------------------------------
extern "C" int orph_synth_reads_device(orph_ctx *ctx, uint64_t seed, uint64_t first_index, uint64_t n_reads, uint32_t read_len,
                                       const uint8_t *d_residues, const uint64_t *d_prot_offsets, uint64_t n_proteins,
                                       const uint64_t *d_syn_table, void *d_packed_out)
{
    if (!ctx || !d_residues || !d_prot_offsets || !d_syn_table || !d_packed_out || n_proteins == 0)
        return fail(ORPH_E_INVALID, "orph_synth_reads_device: NULL argument");
    if (read_len < 3 || read_len > kSynthMaxLen) return fail(ORPH_E_INVALID, "orph_synth_reads_device: read_len %u not in 3..%u", read_len, kSynthMaxLen);
    if (n_reads == 0) return ORPH_OK;
    DeviceGuard guard(ctx->device);
    const uint64_t n_words = (n_reads * read_len + 31) / 32 + 1;
    CUDA_TRY(cudaMemsetAsync(d_packed_out, 0, n_words * 8, ctx->stream));
    synth_reads_kernel<<<grid_for(n_reads, 256, ctx->sm_count, 16), 256, 0, ctx->stream>>>(seed, first_index, n_reads, read_len, d_residues, d_prot_offsets,
                                                                                        n_proteins, d_syn_table, (uint32_t *)d_packed_out);
    CUDA_TRY(cudaGetLastError());
    return ORPH_OK;
}

extern "C" int orph_bench_gather_peak(orph_ctx *ctx, uint64_t footprint_bytes, uint32_t loads_per_thread, int persist_l2,
                                      double *sectors_per_second)
{
    if (!ctx || !sectors_per_second || footprint_bytes < 4096 || loads_per_thread < 4)
        return fail(ORPH_E_INVALID, "orph_bench_gather_peak: bad arguments");
    DeviceGuard guard(ctx->device);
    cudaStream_t s = ctx->stream;
    uint8_t *buf = nullptr;
    uint32_t *sink = nullptr;
    CUDA_TRY(cudaMalloc((void **)&buf, footprint_bytes));
    CUDA_TRY(cudaMalloc((void **)&sink, 4));
    CUDA_TRY(cudaMemsetAsync(buf, 1, footprint_bytes, s));
    if (persist_l2 && ctx->l2_persist_max && ctx->l2_window_max) {
        size_t want = std::min<size_t>(footprint_bytes, ctx->l2_persist_max);  // (the device-wide carve-out stays as orph_ctx_create set it)
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        attr.accessPolicyWindow.base_ptr = buf;
        attr.accessPolicyWindow.num_bytes = std::min<size_t>(footprint_bytes, ctx->l2_window_max);
        attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)want / (double)attr.accessPolicyWindow.num_bytes);
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        if (cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr) != cudaSuccess) cudaGetLastError();
        ctx->l2_window_base = nullptr;  // the filter window must be re-applied afterwards
    }
    loads_per_thread &= ~3u;
    const int block = 256, grid = ctx->sm_count * 8;
    cudaEvent_t e0, e1;
    CUDA_TRY(cudaEventCreate(&e0));
    CUDA_TRY(cudaEventCreate(&e1));
    for (int warm = 0; warm < 2; warm++) gather_peak_kernel<<<grid, block, 0, s>>>(buf, footprint_bytes, loads_per_thread, sink);
    const int reps = 5;
    CUDA_TRY(cudaEventRecord(e0, s));
    for (int i = 0; i < reps; i++) gather_peak_kernel<<<grid, block, 0, s>>>(buf, footprint_bytes, loads_per_thread, sink);
    CUDA_TRY(cudaEventRecord(e1, s));
    CUDA_TRY(cudaEventSynchronize(e1));
    ctx->launches += reps + 2;
    float ms = 0;
    CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
    *sectors_per_second = (double)grid * block * loads_per_thread * reps / (ms * 1e-3);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (persist_l2) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);
        cudaCtxResetPersistingL2Cache();
        cudaGetLastError();
    }
    cudaFree(buf);
    cudaFree(sink);
    return ORPH_OK;
}
