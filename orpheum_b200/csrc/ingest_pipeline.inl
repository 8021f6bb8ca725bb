// Host-buffer ingest of orph_score_reads (included by orpheum_b200.cu, same translation unit).  SURVEY.md 8f rank 1; the
// reference's counterpart is the serial loop over screed records in orpheum/translate.py:367-372.
//
// The batch is consumed from both ends by two lanes that meet wherever their speeds put the meeting point:
//   lane 0 (the caller's thread)  claims chunks from the FRONT and hands them to the copy engine as they are (ASCII or
//           already 2-bit) -- no host work beyond the enqueue;
//   lane 1 (a packer thread + the host pool, ASCII input only)  claims chunks from the BACK, packs them to 2 bits per
//           base into pinned staging (a quarter of the PCIe bytes) and ships those.
// An ASCII batch is PCIe-bound at one byte per base and host-memory-bound at two bits per base, so neither lane alone is
// fastest; claiming from opposite ends balances them with no planner and no rate estimates.  Every chunk then flows
// H2D -> kernels -> D2H on three streams (H2D per lane, one kernel stream, one D2H stream), kLaneSlots chunks in flight
// per lane.  Measured on the B200 boxes (tools/copy_sweep.py, tools/e2e_ab.py): an H2D copy costs ~75 us before its first
// byte moves, so chunks are tens of megabytes and carry ONE copy each -- offsets of equal-length reads are generated on
// the device, offsets of host-packed chunks ride behind the packed words as 32-bit deltas; chunk sizes ramp up at the
// start (the first kernels start early) and down at the end (the last kernels and D2H are short).
// Reads that cannot be 2-bit packed (N, lower case, ...) are scored a second time from their bytes (`exceptions`) and
// overwrite the record of the first pass; a chunk where more than 1/8 of the reads are of that kind is sent as ASCII.
#pragma once

#include <mutex>

// max and min over i of off[i+1] - off[i] for i in [0, m); a decreasing pair wraps to >= 2^63.
static void scan_read_lengths(const uint64_t *off, uint64_t m, uint64_t *longest, uint64_t *shortest)
{
    uint64_t m0 = 0, m1 = 0, m2 = 0, m3 = 0, s0 = ~0ull, s1 = ~0ull, s2 = ~0ull, s3 = ~0ull;
    uint64_t i = 0;
    for (; i + 4 <= m; i += 4) {
        const uint64_t a = off[i + 1] - off[i], b = off[i + 2] - off[i + 1], c = off[i + 3] - off[i + 2], d = off[i + 4] - off[i + 3];
        m0 = std::max(m0, a); m1 = std::max(m1, b); m2 = std::max(m2, c); m3 = std::max(m3, d);
        s0 = std::min(s0, a); s1 = std::min(s1, b); s2 = std::min(s2, c); s3 = std::min(s3, d);
    }
    for (; i < m; i++) {
        m0 = std::max(m0, off[i + 1] - off[i]);
        s0 = std::min(s0, off[i + 1] - off[i]);
    }
    *longest = std::max(std::max(m0, m1), std::max(m2, m3));
    *shortest = std::min(std::min(s0, s1), std::min(s2, s3));
}

namespace {

struct ChunkTrace {
    cudaEvent_t e[6] = {};
    double host_enqueue_s = 0, host_pack_s = 0;
    int lane = 0, packed = 0;
    uint64_t lo = 0, m = 0;
};

// state of one orph_score_reads call shared by the two lanes
struct PipeCall {
    orph_ctx *ctx;
    const orph_filter *f;
    const uint8_t *reads;
    int reads_format;
    const uint64_t *offsets;
    uint64_t n_reads;
    const uint8_t *lut;
    double thr;
    orph_read_result *out;
    std::vector<uint64_t> *exceptions;
    std::mutex claim_mu, stats_mu;
    orph_host::Pool *scan_pool = nullptr;  // free for the front lane's offset scans (no packing lane in this call)
    bool share_h2d = false;                // both lanes are running
    uint64_t front = 0, back = 0;  // reads [front, back) are unclaimed
    int n_front_claims = 0;
    int error = ORPH_OK;
    char error_text[512] = "";
    bool trace = false;
    std::chrono::steady_clock::time_point t_call;
    std::vector<ChunkTrace> traces;
    uint64_t max_chunk_reads[2], max_chunk_bases[2], min_chunk_reads = 1u << 16;

    void set_error(int rc)
    {
        std::lock_guard<std::mutex> lk(claim_mu);
        if (error == ORPH_OK) {
            error = rc;
            snprintf(error_text, sizeof(error_text), "%s", g_err);
        }
    }
    // next chunk of a lane: lane 0 takes [front, front + sz), lane 1 takes [back - sz, back)
    bool claim(int lane, uint64_t *lo, uint64_t *hi)
    {
        std::lock_guard<std::mutex> lk(claim_mu);
        if (error != ORPH_OK || front >= back) return false;
        const uint64_t remaining = back - front;
        // sizes shrink as the lanes approach each other (and ramp up at the very start of lane 0), so that the pipeline
        // fills quickly and whatever finishes last is small
        uint64_t sz = std::min<uint64_t>(max_chunk_reads[lane], std::max<uint64_t>(min_chunk_reads, remaining / 3));
        if (lane == 0 && n_front_claims < 3) sz = std::min<uint64_t>(sz, std::max<uint64_t>(min_chunk_reads, max_chunk_reads[0] >> (3 - n_front_claims)));
        sz = std::max<uint64_t>(1, std::min(sz, remaining));
        if (remaining - sz < min_chunk_reads / 2) sz = remaining;  // no crumbs
        if (lane == 0) {
            uint64_t h = front + sz;
            if (offsets[h] >= offsets[front] && offsets[h] - offsets[front] > max_chunk_bases[0]) {
                h = (uint64_t)(std::upper_bound(offsets + front + 1, offsets + h + 1, offsets[front] + max_chunk_bases[0]) - offsets) - 1;
                if (h <= front) h = front + 1;
            }
            *lo = front;
            *hi = h;
            front = h;
            n_front_claims++;
        } else {
            uint64_t l = back - sz;
            if (offsets[back] >= offsets[l] && offsets[back] - offsets[l] > max_chunk_bases[1]) {
                const uint64_t floor_base = offsets[back] - max_chunk_bases[1];
                l = (uint64_t)(std::lower_bound(offsets + l, offsets + back, floor_base) - offsets);
                if (l >= back) l = back - 1;
            }
            *lo = l;
            *hi = back;
            back = l;
        }
        return true;
    }
};

}  // namespace

// Host wait for an event by polling: a blocking-sync wait costs ~0.5 ms of wake-up latency on these hosts (measured in the
// per-chunk timelines: every chunk that really had to wait for its slot was enqueued that much after the slot was free),
// which is a third of a chunk's copy time.  Yields the core once the wait gets long.
static cudaError_t wait_event_polling(cudaEvent_t e)
{
    for (uint32_t spins = 0;; spins++) {
        const cudaError_t rc = cudaEventQuery(e);
        if (rc != cudaErrorNotReady) return rc;
        if (spins > 2000) sched_yield();
    }
}

static bool input_is_pinned(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) == cudaSuccess) return attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    return false;
}

static int pipe_setup(orph_ctx *ctx)
{
    if (ctx->pipe_ready) return ORPH_OK;
    // ONE H2D stream for both lanes: copies leave in the order the lanes submit them.  With a stream per lane the copy
    // engine served the stream that always had a copy queued and starved the other until the first ran dry.
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_in[0], cudaStreamNonBlocking));
    ctx->s_in[1] = ctx->s_in[0];
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_out[0], cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_out[1], cudaStreamNonBlocking));
    for (auto &lane : ctx->lane_slots)
        for (PipeSlot &q : lane) {
            CUDA_TRY(cudaEventCreateWithFlags(&q.ev_in, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&q.ev_kernels, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&q.ev_out, cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreate(&q.ev_h2d_begin));
            CUDA_TRY(cudaEventCreate(&q.ev_h2d_end));
        }
    ctx->pipe_ready = true;
    return ORPH_OK;
}

// One chunk through its three stages.  `packed_here`: the chunk's 2-bit words (and, for reads of unequal length, its
// 32-bit offset deltas behind them) are in the slot's pinned staging; otherwise the bytes come from the caller's buffer.
static int pipe_enqueue_chunk(PipeCall &pc, int lane, PipeSlot &q, uint64_t lo, uint64_t hi, uint32_t max_len, uint64_t fixed_len, bool packed_here,
                              double host_pack_s)
{
    orph_ctx *ctx = pc.ctx;
    // lane 1 launches on its own context: an in-order kernel stream shared by the lanes would make the chunks of one
    // lane wait for the copies of the other
    orph_ctx *kctx = lane == 1 && ctx->lane_ctx ? ctx->lane_ctx : ctx;
    cudaStream_t s = kctx->stream, s_in = ctx->s_in[lane], s_out = ctx->s_out[lane];
    const uint64_t *offsets = pc.offsets;
    const uint64_t m = hi - lo, b0 = offsets[lo], b1 = offsets[hi];
    int rc;
    static const char *skip_env = getenv("ORPHEUM_B200_AB_SKIP");  // development A/B: "kernels", "d2h"
    static const bool skip_kernels = skip_env && strstr(skip_env, "kernels"), skip_d2h = skip_env && strstr(skip_env, "d2h");
    if ((rc = q.in_offsets.ensure((m + 1) * 8)) != ORPH_OK) return rc;
    if ((rc = q.out_results.ensure(m * sizeof(orph_read_result))) != ORPH_OK) return rc;
    ChunkTrace tr;
    auto mark = [&](int which, cudaStream_t st) {
        if (pc.trace) { cudaEventCreate(&tr.e[which]); cudaEventRecord(tr.e[which], st); }
    };
    tr.lane = lane;
    tr.packed = packed_here;
    tr.lo = lo;
    tr.m = m;
    tr.host_pack_s = host_pack_s;
    tr.host_enqueue_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - pc.t_call).count();
    const uint8_t *d_ascii = nullptr;
    const uint64_t *d_packed = nullptr;
    uint64_t pbase0 = 0, n_ascii = 0;
    const uint32_t *d_rel = nullptr;  // 32-bit offset deltas on the device (host-packed chunks of unequal reads)
    // ---- H2D
    if (q.h2d_bytes) {  // rate achieved by the slot's previous chunk
        float ms = 0;
        if (q.h2d_bytes >= (1u << 20) && cudaEventElapsedTime(&ms, q.ev_h2d_begin, q.ev_h2d_end) == cudaSuccess && ms > 0)
            ctx->h2d_bytes_per_s = 0.75 * ctx->h2d_bytes_per_s + 0.25 * (double)q.h2d_bytes / (1e-3 * ms);
        else cudaGetLastError();
    }
    mark(0, s_in);
    CUDA_TRY(cudaEventRecord(q.ev_h2d_begin, s_in));
    if (packed_here) {
        const uint64_t n_words = (b1 - b0 + 31) / 32;
        const uint64_t bytes = n_words * 8 + (fixed_len ? 0 : (m + 1) * 4);
        if ((rc = q.in_reads.ensure(bytes + 16)) != ORPH_OK) return rc;
        CUDA_TRY(cudaMemcpyAsync(q.in_reads.p, q.h_stage.p, bytes, cudaMemcpyHostToDevice, s_in));
        d_packed = (const uint64_t *)q.in_reads.p;
        d_rel = fixed_len ? nullptr : (const uint32_t *)((const uint8_t *)q.in_reads.p + n_words * 8);
        pbase0 = b0;
        q.h2d_bytes = bytes;
    } else if (pc.reads_format == ORPH_READS_ASCII) {
        n_ascii = b1 - b0;
        if ((rc = q.in_reads.ensure(n_ascii + 64)) != ORPH_OK) return rc;
        if (n_ascii) CUDA_TRY(cudaMemcpyAsync(q.in_reads.p, pc.reads + b0, n_ascii, cudaMemcpyHostToDevice, s_in));
        d_ascii = (const uint8_t *)q.in_reads.p;
        q.h2d_bytes = n_ascii;
    } else {
        const uint64_t w0 = b0 / 32, w1 = (b1 + 31) / 32;
        if ((rc = q.in_reads.ensure((w1 - w0 + 2) * 8)) != ORPH_OK) return rc;
        if (w1 > w0) CUDA_TRY(cudaMemcpyAsync(q.in_reads.p, (const uint64_t *)pc.reads + w0, (w1 - w0) * 8, cudaMemcpyHostToDevice, s_in));
        d_packed = (const uint64_t *)q.in_reads.p;
        pbase0 = w0 * 32;
        q.h2d_bytes = (w1 - w0) * 8;
    }
    if (!fixed_len && !d_rel) {
        CUDA_TRY(cudaMemcpyAsync(q.in_offsets.p, offsets + lo, (m + 1) * 8, cudaMemcpyHostToDevice, s_in));
        q.h2d_bytes += (m + 1) * 8;
    }
    mark(1, s_in);
    CUDA_TRY(cudaEventRecord(q.ev_h2d_end, s_in));
    CUDA_TRY(cudaEventRecord(q.ev_in, s_in));
    // ---- kernels and D2H: one chunk at a time on the lane's kernel stream (the scratch of score_device_impl is per context)
    {
        CUDA_TRY(cudaStreamWaitEvent(s, q.ev_in, 0));
        mark(2, s);
        if (fixed_len) {
            fill_offsets_kernel<<<grid_for(m + 1, 256, ctx->sm_count, 4), 256, 0, s>>>(b0, fixed_len, m + 1, (uint64_t *)q.in_offsets.p);
            kctx->launches++;
        } else if (d_rel) {
            expand_offsets_kernel<<<grid_for(m + 1, 256, ctx->sm_count, 4), 256, 0, s>>>(d_rel, b0, m + 1, (uint64_t *)q.in_offsets.p);
            kctx->launches++;
        }
        CUDA_TRY(cudaGetLastError());
        if (!skip_kernels) {
            rc = score_device_impl(kctx, pc.f, d_ascii, b0, d_packed, pbase0, n_ascii, (const uint64_t *)q.in_offsets.p, m, max_len, pc.lut, pc.thr,
                                   (orph_read_result *)q.out_results.p);
            if (rc != ORPH_OK) return rc;
        }
        mark(3, s);
        CUDA_TRY(cudaEventRecord(q.ev_kernels, s));
        CUDA_TRY(cudaStreamWaitEvent(s_out, q.ev_kernels, 0));
        mark(4, s_out);
        if (!skip_d2h) CUDA_TRY(cudaMemcpyAsync(pc.out + lo, q.out_results.p, m * sizeof(orph_read_result), cudaMemcpyDeviceToHost, s_out));
        mark(5, s_out);
        CUDA_TRY(cudaEventRecord(q.ev_out, s_out));
        std::lock_guard<std::mutex> lk(pc.stats_mu);
        if (packed_here) ctx->chunks_packed_on_host++;
        else if (pc.reads_format == ORPH_READS_ASCII) ctx->chunks_sent_ascii++;
        if (pc.trace) pc.traces.push_back(tr);
    }
    q.used = true;
    return ORPH_OK;
}

// Validates a claimed chunk: the longest read (clamped to what the kernels are told), the common length when every
// read has the same one (0 otherwise), monotonic offsets.
static int pipe_scan_chunk(PipeCall &pc, uint64_t lo, uint64_t hi, uint32_t *max_len, uint64_t *fixed_len, orph_host::Pool *pool = nullptr)
{
    uint64_t longest, shortest;
    constexpr uint64_t kBlock = 1u << 15;
    if (pool && hi - lo >= 4 * kBlock) {
        // one core reads offsets at ~10 GB/s, a 2-bit chunk of 2 M reads crosses PCIe in 1.4 ms: the scan is spread over the pool
        const uint64_t n_blk = (hi - lo + kBlock - 1) / kBlock;
        std::vector<uint64_t> part(2 * n_blk);
        pool->run(n_blk, [&](uint64_t b) {
            const uint64_t a = lo + b * kBlock;
            scan_read_lengths(pc.offsets + a, std::min(hi, a + kBlock) - a, &part[2 * b], &part[2 * b + 1]);
        });
        longest = 0;
        shortest = ~0ull;
        for (uint64_t b = 0; b < n_blk; b++) {
            longest = std::max(longest, part[2 * b]);
            shortest = std::min(shortest, part[2 * b + 1]);
        }
    } else {
        scan_read_lengths(pc.offsets + lo, hi - lo, &longest, &shortest);
    }
    if (longest >= (1ULL << 63)) {  // a decreasing pair wraps around
        uint64_t at = lo;
        while (at < hi && pc.offsets[at + 1] >= pc.offsets[at]) at++;
        return fail(ORPH_E_INVALID, "orph_score_reads: offsets are not monotonic at read %llu", (unsigned long long)at);
    }
    *max_len = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(longest, ORPH_MAX_READ_LEN));  // over-long reads are reported through the status
    *fixed_len = longest == shortest && longest > 0 && longest <= ORPH_MAX_READ_LEN ? longest : 0;
    return ORPH_OK;
}

// lane 0: chunks as they are, from the front
static void pipe_front_lane(PipeCall &pc)
{
    orph_ctx *ctx = pc.ctx;
    uint64_t lo, hi;
    int c = 0;
    while (pc.claim(0, &lo, &hi)) {
        PipeSlot &q = ctx->lane_slots[0][c % kFrontSlots];
        if (q.used && wait_event_polling(q.ev_out) != cudaSuccess) {  // the slot's previous chunk has left the device
            pc.set_error(fail(ORPH_E_CUDA, "cudaEventQuery: %s", cudaGetErrorString(cudaGetLastError())));
            return;
        }
        if (pc.share_h2d && c >= 2) {
            // the packing lane's copies queue behind this lane's in the shared H2D stream: at most two ASCII chunks there
            PipeSlot &q2 = ctx->lane_slots[0][(c - 2) % kFrontSlots];
            if (q2.used && wait_event_polling(q2.ev_in) != cudaSuccess) {
                pc.set_error(fail(ORPH_E_CUDA, "cudaEventQuery: %s", cudaGetErrorString(cudaGetLastError())));
                return;
            }
        }
        c++;
        uint32_t max_len = 0;
        uint64_t fixed_len = 0;
        int rc = pipe_scan_chunk(pc, lo, hi, &max_len, &fixed_len, pc.scan_pool);
        if (rc == ORPH_OK) rc = pipe_enqueue_chunk(pc, 0, q, lo, hi, max_len, fixed_len, false, 0.0);
        if (rc != ORPH_OK) {
            pc.set_error(rc);
            return;
        }
    }
}

// lane 1: chunks packed to 2 bits per base on the host pool, from the back
static void pipe_back_lane(PipeCall &pc)
{
    using clk = std::chrono::steady_clock;
    orph_ctx *ctx = pc.ctx;
    cudaSetDevice(ctx->device);
    std::vector<uint8_t> needs;
    uint64_t lo, hi;
    int c = 0, ascii_streak = 0;
    while (pc.claim(1, &lo, &hi)) {
        PipeSlot &q = ctx->lane_slots[1][c++ % kLaneSlots];
        if (q.used && wait_event_polling(q.ev_out) != cudaSuccess) {
            pc.set_error(fail(ORPH_E_CUDA, "cudaEventQuery: %s", cudaGetErrorString(cudaGetLastError())));
            return;
        }
        uint32_t max_len = 0;
        uint64_t fixed_len = 0;
        int rc = pipe_scan_chunk(pc, lo, hi, &max_len, &fixed_len, ctx->pool);  // this lane owns the pool
        if (rc != ORPH_OK) {
            pc.set_error(rc);
            return;
        }
        const uint64_t m = hi - lo, b0 = pc.offsets[lo], b1 = pc.offsets[hi];
        const uint64_t n_words = (b1 - b0 + 31) / 32;
        bool packed_here = ascii_streak == 0 && b1 > b0;
        double pack_s = 0;
        if (ascii_streak > 0) ascii_streak--;
        if (packed_here && q.h_stage.ensure(n_words * 8 + (m + 1) * 4 + 64) != ORPH_OK) {
            cudaGetLastError();  // no pinned memory to be had: the chunk goes out as it is
            packed_here = false;
        }
        if (packed_here) {
            needs.assign(m, 0);
            const auto t0 = clk::now();
            orph_host::pack_stream(ctx->pool, pc.reads + b0, b0, b1 - b0, pc.offsets + lo, m, (uint64_t *)q.h_stage.p, needs.data());
            if (!fixed_len) {  // 32-bit offset deltas behind the words: the chunk is one copy
                uint32_t *rel = (uint32_t *)((uint8_t *)q.h_stage.p + n_words * 8);
                const uint64_t *off = pc.offsets + lo;
                for (uint64_t i = 0; i <= m; i++) rel[i] = (uint32_t)(off[i] - b0);
            }
            pack_s = std::chrono::duration<double>(clk::now() - t0).count();
            const double per_byte = pack_s / (double)(b1 - b0);
            ctx->pack_s_per_byte = ctx->pack_s_per_byte == 0 ? per_byte : 0.75 * ctx->pack_s_per_byte + 0.25 * per_byte;
            uint64_t n_flagged = 0;
            for (uint64_t i = 0; i < m; i++) n_flagged += needs[i];
            if (n_flagged * 8 > m) {  // mostly unpackable input (soft-masked genome, ...): ASCII is the cheaper route
                packed_here = false;
                ascii_streak = 8;
            } else if (n_flagged && pc.exceptions) {
                std::lock_guard<std::mutex> lk(pc.claim_mu);
                for (uint64_t i = 0; i < m; i++)
                    if (needs[i]) pc.exceptions->push_back(lo + i);
            }
        }
        rc = pipe_enqueue_chunk(pc, 1, q, lo, hi, max_len, fixed_len, packed_here, pack_s);
        if (rc != ORPH_OK) {
            pc.set_error(rc);
            return;
        }
    }
}

static int score_host_pipeline(orph_ctx *ctx, const orph_filter *f, const void *reads, int reads_format, const uint64_t *offsets,
                               uint64_t n_reads, const uint8_t lut[256], double jaccard_threshold, orph_read_result *out,
                               bool allow_pack, std::vector<uint64_t> *exceptions)
{
    using clk = std::chrono::steady_clock;
    int rc;
    cudaStream_t s = ctx->stream;
    if ((rc = pipe_setup(ctx)) != ORPH_OK) return rc;
    // every early return below may leave copies / kernels of earlier chunks in flight on the streams, reading the
    // caller's host buffers: drain the device before handing control back
    struct DrainOnError {
        bool armed = true;
        ~DrainOnError()
        {
            if (armed) cudaDeviceSynchronize();
        }
    } drain;
    const bool want_pool = allow_pack || (ctx->host_threads != 0 && n_reads >= (1u << 20));
    if (want_pool && !ctx->pool) {
        int n = ctx->host_threads;
        if (n <= 0) {  // auto: the cores this process may run on, shared between the ranks of one node
            cpu_set_t set;
            n = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
            const char *lws = getenv("LOCAL_WORLD_SIZE");
            const int ranks = lws ? atoi(lws) : 1;
            if (ranks > 1) n /= ranks;
            n = std::max(1, std::min(n, 64));
        }
        ctx->pool = new (std::nothrow) orph_host::Pool(n);
        if (!ctx->pool) return fail(ORPH_E_NOMEM, "out of host memory");
    }
    const bool input_pinned = input_is_pinned(reads);
    if (allow_pack && !ctx->lane_ctx) {
        if ((rc = orph_ctx_create(ctx->device, nullptr, &ctx->lane_ctx)) != ORPH_OK) return rc;
    }
    if (ctx->lane_ctx) {
        ctx->lane_ctx->force_exact = ctx->force_exact;
        ctx->lane_ctx->disable_task_kernel = ctx->disable_task_kernel;
    }
    PipeCall pc;
    pc.ctx = ctx;
    pc.scan_pool = allow_pack ? nullptr : ctx->pool;
    pc.f = f;
    pc.reads = (const uint8_t *)reads;
    pc.reads_format = reads_format;
    pc.offsets = offsets;
    pc.n_reads = n_reads;
    pc.lut = lut;
    pc.thr = jaccard_threshold;
    pc.out = out;
    pc.exceptions = exceptions;
    pc.front = 0;
    pc.back = n_reads;
    pc.trace = getenv("ORPHEUM_B200_TRACE") != nullptr;  // per-chunk timeline on stderr (development aid)
    pc.t_call = clk::now();
    // chunk geometry: ASCII chunks of 2^19 reads / 2^27 bases are ~80 MB copies; 2-bit chunks carry four times the reads
    static const int chunk_log2 = [] { const char *v = getenv("ORPHEUM_B200_CHUNK_LOG2"); const int n = v ? atoi(v) : 0; return n < 12 ? 0 : n > 24 ? 24 : n; }();
    const int log2_plain = chunk_log2 ? chunk_log2 : (reads_format == ORPH_READS_ASCII ? 19 : 21);
    pc.max_chunk_reads[0] = 1ULL << log2_plain;
    pc.max_chunk_bases[0] = 1ULL << (log2_plain + 8);
    pc.max_chunk_reads[1] = 1ULL << 19;
    pc.max_chunk_bases[1] = 1ULL << 27;
    pc.min_chunk_reads = reads_format == ORPH_READS_ASCII ? 1u << 16 : 1u << 17;  // a chunk costs ~50 us of launches and copy set-up
    // a pageable ASCII buffer cannot be shipped asynchronously: the packing lane takes all of it (the CPU reads it faster
    // than a staged copy would)
    const bool lane0 = !(allow_pack && !input_pinned);
    const bool lane1 = allow_pack;
    pc.share_h2d = lane0 && lane1;
    std::thread packer;
    if (lane1) packer = std::thread(pipe_back_lane, std::ref(pc));
    if (lane0) pipe_front_lane(pc);
    if (packer.joinable()) packer.join();
    if (pc.error != ORPH_OK) {
        cudaDeviceSynchronize();
        drain.armed = false;
        for (auto &lane : ctx->lane_slots)
            for (PipeSlot &q : lane) q.used = false;
        return fail(pc.error, "%s", pc.error_text);
    }
    ScoreCounters counters, counters1;
    counters1.status = 0;
    CUDA_TRY(cudaMemcpyAsync(&counters, ctx->counters.p, sizeof(counters), cudaMemcpyDeviceToHost, s));
    if (lane1) {
        cudaStream_t s1 = ctx->lane_ctx->stream;
        CUDA_TRY(cudaMemcpyAsync(&counters1, ctx->lane_ctx->counters.p, sizeof(counters1), cudaMemcpyDeviceToHost, s1));
        CUDA_TRY(cudaMemsetAsync((uint8_t *)ctx->lane_ctx->counters.p + offsetof(ScoreCounters, status), 0, sizeof(uint32_t), s1));
        CUDA_TRY(cudaStreamSynchronize(s1));
        ctx->launches += ctx->lane_ctx->launches;
        ctx->lane_ctx->launches = 0;
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    for (cudaStream_t st : {ctx->s_out[0], ctx->s_out[1], ctx->s_in[0]}) CUDA_TRY(cudaStreamSynchronize(st));
    counters.status |= counters1.status;
    for (auto &lane : ctx->lane_slots)
        for (PipeSlot &q : lane) q.used = false;  // everything has left the device: the next call starts with free slots
    if (pc.trace && !pc.traces.empty()) {
        fprintf(stderr, "[orph trace] call of %llu reads, %.3f ms on the host clock\n", (unsigned long long)n_reads,
                1e3 * std::chrono::duration<double>(clk::now() - pc.t_call).count());
        fprintf(stderr, "[orph trace] lane first_read reads mode host_enq host_pack | h2d_start h2d_end | k_start k_end | d2h_start d2h_end (ms, device times relative to the first chunk's h2d_start)\n");
        for (size_t i = 0; i < pc.traces.size(); i++) {
            const ChunkTrace &ct = pc.traces[i];
            float t[6];
            for (int k = 0; k < 6; k++) cudaEventElapsedTime(&t[k], pc.traces[0].e[0], ct.e[k]);
            fprintf(stderr, "[orph trace] %d %9llu %8llu %s %8.3f %8.3f | %8.3f %8.3f | %8.3f %8.3f | %8.3f %8.3f\n", ct.lane, (unsigned long long)ct.lo,
                    (unsigned long long)ct.m, ct.packed ? "pack " : "plain", 1e3 * ct.host_enqueue_s, 1e3 * ct.host_pack_s, t[0], t[1], t[2], t[3], t[4], t[5]);
        }
        for (auto &ct : pc.traces)
            for (int k = 0; k < 6; k++) cudaEventDestroy(ct.e[k]);
    }
    CUDA_TRY(cudaMemsetAsync((uint8_t *)ctx->counters.p + offsetof(ScoreCounters, status), 0, sizeof(uint32_t), s));
    drain.armed = false;  // all streams were synchronised above
    return status_to_code(counters.status);
}

extern "C" int orph_score_reads(orph_ctx *ctx, const orph_filter *f, const void *reads, int reads_format,
                                const uint64_t *offsets, uint64_t n_reads, const uint8_t lut[256], double jaccard_threshold,
                                orph_read_result *out)
{
    int rc = check_score_args(ctx, f, reads, reads_format, offsets, n_reads, lut, jaccard_threshold, out);
    if (rc != ORPH_OK) return rc;
    if (n_reads == 0) return ORPH_OK;
    DeviceGuard guard(ctx->device);
    // host packing pays once a batch is large enough to keep the pool busy; small batches go out as they are
    bool allow_pack = reads_format == ORPH_READS_ASCII && ctx->host_threads != 0 && offsets[n_reads] >= offsets[0] &&
                      offsets[n_reads] - offsets[0] >= (1ULL << 22);
    // A host-packed chunk costs the host's memory 1.25-1.5 bytes per base (read, write 2 bits, DMA reads them) against 1 for a
    // chunk the copy engine takes as it is.  That is a loss only when the host's memory system is what everything waits for:
    // many ranks sharing one host -- measured with eight ranks of four threads each: links at 20 GB/s instead of 50, the pool
    // packing 12 GB/s next to them, the call 5 % faster without the packing lane.  (With the same four threads and the link at
    // full rate the lane is worth 25 %.)  Then the lane stays off, and is tried again every 16th call.
    if (allow_pack && input_is_pinned(reads) && ctx->pack_s_per_byte > 0 && ctx->h2d_bytes_per_s < 25e9 &&
        1.0 / ctx->pack_s_per_byte < 0.75 * ctx->h2d_bytes_per_s && (++ctx->pack_lane_skips & 15) != 0)
        allow_pack = false;
    std::vector<uint64_t> exceptions;
    rc = score_host_pipeline(ctx, f, reads, reads_format, offsets, n_reads, lut, jaccard_threshold, out, allow_pack, &exceptions);
    if (rc != ORPH_OK && rc != ORPH_E_EMPTY_READ && rc != ORPH_E_TOO_LONG) return rc;
    if (!exceptions.empty()) {
        // second pass over the reads the 2-bit stream cannot represent: their bytes, gathered into one small ASCII batch
        std::sort(exceptions.begin(), exceptions.end());
        const uint64_t n_exc = exceptions.size();
        std::vector<uint64_t> eoff(n_exc + 1);
        eoff[0] = 0;
        for (uint64_t i = 0; i < n_exc; i++) eoff[i + 1] = eoff[i] + (offsets[exceptions[i] + 1] - offsets[exceptions[i]]);
        std::vector<uint8_t> ebytes(eoff[n_exc] + 64);
        for (uint64_t i = 0; i < n_exc; i++)
            memcpy(ebytes.data() + eoff[i], (const uint8_t *)reads + offsets[exceptions[i]], eoff[i + 1] - eoff[i]);
        std::vector<orph_read_result> eout(n_exc);
        const int rc2 = score_host_pipeline(ctx, f, ebytes.data(), ORPH_READS_ASCII, eoff.data(), n_exc, lut, jaccard_threshold, eout.data(), false, nullptr);
        if (rc2 != ORPH_OK && rc2 != ORPH_E_EMPTY_READ && rc2 != ORPH_E_TOO_LONG) return rc2;
        for (uint64_t i = 0; i < n_exc; i++) out[exceptions[i]] = eout[i];
        if (rc == ORPH_OK) rc = rc2;
    }
    if (g_guard) {
        const int rg = verify_guards(ctx);
        if (rg != ORPH_OK) return rg;
    }
    return rc;
}

extern "C" int orph_ctx_set_host_threads(orph_ctx *ctx, int n_threads)
{
    if (!ctx) return fail(ORPH_E_INVALID, "ctx is NULL");
    if (ctx->pool && n_threads != ctx->host_threads) {  // re-created at the new size on the next large batch
        delete ctx->pool;
        ctx->pool = nullptr;
    }
    ctx->host_threads = n_threads;
    ctx->pack_s_per_byte = 0;
    return ORPH_OK;
}

extern "C" int orph_ctx_ingest_stats(orph_ctx *ctx, uint64_t *chunks_packed_on_host, uint64_t *chunks_sent_ascii, int *host_threads,
                                     double *pack_bytes_per_s, double *h2d_bytes_per_s)
{
    if (!ctx) return fail(ORPH_E_INVALID, "ctx is NULL");
    if (chunks_packed_on_host) *chunks_packed_on_host = ctx->chunks_packed_on_host;
    if (chunks_sent_ascii) *chunks_sent_ascii = ctx->chunks_sent_ascii;
    if (host_threads) *host_threads = ctx->pool ? ctx->pool->size() : ctx->host_threads;
    if (pack_bytes_per_s) *pack_bytes_per_s = ctx->pack_s_per_byte > 0 ? 1.0 / ctx->pack_s_per_byte : 0.0;
    if (h2d_bytes_per_s) *h2d_bytes_per_s = ctx->h2d_bytes_per_s;
    return ORPH_OK;
}
