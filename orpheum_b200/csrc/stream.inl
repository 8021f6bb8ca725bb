// orph_stream: sequence file in, scored records out (included by orpheum_b200.cu, same translation unit).
//
// One background thread per stream runs the whole path: it asks orph_ingest for the next batch (parallel parse + 2-bit
// pack into pinned memory, fused_ingest.cpp), enqueues that batch's copies and kernels on the stream's private context,
// and only then waits for the batch before it -- so the host parses batch k + 1 while the GPU scores batch k.  Reads
// the 2-bit stream cannot represent (N, lower case, ...) are gathered from the text and scored from their bytes in the
// same enqueue; their records replace the ones the packed pass produced.  The reference's counterpart is the serial
// loop over screed records in Translate.score_reads_per_file (orpheum/translate.py:367-372).
#pragma once

#include <condition_variable>
#include <deque>
#include <mutex>
#include <string>

#include "ingest_internal.h"

namespace {

void *pinned_alloc(size_t n)
{
    size_t got = 0;
    return g_pinned.get(n ? n : 1, &got);
}
void pinned_release(void *p) { g_pinned.put(p); }

constexpr int kStreamSlots = 4;  // being parsed + on the GPU + finished + held by the caller

// Everything of a stream that is expensive to set up and independent of the file: the private context (CUDA stream,
// kernel scratch, L2 window), the device staging and the pinned result buffers.  Closed streams hand theirs to a
// process-wide free list (two per device are kept), so opening a stream per file costs no allocation after the first.
struct StreamRes {
    int device = 0;
    orph_ctx *ctx = nullptr;  // private: own CUDA stream and scratch, so the caller's context stays free
    struct Job {
        IngestSlot *slot = nullptr;
        PinnedBuf results, side_ascii, side_offsets, side_results, status_word;
        cudaEvent_t done = nullptr;
        uint64_t n = 0, n_side = 0;
        int status = ORPH_OK;
    } jobs[kStreamSlots];
    DevBuf d_packed, d_off32, d_off64, d_results, d_side_ascii, d_side_offsets, d_side_results;
};

std::mutex g_stream_res_mu;
std::vector<StreamRes *> g_stream_res_free;

void stream_res_destroy(StreamRes *r)
{
    if (!r) return;
    DeviceGuard guard(r->device);
    if (r->ctx) cudaStreamSynchronize(r->ctx->stream);
    for (DevBuf *b : {&r->d_packed, &r->d_off32, &r->d_off64, &r->d_results, &r->d_side_ascii, &r->d_side_offsets, &r->d_side_results}) b->release();
    for (auto &job : r->jobs) {
        for (PinnedBuf *b : {&job.results, &job.side_ascii, &job.side_offsets, &job.side_results, &job.status_word}) b->release();
        if (job.done) cudaEventDestroy(job.done);
    }
    if (r->ctx) orph_ctx_destroy(r->ctx);
    delete r;
}

int stream_res_acquire(int device, StreamRes **out)
{
    {
        std::lock_guard<std::mutex> lk(g_stream_res_mu);
        for (size_t i = 0; i < g_stream_res_free.size(); i++)
            if (g_stream_res_free[i]->device == device) {
                *out = g_stream_res_free[i];
                g_stream_res_free.erase(g_stream_res_free.begin() + (long)i);
                return ORPH_OK;
            }
    }
    StreamRes *r = new (std::nothrow) StreamRes();
    if (!r) return fail(ORPH_E_NOMEM, "out of host memory");
    r->device = device;
    int rc = orph_ctx_create(device, nullptr, &r->ctx);
    if (rc == ORPH_OK)
        for (auto &job : r->jobs)
            if (cudaEventCreateWithFlags(&job.done, cudaEventDisableTiming | cudaEventBlockingSync) != cudaSuccess) rc = fail(ORPH_E_CUDA, "cudaEventCreate failed");
    if (rc != ORPH_OK) {
        stream_res_destroy(r);
        return rc;
    }
    *out = r;
    return ORPH_OK;
}

void stream_res_release(StreamRes *r)
{
    if (!r) return;
    {
        DeviceGuard guard(r->device);
        cudaStreamSynchronize(r->ctx->stream);
    }
    for (auto &job : r->jobs) job.slot = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_stream_res_mu);
        int same = 0;
        for (StreamRes *o : g_stream_res_free) same += o->device == r->device;
        if (same < 2) {
            g_stream_res_free.push_back(r);
            return;
        }
    }
    stream_res_destroy(r);
}

}  // namespace

struct orph_stream {
    // one entry per device lane (the same device may appear twice: two private contexts); batch k runs on lane k % n
    std::vector<StreamRes *> rs;
    std::vector<const orph_filter *> fs;
    orph_ingest *ing = nullptr;
    uint8_t lut[256];
    double thr = 0;
    uint64_t batch_reads = 0;
    std::thread worker;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<uint64_t> finished;  // batch numbers, in file order
    int outstanding = 0;  // batches handed to the parser and not yet released by the caller
    int max_outstanding = kStreamSlots;
    bool holding = false;
    uint64_t held = 0;
    bool eof = false, stop = false;
    int error = ORPH_OK;
    std::string error_text;
    // orph_stream_emit: its own small pool (the parser's pool is busy with the next batch) and output buffers
    orph_host::Pool *emit_pool = nullptr;
    std::vector<uint8_t> emit_out[4], emit_reads, emit_peptides;
    std::vector<uint64_t> emit_read_off, emit_item_read, emit_pep_off;
    std::vector<int8_t> emit_item_frame;
    size_t lanes() const { return rs.size(); }
    StreamRes *res_of(uint64_t k) const { return rs[k % rs.size()]; }
    StreamRes::Job &job_of(uint64_t k) const { return rs[k % rs.size()]->jobs[(k / rs.size()) % kStreamSlots]; }
};

static int stream_enqueue(orph_stream *st, uint64_t k)
{
    StreamRes *r = st->res_of(k);
    StreamRes::Job &job = st->job_of(k);
    const orph_filter *f = st->fs[k % st->lanes()];
    orph_ctx *ctx = r->ctx;
    DeviceGuard guard(r->device);
    cudaStream_t s = ctx->stream;
    const orph_host::FusedBatch &b = job.slot->b;
    const uint64_t n = b.n_reads, n_words = (b.total_bases + 31) / 32;
    int rc;
    job.n = n;
    job.n_side = b.flagged.size();
    if ((rc = r->d_packed.ensure((n_words + 2) * 8)) != ORPH_OK) return rc;
    if ((rc = r->d_off32.ensure((n + 1) * 4)) != ORPH_OK) return rc;
    if ((rc = r->d_off64.ensure((n + 1) * 8)) != ORPH_OK) return rc;
    if ((rc = r->d_results.ensure(n * sizeof(orph_read_result))) != ORPH_OK) return rc;
    if ((rc = job.results.ensure(n * sizeof(orph_read_result))) != ORPH_OK) return rc;
    if ((rc = job.status_word.ensure(sizeof(ScoreCounters))) != ORPH_OK) return rc;
    if (n_words) CUDA_TRY(cudaMemcpyAsync(r->d_packed.p, b.packed, n_words * 8, cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(r->d_off32.p, b.off32, (n + 1) * 4, cudaMemcpyHostToDevice, s));
    expand_offsets_kernel<<<grid_for(n + 1, 256, ctx->sm_count, 8), 256, 0, s>>>((const uint32_t *)r->d_off32.p, 0, n + 1, (uint64_t *)r->d_off64.p);
    ctx->launches++;
    CUDA_TRY(cudaGetLastError());
    const uint32_t max_len = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(b.max_len, ORPH_MAX_READ_LEN));
    rc = score_device_impl(ctx, f, nullptr, 0, (const uint64_t *)r->d_packed.p, 0, 0, (const uint64_t *)r->d_off64.p, n, max_len, st->lut, st->thr,
                           (orph_read_result *)r->d_results.p);
    if (rc != ORPH_OK) return rc;
    CUDA_TRY(cudaMemcpyAsync(job.results.p, r->d_results.p, n * sizeof(orph_read_result), cudaMemcpyDeviceToHost, s));
    if (job.n_side) {
        // the reads the 2-bit stream cannot hold: their bytes, gathered from the text, through the byte-exact kernels
        const uint64_t m = job.n_side;
        if ((rc = job.side_offsets.ensure((m + 1) * 8)) != ORPH_OK) return rc;
        uint64_t *eoff = (uint64_t *)job.side_offsets.p;
        eoff[0] = 0;
        uint32_t side_max = 1;
        for (uint64_t i = 0; i < m; i++) {
            const uint32_t len = b.seq_len[b.flagged[i]];
            eoff[i + 1] = eoff[i] + len;
            side_max = std::max(side_max, len);
        }
        const uint64_t total = eoff[m];
        if ((rc = job.side_ascii.ensure(total + 64)) != ORPH_OK) return rc;
        if ((rc = job.side_results.ensure(m * sizeof(orph_read_result))) != ORPH_OK) return rc;
        uint8_t *bytes = (uint8_t *)job.side_ascii.p;
        const uint8_t *text = job.slot->text;
        auto gather = [&](uint64_t blk) {
            const uint64_t i1 = std::min(m, (blk + 1) * 4096);
            for (uint64_t i = blk * 4096; i < i1; i++) memcpy(bytes + eoff[i], text + b.seq_start[b.flagged[i]], b.seq_len[b.flagged[i]]);
        };
        const uint64_t n_blk = (m + 4095) / 4096;
        if (st->ing->pool && n_blk > 4) st->ing->pool->run(n_blk, gather);
        else
            for (uint64_t blk = 0; blk < n_blk; blk++) gather(blk);
        if ((rc = r->d_side_ascii.ensure(total + 64)) != ORPH_OK) return rc;
        if ((rc = r->d_side_offsets.ensure((m + 1) * 8)) != ORPH_OK) return rc;
        if ((rc = r->d_side_results.ensure(m * sizeof(orph_read_result))) != ORPH_OK) return rc;
        if (total) CUDA_TRY(cudaMemcpyAsync(r->d_side_ascii.p, bytes, total, cudaMemcpyHostToDevice, s));
        CUDA_TRY(cudaMemcpyAsync(r->d_side_offsets.p, eoff, (m + 1) * 8, cudaMemcpyHostToDevice, s));
        for (uint64_t lo = 0; lo < m; lo += 1ull << 26) {
            const uint64_t cnt = std::min<uint64_t>(1ull << 26, m - lo);
            rc = score_device_impl(ctx, f, (const uint8_t *)r->d_side_ascii.p + eoff[lo], eoff[lo], nullptr, 0, eoff[lo + cnt] - eoff[lo],
                                   (const uint64_t *)r->d_side_offsets.p + lo, cnt, std::min<uint32_t>(side_max, ORPH_MAX_READ_LEN), st->lut, st->thr,
                                   (orph_read_result *)r->d_side_results.p + lo);
            if (rc != ORPH_OK) return rc;
        }
        CUDA_TRY(cudaMemcpyAsync(job.side_results.p, r->d_side_results.p, m * sizeof(orph_read_result), cudaMemcpyDeviceToHost, s));
    }
    // the sticky status of this batch (empty / over-long reads), then clear it for the next one
    CUDA_TRY(cudaMemcpyAsync(job.status_word.p, ctx->counters.p, sizeof(ScoreCounters), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaMemsetAsync((uint8_t *)ctx->counters.p + offsetof(ScoreCounters, status), 0, sizeof(uint32_t), s));
    CUDA_TRY(cudaEventRecord(job.done, s));
    return ORPH_OK;
}

static int stream_complete(orph_stream *st, uint64_t k)
{
    StreamRes::Job &job = st->job_of(k);
    DeviceGuard guard(st->res_of(k)->device);
    CUDA_TRY(cudaEventSynchronize(job.done));
    if (job.n_side) {
        orph_read_result *res = (orph_read_result *)job.results.p;
        const orph_read_result *side = (const orph_read_result *)job.side_results.p;
        const std::vector<uint32_t> &flagged = job.slot->b.flagged;
        for (uint64_t i = 0; i < job.n_side; i++) res[flagged[i]] = side[i];
    }
    const uint32_t bits = ((const ScoreCounters *)job.status_word.p)->status;
    job.status = bits & kStatusEmptyRead ? ORPH_E_EMPTY_READ : bits & kStatusTooLong ? ORPH_E_TOO_LONG : ORPH_OK;
    return ORPH_OK;
}

static void stream_worker(orph_stream *st)
{
    const uint64_t lanes = st->lanes();
    uint64_t k = 0, completed = 0;  // batches [completed, k) are on the GPUs
    static const bool trace = getenv("ORPHEUM_B200_TRACE") != nullptr;
    auto fail_with = [&](int rc) {
        for (StreamRes *r : st->rs) {
            DeviceGuard guard(r->device);
            cudaStreamSynchronize(r->ctx->stream);
        }
        std::lock_guard<std::mutex> lk(st->mu);
        st->error = rc;
        st->error_text = g_err;  // thread-local message of this worker thread
        st->eof = true;
        st->cv.notify_all();
    };
    auto complete_next = [&]() -> int {
        const int rc = stream_complete(st, completed);
        if (rc != ORPH_OK) return rc;
        {
            std::lock_guard<std::mutex> lk(st->mu);
            st->finished.push_back(completed);
            st->cv.notify_all();
        }
        completed++;
        return ORPH_OK;
    };
    for (;;) {
        {
            std::unique_lock<std::mutex> lk(st->mu);
            st->cv.wait(lk, [&] { return st->stop || st->outstanding < st->max_outstanding; });
            if (st->stop) break;
            st->outstanding++;
        }
        StreamRes::Job &job = st->job_of(k);
        const auto t0 = std::chrono::steady_clock::now();
        int rc = orph_ingest_next_slot(st->ing, st->batch_reads, &job.slot);
        const auto t1 = std::chrono::steady_clock::now();
        if (rc == ORPH_OK && job.slot->b.n_reads == 0) {  // end of file: drain what is on the GPUs, in order
            while (rc == ORPH_OK && completed < k) rc = complete_next();
            if (rc != ORPH_OK) { fail_with(rc); return; }
            std::lock_guard<std::mutex> lk(st->mu);
            st->outstanding--;
            st->eof = true;
            st->cv.notify_all();
            return;
        }
        if (rc == ORPH_OK) rc = stream_enqueue(st, k);
        if (rc != ORPH_OK) { fail_with(rc); return; }
        const auto t2 = std::chrono::steady_clock::now();
        k++;
        // keep one batch per lane on the GPUs: the parser works on the next one while they run
        while (k - completed > lanes) {
            rc = complete_next();
            if (rc != ORPH_OK) { fail_with(rc); return; }
        }
        if (trace) {
            const auto t3 = std::chrono::steady_clock::now();
            auto ms = [](auto a, auto b) { return 1e3 * std::chrono::duration<double>(b - a).count(); };
            fprintf(stderr, "[orph trace] stream batch %llu (lane %llu): %llu reads (%llu byte-path), parse+pack %.2f ms, enqueue %.2f ms, wait for the batch before %.2f ms\n",
                    (unsigned long long)(k - 1), (unsigned long long)((k - 1) % lanes), (unsigned long long)job.n, (unsigned long long)job.n_side, ms(t0, t1), ms(t1, t2), ms(t2, t3));
        }
    }
    for (StreamRes *r : st->rs) {
        DeviceGuard guard(r->device);
        cudaStreamSynchronize(r->ctx->stream);
    }
}

extern "C" int orph_stream_open_multi(orph_ctx *const *ctxs, const orph_filter *const *filters, int n_lanes, const char *path, const uint8_t lut[256],
                                      double jaccard_threshold, uint64_t batch_reads, int n_threads, orph_stream **out)
{
    if (!ctxs || !filters || !path || !lut || !out || n_lanes < 1 || n_lanes > 64) return fail(ORPH_E_INVALID, "orph_stream_open: bad argument");
    *out = nullptr;
    for (int i = 0; i < n_lanes; i++) {
        if (!ctxs[i] || !filters[i]) return fail(ORPH_E_INVALID, "orph_stream_open: NULL context / filter");
        if (filters[i]->device != ctxs[i]->device)
            return fail(ORPH_E_INVALID, "orph_stream_open: filter lives on device %d, context on %d", filters[i]->device, ctxs[i]->device);
        if (filters[i]->ksize != filters[0]->ksize || filters[i]->n_tables != filters[0]->n_tables || filters[i]->n_bytes != filters[0]->n_bytes)
            return fail(ORPH_E_INVALID, "orph_stream_open: the filters of the lanes are not replicas of one another");
    }
    if (std::isnan(jaccard_threshold)) return fail(ORPH_E_INVALID, "orph_stream_open: jaccard_threshold is NaN");
    if (n_threads <= 0) {  // the cores this process may run on, shared between the ranks of one node
        cpu_set_t set;
        n_threads = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
        const char *lws = getenv("LOCAL_WORLD_SIZE");
        const int ranks = lws ? atoi(lws) : 1;
        if (ranks > 1) n_threads /= ranks;
        n_threads = std::max(1, std::min(n_threads, 64));
    }
    orph_ingest *ing = nullptr;
    // slots of the parser ring: one per lane on the GPUs + being parsed + finished + held by the caller
    int rc;
    {
        DeviceGuard guard(ctxs[0]->device);  // pinned allocations
        rc = orph_ingest_open_with(path, n_threads, n_lanes + 3, IngestAlloc{pinned_alloc, pinned_release}, &ing);
    }
    if (rc != ORPH_OK) return rc;
    orph_stream *st = new (std::nothrow) orph_stream();
    if (!st) {
        orph_ingest_close(ing);
        return fail(ORPH_E_NOMEM, "out of host memory");
    }
    st->ing = ing;
    st->max_outstanding = n_lanes + 3;
    memcpy(st->lut, lut, 256);
    st->thr = jaccard_threshold;
    st->batch_reads = batch_reads ? batch_reads : (1u << 20);
    for (int i = 0; i < n_lanes && rc == ORPH_OK; i++) {
        StreamRes *r = nullptr;
        DeviceGuard guard(ctxs[i]->device);
        rc = stream_res_acquire(ctxs[i]->device, &r);
        if (rc != ORPH_OK) break;
        r->ctx->force_exact = ctxs[i]->force_exact;
        r->ctx->disable_task_kernel = ctxs[i]->disable_task_kernel;
        st->rs.push_back(r);
        st->fs.push_back(filters[i]);
    }
    if (rc != ORPH_OK) {
        orph_stream_close(st);
        return rc;
    }
    st->worker = std::thread(stream_worker, st);
    *out = st;
    return ORPH_OK;
}

extern "C" int orph_stream_open(orph_ctx *ctx, const orph_filter *f, const char *path, const uint8_t lut[256], double jaccard_threshold,
                                uint64_t batch_reads, int n_threads, orph_stream **out)
{
    if (!ctx || !f) return fail(ORPH_E_INVALID, "orph_stream_open: NULL argument");
    return orph_stream_open_multi(&ctx, &f, 1, path, lut, jaccard_threshold, batch_reads, n_threads, out);
}

extern "C" int orph_stream_next(orph_stream *st, orph_stream_batch *out)
{
    if (!st || !out) return fail(ORPH_E_INVALID, "orph_stream_next: NULL argument");
    memset(out, 0, sizeof(*out));
    std::unique_lock<std::mutex> lk(st->mu);
    if (st->holding) {  // the caller is done with the previous batch: its slot goes back to the parser
        st->holding = false;
        st->outstanding--;
        st->cv.notify_all();
    }
    st->cv.wait(lk, [&] { return !st->finished.empty() || st->eof; });
    if (st->finished.empty()) {
        if (st->error != ORPH_OK) return fail(st->error, "%s", st->error_text.c_str());
        return ORPH_OK;  // end of file: n_reads == 0
    }
    const uint64_t k = st->finished.front();
    st->finished.pop_front();
    st->holding = true;
    st->held = k;
    lk.unlock();
    const StreamRes::Job &job = st->job_of(k);
    const orph_host::FusedBatch &b = job.slot->b;
    out->n_reads = job.n;
    out->results = (const orph_read_result *)job.results.p;
    out->names = b.names;
    out->name_offsets = b.name_off;
    out->text = job.slot->text;
    out->seq_start = b.seq_start;
    out->seq_len = b.seq_len;
    out->n_byte_path = job.n_side;
    out->status = job.status;
    return ORPH_OK;
}

// The FASTA records Translate writes for the batch the caller holds (translate.py:244-298), formatted in the library: the
// frames that emit something are found in the records, the reads that carry them are gathered from the text, their
// peptides come from ONE orph_translate_frames call on `ctx`, and the text of every requested kind (see
// orph_format_fasta) is formatted block-parallel.  Nothing here runs per read in the caller's language.
extern "C" int orph_stream_emit(orph_stream *st, orph_ctx *ctx, uint32_t kinds_mask, const uint8_t *out[4], uint64_t out_len[4])
{
    if (!st || !ctx || !out || !out_len) return fail(ORPH_E_INVALID, "orph_stream_emit: NULL argument");
    for (int k = 0; k < 4; k++) {
        out[k] = nullptr;
        out_len[k] = 0;
    }
    if (!st->holding) return fail(ORPH_E_INVALID, "orph_stream_emit: no batch is held (call orph_stream_next first)");
    const StreamRes::Job &job = st->job_of(st->held);
    const orph_host::FusedBatch &b = job.slot->b;
    const uint64_t n = job.n;
    const orph_read_result *res = (const orph_read_result *)job.results.p;
    const bool with_lcp = (kinds_mask & 8u) != 0;
    if (n == 0 || (kinds_mask & 15u) == 0) return ORPH_OK;
    if (!st->emit_pool) {
        cpu_set_t set;
        int cores = sched_getaffinity(0, sizeof(set), &set) == 0 ? CPU_COUNT(&set) : (int)std::thread::hardware_concurrency();
        const char *lws = getenv("LOCAL_WORLD_SIZE");
        if (lws && atoi(lws) > 1) cores /= atoi(lws);
        st->emit_pool = new (std::nothrow) orph_host::Pool(std::max(1, std::min(cores / 2, 16)));
        if (!st->emit_pool) return fail(ORPH_E_NOMEM, "out of host memory");
    }
    orph_host::Pool *pool = st->emit_pool;
    constexpr uint64_t kBlock = 4096;
    const uint64_t n_blk = (n + kBlock - 1) / kBlock;
    auto has_pep = [&](unsigned cat) { return cat == ORPH_CAT_CODING || (with_lcp && cat == ORPH_CAT_LOW_COMPLEXITY_PEPTIDE); };
    // ---- pass A: per block, the peptide-bearing frames, the reads that carry one, their bases and residues
    struct BlockCount { uint64_t items, reads, bases, residues, out_bound[4]; };
    std::vector<BlockCount> cnt(n_blk + 1);
    pool->run(n_blk, [&](uint64_t blk) {
        BlockCount c{};
        const uint64_t r1 = std::min(n, (blk + 1) * kBlock);
        for (uint64_t r = blk * kBlock; r < r1; r++) {
            const uint64_t name_len = b.name_off[r + 1] - b.name_off[r];
            const uint32_t L = b.seq_len[r];
            uint32_t here = 0;
            for (int f = 0; f < 6; f++) {
                const unsigned cat = res[r].category[f];
                const uint32_t phase = f % 3;
                const uint64_t plen = L >= phase ? (L - phase) / 3 : 0;
                if (has_pep(cat)) {
                    here++;
                    c.residues += plen;
                }
                const uint64_t head = 2 * name_len + 96;  // '>' + name (spaces doubled) + frame / jaccard decoration + newlines
                if (cat == ORPH_CAT_CODING) {
                    c.out_bound[0] += head + plen;
                    c.out_bound[1] += head + L;
                } else if (cat == ORPH_CAT_NON_CODING) {
                    c.out_bound[2] += head + L;
                } else if (cat == ORPH_CAT_LOW_COMPLEXITY_PEPTIDE) {
                    c.out_bound[3] += head + plen;
                }
            }
            if (here) {
                c.items += here;
                c.reads++;
                c.bases += L;
            }
        }
        cnt[blk] = c;
    });
    std::vector<BlockCount> at(n_blk + 1);
    BlockCount run{};
    for (uint64_t blk = 0; blk <= n_blk; blk++) {
        at[blk] = run;
        if (blk == n_blk) break;
        run.items += cnt[blk].items;
        run.reads += cnt[blk].reads;
        run.bases += cnt[blk].bases;
        run.residues += cnt[blk].residues;
        for (int k = 0; k < 4; k++) run.out_bound[k] += cnt[blk].out_bound[k];
    }
    const uint64_t n_items = run.items, m = run.reads;
    // ---- pass B: gather the emitting reads, list the (read, frame) items with their peptide offsets
    const bool need_peptides = (kinds_mask & 9u) != 0 && n_items > 0;
    if (need_peptides) {
        st->emit_reads.resize(run.bases + 64);
        st->emit_read_off.resize(m + 1);
        st->emit_item_read.resize(n_items);
        st->emit_item_frame.resize(n_items);
        st->emit_pep_off.resize(n_items + 1);
        st->emit_peptides.resize(run.residues + 64);
        const uint8_t *text = job.slot->text;
        static const int8_t kFrameKeys[6] = {1, 2, 3, -1, -2, -3};
        pool->run(n_blk, [&](uint64_t blk) {
            uint64_t item = at[blk].items, rd = at[blk].reads, base = at[blk].bases, pep = at[blk].residues;
            const uint64_t r1 = std::min(n, (blk + 1) * kBlock);
            for (uint64_t r = blk * kBlock; r < r1; r++) {
                const uint32_t L = b.seq_len[r];
                bool any = false;
                for (int f = 0; f < 6; f++) {
                    if (!has_pep(res[r].category[f])) continue;
                    const uint32_t phase = f % 3;
                    st->emit_item_read[item] = rd;
                    st->emit_item_frame[item] = kFrameKeys[f];
                    st->emit_pep_off[item] = pep;
                    pep += L >= phase ? (L - phase) / 3 : 0;
                    item++;
                    any = true;
                }
                if (any) {
                    st->emit_read_off[rd++] = base;
                    memcpy(st->emit_reads.data() + base, text + b.seq_start[r], L);
                    base += L;
                }
            }
        });
        st->emit_read_off[m] = run.bases;
        st->emit_pep_off[n_items] = run.residues;
        const int rc = orph_translate_frames(ctx, st->emit_reads.data(), st->emit_read_off.data(), m, st->emit_item_read.data(), st->emit_item_frame.data(),
                                             n_items, st->emit_pep_off.data(), st->emit_peptides.data());
        if (rc != ORPH_OK) return rc;
    }
    // ---- pass C: the text of each requested kind, block-parallel into bounded segments, then closed up
    OrphFastaSource src{};
    src.names = b.names;
    src.name_offsets = b.name_off;
    src.results = res;
    src.text = job.slot->text;
    src.seq_start = b.seq_start;
    src.seq_len = b.seq_len;
    src.peptides = st->emit_peptides.data();
    src.pep_offsets = st->emit_pep_off.data();
    src.peptides_include_low_complexity = with_lcp ? 1 : 0;
    for (int kind = 0; kind < 4; kind++) {
        if (!(kinds_mask & (1u << kind)) || run.out_bound[kind] == 0) continue;
        std::vector<uint8_t> &buf = st->emit_out[kind];
        buf.resize(run.out_bound[kind]);
        std::vector<int64_t> len(n_blk);
        pool->run(n_blk, [&](uint64_t blk) {
            len[blk] = orph_format_fasta_range(kind, src, blk * kBlock, std::min(n, (blk + 1) * kBlock), at[blk].items, buf.data() + at[blk].out_bound[kind],
                                               cnt[blk].out_bound[kind]);
        });
        uint64_t total = 0;
        for (uint64_t blk = 0; blk < n_blk; blk++) {
            if (len[blk] < 0 || (uint64_t)len[blk] > cnt[blk].out_bound[kind]) return fail(ORPH_E_INVALID, "orph_stream_emit: internal size bound violated");
            if (len[blk] && at[blk].out_bound[kind] != total) memmove(buf.data() + total, buf.data() + at[blk].out_bound[kind], (size_t)len[blk]);
            total += (uint64_t)len[blk];
        }
        out[kind] = buf.data();
        out_len[kind] = total;
    }
    return ORPH_OK;
}

extern "C" void orph_stream_close(orph_stream *st)
{
    if (!st) return;
    {
        std::lock_guard<std::mutex> lk(st->mu);
        st->stop = true;
        st->cv.notify_all();
    }
    if (st->worker.joinable()) st->worker.join();
    for (StreamRes *r : st->rs) {
        DeviceGuard guard(r->device);
        cudaStreamSynchronize(r->ctx->stream);
    }
    if (st->ing) {
        DeviceGuard guard(st->rs.empty() ? 0 : st->rs[0]->device);
        orph_ingest_close(st->ing);  // pinned slot memory: while a device is current
        st->ing = nullptr;
    }
    for (StreamRes *r : st->rs) stream_res_release(r);
    delete st->emit_pool;
    delete st;
}
