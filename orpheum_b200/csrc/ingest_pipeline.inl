// Host-buffer ingest of orph_score_reads (included by orpheum_b200.cu, same translation unit): the chunked
// three-stream pipeline (H2D / kernels / D2H over kPipeSlots staging slots), the planner that decides per chunk
// between packing to 2 bits on the host threads and shipping the ASCII as it is, the second pass over reads the
// 2-bit stream cannot represent, and the ingest statistics.  SURVEY.md 8f rank 1; the reference's counterpart is the
// serial loop over screed records in orpheum/translate.py:367-372.
#pragma once

// max over i of off[i+1] - off[i] for i in [0, m); a decreasing pair wraps to >= 2^63.  Four independent accumulators
// per thread, spread over the pool when there is one (1 M offsets in well under 0.1 ms).
static uint64_t scan_read_lengths(orph_host::Pool *pool, const uint64_t *off, uint64_t m)
{
    auto part = [&](uint64_t a, uint64_t b) {
        uint64_t m0 = 0, m1 = 0, m2 = 0, m3 = 0;
        uint64_t i = a;
        for (; i + 4 <= b; i += 4) {
            m0 = std::max(m0, off[i + 1] - off[i]);
            m1 = std::max(m1, off[i + 2] - off[i + 1]);
            m2 = std::max(m2, off[i + 3] - off[i + 2]);
            m3 = std::max(m3, off[i + 4] - off[i + 3]);
        }
        for (; i < b; i++) m0 = std::max(m0, off[i + 1] - off[i]);
        return std::max(std::max(m0, m1), std::max(m2, m3));
    };
    constexpr uint64_t kBlock = 1 << 16;
    if (!pool || m <= 2 * kBlock) return part(0, m);
    const uint64_t n_blocks = (m + kBlock - 1) / kBlock;
    std::vector<uint64_t> partial(n_blocks, 0);
    pool->run(n_blocks, [&](uint64_t blk) { partial[blk] = part(blk * kBlock, std::min(m, (blk + 1) * kBlock)); });
    uint64_t best = 0;
    for (uint64_t v : partial) best = std::max(best, v);
    return best;
}

// Host-buffer pipeline.  The batch is cut into chunks that flow through three stages on three streams -- H2D of
// chunk c+1, kernels of chunk c and D2H of chunk c-1 overlap (PCIe is full duplex) -- over kPipeSlots staging slots.
//
// ASCII batches are PCIe-bound (150 B per 150-nt read against 37.5 B as 2-bit codes), so when `allow_pack` is set the
// host cores pack part of the chunks (SIMD, thread pool) while the copy engine ships the others as ASCII: before
// each chunk the H2D backlog is compared with the time the host needs to pack the chunk, and the chunk goes out as
// ASCII whenever the copy engine would otherwise run dry.  Reads that cannot be 2-bit packed (N, lower case, ...)
// are scored a second time from their bytes (`exceptions`, filled with their indices) and overwrite the record of the
// first pass; a chunk where more than 1/8 of the reads are of that kind is sent as ASCII instead.
static int score_host_pipeline(orph_ctx *ctx, const orph_filter *f, const void *reads, int reads_format, const uint64_t *offsets,
                               uint64_t n_reads, const uint8_t lut[256], double jaccard_threshold, orph_read_result *out,
                               bool allow_pack_in, std::vector<uint64_t> *exceptions)
{
    bool allow_pack = allow_pack_in;
    using clk = std::chrono::steady_clock;
    int rc;
    cudaStream_t s = ctx->stream;
    if (!ctx->pipe_ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking));
        for (int b = 0; b < kPipeSlots; b++) {
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_in[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_kernels[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&ctx->ev_out[b], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreate(&ctx->ev_h2d_begin[b]));
            CUDA_TRY(cudaEventCreate(&ctx->ev_h2d_end[b]));
        }
        ctx->pipe_ready = true;
    }
    // every early return below leaves copies / kernels of earlier chunks in flight on the three streams, reading the
    // caller's host buffers: drain the device before handing control back
    struct DrainOnError {
        bool armed = true;
        ~DrainOnError()
        {
            if (armed) cudaDeviceSynchronize();
        }
    } drain;
    bool input_pinned = false;
    if (!ctx->pool && (allow_pack || n_reads >= (1ULL << 18))) {  // the pool also scans the offsets of large batches
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
    // chunk boundaries: <= 2^19 reads and <= 2^27 bases each
    static const int chunk_log2 = [] { const char *v = getenv("ORPHEUM_B200_CHUNK_LOG2"); const int n = v ? atoi(v) : 19; return n < 12 ? 12 : n > 24 ? 24 : n; }();
    static const char *skip_env = getenv("ORPHEUM_B200_AB_SKIP");  // development A/B: "kernels", "d2h", "h2d"
    static const bool skip_kernels = skip_env && strstr(skip_env, "kernels"), skip_d2h = skip_env && strstr(skip_env, "d2h"), skip_h2d = skip_env && strstr(skip_env, "h2d");
    const uint64_t kMaxChunkReads = 1ULL << chunk_log2, kMaxChunkBases = 1ULL << (chunk_log2 + 8);
    if (allow_pack) {
        // pinned staging for host-packed chunks: every slot sized now, for the largest chunk this call can form (pinned
        // allocations cost milliseconds and synchronise the device, and nothing is in flight at this point)
        const uint64_t max_chunk_words = std::min<uint64_t>(kMaxChunkBases, offsets[n_reads] - offsets[0]) / 32 + 4;
        for (int q = 0; q < kPipeSlots && allow_pack; q++)
            if (ctx->h_packed[q].ensure(max_chunk_words * 8) != ORPH_OK) {
                cudaGetLastError();  // no pinned memory to be had (locked-memory limit): every chunk goes out as it is
                allow_pack = false;
            }
    }
    if (allow_pack) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, reads) == cudaSuccess) input_pinned = attr.type == cudaMemoryTypeHost;
        else cudaGetLastError();
    }
    // ORPHEUM_B200_TRACE=1: per-chunk timeline of the three stages on stderr (development aid)
    static const bool trace = getenv("ORPHEUM_B200_TRACE") != nullptr;
    struct ChunkTrace { cudaEvent_t e[6]; double host_enqueue_s, host_pack_s; int packed; uint64_t m; };
    std::vector<ChunkTrace> tr;
    const auto t_call = clk::now();
    auto mark = [&](int which, cudaStream_t st) {
        if (trace) { cudaEventCreate(&tr.back().e[which]); cudaEventRecord(tr.back().e[which], st); }
    };
    // the H2D rate actually achieved by a finished slot (PCIe generation, NUMA placement and contention with the
    // running kernels all move it) replaces the planner's estimate
    auto calibrate_h2d = [&](int slot) {
        float ms = 0;
        if (ctx->h2d_bytes[slot] >= (1u << 20) && cudaEventElapsedTime(&ms, ctx->ev_h2d_begin[slot], ctx->ev_h2d_end[slot]) == cudaSuccess && ms > 0) {
            const double rate = (double)ctx->h2d_bytes[slot] / (1e-3 * ms);
            ctx->h2d_bytes_per_s = 0.75 * ctx->h2d_bytes_per_s + 0.25 * rate;
        } else {
            cudaGetLastError();
        }
        ctx->h2d_bytes[slot] = 0;
    };
    std::vector<uint8_t> needs;
    uint64_t lo = 0;
    int c = 0;
    bool used[kPipeSlots] = {};
    int ascii_streak = 0;  // chunks still to be sent as ASCII after one with many unpackable reads
    clk::time_point h2d_free_at = clk::now();  // planning estimate: when the copy engine runs out of queued work
    while (lo < n_reads) {
        // this chunk's extent, validity and longest read (overlaps the GPU work of the previous chunks)
        // the first chunks are small so that the pipeline fills quickly (nothing overlaps the first H2D)
        const int ramp = std::min(c, 3);
        const uint64_t chunk_reads = kMaxChunkReads >> (3 - ramp), chunk_bases = kMaxChunkBases >> (3 - ramp);
        uint64_t hi = std::min(n_reads, lo + chunk_reads);
        uint64_t longest = scan_read_lengths(ctx->pool, offsets + lo, hi - lo);
        if (longest < (1ULL << 63) && offsets[hi] - offsets[lo] > chunk_bases) {
            hi = (uint64_t)(std::upper_bound(offsets + lo + 1, offsets + hi + 1, offsets[lo] + chunk_bases) - offsets) - 1;
            if (hi <= lo) hi = lo + 1;
            longest = scan_read_lengths(ctx->pool, offsets + lo, hi - lo);
        }
        if (longest >= (1ULL << 63)) {  // a decreasing pair wraps around
            uint64_t at = lo;
            while (at < hi && offsets[at + 1] >= offsets[at]) at++;
            cudaDeviceSynchronize();
            return fail(ORPH_E_INVALID, "orph_score_reads: offsets are not monotonic at read %llu", (unsigned long long)at);
        }
        uint32_t max_len = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(longest, ORPH_MAX_READ_LEN + 1));
        if (max_len > ORPH_MAX_READ_LEN) max_len = ORPH_MAX_READ_LEN;  // over-long reads are reported through the status
        const uint64_t m = hi - lo;
        const int b = c % kPipeSlots;
        const uint64_t b0 = offsets[lo], b1 = offsets[hi];
        // the staging buffers of slot b are free once the kernels (inputs) and the D2H (results) of its previous chunk are done
        if (used[b]) {
            CUDA_TRY(cudaStreamWaitEvent(ctx->s_in, ctx->ev_kernels[b], 0));
            CUDA_TRY(cudaEventSynchronize(ctx->ev_out[b]));  // also bounds how far the host runs ahead
            calibrate_h2d(b);
        }
        if ((rc = ctx->in_offsets2[b].ensure((m + 1) * 8)) != ORPH_OK) return rc;
        if ((rc = ctx->out_results2[b].ensure(m * sizeof(orph_read_result))) != ORPH_OK) return rc;
        const uint8_t *d_ascii = nullptr;
        const uint64_t *d_packed = nullptr;
        uint64_t pbase0 = 0, n_ascii = 0;
        bool pack_here = false;
        if (trace) { tr.push_back(ChunkTrace{}); tr.back().m = m; tr.back().host_pack_s = 0; }
        if (reads_format == ORPH_READS_ASCII && allow_pack && b1 > b0 && ascii_streak == 0) {
            // pack on the host unless the copy engine would run dry meanwhile (only a pinned buffer can be sent asynchronously)
            const double backlog = std::chrono::duration<double>(h2d_free_at - clk::now()).count();
            const double pack_s = ctx->pack_s_per_byte * (double)(b1 - b0);
            if (!input_pinned) pack_here = true;                       // pageable input: the CPU reads it faster than a staged copy
            else if (ctx->pack_s_per_byte == 0) pack_here = c > 0;    // rate not measured yet: chunk 0 feeds the copy engine
            else if (ctx->pack_s_per_byte * ctx->h2d_bytes_per_s > 1.0 && (c & 15) != 15)
                pack_here = false;  // the host packs slower than the wire carries (many ranks sharing few cores and one memory
                                    // system): packing only adds host memory traffic; one chunk in 16 re-measures the rate
            else pack_here = backlog >= pack_s;
        }
        if (ascii_streak > 0) ascii_streak--;
        if (pack_here && ((b1 - b0 + 31) / 32 + 2) * 8 > ctx->h_packed[b].cap) pack_here = false;  // one absurdly long read: no staging for it
        if (pack_here) {
            const uint64_t n_words = (b1 - b0 + 31) / 32;
            needs.assign(m, 0);
            const auto t0 = clk::now();
            orph_host::pack_stream(ctx->pool, (const uint8_t *)reads + b0, b0, b1 - b0, offsets + lo, m, (uint64_t *)ctx->h_packed[b].p, needs.data());
            const double dt = std::chrono::duration<double>(clk::now() - t0).count() / (double)(b1 - b0);
            ctx->pack_s_per_byte = ctx->pack_s_per_byte == 0 ? dt : 0.75 * ctx->pack_s_per_byte + 0.25 * dt;
            if (trace) tr.back().host_pack_s = dt * (double)(b1 - b0);
            uint64_t n_flagged = 0;
            for (uint64_t i = 0; i < m; i++) n_flagged += needs[i];
            if (n_flagged * 8 > m) {  // mostly unpackable input (soft-masked genome, ...): ASCII is the cheaper route
                pack_here = false;
                ascii_streak = 8;
            } else {
                if (n_flagged && exceptions)
                    for (uint64_t i = 0; i < m; i++)
                        if (needs[i]) exceptions->push_back(lo + i);
                if ((rc = ctx->in_reads2[b].ensure((n_words + 2) * 8)) != ORPH_OK) return rc;
                d_packed = (const uint64_t *)ctx->in_reads2[b].p;
                pbase0 = b0;
                ctx->chunks_packed_on_host++;
                h2d_free_at = std::max(h2d_free_at, clk::now()) +
                              std::chrono::duration_cast<clk::duration>(std::chrono::duration<double>((double)(n_words * 8 + (m + 1) * 8) / ctx->h2d_bytes_per_s));
            }
        }
        if (trace) { tr.back().packed = pack_here; tr.back().host_enqueue_s = std::chrono::duration<double>(clk::now() - t_call).count(); }
        mark(0, ctx->s_in);
        CUDA_TRY(cudaEventRecord(ctx->ev_h2d_begin[b], ctx->s_in));
        ctx->h2d_bytes[b] = (m + 1) * 8 + (pack_here ? ((b1 - b0 + 31) / 32) * 8 : reads_format == ORPH_READS_ASCII ? b1 - b0 : (b1 - b0) / 4);
        if (pack_here) {
            const uint64_t n_words = (b1 - b0 + 31) / 32;
            CUDA_TRY(cudaMemcpyAsync(ctx->in_reads2[b].p, ctx->h_packed[b].p, n_words * 8, cudaMemcpyHostToDevice, ctx->s_in));
        }
        CUDA_TRY(cudaMemcpyAsync(ctx->in_offsets2[b].p, offsets + lo, (m + 1) * 8, cudaMemcpyHostToDevice, ctx->s_in));
        if (!pack_here) {
            if (reads_format == ORPH_READS_ASCII) {
                n_ascii = b1 - b0;
                if ((rc = ctx->in_reads2[b].ensure(n_ascii + 64)) != ORPH_OK) return rc;
                if (n_ascii) CUDA_TRY(cudaMemcpyAsync(ctx->in_reads2[b].p, (const uint8_t *)reads + b0, n_ascii, cudaMemcpyHostToDevice, ctx->s_in));
                d_ascii = (const uint8_t *)ctx->in_reads2[b].p;
                ctx->chunks_sent_ascii++;
                h2d_free_at = std::max(h2d_free_at, clk::now()) +
                              std::chrono::duration_cast<clk::duration>(std::chrono::duration<double>((double)(n_ascii + (m + 1) * 8) / ctx->h2d_bytes_per_s));
            } else {
                const uint64_t w0 = b0 / 32, w1 = (b1 + 31) / 32;
                if ((rc = ctx->in_reads2[b].ensure((w1 - w0 + 2) * 8)) != ORPH_OK) return rc;
                if (w1 > w0 && !(skip_h2d && c >= 2 * kPipeSlots)) CUDA_TRY(cudaMemcpyAsync(ctx->in_reads2[b].p, (const uint64_t *)reads + w0, (w1 - w0) * 8, cudaMemcpyHostToDevice, ctx->s_in));
                d_packed = (const uint64_t *)ctx->in_reads2[b].p;
                pbase0 = w0 * 32;
            }
        }
        mark(1, ctx->s_in);
        CUDA_TRY(cudaEventRecord(ctx->ev_h2d_end[b], ctx->s_in));
        CUDA_TRY(cudaEventRecord(ctx->ev_in[b], ctx->s_in));
        CUDA_TRY(cudaStreamWaitEvent(s, ctx->ev_in[b], 0));
        if (used[b]) CUDA_TRY(cudaStreamWaitEvent(s, ctx->ev_out[b], 0));
        mark(2, s);
        rc = skip_kernels ? ORPH_OK : score_device_impl(ctx, f, d_ascii, b0, d_packed, pbase0, n_ascii, (const uint64_t *)ctx->in_offsets2[b].p, m, max_len, lut,
                               jaccard_threshold, (orph_read_result *)ctx->out_results2[b].p);
        if (rc != ORPH_OK) {
            cudaDeviceSynchronize();
            return rc;
        }
        mark(3, s);
        CUDA_TRY(cudaEventRecord(ctx->ev_kernels[b], s));
        CUDA_TRY(cudaStreamWaitEvent(ctx->s_out, ctx->ev_kernels[b], 0));
        mark(4, ctx->s_out);
        if (!skip_d2h) CUDA_TRY(cudaMemcpyAsync(out + lo, ctx->out_results2[b].p, m * sizeof(orph_read_result), cudaMemcpyDeviceToHost, ctx->s_out));
        mark(5, ctx->s_out);
        CUDA_TRY(cudaEventRecord(ctx->ev_out[b], ctx->s_out));
        used[b] = true;
        lo = hi;
        c++;
    }
    ScoreCounters counters;
    CUDA_TRY(cudaMemcpyAsync(&counters, ctx->counters.p, sizeof(counters), cudaMemcpyDeviceToHost, s));
    CUDA_TRY(cudaStreamSynchronize(s));
    CUDA_TRY(cudaStreamSynchronize(ctx->s_out));
    CUDA_TRY(cudaStreamSynchronize(ctx->s_in));
    for (int q = 0; q < kPipeSlots; q++)
        if (used[q]) calibrate_h2d(q);
    if (trace && !tr.empty()) {
        fprintf(stderr, "[orph trace] call of %llu reads, %.3f ms on the host clock\n", (unsigned long long)n_reads,
                1e3 * std::chrono::duration<double>(clk::now() - t_call).count());
        fprintf(stderr, "[orph trace] chunk reads mode host_enq host_pack | h2d_start h2d_end | k_start k_end | d2h_start d2h_end (ms, device times relative to chunk 0's h2d_start)\n");
        for (size_t i = 0; i < tr.size(); i++) {
            float t[6];
            for (int k = 0; k < 6; k++) cudaEventElapsedTime(&t[k], tr[0].e[0], tr[i].e[k]);
            fprintf(stderr, "[orph trace] %3zu %8llu %s %8.3f %8.3f | %8.3f %8.3f | %8.3f %8.3f | %8.3f %8.3f\n", i, (unsigned long long)tr[i].m,
                    tr[i].packed ? "pack " : "plain", 1e3 * tr[i].host_enqueue_s, 1e3 * tr[i].host_pack_s, t[0], t[1], t[2], t[3], t[4], t[5]);
        }
        for (auto &ct : tr)
            for (int k = 0; k < 6; k++) cudaEventDestroy(ct.e[k]);
    }
    CUDA_TRY(cudaMemsetAsync((uint8_t *)ctx->counters.p + offsetof(ScoreCounters, status), 0, sizeof(uint32_t), s));
    drain.armed = false;  // all three streams were synchronised above
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
    const bool allow_pack = reads_format == ORPH_READS_ASCII && ctx->host_threads != 0 && offsets[n_reads] - offsets[0] >= (1ULL << 22);
    std::vector<uint64_t> exceptions;
    rc = score_host_pipeline(ctx, f, reads, reads_format, offsets, n_reads, lut, jaccard_threshold, out, allow_pack, &exceptions);
    if (rc != ORPH_OK && rc != ORPH_E_EMPTY_READ && rc != ORPH_E_TOO_LONG) return rc;
    if (!exceptions.empty()) {
        // second pass over the reads the 2-bit stream cannot represent: their bytes, gathered into one small ASCII batch
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

