// Host ingest: FASTA / FASTQ (plain or gzip) -> the batch layout the scoring ABI takes (concatenated
// sequence bytes + u64 offsets) plus the record names.  Record semantics are screed's, which is what the
// reference reads its inputs with (orpheum/translate.py:370, orpheum/index.py:83): the name is the whole
// header line without the leading '>' / '@' (no split at white space), sequence lines are concatenated
// as they are (no upper-casing), FASTQ may wrap its sequence over several lines, an empty file has no
// records, and a file that starts with anything else is not a sequence file.
#include <fcntl.h>
#include <mutex>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/statvfs.h>
#include <unistd.h>
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <thread>

#include "../../include/orpheum_b200.h"
#include "hostpack.h"
#include "ingest_internal.h"

extern int orph_set_error(int code, const char *fmt, ...);  // defined in orpheum_b200.cu

// Text window of the input.  gzip: a malloc'ed buffer refilled through gzread (grown with realloc, never
// zero-filled).  Plain files: the whole file mapped read-only, so that the parallel parser touches the page cache
// directly (no copy, page faults spread over the threads).
struct TextBuf {
    char *p = nullptr;
    size_t cap = 0;
    bool mapped = false;
    char *data() const { return p; }
    size_t size() const { return cap; }
    bool grow(size_t want)
    {
        if (mapped) return false;
        char *q = (char *)realloc(p, want);
        if (!q) return false;
        p = q;
        cap = want;
        return true;
    }
    void release()
    {
        if (mapped && p) munmap(p, cap);
        else free(p);
        p = nullptr;
        cap = 0;
    }
};

struct orph_fastx {
    gzFile gz = nullptr;
    char kind = 0;  // '>' FASTA, '@' FASTQ, 0 empty file
    TextBuf buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    double bytes_per_record = 400;  // running estimate, sizes the window of the FASTQ fast path
    std::string pending;  // FASTA: header line of the next record, already consumed
    bool have_pending = false;
    // current batch
    std::vector<uint8_t> seq, names;
    std::vector<uint64_t> seq_off, name_off;
    std::string line;  // scratch
    // fast path for strict four-line FASTQ (parallel over the host cores); any other layout takes the general parser
    bool fast_ok = true;
    orph_host::Pool *pool = nullptr;
    std::vector<uint64_t> nl;               // newline positions of the buffered text
    std::vector<uint32_t> seq_len, name_len;
    std::vector<uint64_t> seq_src, name_src;
    ~orph_fastx() { delete pool; buf.release(); }
};

static bool fill(orph_fastx *fx)
{
    if (fx->eof) return false;
    if (fx->pos > 0 && fx->pos < fx->end) memmove(fx->buf.data(), fx->buf.data() + fx->pos, fx->end - fx->pos);
    fx->end -= fx->pos;
    fx->pos = 0;
    if (fx->end == fx->buf.size() && !fx->buf.grow(fx->buf.size() * 2)) {
        fx->eof = true;  // out of memory: stop reading; the callers report a truncated record
        return false;
    }
    const int n = gzread(fx->gz, fx->buf.data() + fx->end, (unsigned)(fx->buf.size() - fx->end));
    if (n <= 0) {
        fx->eof = true;
        return false;
    }
    fx->end += (size_t)n;
    return true;
}

// next line without its terminator; false at end of file
static bool next_line(orph_fastx *fx, const char **s, size_t *len)
{
    for (;;) {
        const char *base = fx->buf.data() + fx->pos;
        const char *nl = (const char *)memchr(base, '\n', fx->end - fx->pos);
        if (nl) {
            *s = base;
            *len = (size_t)(nl - base);
            fx->pos += *len + 1;
            return true;
        }
        if (!fill(fx)) {
            if (fx->pos < fx->end) {  // last line without a newline
                *s = fx->buf.data() + fx->pos;
                *len = fx->end - fx->pos;
                fx->pos = fx->end;
                return true;
            }
            return false;
        }
    }
}

static void strip(const char **s, size_t *len)
{
    while (*len && ((*s)[*len - 1] == '\r' || (*s)[*len - 1] == ' ' || (*s)[*len - 1] == '\t' || (*s)[*len - 1] == '\n')) (*len)--;
    while (*len && (**s == ' ' || **s == '\t' || **s == '\r')) {
        (*s)++;
        (*len)--;
    }
}

extern "C" int orph_fastx_open(const char *path, orph_fastx **out)
{
    if (!path || !out) return orph_set_error(ORPH_E_INVALID, "orph_fastx_open: NULL argument");
    *out = nullptr;
    gzFile gz = gzopen(path, "rb");
    if (!gz) return orph_set_error(ORPH_E_IO, "cannot open '%s'", path);
    gzbuffer(gz, 1 << 20);
    orph_fastx *fx = new orph_fastx();
    fx->gz = gz;
    bool mapped = false;
    if (gzdirect(gz)) {  // not compressed: map the file instead of copying it through gzread
        const int fd = open(path, O_RDONLY);
        struct stat st;
        if (fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (m != MAP_FAILED) {
                madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                fx->buf.p = (char *)m;
                fx->buf.cap = (size_t)st.st_size;
                fx->buf.mapped = true;
                fx->end = (size_t)st.st_size;
                fx->eof = true;
                mapped = true;
            }
        }
        if (fd >= 0) close(fd);
    }
    if (!mapped) {
        if (!fx->buf.grow(4 << 20)) {
            gzclose(gz);
            delete fx;
            return orph_set_error(ORPH_E_NOMEM, "out of host memory");
        }
        fill(fx);
    }
    if (fx->end == 0) {
        fx->kind = 0;  // empty file: no records
    } else if (fx->buf.p[0] == '>' || fx->buf.p[0] == '@') {
        fx->kind = fx->buf.p[0];
    } else {
        gzclose(gz);
        delete fx;
        return orph_set_error(ORPH_E_INVALID, "unknown file format for '%s'", path);
    }
    *out = fx;
    return ORPH_OK;
}

// internal (ingest.cpp): the general sequential parser from the middle of a file -- at byte `byte_pos` of a plain
// (mapped) file, or behind the first `skip_records` records of a compressed one
int orph_fastx_open_general(const char *path, uint64_t byte_pos, uint64_t skip_records, orph_fastx **out)
{
    const int rc = orph_fastx_open(path, out);
    if (rc != ORPH_OK) return rc;
    orph_fastx *fx = *out;
    fx->fast_ok = false;
    if (fx->buf.mapped) {
        fx->pos = std::min<uint64_t>(byte_pos, fx->end);
        return ORPH_OK;
    }
    while (skip_records > 0) {
        uint64_t n = 0;
        const int rc2 = orph_fastx_next_batch(fx, std::min<uint64_t>(skip_records, 1u << 16), 1ull << 40, &n, nullptr, nullptr, nullptr, nullptr);
        if (rc2 != ORPH_OK || n == 0) break;
        skip_records -= n;
    }
    return ORPH_OK;
}

extern "C" void orph_fastx_close(orph_fastx *fx)
{
    if (!fx) return;
    if (fx->gz) gzclose(fx->gz);
    delete fx;
}

static void begin_record(orph_fastx *fx, const char *name, size_t len)
{
    strip(&name, &len);
    fx->names.insert(fx->names.end(), (const uint8_t *)name, (const uint8_t *)name + len);
    fx->name_off.push_back(fx->names.size());
}

// ------------------------------------------------------------------------------------------------
// Strict four-line FASTQ, parsed in parallel: index the newlines of the buffered text, check every record
// (line 0 starts with '@', line 2 with '+', sequence non-empty and not starting with '+' / '#', quality as long
// as the sequence), then copy names and sequences to their places.  Returns false -- with nothing consumed --
// whenever the text is not of that shape (wrapped records, blank lines, a truncated tail, ...): the general
// parser below then produces the records or the error.
// ------------------------------------------------------------------------------------------------
static inline void strip_range(const char *base, uint64_t &a, uint64_t &b)
{
    while (b > a && (base[b - 1] == '\r' || base[b - 1] == ' ' || base[b - 1] == '\t' || base[b - 1] == '\n')) b--;
    while (b > a && (base[a] == ' ' || base[a] == '\t' || base[a] == '\r')) a++;
}

#include <chrono>
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static bool fastq_fast_batch(orph_fastx *fx, uint64_t max_records, uint64_t max_bases, uint64_t *n_out)
{
    static const bool trace = getenv("ORPHEUM_B200_TRACE") != nullptr;
    const double t_begin = now_s();
    // a window of text that should hold the batch, sized from the bytes per record seen so far
    uint64_t want = (uint64_t)std::min<double>(1024.0 * (1 << 20), std::min<double>((double)max_records * fx->bytes_per_record, (double)max_bases * 3.0)) + (1 << 16);
    while (!fx->eof && fx->end - fx->pos < want)
        if (!fill(fx)) break;
    const char *base = fx->buf.data();
    const uint64_t lo = fx->pos;
    if (fx->end <= lo) return false;
    const uint64_t hi = std::min<uint64_t>(fx->end, lo + want);  // the rest of a mapped file waits for the next batches
    const bool at_end = fx->eof && hi == fx->end;
    const uint64_t n_bytes = hi - lo;
    const double t_filled = now_s();
    if (!fx->pool && n_bytes >= (8u << 20)) {
        const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
        fx->pool = new (std::nothrow) orph_host::Pool((int)std::min(hw, 32u));
    }
    auto parallel_for = [&](uint64_t n_items, const std::function<void(uint64_t)> &fn) {
        if (fx->pool) fx->pool->run(n_items, fn);
        else
            for (uint64_t i = 0; i < n_items; i++) fn(i);
    };
    // ---- newline index
    constexpr uint64_t kSeg = 4u << 20;
    const uint64_t n_seg = (n_bytes + kSeg - 1) / kSeg;
    std::vector<uint64_t> seg_count(n_seg + 1, 0);
    parallel_for(n_seg, [&](uint64_t sgm) {
        const char *p = base + lo + sgm * kSeg, *e = base + std::min(hi, lo + (sgm + 1) * kSeg);
        uint64_t c = 0;
        while (p < e && (p = (const char *)memchr(p, '\n', (size_t)(e - p))) != nullptr) { c++; p++; }
        seg_count[sgm + 1] = c;
    });
    for (uint64_t sgm = 0; sgm < n_seg; sgm++) seg_count[sgm + 1] += seg_count[sgm];
    uint64_t n_lines = seg_count[n_seg];
    const bool tail_line = at_end && base[hi - 1] != '\n';  // last line of the file without a terminator
    fx->nl.resize(n_lines + 1);
    parallel_for(n_seg, [&](uint64_t sgm) {
        const char *p = base + lo + sgm * kSeg, *e = base + std::min(hi, lo + (sgm + 1) * kSeg);
        uint64_t at = seg_count[sgm];
        while (p < e && (p = (const char *)memchr(p, '\n', (size_t)(e - p))) != nullptr) { fx->nl[at++] = (uint64_t)(p - base); p++; }
    });
    if (tail_line) fx->nl[n_lines++] = hi;
    const double t_indexed = now_s();
    if (at_end && n_lines % 4 != 0) return false;  // odd tail: let the general parser decide
    uint64_t n_avail = n_lines / 4;
    if (n_avail == 0) {
        // not even one record in the window (long reads): widen it, unless it already spans everything there is
        if (at_end || want >= (1ull << 30)) return false;
        fx->bytes_per_record = std::max(8.0 * fx->bytes_per_record, 8.0 * (double)want / (double)std::max<uint64_t>(1, max_records));
        return fastq_fast_batch(fx, max_records, max_bases, n_out);
    }
    uint64_t n_take = std::min(n_avail, max_records);
    const uint64_t *nl = fx->nl.data();
    auto line_begin = [&](uint64_t k) { return k == 0 ? lo : nl[k - 1] + 1; };
    // ---- validate + measure
    fx->seq_len.resize(n_take);
    fx->name_len.resize(n_take);
    fx->seq_src.resize(n_take);
    fx->name_src.resize(n_take);
    constexpr uint64_t kRecBlock = 1 << 14;
    const uint64_t n_blk = (n_take + kRecBlock - 1) / kRecBlock;
    std::vector<uint8_t> blk_bad(n_blk, 0);
    parallel_for(n_blk, [&](uint64_t blk) {
        const uint64_t r1 = std::min(n_take, (blk + 1) * kRecBlock);
        for (uint64_t r = blk * kRecBlock; r < r1; r++) {
            uint64_t h0 = line_begin(4 * r), h1 = nl[4 * r];
            uint64_t s0 = h1 + 1, s1 = nl[4 * r + 1];
            const uint64_t p0 = s1 + 1, p1 = nl[4 * r + 2];
            uint64_t q0 = p1 + 1, q1 = nl[4 * r + 3];
            if (h1 == h0 || base[h0] != '@' || p1 == p0 || base[p0] != '+' || s1 == s0 || base[s0] == '+' || base[s0] == '#') { blk_bad[blk] = 1; return; }
            h0++;
            strip_range(base, h0, h1);
            strip_range(base, s0, s1);
            strip_range(base, q0, q1);
            if (s1 == s0 || q1 - q0 != s1 - s0 || s1 - s0 > 0xFFFFFFFFull || h1 - h0 > 0xFFFFFFFFull) { blk_bad[blk] = 1; return; }
            fx->seq_src[r] = s0;
            fx->seq_len[r] = (uint32_t)(s1 - s0);
            fx->name_src[r] = h0;
            fx->name_len[r] = (uint32_t)(h1 - h0);
        }
    });
    for (uint64_t blk = 0; blk < n_blk; blk++)
        if (blk_bad[blk]) return false;
    const double t_validated = now_s();
    // ---- offsets (records are taken while the bases collected so far are below max_bases, like the general parser)
    fx->seq_off.resize(n_take + 1);
    fx->name_off.resize(n_take + 1);
    fx->seq_off[0] = fx->name_off[0] = 0;
    uint64_t n = 0;
    while (n < n_take && fx->seq_off[n] < max_bases) {
        fx->seq_off[n + 1] = fx->seq_off[n] + fx->seq_len[n];
        fx->name_off[n + 1] = fx->name_off[n] + fx->name_len[n];
        n++;
    }
    n_take = n;
    fx->seq_off.resize(n_take + 1);
    fx->name_off.resize(n_take + 1);
    fx->seq.resize(fx->seq_off[n_take]);
    fx->names.resize(fx->name_off[n_take]);
    // ---- copy
    const uint64_t n_cblk = (n_take + kRecBlock - 1) / kRecBlock;
    parallel_for(n_cblk, [&](uint64_t blk) {
        const uint64_t r1 = std::min(n_take, (blk + 1) * kRecBlock);
        for (uint64_t r = blk * kRecBlock; r < r1; r++) {
            memcpy(fx->seq.data() + fx->seq_off[r], base + fx->seq_src[r], fx->seq_len[r]);
            memcpy(fx->names.data() + fx->name_off[r], base + fx->name_src[r], fx->name_len[r]);
        }
    });
    fx->pos = std::min<uint64_t>(hi, nl[4 * n_take - 1] + 1);
    fx->bytes_per_record = 1.05 * (double)(fx->pos - lo) / (double)n_take + 8;
    *n_out = n_take;
    if (trace)
        fprintf(stderr, "[orph trace] fastq batch: %llu records from %.1f MB of text: fill %.1f ms, newline index %.1f ms, validate %.1f ms, offsets + copy %.1f ms\n",
                (unsigned long long)n_take, n_bytes / 1e6, 1e3 * (t_filled - t_begin), 1e3 * (t_indexed - t_filled), 1e3 * (t_validated - t_indexed),
                1e3 * (now_s() - t_validated));
    return true;
}

extern "C" int orph_fastx_next_batch(orph_fastx *fx, uint64_t max_records, uint64_t max_bases, uint64_t *n_records,
                                     const uint8_t **seq_bytes, const uint64_t **seq_offsets, const uint8_t **name_bytes,
                                     const uint64_t **name_offsets)
{
    if (!fx || !n_records) return orph_set_error(ORPH_E_INVALID, "orph_fastx_next_batch: NULL argument");
    fx->seq.clear();
    fx->names.clear();
    fx->seq_off.assign(1, 0);
    fx->name_off.assign(1, 0);
    uint64_t n = 0;
    const char *s;
    size_t len;
    if (fx->kind == '>') {
        while (n < max_records && fx->seq.size() < max_bases) {
            if (!fx->have_pending) {
                if (!next_line(fx, &s, &len)) break;
                if (len == 0 || s[0] != '>') continue;  // stray line before the first header
                fx->pending.assign(s + 1, len - 1);
                fx->have_pending = true;
            }
            begin_record(fx, fx->pending.data(), fx->pending.size());
            fx->have_pending = false;
            while (next_line(fx, &s, &len)) {
                if (len && s[0] == '>') {
                    fx->pending.assign(s + 1, len - 1);
                    fx->have_pending = true;
                    break;
                }
                strip(&s, &len);
                fx->seq.insert(fx->seq.end(), (const uint8_t *)s, (const uint8_t *)s + len);
            }
            fx->seq_off.push_back(fx->seq.size());
            n++;
        }
    } else if (fx->kind == '@' && fx->fast_ok && fastq_fast_batch(fx, max_records, max_bases, &n)) {
        // done: strict four-line records, parsed in parallel
    } else if (fx->kind == '@') {
        fx->fast_ok = false;  // some other FASTQ layout (or the tail of the file): general parser from here on
        fx->seq.clear();
        fx->names.clear();
        fx->seq_off.assign(1, 0);
        fx->name_off.assign(1, 0);
        while (n < max_records && fx->seq.size() < max_bases) {
            if (!next_line(fx, &s, &len)) break;
            {
                const char *t = s;
                size_t tl = len;
                strip(&t, &tl);
                if (tl == 0) continue;  // blank line (possibly just "\r") between records
            }
            if (s[0] != '@') return orph_set_error(ORPH_E_IO, "bad FASTQ: no '@' at the beginning of a record");
            begin_record(fx, s + 1, len - 1);
            const size_t seq_start = fx->seq.size();
            bool plus = false;
            while (next_line(fx, &s, &len)) {
                if (len && (s[0] == '+' || s[0] == '#')) {
                    plus = true;
                    break;
                }
                strip(&s, &len);
                fx->seq.insert(fx->seq.end(), (const uint8_t *)s, (const uint8_t *)s + len);
            }
            const size_t seq_len = fx->seq.size() - seq_start;
            size_t qual_len = 0;
            while (plus && qual_len < seq_len && next_line(fx, &s, &len)) {
                strip(&s, &len);
                qual_len += len;
            }
            if (qual_len != seq_len) return orph_set_error(ORPH_E_IO, "bad FASTQ: quality and sequence lengths differ");
            fx->seq_off.push_back(fx->seq.size());
            n++;
        }
    }
    *n_records = n;
    if (seq_bytes) *seq_bytes = fx->seq.data();
    if (seq_offsets) *seq_offsets = fx->seq_off.data();
    if (name_bytes) *name_bytes = fx->names.data();
    if (name_offsets) *name_offsets = fx->name_off.data();
    return ORPH_OK;
}

// Copies the current batch's sequence bytes and offsets into caller-owned arrays (numpy buffers on the Python side),
// spread over the pool: a fresh 150 MB destination is mostly page faults, which parallelise.
extern "C" int orph_fastx_copy_batch(orph_fastx *fx, uint8_t *seq_dst, uint64_t *seq_offsets_dst)
{
    if (!fx || (!seq_dst && !fx->seq.empty()) || !seq_offsets_dst) return orph_set_error(ORPH_E_INVALID, "orph_fastx_copy_batch: NULL argument");
    memcpy(seq_offsets_dst, fx->seq_off.data(), fx->seq_off.size() * sizeof(uint64_t));
    const uint64_t total = fx->seq.size();
    constexpr uint64_t kBlock = 4u << 20;
    const uint64_t n_blk = (total + kBlock - 1) / kBlock;
    auto work = [&](uint64_t b) { memcpy(seq_dst + b * kBlock, fx->seq.data() + b * kBlock, (size_t)std::min<uint64_t>(kBlock, total - b * kBlock)); };
    if (fx->pool && n_blk > 1) fx->pool->run(n_blk, work);
    else
        for (uint64_t b = 0; b < n_blk; b++) work(b);
    return ORPH_OK;
}

// ------------------------------------------------------------------------------------------------
// Score CSV writer: the rows Translate.score_reads_per_record builds (translate.py:342-365) and
// CreateSaveSummary.maybe_write_csv writes with csv.writer(lineterminator="\n")
// (create_save_summary.py:52-64), produced straight from the result records.
// ------------------------------------------------------------------------------------------------
#include <charconv>

// Python's repr(float) for a finite double: shortest round-trip digits; fixed notation when the decimal
// exponent e satisfies -4 <= e < 16, otherwise d.ddde[+-]XX; integers get a trailing ".0".
static size_t py_repr_double(double v, char *out)
{
    char sci[40];
    auto res = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);  // d[.ddd]e[+-]XX, shortest
    *res.ptr = 0;
    char *e = strchr(sci, 'e');
    const int exp10 = atoi(e + 1);
    char digits[32];
    size_t nd = 0;
    for (char *p = sci; p < e; p++)
        if (*p >= '0' && *p <= '9') digits[nd++] = *p;
    char *o = out;
    if (v < 0) *o++ = '-';
    if (exp10 >= -4 && exp10 < 16) {
        if (exp10 < 0) {
            *o++ = '0';
            *o++ = '.';
            for (int i = 0; i < -exp10 - 1; i++) *o++ = '0';
            for (size_t i = 0; i < nd; i++) *o++ = digits[i];
        } else {
            for (size_t i = 0; i < nd || (int)i <= exp10; i++) {
                if ((int)i == exp10 + 1) *o++ = '.';
                *o++ = i < nd ? digits[i] : '0';
            }
            if ((int)nd <= exp10 + 1) {
                *o++ = '.';
                *o++ = '0';
            }
        }
    } else {
        *o++ = digits[0];
        if (nd > 1) {
            *o++ = '.';
            for (size_t i = 1; i < nd; i++) *o++ = digits[i];
        }
        *o++ = 'e';
        *o++ = exp10 < 0 ? '-' : '+';
        const int a = exp10 < 0 ? -exp10 : exp10;
        if (a < 10) *o++ = '0';
        o += snprintf(o, 8, "%d", a);
    }
    *o = 0;
    return (size_t)(o - out);
}

extern "C" int orph_format_float_repr(double v, char *out32)
{
    if (!out32) return orph_set_error(ORPH_E_INVALID, "orph_format_float_repr: NULL buffer");
    if (v != v) {
        strcpy(out32, "nan");
        return ORPH_OK;
    }
    py_repr_double(v, out32);
    return ORPH_OK;
}

extern "C" int orph_csv_write_scores(const char *path, int append, const uint8_t *name_fields, const uint64_t *name_offsets,
                                     uint64_t n_reads, const orph_read_result *results, const char *const *category_text,
                                     const char *filename_field, int n_threads)
{
    if (!path || !category_text || !filename_field || (n_reads && (!name_fields || !name_offsets || !results)))
        return orph_set_error(ORPH_E_INVALID, "orph_csv_write_scores: NULL argument");
    // (no O_APPEND: pwrite would ignore its offset).  Read-write when possible: the formatted blocks are then copied into a
    // shared mapping of the file's new tail by all threads at once -- pwrite()s to one file queue behind its inode lock
    // (about 2 GB/s on tmpfs), page faults into a mapping do not.
    int fd = open(path, O_RDWR | O_CREAT | (append ? 0 : O_TRUNC), 0666);
    bool can_map = fd >= 0 && !getenv("ORPHEUM_B200_CSV_PWRITE");
    if (fd < 0) fd = open(path, O_WRONLY | O_CREAT | (append ? 0 : O_TRUNC), 0666);
    if (fd < 0) return orph_set_error(ORPH_E_IO, "cannot open '%s' for writing", path);
    const uint64_t page = (uint64_t)sysconf(_SC_PAGESIZE);
    struct stat st;
    uint64_t file_at = 0;
    if (append && fstat(fd, &st) == 0) file_at = (uint64_t)st.st_size;
    // everything of a row behind the two numbers, per (category, frame): ",<category>,<frame>,<filename>\n"
    static const char *kFrames[6] = {"1", "2", "3", "-1", "-2", "-3"};
    std::string tail[6][6];
    size_t max_tail = 0;
    for (int c = 0; c < 6; c++)
        for (int fr = 0; fr < 6; fr++) {
            if (!category_text[c]) continue;
            tail[c][fr] = std::string(",") + category_text[c] + "," + kFrames[fr] + "," + filename_field + "\n";
            max_tail = std::max(max_tail, tail[c][fr].size());
        }
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<uint64_t>((uint64_t)std::min(n_threads, 64), std::max<uint64_t>(1, n_reads / 4096));
    orph_host::Pool pool(n_threads);
    constexpr uint64_t kBlock = 8192;
    const uint64_t n_blocks = (n_reads + kBlock - 1) / kBlock, group = (uint64_t)n_threads * 4;
    // block buffers are kept between calls (the CLI calls this once per batch): fresh ones cost a page fault per 4 KB
    static std::mutex cache_mutex;
    static std::vector<std::vector<char>> cache;
    std::unique_lock<std::mutex> cache_lock(cache_mutex, std::try_to_lock);
    std::vector<std::vector<char>> local;
    std::vector<std::vector<char>> &bufs = cache_lock.owns_lock() ? cache : local;
    if (bufs.size() < std::min(group, std::max<uint64_t>(1, n_blocks))) bufs.resize(std::min(group, std::max<uint64_t>(1, n_blocks)));
    std::vector<uint64_t> used(bufs.size(), 0), at(bufs.size(), 0);
    std::atomic<uint64_t> bad_read{~0ull};
    std::atomic<bool> io_ok{true};
    for (uint64_t g0 = 0; g0 < n_blocks && io_ok.load() && bad_read.load() == ~0ull; g0 += group) {
        const uint64_t g1 = std::min(n_blocks, g0 + group);
        pool.run(g1 - g0, [&](uint64_t k) {
            const uint64_t r0 = (g0 + k) * kBlock, r1 = std::min(n_reads, r0 + kBlock);
            std::vector<char> &buf = bufs[k];
            const uint64_t need = 6 * ((name_offsets[r1] - name_offsets[r0]) + (r1 - r0) * (40 + max_tail));
            if (buf.size() < need) buf.resize(need);
            char *o = buf.data();
            // (n_hits, n_kmers) -> repr(n_hits / n_kmers): the same few quotients come back all the time
            struct Memo { uint32_t key; uint32_t len; char txt[24]; };
            static thread_local Memo memo[1024];
            static thread_local bool memo_ready = false;
            if (!memo_ready) {
                for (auto &e : memo) e.key = ~0u;
                memo_ready = true;
            }
            for (uint64_t r = r0; r < r1; r++) {
                const uint8_t *name = name_fields + name_offsets[r];
                const size_t name_len = (size_t)(name_offsets[r + 1] - name_offsets[r]);
                const orph_read_result &res = results[r];
                for (int fr = 0; fr < 6; fr++) {
                    const unsigned cat = res.category[fr];
                    if (cat > 5 || !category_text[cat]) {
                        uint64_t expect = ~0ull;
                        bad_read.compare_exchange_strong(expect, r);
                        used[k] = 0;
                        return;
                    }
                    memcpy(o, name, name_len);
                    o += name_len;
                    *o++ = ',';
                    const bool scored = cat == ORPH_CAT_CODING || cat == ORPH_CAT_NON_CODING;
                    if (scored) {
                        const uint32_t key = ((uint32_t)res.n_hits[fr] << 16) | res.n_kmers[fr];
                        Memo &e = memo[(key * 2654435761u) >> 22];
                        if (e.key != key) {
                            e.key = key;
                            e.len = (uint32_t)py_repr_double((double)res.n_hits[fr] / (double)res.n_kmers[fr], e.txt);
                        }
                        memcpy(o, e.txt, e.len);
                        o += e.len;
                    } else {
                        memcpy(o, "nan", 3);
                        o += 3;
                    }
                    *o++ = ',';
                    if (scored || cat == ORPH_CAT_LOW_COMPLEXITY_PEPTIDE) {
                        auto res_n = std::to_chars(o, o + 8, (unsigned)res.n_kmers[fr]);
                        o = res_n.ptr;
                    } else {
                        memcpy(o, "nan", 3);
                        o += 3;
                    }
                    const std::string &t = tail[cat][fr];
                    memcpy(o, t.data(), t.size());
                    o += t.size();
                }
            }
            used[k] = (uint64_t)(o - buf.data());
        });
        if (bad_read.load() != ~0ull) break;
        for (uint64_t k = 0; k < g1 - g0; k++) {
            at[k] = file_at;
            file_at += used[k];
        }
        if (can_map && file_at > at[0]) {
            // a store into a mapping of a file system that has run out of space is a SIGBUS, a pwrite() an error code: the
            // mapping is used only while the file system has room to spare
            struct statvfs vfs;
            if (fstatvfs(fd, &vfs) != 0 || (uint64_t)vfs.f_bavail * (uint64_t)vfs.f_frsize < 2 * (file_at - at[0]) + (64u << 20)) can_map = false;
        }
        if (can_map && file_at > at[0]) {
            const uint64_t map_from = at[0] / page * page;
            void *m = MAP_FAILED;
            if (ftruncate(fd, (off_t)file_at) == 0) m = mmap(nullptr, (size_t)(file_at - map_from), PROT_READ | PROT_WRITE, MAP_SHARED, fd, (off_t)map_from);
            if (m != MAP_FAILED) {
                char *base = static_cast<char *>(m) - map_from;  // base + file offset
                pool.run(g1 - g0, [&](uint64_t k) { memcpy(base + at[k], bufs[k].data(), (size_t)used[k]); });
                munmap(m, (size_t)(file_at - map_from));
                continue;
            }
            can_map = false;  // a file system without shared mappings: pwrite from here on (it extends the file itself)
        }
        pool.run(g1 - g0, [&](uint64_t k) {
            uint64_t done = 0;
            while (done < used[k]) {
                const ssize_t w = pwrite(fd, bufs[k].data() + done, (size_t)(used[k] - done), (off_t)(at[k] + done));
                if (w <= 0) {
                    io_ok.store(false);
                    return;
                }
                done += (uint64_t)w;
            }
        });
    }
    const bool closed = close(fd) == 0;
    if (bad_read.load() != ~0ull)
        return orph_set_error(ORPH_E_INVALID, "orph_csv_write_scores: read %llu has a frame category without a text", (unsigned long long)bad_read.load());
    return io_ok.load() && closed ? ORPH_OK : orph_set_error(ORPH_E_IO, "write to '%s' failed", path);
}

// ------------------------------------------------------------------------------------------------
// FASTA records Translate emits for one batch (translate.py:244-298), in read order then frame order, formatted
// straight from the result records.
// ------------------------------------------------------------------------------------------------
// Records of reads [r0, r1) of a batch; `item0` = number of peptide-bearing frames before read r0.  Writes at most `cap`
// bytes and returns the full length.  -1: a peptide was needed and is missing.
int64_t orph_format_fasta_range(int kind, const OrphFastaSource &src, uint64_t r0, uint64_t r1, uint64_t item0, uint8_t *out, uint64_t out_cap)
{
    static const int kFrameKeys[6] = {1, 2, 3, -1, -2, -3};
    static const char kLowSuffix[] = " translation_frame: {}.format(frame)";  // verbatim, translate.py:247
    uint64_t at = 0, item = item0;
    char num[48], mid[80];
    auto put = [&](const void *p, uint64_t n) {
        if (at + n <= out_cap && out) memcpy(out + at, p, n);
        at += n;
    };
    for (uint64_t r = r0; r < r1; r++) {
        const orph_read_result &res = src.results[r];
        const uint8_t *name = src.names + src.name_offsets[r];
        const uint64_t name_len = src.name_offsets[r + 1] - src.name_offsets[r];
        for (int f = 0; f < 6; f++) {
            const unsigned cat = res.category[f];
            const bool has_pep = cat == ORPH_CAT_CODING || (src.peptides_include_low_complexity && cat == ORPH_CAT_LOW_COMPLEXITY_PEPTIDE);
            const uint64_t my_item = item;
            if (has_pep) item++;
            const bool wanted = kind == 0 ? cat == ORPH_CAT_CODING : kind == 1 ? cat == ORPH_CAT_CODING
                                : kind == 2 ? cat == ORPH_CAT_NON_CODING : cat == ORPH_CAT_LOW_COMPLEXITY_PEPTIDE;
            if (!wanted) continue;
            if ((kind == 0 || kind == 3) && (!has_pep || !src.peptides || !src.pep_offsets)) return -1;
            put(">", 1);
            if (kind == 3) {
                put(name, name_len);
                put(kLowSuffix, sizeof(kLowSuffix) - 1);
            } else {
                // "{name}__translation_frame:{frame}__jaccard:{jaccard}" with every space replaced by "__"
                uint64_t run = 0;
                for (uint64_t i = 0; i < name_len; i++) {
                    if (name[i] == ' ') {
                        if (i > run) put(name + run, i - run);
                        put("__", 2);
                        run = i + 1;
                    }
                }
                if (name_len > run) put(name + run, name_len - run);
                const size_t nj = py_repr_double((double)res.n_hits[f] / (double)res.n_kmers[f], num);
                const int nm = snprintf(mid, sizeof(mid), "__translation_frame:%d__jaccard:", kFrameKeys[f]);
                put(mid, (uint64_t)nm);
                put(num, nj);
            }
            put("\n", 1);
            if (kind == 0 || kind == 3) put(src.peptides + src.pep_offsets[my_item], src.pep_offsets[my_item + 1] - src.pep_offsets[my_item]);
            else if (src.reads) put(src.reads + src.read_offsets[r], src.read_offsets[r + 1] - src.read_offsets[r]);
            else put(src.text + src.seq_start[r], src.seq_len[r]);
            put("\n", 1);
        }
    }
    return (int64_t)at;
}

extern "C" int orph_format_fasta(int kind, const uint8_t *names, const uint64_t *name_offsets, uint64_t n_reads,
                                 const orph_read_result *results, const uint8_t *reads, const uint64_t *read_offsets,
                                 const uint8_t *peptides, const uint64_t *pep_offsets, int peptides_include_low_complexity,
                                 uint8_t *out, uint64_t out_cap, uint64_t *out_len)
{
    if (kind < 0 || kind > 3 || !out_len || (n_reads && (!names || !name_offsets || !results)))
        return orph_set_error(ORPH_E_INVALID, "orph_format_fasta: bad arguments");
    if ((kind == 1 || kind == 2) && n_reads && (!reads || !read_offsets)) return orph_set_error(ORPH_E_INVALID, "orph_format_fasta: reads missing");
    OrphFastaSource src{};
    src.names = names;
    src.name_offsets = name_offsets;
    src.results = results;
    src.reads = reads;
    src.read_offsets = read_offsets;
    src.peptides = peptides;
    src.pep_offsets = pep_offsets;
    src.peptides_include_low_complexity = peptides_include_low_complexity;
    const int64_t n = orph_format_fasta_range(kind, src, 0, n_reads, 0, out, out_cap);
    if (n < 0) return orph_set_error(ORPH_E_INVALID, "orph_format_fasta: peptides missing");
    *out_len = (uint64_t)n;
    return ORPH_OK;
}

// ------------------------------------------------------------------------------------------------
// Are all read ids different?  The JSON summary groups rows by read id (create_save_summary.py:180-222); when every id
// occurs once -- the usual case -- the groups are the reads themselves and no per-id bookkeeping is needed.  Ids are
// compared through 64-bit hashes: "distinct" is certain, "not distinct" may be a hash collision (the caller then takes
// its exact path).
// ------------------------------------------------------------------------------------------------
static inline uint64_t hash_bytes64(const uint8_t *p, size_t len)
{
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    uint64_t h = 0x9e3779b97f4a7c15ULL ^ (len * m);
    for (; len >= 8; p += 8, len -= 8) {
        uint64_t k;
        memcpy(&k, p, 8);
        k *= m; k ^= k >> 47; k *= m;
        h ^= k; h *= m;
    }
    uint64_t k = 0;
    for (size_t i = 0; i < len; i++) k |= (uint64_t)p[i] << (8 * i);
    h ^= k; h *= m;
    h ^= h >> 47; h *= m; h ^= h >> 47;
    return h;
}

extern "C" int orph_hash_names(const uint8_t *names, const uint64_t *name_offsets, uint64_t n, uint64_t *out_hashes, int n_threads)
{
    if (n && (!names || !name_offsets || !out_hashes)) return orph_set_error(ORPH_E_INVALID, "orph_hash_names: NULL argument");
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<uint64_t>((uint64_t)std::min(n_threads, 64), std::max<uint64_t>(1, n / 65536));
    constexpr uint64_t kBlock = 1 << 15;
    auto work = [&](uint64_t blk) {
        const uint64_t r1 = std::min(n, (blk + 1) * kBlock);
        for (uint64_t r = blk * kBlock; r < r1; r++) out_hashes[r] = hash_bytes64(names + name_offsets[r], (size_t)(name_offsets[r + 1] - name_offsets[r]));
    };
    const uint64_t n_blk = (n + kBlock - 1) / kBlock;
    if (n_threads > 1) {
        orph_host::Pool pool(n_threads);
        pool.run(n_blk, work);
    } else {
        for (uint64_t b = 0; b < n_blk; b++) work(b);
    }
    return ORPH_OK;
}

extern "C" int orph_u64_all_distinct(const uint64_t *values, uint64_t n, int n_threads, int *all_distinct)
{
    if (!all_distinct || (n && !values)) return orph_set_error(ORPH_E_INVALID, "orph_u64_all_distinct: NULL argument");
    *all_distinct = 1;
    if (n < 2) return ORPH_OK;
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<uint64_t>((uint64_t)std::min(n_threads, 64), std::max<uint64_t>(1, n / 65536));
    // partition by the top byte (two parallel passes), then sort every bucket on its own
    constexpr uint64_t kBlock = 1 << 16;
    const uint64_t n_blk = (n + kBlock - 1) / kBlock;
    std::vector<uint64_t> hist(n_blk * 256, 0), tmp(n);
    orph_host::Pool pool(n_threads);
    pool.run(n_blk, [&](uint64_t blk) {
        uint64_t *h = hist.data() + blk * 256;
        const uint64_t r1 = std::min(n, (blk + 1) * kBlock);
        for (uint64_t r = blk * kBlock; r < r1; r++) h[values[r] >> 56]++;
    });
    uint64_t bucket_start[257], run = 0;
    for (int b = 0; b < 256; b++) {
        bucket_start[b] = run;
        for (uint64_t blk = 0; blk < n_blk; blk++) {
            const uint64_t c = hist[blk * 256 + b];
            hist[blk * 256 + b] = run;
            run += c;
        }
    }
    bucket_start[256] = run;
    pool.run(n_blk, [&](uint64_t blk) {
        uint64_t *cur = hist.data() + blk * 256;
        const uint64_t r1 = std::min(n, (blk + 1) * kBlock);
        for (uint64_t r = blk * kBlock; r < r1; r++) tmp[cur[values[r] >> 56]++] = values[r];
    });
    std::atomic<int> dup{0};
    pool.run(256, [&](uint64_t b) {
        uint64_t *a = tmp.data() + bucket_start[b], *e = tmp.data() + bucket_start[b + 1];
        if (e - a < 2 || dup.load(std::memory_order_relaxed)) return;
        std::sort(a, e);
        if (std::adjacent_find(a, e) != e) dup.store(1);
    });
    *all_distinct = dup.load() ? 0 : 1;
    return ORPH_OK;
}
