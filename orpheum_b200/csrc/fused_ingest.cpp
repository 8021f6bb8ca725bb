// Fused FASTQ parse + 2-bit pack (see fused_ingest.h).  Record semantics are the ones fastx.cpp implements for strict
// four-line FASTQ (screed's: name = header line without '@', stripped; sequence line stripped; reference call sites
// orpheum/translate.py:370, orpheum/index.py:83).
#include "fused_ingest.h"

#include <immintrin.h>

#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>

namespace orph_host {

namespace {

constexpr uint64_t kSegBytes = 1u << 20;  // text per work item: small enough to stay in a core's L2 between the two phases

struct Rec {
    uint64_t seq_start, name_start;
    uint32_t seq_len, name_len;
};

// running totals up to and including segment s, published in segment order
struct Cum {
    uint64_t reads = 0, bases = 0, name_bytes = 0;
    uint64_t end = 0;         // text position behind the last record of the verified chain
    bool open = true;         // the chain is intact up to the end of this segment: the next segment may append to it
    bool not_strict = false;  // the chain ended at a record that is not strict four-line FASTQ
    bool full = false;        // the chain ended because the next segment did not fit the batch's arrays
};

struct SegState {
    Cum cum;
    std::atomic<int> ready{0};
    // partial words at the two ends of the segment's stretch of the packed stream (shared with the neighbours)
    uint64_t head_w = 0, head_bits = 0, tail_w = 0, tail_bits = 0;
    bool has_head = false, has_tail = false;
    bool written = false;  // phase B ran: the segment's reads are part of the batch
    uint32_t max_len = 0;
    std::vector<uint32_t> flagged;
};

inline bool is_strip_end(uint8_t c) { return c == '\r' || c == ' ' || c == '\t' || c == '\n'; }
inline bool is_strip_begin(uint8_t c) { return c == ' ' || c == '\t' || c == '\r'; }

inline void strip_range(const uint8_t *t, uint64_t &a, uint64_t &b)
{
    while (b > a && is_strip_end(t[b - 1])) b--;
    while (b > a && is_strip_begin(t[a])) a++;
}

// end of the line starting at p: position of its '\n', or hi when the text ends first (found = false)
inline uint64_t line_end(const uint8_t *t, uint64_t p, uint64_t hi, bool &found)
{
    const void *nl = p < hi ? memchr(t + p, '\n', (size_t)(hi - p)) : nullptr;
    found = nl != nullptr;
    return nl ? (uint64_t)((const uint8_t *)nl - t) : hi;
}

enum class RecStatus { ok, incomplete, bad };

// line_end with the newline search inlined (32 bytes per compare): the lines of a short-read record are 2 .. 300 bytes, a
// memchr() call per line costs more than the search itself
__attribute__((target("avx2"))) inline uint64_t line_end_avx2(const uint8_t *t, uint64_t p, uint64_t hi, bool &found)
{
    const __m256i nl = _mm256_set1_epi8('\n');
    uint64_t q = p;
    for (; q + 32 <= hi; q += 32) {
        const uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_loadu_si256(reinterpret_cast<const __m256i *>(t + q)), nl));
        if (m) {
            found = true;
            return q + (uint64_t)__builtin_ctz(m);
        }
    }
    for (; q < hi; q++)
        if (t[q] == '\n') {
            found = true;
            return q;
        }
    found = false;
    return hi;
}

// One strict four-line record starting at p < hi.  On ok, `next` is the start of the following record.
#define ORPH_PARSE_RECORD_BODY(LINE_END)                                                                                       \
    const RecStatus cut = at_eof ? RecStatus::bad : RecStatus::incomplete; /* the text ends inside the record */               \
    bool found;                                                                                                                \
    if (t[p] != '@') return RecStatus::bad;                                                                                    \
    uint64_t h0 = p + 1, h1 = LINE_END(t, p, hi, found);                                                                       \
    if (!found) return cut;                                                                                                    \
    uint64_t s0 = h1 + 1, s1 = LINE_END(t, s0, hi, found);                                                                     \
    if (!found) return cut;                                                                                                    \
    const uint64_t p0 = s1 + 1;                                                                                                \
    if (p0 >= hi) return cut;                                                                                                  \
    if (t[p0] != '+') return RecStatus::bad;                                                                                   \
    uint64_t p1;                                                                                                               \
    if (p0 + 1 < hi && t[p0 + 1] == '\n') p1 = p0 + 1; /* the usual bare "+" */                                                \
    else {                                                                                                                     \
        p1 = LINE_END(t, p0, hi, found);                                                                                       \
        if (!found) return cut;                                                                                                \
    }                                                                                                                          \
    uint64_t q0 = p1 + 1;                                                                                                      \
    const uint64_t q_nl = LINE_END(t, q0, hi, found);                                                                          \
    if (!found && (!at_eof || q0 >= hi)) return cut; /* only the very last line of the file may lack its terminator */         \
    uint64_t q1 = q_nl;                                                                                                        \
    if (s1 == s0 || t[s0] == '+' || t[s0] == '#') return RecStatus::bad;                                                       \
    /* the usual record needs no stripping: test the end bytes before entering the loops */                                    \
    if (h1 > h0 && (is_strip_end(t[h1 - 1]) || is_strip_begin(t[h0]))) strip_range(t, h0, h1);                                 \
    if (is_strip_end(t[s1 - 1]) || is_strip_begin(t[s0])) strip_range(t, s0, s1);                                              \
    if (q1 > q0 && (is_strip_end(t[q1 - 1]) || is_strip_begin(t[q0]))) strip_range(t, q0, q1);                                 \
    if (s1 == s0 || q1 - q0 != s1 - s0 || s1 - s0 > 0xFFFFFFFFull || h1 - h0 > 0xFFFFFFFFull) return RecStatus::bad;           \
    rec.seq_start = s0;                                                                                                        \
    rec.seq_len = (uint32_t)(s1 - s0);                                                                                         \
    rec.name_start = h0;                                                                                                       \
    rec.name_len = (uint32_t)(h1 - h0);                                                                                        \
    next = found ? q_nl + 1 : hi;                                                                                              \
    return RecStatus::ok;

inline RecStatus parse_record(const uint8_t *t, uint64_t p, uint64_t hi, bool at_eof, Rec &rec, uint64_t &next)
{
    ORPH_PARSE_RECORD_BODY(line_end)
}

__attribute__((target("avx2"))) inline RecStatus parse_record_avx2(const uint8_t *t, uint64_t p, uint64_t hi, bool at_eof, Rec &rec, uint64_t &next)
{
    ORPH_PARSE_RECORD_BODY(line_end_avx2)
}
#undef ORPH_PARSE_RECORD_BODY

// The records that start in [p, seg_hi), appended to `recs`; p ends behind the last complete one.
#define ORPH_SCAN_RECORDS_BODY(PARSE)                                  \
    while (p < seg_hi) {                                               \
        Rec rec;                                                       \
        uint64_t next;                                                 \
        const RecStatus st = PARSE(t, p, hi, at_eof, rec, next);       \
        if (st == RecStatus::incomplete) break;                        \
        if (st == RecStatus::bad) {                                    \
            bad = true;                                                \
            break;                                                     \
        }                                                              \
        recs.push_back(rec);                                           \
        n_bases += rec.seq_len;                                        \
        n_name += rec.name_len;                                        \
        max_len = std::max(max_len, rec.seq_len);                      \
        p = next;                                                      \
    }

void scan_records(const uint8_t *t, uint64_t &p, uint64_t seg_hi, uint64_t hi, bool at_eof, std::vector<Rec> &recs, uint64_t &n_bases,
                  uint64_t &n_name, uint32_t &max_len, bool &bad)
{
    ORPH_SCAN_RECORDS_BODY(parse_record)
}

__attribute__((target("avx2"))) void scan_records_avx2(const uint8_t *t, uint64_t &p, uint64_t seg_hi, uint64_t hi, bool at_eof, std::vector<Rec> &recs,
                                                        uint64_t &n_bases, uint64_t &n_name, uint32_t &max_len, bool &bad)
{
    ORPH_SCAN_RECORDS_BODY(parse_record_avx2)
}
#undef ORPH_SCAN_RECORDS_BODY

// First position >= from that looks like the start of a record: a line start holding '@' whose next-but-one line
// starts with '+'.  (A quality line may start with '@', but the line two below it is a sequence line, which strict
// records never start with '+'.)  When the text ends before the check can be made the candidate is returned as it is;
// the chain verification decides.  Returns hi when there is no candidate.
uint64_t find_record_start(const uint8_t *t, uint64_t lo, uint64_t from, uint64_t hi)
{
    uint64_t p = from;
    bool found;
    if (p > lo && t[p - 1] != '\n') {
        p = line_end(t, p, hi, found);
        if (!found) return hi;
        p++;
    }
    while (p < hi) {
        const uint64_t e0 = line_end(t, p, hi, found);
        if (t[p] == '@') {
            if (!found) return p;
            const uint64_t e1 = line_end(t, e0 + 1, hi, found);
            if (!found || e1 + 1 >= hi) return p;
            if (t[e1 + 1] == '+') return p;
        }
        if (e0 >= hi) return hi;
        p = e0 + 1;
    }
    return hi;
}

// ---- 32 bases -> 64 bits.  bad: bit i set when byte i is not upper-case ACGT -------------------------------------
inline uint64_t pack32_scalar(const uint8_t *p, uint32_t count, uint32_t &bad)
{
    uint64_t bits = 0;
    uint32_t b = 0;
    for (uint32_t i = 0; i < count; i++) {
        const uint8_t c = p[i];
        const uint32_t code = (c >> 1) & 3u;
        bits |= (uint64_t)code << (2 * i);
        b |= (uint32_t)(c != (uint8_t)"ACTG"[code]) << i;
    }
    bad = b;
    return bits;
}

__attribute__((target("avx2"))) inline uint64_t pack32_avx2(const uint8_t *p, uint32_t &bad)
{
    const __m256i v = _mm256_loadu_si256(reinterpret_cast<const __m256i *>(p));
    const __m256i code = _mm256_and_si256(_mm256_srli_epi16(v, 1), _mm256_set1_epi8(3));
    const __m256i lut = _mm256_setr_epi8('A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 'A', 'C', 'T', 'G', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    bad = ~(uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, code), v));
    const __m256i gather = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1,
                                            -1, -1, -1, -1);
    const __m256i g = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(code, _mm256_set1_epi16(0x0401)), _mm256_set1_epi32(0x00100001)), gather);
    return (uint64_t)(uint32_t)_mm256_extract_epi32(g, 0) | ((uint64_t)(uint32_t)_mm256_extract_epi32(g, 4) << 32);
}

// Appends bit groups to the packed stream from an arbitrary bit position.  The first and the last word of a segment's
// stretch are shared with the neighbouring segments: they are handed back instead of stored.
struct BitWriter {
    uint64_t *dst;
    uint64_t w, acc = 0;
    uint32_t pend;
    bool first_partial;
    SegState *seg;
    BitWriter(uint64_t *stream, uint64_t bitpos, SegState *s) : dst(stream), w(bitpos >> 6), pend((uint32_t)(bitpos & 63)), first_partial((bitpos & 63) != 0), seg(s) {}
    inline void emit(uint64_t v)
    {
        if (first_partial) {
            seg->head_w = w;
            seg->head_bits = v;
            seg->has_head = true;
            first_partial = false;
        } else {
            _mm_stream_si64(reinterpret_cast<long long *>(dst + w), (long long)v);  // written once, read by the DMA engine
        }
        w++;
    }
    // bits: only the low nbits (2 .. 64, even) may be set
    inline void put(uint64_t bits, uint32_t nbits)
    {
        acc |= bits << pend;
        const uint32_t tot = pend + nbits;
        if (tot >= 64) {
            emit(acc);
            acc = pend ? bits >> (64 - pend) : 0;
            pend = tot - 64;
        } else {
            pend = tot;
        }
    }
    void finish()
    {
        if (pend) {
            seg->tail_w = w;
            seg->tail_bits = acc;
            seg->has_tail = true;
        }
    }
};

__attribute__((target("avx2"))) void pack_records_avx2(const uint8_t *t, uint64_t text_end, const Rec *recs, uint64_t n, uint64_t read0,
                                                             BitWriter &bw, std::vector<uint32_t> &flagged)
{
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *s = t + recs[i].seq_start;
        const uint32_t L = recs[i].seq_len;
        uint32_t any_bad = 0, bad, g = 0;
        for (; g + 32 <= L; g += 32) {
            bw.put(pack32_avx2(s + g, bad), 64);
            any_bad |= bad;
        }
        if (g < L) {
            const uint32_t c = L - g;
            uint64_t bits;
            if (recs[i].seq_start + g + 32 <= text_end) {  // reading past the read's end is harmless inside the text
                bits = pack32_avx2(s + g, bad) & ((1ull << (2 * c)) - 1);
                bad &= (1u << c) - 1;
            } else {
                bits = pack32_scalar(s + g, c, bad);
            }
            bw.put(bits, 2 * c);
            any_bad |= bad;
        }
        if (any_bad) flagged.push_back((uint32_t)(read0 + i));
    }
}

void pack_records_scalar(const uint8_t *t, uint64_t, const Rec *recs, uint64_t n, uint64_t read0, BitWriter &bw,
                              std::vector<uint32_t> &flagged)
{
    for (uint64_t i = 0; i < n; i++) {
        const uint8_t *s = t + recs[i].seq_start;
        const uint32_t L = recs[i].seq_len;
        uint32_t any_bad = 0, bad;
        for (uint32_t g = 0; g < L; g += 32) {
            const uint32_t c = std::min<uint32_t>(32, L - g);
            bw.put(pack32_scalar(s + g, c, bad), 2 * c);
            any_bad |= bad;
        }
        if (any_bad) flagged.push_back((uint32_t)(read0 + i));
    }
}

bool have_avx2()
{
    static const bool yes = [] {
        __builtin_cpu_init();
        return (bool)__builtin_cpu_supports("avx2");
    }();
    return yes;
}

}  // namespace

FusedResult parse_fastq_window(Pool *pool, const uint8_t *text, uint64_t lo, uint64_t hi, uint64_t text_end, bool at_eof, FusedBatch &batch)
{
    FusedResult result;
    result.consumed_to = lo;
    batch.n_reads = batch.total_bases = batch.names_bytes = 0;
    batch.max_len = 0;
    batch.flagged.clear();
    if (hi <= lo) return result;
    const uint64_t n_seg = (hi - lo + kSegBytes - 1) / kSegBytes;
    std::vector<SegState> segs(n_seg);
    const bool avx2 = have_avx2();

    auto work = [&](uint64_t s) {
        SegState &me = segs[s];
        const uint64_t seg_lo = lo + s * kSegBytes, seg_hi = std::min(hi, seg_lo + kSegBytes);
        // ---- phase A: the records that start in [seg_lo, seg_hi)
        static thread_local std::vector<Rec> recs;
        recs.clear();
        const uint64_t first = s == 0 ? lo : find_record_start(text, lo, seg_lo, hi);
        uint64_t p = first, n_bases = 0, n_name = 0;
        uint32_t max_len = 0;
        bool bad = false;
        if (avx2) scan_records_avx2(text, p, seg_hi, hi, at_eof, recs, n_bases, n_name, max_len, bad);
        else scan_records(text, p, seg_hi, hi, at_eof, recs, n_bases, n_name, max_len, bad);
        // ---- running totals, in segment order (the segments before this one were handed out earlier)
        Cum prev;
        prev.end = lo;
        if (s > 0) {
            while (!segs[s - 1].ready.load(std::memory_order_acquire)) std::this_thread::yield();
            prev = segs[s - 1].cum;
        }
        Cum cum = prev;
        const bool joined = prev.open && first == prev.end;  // this segment continues the sequential parse
        uint64_t n = recs.size();
        if (joined) {
            // arrays full: the batch ends in front of this segment
            const uint64_t words_needed = (prev.bases + n_bases + 31) / 32 + 1;
            if (prev.reads + n > batch.cap_reads || prev.name_bytes + n_name > batch.names_cap || words_needed > batch.packed_cap_words ||
                prev.bases + n_bases > 0xFFFFFFFFull) {
                cum.open = false;
                cum.full = true;
                n = 0;
            } else {
                cum.reads += n;
                cum.bases += n_bases;
                cum.name_bytes += n_name;
                cum.end = p;
                if (bad) { cum.open = false; cum.not_strict = true; }
                else if (p < seg_hi) cum.open = false;  // an incomplete record: the window ends here
            }
        } else {
            cum.open = false;  // a wrong guess of the segment's first record (or an earlier stop): the batch ends at prev.end
            n = 0;
        }
        me.cum = cum;
        me.ready.store(1, std::memory_order_release);
        if (!joined || n == 0) return;
        // ---- phase B: everything at its final place
        me.written = true;
        me.max_len = max_len;
        const uint64_t r0 = prev.reads;
        uint64_t base = prev.bases, nb = prev.name_bytes;
        for (uint64_t i = 0; i < n; i++) {
            const Rec &rec = recs[i];
            batch.seq_start[r0 + i] = rec.seq_start;
            batch.seq_len[r0 + i] = rec.seq_len;
            batch.off32[r0 + i] = (uint32_t)base;
            batch.name_off[r0 + i] = nb;
            memcpy(batch.names + nb, text + rec.name_start, rec.name_len);
            base += rec.seq_len;
            nb += rec.name_len;
        }
        BitWriter bw(batch.packed, 2 * prev.bases, &me);
        if (avx2) pack_records_avx2(text, text_end, recs.data(), n, r0, bw, me.flagged);
        else pack_records_scalar(text, text_end, recs.data(), n, r0, bw, me.flagged);
        bw.finish();
    };
    if (pool && n_seg > 1) pool->run(n_seg, work);
    else
        for (uint64_t s = 0; s < n_seg; s++) work(s);
    _mm_sfence();

    // ---- the batch is the chain of joined segments
    const Cum &last = segs[n_seg - 1].cum;
    batch.n_reads = last.reads;
    batch.total_bases = last.bases;
    batch.names_bytes = last.name_bytes;
    result.consumed_to = last.end;
    result.not_strict = last.not_strict;
    if (batch.n_reads) {
        batch.off32[batch.n_reads] = (uint32_t)last.bases;
        batch.name_off[batch.n_reads] = last.name_bytes;
    }
    // shared words between neighbouring segments: zero them, then OR the contributions in
    for (const SegState &sg : segs) {
        if (!sg.written) continue;
        if (sg.has_head) batch.packed[sg.head_w] = 0;
        if (sg.has_tail) batch.packed[sg.tail_w] = 0;
    }
    for (const SegState &sg : segs) {
        if (!sg.written) continue;
        if (sg.has_head) batch.packed[sg.head_w] |= sg.head_bits;
        if (sg.has_tail) batch.packed[sg.tail_w] |= sg.tail_bits;
        batch.max_len = std::max(batch.max_len, sg.max_len);
        batch.flagged.insert(batch.flagged.end(), sg.flagged.begin(), sg.flagged.end());
    }
    result.out_of_capacity = last.full;
    return result;
}

}  // namespace orph_host
