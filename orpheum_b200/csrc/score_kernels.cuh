// Scoring kernels: everything under Translate.maybe_score_single_read (orpheum/translate.py:323-331).
//
//   pack_ascii_kernel      ASCII -> 2-bit stream, flags reads that need the byte-exact path
//   prefilter_kernel       one THREAD per read on the 2-bit stream: fastp complexity
//                          (translate.py:56-84), stop codons in all six frames by bit-parallel
//                          matching (translate.py:218), too-short test (translate.py:227); writes the
//                          whole result record and queues reads that still have frames to score
//   score_frames_kernel    one WARP per queued read: translates only the open frames through a
//                          shared-memory codon table (translate_single_seq.py:21-28 fused with
//                          encode_peptide, sequence_encodings.py:457-468), then score_frame()
//   score_bytes_kernel     one WARP per read with the reference's byte semantics (N, lower case,
//                          IUPAC, over-long reads); used for the reads the two kernels above skip
//
// score_frame(): distinct k-mers (compare_kmer_content.py:116-118) via a per-warp shared-memory hash
// set of [hash tag | k-mer position] entries, every tag match confirmed on the k-mer bytes (exact),
// Bloom probes with early exit (translate.py:187-191).  The short-read hot kernel is task_kernels.cuh.
#pragma once
#include "device_common.cuh"

namespace orph {

constexpr int kWarp = 32;
constexpr uint32_t kFastMaxLen = 1536;   // longest read the packed fast path handles
constexpr uint32_t kBytesMaxLen = 6144;  // longest read the warp-per-read byte path stages in shared memory; beyond: block per read
constexpr uint32_t kStatusEmptyRead = 1u;
constexpr uint32_t kStatusTooLong = 2u;

constexpr uint32_t kLenBuckets = 128;  // peptide lengths of the thread-per-frame tiers (<= 320 / 3) fit

struct ScoreParams {
    FilterDev F;
    uint8_t lut[256];          // encode_peptide as a byte table
    double jaccard_threshold;
    uint32_t k;                // peptide k-mer size (== F.ksize)
    uint32_t max_len;          // declared upper bound of read length in this batch
    uint32_t fast_max_len;     // reads longer than this skip the packed fast path
    uint32_t thread_max_len;   // reads up to this length are scored one thread per frame (0: path disabled)
    uint32_t thread_short_len; // of those, reads up to this length use the short tier (smaller seen filter)
    uint32_t seen_words;       // 64-bit words of the per-thread seen filter in this launch (power of two)
    uint32_t task_capacity;    // entries in the task queue: the short tier fills it from the front, the long tier from the back
    uint32_t slots;            // dedup table slots per warp (power of two >= 128, >= 2 * max k-mers per frame)
    uint32_t pep_bytes;        // peptide buffer bytes per warp (multiple of 4, incl. kPepPad)
    uint32_t read_words;       // staged read words per warp (fast path); per thread: height of a task's strand / residue column
    uint32_t stage_words;      // score_tasks_kernel: words of that column the 2-bit strand takes
    uint32_t force_exact;      // test hook: always take the exact all-pairs dedup path
};

// device counters shared by one scoring call
struct ScoreCounters {
    uint32_t n_queue;      // reads queued for score_frames_kernel (warp per read)
    uint32_t next_queue;   // dynamic work fetch
    uint32_t n_bytes;      // reads queued for score_bytes_kernel
    uint32_t next_bytes;
    uint32_t n_tasks;      // (read, frame) tasks queued for score_tasks_kernel (thread per frame), short tier
    uint32_t next_task;
    uint32_t n_tasks_long; // same, long tier (reads of kThreadMaxLenShort+1 .. kThreadMaxLenLong bases)
    uint32_t next_task_long;
    uint32_t sort_tier[2];        // set by task_scan_kernel: this tier's tasks are consumed from the length-sorted copy
    uint32_t hist[2][kLenBuckets];   // tasks per peptide length, per tier (filled by prefilter_kernel)
    uint32_t cursor[2][kLenBuckets]; // scatter cursors of the counting sort (longest peptides first)
    uint32_t n_long;       // reads queued for score_long_kernel (block per read), from the back of the byte queue
    uint32_t next_long;
    uint32_t n_tasks_n;    // tasks of reads that hold N (and nothing else unusual): thread per frame with an N mask
    uint32_t next_task_n;
    uint32_t n_reads_n;    // such reads, queued by prefilter_kernel for prefilter_n_kernel
    uint32_t pad[2];
    uint32_t status;       // kStatus* bits; sticky until orph_ctx_check -- keep it last
};

// ------------------------------------------------------------------------------------------------
// score_frame: all lanes of one warp.  enc = encoded peptide bytes in shared memory (4-byte aligned,
// kPepPad readable bytes behind plen), plen > k.  Returns warp-uniform (n_distinct, n_hits).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool kmer_bytes_equal(const uint8_t *enc, uint32_t a, uint32_t b, uint32_t k)
{
    for (uint32_t i = 0; i < k; i++)
        if (enc[a + i] != enc[b + i]) return false;
    return true;
}

// Exact reference semantics with no hashing shortcuts for the de-duplication: k-mer j counts iff no
// earlier position holds the same bytes.  O(n^2 k); only reached on a genuine 64-bit hash collision
// (or through the force_exact test hook).
__device__ __noinline__ void score_frame_exact(const uint8_t *enc, uint32_t plen, uint32_t k, const FilterDev &F,
                                               int lane, uint32_t &n_distinct, uint32_t &n_hits)
{
    const uint32_t n = plen - k + 1;
    uint32_t nd = 0, nh = 0;
    const SmemKeyWords base{reinterpret_cast<const uint32_t *>(enc), 0, k};
    for (uint32_t j0 = 0; j0 < n; j0 += kWarp) {
        const uint32_t j = j0 + lane;
        bool first = j < n;
        for (uint32_t e = 0; first && e < j; e++)
            if (kmer_bytes_equal(enc, e, j, k)) first = false;
        bool hit = false;
        if (first) {
            SmemKeyWords kw = base;
            kw.j = j;
            hit = filter_contains(F, murmur3_h1(kw, k));
        }
        nd += __popc(__ballot_sync(0xffffffffu, first));
        nh += __popc(__ballot_sync(0xffffffffu, hit));
    }
    n_distinct = nd;
    n_hits = nh;
}

// De-duplication table entry: [ 21-bit tag of the k-mer's hash | 11-bit k-mer position ].  A slot whose
// tag matches is only a duplicate if the k-mer bytes at the stored position equal ours, so the set is
// exact with no fallback; 0xFFFFFFFF (position 2047 cannot occur, see kMaxPeptide) marks an empty slot.
constexpr uint32_t kPosBits = 11;
constexpr uint32_t kEmptySlot = 0xFFFFFFFFu;
constexpr uint32_t kMaxPeptide = kBytesMaxLen / 3;  // 2048 residues: positions <= 2048 - k < 2047

template <int K>  // K > 0: compile-time k-mer size; K == 0: runtime k
__device__ __forceinline__ void score_frame(const uint8_t *enc, uint32_t plen, uint32_t k_rt, const FilterDev &F,
                                            uint32_t *tab, uint32_t slots, int lane, bool force_exact,
                                            uint32_t &n_distinct, uint32_t &n_hits)
{
    const uint32_t k = K > 0 ? (uint32_t)K : k_rt;
    const uint32_t n = plen - k + 1;
    if (force_exact) {  // test hook: the all-pairs restatement instead of the hash set
        score_frame_exact(enc, plen, k, F, lane, n_distinct, n_hits);
        return;
    }
    // smallest power-of-two table with load <= 1/4 for this frame (never larger than the carve-out,
    // which the host sizes for load <= 1/2 at the longest admissible peptide)
    uint32_t mask = 127;
    while (mask + 1 < 4 * n && mask + 1 < slots) mask = 2 * mask + 1;
    {
        uint4 *t4 = reinterpret_cast<uint4 *>(tab);  // 16-byte aligned, slot count is a multiple of 128
        for (uint32_t s = lane; s <= (mask >> 2); s += kWarp) t4[s] = make_uint4(kEmptySlot, kEmptySlot, kEmptySlot, kEmptySlot);
    }
    __syncwarp();

    uint32_t nd = 0, nh = 0;
    const uint32_t *encw = reinterpret_cast<const uint32_t *>(enc);
    for (uint32_t j0 = 0; j0 < n; j0 += kWarp) {
        const uint32_t j = j0 + lane;
        bool winner = false;
        uint64_t h = 0;
        uint32_t v0 = 0, sh0 = 0;
        if (j < n) {
            h = murmur3_h1(SmemKeyWords{encw, j, k}, k);
            // table-0 probe first: its L2 latency overlaps the de-duplication below
            const uint64_t bin0 = filter_bin(F, 0, h);
            v0 = ld_table_byte(F.table[0] + (bin0 >> 3));
            sh0 = (uint32_t)bin0 & 7u;
            const uint32_t tag = (uint32_t)(h >> 40) & ((1u << (32 - kPosBits)) - 1);
            const uint32_t mine = (tag << kPosBits) | j;
            uint32_t slot = (uint32_t)h & mask;
            const uint32_t step = ((uint32_t)(h >> 20) & mask) | 1u;  // odd: visits every slot
            while (true) {
                const uint32_t old = atomicCAS(&tab[slot], kEmptySlot, mine);
                if (old == kEmptySlot) { winner = true; break; }
                if ((old >> kPosBits) == tag && kmer_bytes_equal(enc, old & ((1u << kPosBits) - 1), j, k)) break;
                slot = (slot + step) & mask;
            }
        }
        bool hit = winner && ((v0 >> sh0) & 1u);
        if (hit && F.n_tables > 1) hit = filter_contains_rest(F, h);
        nd += __popc(__ballot_sync(0xffffffffu, winner));
        nh += __popc(__ballot_sync(0xffffffffu, hit));
    }
    n_distinct = nd;
    n_hits = nh;
}

// category of a scored frame (translate.py:241-298): low-complexity test first, then the threshold
__device__ __forceinline__ uint8_t scored_category(uint32_t n_distinct, uint32_t n_hits, uint32_t plen, uint32_t k,
                                                   double thr)
{
    // evaluate_is_kmer_low_complexity: n_kmers <= (len - k + 1) / 2   (translate.py:98-100)
    if ((double)n_distinct <= (double)(plen - k + 1) / 2.0) return ORPH_CAT_LOW_COMPLEXITY_PEPTIDE;
    const double jaccard = (double)n_hits / (double)n_distinct;  // translate.py:204
    return jaccard > thr ? ORPH_CAT_CODING : ORPH_CAT_NON_CODING;
}

// smallest hit count h in [0, n] with (double)h / (double)n > thr, or n + 1 if there is none: the
// reference's `jaccard > threshold` (translate.py:204,271) as an integer compare, decided by the very
// same IEEE division (monotone in h, so the first h that passes is exact).
__device__ __forceinline__ uint32_t min_hits_for(uint32_t n, double thr)
{
    const double dn = (double)n;
    double est = floor(thr * dn) - 2.0;
    uint32_t h = est > 0.0 ? (est < dn ? (uint32_t)est : n) : 0u;
    while (h <= n && !((double)h / dn > thr)) h++;
    return h;
}

struct BestFrame {
    uint32_t hits = 0, n = 0;
    int frame = 0;
    // exact rational comparison hits/n > best.hits/best.n; ties keep the first frame in order
    __device__ __forceinline__ void offer(uint32_t h, uint32_t nk, uint8_t cat, int f)
    {
        if (cat != ORPH_CAT_CODING && cat != ORPH_CAT_NON_CODING) return;
        if (frame == 0 || (uint64_t)h * n > (uint64_t)hits * nk) { hits = h; n = nk; frame = f; }
    }
};

__device__ __forceinline__ int frame_key(int f) { return f < 3 ? f + 1 : -(f - 2); }  // 0..5 -> 1,2,3,-1,-2,-3

// ------------------------------------------------------------------------------------------------
// pack_ascii_kernel: thread t packs bases [32t, 32t+32) of the concatenated ASCII stream.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t find_read(const uint64_t *offsets, uint64_t n_reads, uint64_t g)
{
    // largest r with offsets[r] <= g  (g is an absolute base index)
    uint64_t lo = 0, hi = n_reads;
    while (lo + 1 < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= g) lo = mid; else hi = mid;
    }
    return (uint32_t)lo;
}

// `ascii` points at the byte of absolute base index abase0; the packed stream starts at absolute base pbase0.
// Besides the 2-bit words it writes an N plane -- bit i of nmask[w] is set where base 32 w + i is an upper-case 'N' -- and
// classifies the reads: needs_bytes[r] = 1 if read r holds a byte that is neither ACGT nor N (lower case, IUPAC codes, ...:
// only the byte-exact kernels can score it), has_n[r] = 1 if it holds an N (scored on the 2-bit path with the N plane).
__global__ void pack_ascii_kernel(const uint8_t *__restrict__ ascii, uint64_t abase0, const uint64_t *__restrict__ offsets,
                                  uint64_t n_reads, uint64_t pbase0, uint64_t total_bases,
                                  uint64_t *__restrict__ packed, uint8_t *__restrict__ needs_bytes, uint32_t *__restrict__ nmask,
                                  uint8_t *__restrict__ has_n)
{
    const uint8_t *reads = ascii + (pbase0 - abase0);  // byte of the first base to pack
    const uint64_t n_words = (total_bases + 31) / 32;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < n_words;
         w += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g0 = 32 * w;  // base index of this word's first base, relative to pbase0
        const uint32_t count = (uint32_t)((total_bases - 32 * w) < 32 ? (total_bases - 32 * w) : 32);
        uint64_t bits = 0;
        uint32_t bad = 0;  // bit i: byte i is not upper-case ACGT
        uint32_t isn = 0;  // bit i: byte i is 'N'
        if (count == 32 && ((reinterpret_cast<uintptr_t>(reads + g0) & 15) == 0)) {
            const uint4 *p = reinterpret_cast<const uint4 *>(reads + g0);
            const uint4 v0 = __ldg(p), v1 = __ldg(p + 1);
            const uint32_t x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int q = 0; q < 8; q++) {
                const uint32_t c0 = (x[q] >> 1) & 0x01010101u, c1 = (x[q] >> 2) & 0x01010101u;
                // expected byte for each 2-bit code: A 0x41, C 0x43, T 0x54, G 0x47
                const uint32_t e = 0x41414141u + (c0 << 1) + c1 * 0x13u - (c0 & c1) * 0x0Fu;
                uint32_t d = e ^ x[q];
                d |= d >> 4; d |= d >> 2; d |= d >> 1; d &= 0x01010101u;                // any bit differs, per byte
                bad |= ((d | (d >> 7) | (d >> 14) | (d >> 21)) & 0xFu) << (4 * q);
                uint32_t dn = x[q] ^ 0x4E4E4E4Eu;                                       // zero byte where the byte is 'N'
                dn |= dn >> 4; dn |= dn >> 2; dn |= dn >> 1; dn = ~dn & 0x01010101u;
                isn |= ((dn | (dn >> 7) | (dn >> 14) | (dn >> 21)) & 0xFu) << (4 * q);
                const uint32_t c = (x[q] >> 1) & 0x03030303u;
                const uint32_t b8 = (c | (c >> 6) | (c >> 12) | (c >> 18)) & 0xFFu;
                bits |= (uint64_t)b8 << (8 * q);
            }
        } else {
            for (uint32_t i = 0; i < count; i++) {
                const uint8_t b = reads[g0 + i];
                const uint32_t c = (b >> 1) & 3u;
                if (b != (uint8_t)"ACTG"[c]) bad |= 1u << i;
                if (b == 'N') isn |= 1u << i;
                bits |= (uint64_t)c << (2 * i);
            }
        }
        packed[w] = bits;
        nmask[w] = isn;
        while (bad) {
            const int i = __ffs(bad) - 1;
            bad &= bad - 1;
            const uint32_t r = find_read(offsets, n_reads, pbase0 + g0 + i);
            if ((isn >> i) & 1u) has_n[r] = 1;
            else needs_bytes[r] = 1;
        }
    }
}

// 32-bit batch-relative offsets (what the host ingest ships) -> the 64-bit absolute offsets the kernels index with
__global__ void expand_offsets_kernel(const uint32_t *__restrict__ rel, uint64_t base, uint64_t n, uint64_t *__restrict__ out)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = base + rel[i];
}

// offsets of a batch of equal-length reads, generated instead of shipped: out[i] = base + i * len
__global__ void fill_offsets_kernel(uint64_t base, uint64_t len, uint64_t n, uint64_t *__restrict__ out)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) out[i] = base + i * len;
}

// ------------------------------------------------------------------------------------------------
// prefilter_kernel
// ------------------------------------------------------------------------------------------------
constexpr uint64_t kLo = 0x5555555555555555ULL;

// per-base predicates on a word of 32 2-bit codes (result bit at the even bit of each field)
struct Planes {
    uint64_t isA, isC, isT, isG;
    __device__ __forceinline__ explicit Planes(uint64_t x)
    {
        const uint64_t lo = x & kLo, hi = (x >> 1) & kLo;
        isA = ~(hi | lo) & kLo;
        isC = ~hi & lo;
        isT = hi & ~lo;
        isG = hi & lo;
    }
};

__global__ void __launch_bounds__(256)
prefilter_kernel(const uint64_t *__restrict__ packed, const uint64_t *__restrict__ offsets, uint64_t base0,
                 uint64_t n_reads, const uint8_t *__restrict__ needs_bytes, uint32_t k, uint32_t fast_max_len,
                 uint32_t thread_max_len, uint32_t thread_short_len, uint32_t task_capacity, uint32_t max_len,
                 uint32_t bytes_max_len, orph_read_result *__restrict__ out,
                 uint32_t *__restrict__ queue, uint32_t *__restrict__ task_queue, uint32_t *__restrict__ bytes_queue,
                 ScoreCounters *__restrict__ counters, const uint32_t *__restrict__ nmask, const uint8_t *__restrict__ has_n,
                 uint32_t *__restrict__ task_queue_n)
{
    const int lane = threadIdx.x & 31;
    __shared__ uint32_t s_hist[2][kLenBuckets];
    for (uint32_t i = threadIdx.x; i < 2 * kLenBuckets; i += blockDim.x) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const uint64_t first = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) & ~31ull;  // warp-uniform loop bound
    for (uint64_t r0 = first; r0 < n_reads; r0 += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t r = r0 + lane;
        const bool in_range = r < n_reads;
        bool to_queue = false, to_bytes = false, to_long = false, to_n = false;
        uint32_t open_bits = 0, n_my_tasks = 0, n_my_tasks_long = 0;
        if (in_range) {
            const uint64_t o0 = offsets[r], o1 = offsets[r + 1];
            const uint64_t L = o1 - o0;
            __align__(16) orph_read_result rec;
#pragma unroll
            for (int f = 0; f < 6; f++) { rec.n_kmers[f] = 0; rec.n_hits[f] = 0; rec.category[f] = ORPH_CAT_INVALID; }
            rec.best_frame = 0;
            rec.flags = 0;
            if (L == 0) {
                atomicOr(&counters->status, kStatusEmptyRead);
            } else if (L > max_len && (max_len < ORPH_MAX_READ_LEN || L > 0xFFFFFFF0ull)) {
                atomicOr(&counters->status, kStatusTooLong);  // longer than the caller declared
            } else if (L > bytes_max_len) {
                // (beyond ORPH_MAX_READ_LEN too: score_long_kernel classifies such a read from its stop codons alone)
                to_long = true;
                rec.flags = ORPH_FLAG_BYTE_PATH;
            } else if ((needs_bytes && needs_bytes[r]) || L > fast_max_len || (has_n && has_n[r] && (L > thread_max_len || !task_queue_n))) {
                to_bytes = true;
                rec.flags = ORPH_FLAG_BYTE_PATH;
            } else if (has_n && has_n[r]) {
                to_n = true;  // prefilter_n_kernel classifies it (and writes its record)
            } else {
                // the read as aligned 32-bit words: A[j] = bases [16j, 16j+16), one funnel shift each
                const uint64_t bitpos = 2 * (o0 - base0);
                const uint32_t *src = reinterpret_cast<const uint32_t *>(packed) + (bitpos >> 5);
                const uint32_t sh = (uint32_t)(bitpos & 31);
                const uint32_t n_src = (sh + 2 * (uint32_t)L + 31) / 32;  // source words holding bases of this read
                const uint32_t n_chunks = (uint32_t)((L + 31) / 32);      // 64-bit chunks of 32 bases
                uint32_t s_prev = src[0];
                uint32_t next_i = 1;
                auto next_aligned = [&]() {
                    const uint32_t s_next = next_i < n_src ? src[next_i] : 0u;
                    const uint32_t a = next_i <= n_src ? __funnelshift_r(s_prev, s_next, sh) : 0u;
                    s_prev = s_next;
                    next_i++;
                    return a;
                };
                uint32_t a0 = next_aligned(), a1 = next_aligned();
                uint32_t ndiff = 0;
                // stop hits by RELATIVE class: before chunk m, rel[c] collects positions p with p mod 3 == (c + 2m) mod 3
                uint64_t fr0 = 0, fr1 = 0, fr2 = 0, rr0 = 0, rr1 = 0, rr2 = 0;
                constexpr uint64_t kPh0 = 0x1041041041041041ULL;  // fields t = 0, 3, 6, ... of a chunk
                for (uint32_t m = 0; m < n_chunks; m++) {
                    const uint32_t a2 = next_aligned(), a3 = next_aligned();
                    uint64_t cur = (uint64_t)a0 | ((uint64_t)a1 << 32);
                    uint64_t nxt = (uint64_t)a2 | ((uint64_t)a3 << 32);
                    a0 = a2;
                    a1 = a3;
                    const uint32_t remaining = (uint32_t)L - 32u * m;  // bases from this chunk's start
                    uint64_t m2 = kLo, m3 = kLo;                       // positions valid for a pair / a codon
                    if (remaining < 66) {                              // only the last two chunks need masks
                        if (remaining < 32) cur &= (1ull << (2 * remaining)) - 1;
                        nxt = remaining > 32 ? (remaining < 64 ? nxt & ((1ull << (2 * (remaining - 32))) - 1) : nxt) : 0ull;
                        const uint32_t v2 = remaining > 1 ? (remaining - 1 < 32 ? remaining - 1 : 32) : 0;
                        const uint32_t v3 = remaining > 2 ? (remaining - 2 < 32 ? remaining - 2 : 32) : 0;
                        m2 = v2 >= 32 ? kLo : (((1ull << (2 * v2)) - 1) & kLo);
                        m3 = v3 >= 32 ? kLo : (((1ull << (2 * v3)) - 1) & kLo);
                    }
                    const uint64_t x1 = (cur >> 2) | (nxt << 62);
                    const uint64_t x2 = (cur >> 4) | (nxt << 60);
                    const uint64_t d = cur ^ x1;
                    ndiff += __popcll((d | (d >> 1)) & m2);
                    const Planes b0(cur), b1(x1), b2(x2);
                    // forward stops TAA TAG TGA
                    const uint64_t fs = b0.isT & ((b1.isA & (b2.isA | b2.isG)) | (b1.isG & b2.isA)) & m3;
                    // triples whose reverse complement is a stop: TTA CTA TCA
                    const uint64_t rs = b2.isA & ((b1.isT & (b0.isT | b0.isC)) | (b1.isC & b0.isT)) & m3;
                    // field t of chunk m is position 32m + t, and 32 = 2 (mod 3): rotate the accumulators instead
                    // of computing the class of every field
                    const uint64_t f0 = fr0 | (fs & kPh0), f1 = fr1 | (fs & (kPh0 << 2)), f2 = fr2 | (fs & (kPh0 << 4));
                    const uint64_t r0 = rr0 | (rs & kPh0), r1 = rr1 | (rs & (kPh0 << 2)), r2 = rr2 | (rs & (kPh0 << 4));
                    fr0 = f2; fr1 = f0; fr2 = f1;
                    rr0 = r2; rr1 = r0; rr2 = r1;
                }
                // after M chunks rel[c] holds class (c + 2M) mod 3, i.e. class k sits at index (k + M) mod 3
                uint64_t fwd[3], rev[3];
                {
                    const uint32_t rot = n_chunks % 3;
                    fwd[0] = rot == 0 ? fr0 : rot == 1 ? fr1 : fr2;
                    fwd[1] = rot == 0 ? fr1 : rot == 1 ? fr2 : fr0;
                    fwd[2] = rot == 0 ? fr2 : rot == 1 ? fr0 : fr1;
                    rev[0] = rot == 0 ? rr0 : rot == 1 ? rr1 : rr2;
                    rev[1] = rot == 0 ? rr1 : rot == 1 ? rr2 : rr0;
                    rev[2] = rot == 0 ? rr2 : rot == 1 ? rr0 : rr1;
                }
                // evaluate_is_fastp_low_complexity: ndiff / len < 0.3  (translate.py:75-84).  In integers: double(0.3) lies
                // 1.1e-17 below 3/10 and the quotients of integers below 2^18 nearest to 3/10 are 1/(10 L) > 3.8e-7 away, so
                // the IEEE compare is true exactly when 10 ndiff < 3 L (equality gives fl(3/10) < fl(3/10): false both ways)
                if (10ull * ndiff < 3ull * L) {
                    rec.flags = ORPH_FLAG_LOW_COMPLEXITY;
#pragma unroll
                    for (int f = 0; f < 6; f++) rec.category[f] = ORPH_CAT_LOW_COMPLEXITY_NUCLEOTIDE;
                } else {
                    uint32_t open = 0;
#pragma unroll
                    for (int f = 0; f < 3; f++) {
                        const uint32_t plen = (uint32_t)((L - f) / 3);  // L >= 1; (L - f) may wrap for L < f: guard
                        const uint32_t plen_ok = L >= (uint64_t)f ? plen : 0;
                        // forward frame f: codons at positions p = f (mod 3)
                        if (fwd[f]) rec.category[f] = ORPH_CAT_STOP_CODONS;
                        else if (plen_ok <= k) rec.category[f] = ORPH_CAT_TOO_SHORT_PEPTIDE;
                        else open |= 1u << f;
                        // reverse frame f: codons at forward positions p = (L - f) (mod 3)
                        const uint32_t cls = (uint32_t)((L + 3 - f) % 3);
                        const uint64_t rev_hits = cls == 0 ? rev[0] : cls == 1 ? rev[1] : rev[2];
                        if (rev_hits) rec.category[3 + f] = ORPH_CAT_STOP_CODONS;
                        else if (plen_ok <= k) rec.category[3 + f] = ORPH_CAT_TOO_SHORT_PEPTIDE;
                        else open |= 8u << f;
                    }
                    rec.flags = (uint8_t)open;
                    open_bits = open;
                    if (L <= thread_max_len) {  // one thread per open frame
                        const int tier = L <= thread_short_len ? 0 : 1;
                        if (tier == 0) n_my_tasks = __popc(open);
                        else n_my_tasks_long = __popc(open);
                        for (uint32_t bits = open; bits; bits &= bits - 1) {  // peptide-length histogram for the task sort
                            const uint32_t f = __ffs(bits) - 1;
                            atomicAdd(&s_hist[tier][min((uint32_t)((L - (f >= 3 ? f - 3 : f)) / 3), kLenBuckets - 1)], 1u);
                        }
                    } else {
                        to_queue = open != 0;  // one warp per read
                    }
                }
            }
            // one 32-byte record per read
            if (!to_n) {
                uint4 *dst = reinterpret_cast<uint4 *>(out + r);
                const uint4 *src = reinterpret_cast<const uint4 *>(&rec);
                dst[0] = src[0];
                dst[1] = src[1];
            }
        }
        // warp-aggregated queue pushes
        const uint32_t mq = __ballot_sync(0xffffffffu, to_queue);
        const uint32_t mb = __ballot_sync(0xffffffffu, to_bytes);
        uint32_t base_q = 0, base_b = 0;
        if (lane == 0) {
            if (mq) base_q = atomicAdd(&counters->n_queue, __popc(mq));
            if (mb) base_b = atomicAdd(&counters->n_bytes, __popc(mb));
        }
        base_q = __shfl_sync(0xffffffffu, base_q, 0);
        base_b = __shfl_sync(0xffffffffu, base_b, 0);
        const uint32_t below = (1u << lane) - 1;
        if (to_queue) queue[base_q + __popc(mq & below)] = ((uint32_t)r << 6) | open_bits;
        if (to_bytes) bytes_queue[base_b + __popc(mb & below)] = (uint32_t)r;
        if (to_long)  // rare: the byte queue is filled from the back, one atomic per read
            bytes_queue[(uint32_t)n_reads - 1 - atomicAdd(&counters->n_long, 1u)] = (uint32_t)r;
        if (to_n)     // reads with N: listed behind the 6 n task slots of their queue
            task_queue_n[task_capacity + atomicAdd(&counters->n_reads_n, 1u)] = (uint32_t)r;
        // per-frame tasks: warp-wide scan of the per-read task counts, one atomic per warp and tier
#pragma unroll
        for (int tier = 0; tier < 2; tier++) {
            const uint32_t mine = tier == 0 ? n_my_tasks : n_my_tasks_long;
            uint32_t incl = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += up;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
            if (total) {
                uint32_t base_t = 0;
                if (lane == 0) base_t = atomicAdd(tier == 0 ? &counters->n_tasks : &counters->n_tasks_long, total);
                base_t = __shfl_sync(0xffffffffu, base_t, 0) + incl - mine;
                uint32_t bits = mine ? open_bits : 0u;
                while (bits) {
                    const uint32_t f = __ffs(bits) - 1;
                    bits &= bits - 1;
                    // the short tier grows from the front of the queue, the long tier from the back
                    task_queue[tier == 0 ? base_t : task_capacity - 1 - base_t] = ((uint32_t)r << 3) | f;
                    base_t++;
                }
            }
        }
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < 2 * kLenBuckets; i += blockDim.x) {
        const uint32_t v = (&s_hist[0][0])[i];
        if (v) atomicAdd(&counters->hist[0][0] + i, v);
    }
}

// ------------------------------------------------------------------------------------------------
// prefilter_n_kernel: one WARP per read that holds N (and no other unusual byte, <= thread_max_len bases): the tests of
// prefilter_kernel with the reference's string semantics for N -- N differs from every base and equals N in the complexity
// count (translate.py:79-84); a codon that holds an N is never a stop codon -- on the 2-bit words and the N plane.  Writes
// the read's record and queues its open frames for score_tasks_kernel<K, true>.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prefilter_n_kernel(const uint64_t *__restrict__ packed, const uint32_t *__restrict__ nmask, const uint64_t *__restrict__ offsets, uint64_t base0,
                   uint32_t k, uint32_t task_capacity, uint32_t *__restrict__ task_queue_n, ScoreCounters *__restrict__ counters,
                   orph_read_result *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const uint32_t n = counters->n_reads_n;
    const uint32_t *p32 = reinterpret_cast<const uint32_t *>(packed);
    for (uint32_t item = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; item < n; item += (gridDim.x * blockDim.x) >> 5) {
        const uint64_t r = task_queue_n[task_capacity + item];
        const uint64_t o0 = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - o0);
        const uint64_t g0 = o0 - base0;
        auto code = [&](uint32_t i) { const uint64_t b = 2 * (g0 + i); return (p32[b >> 5] >> (b & 31)) & 3u; };
        auto is_n = [&](uint32_t i) { const uint64_t g = g0 + i; return (nmask[g >> 5] >> (g & 31)) & 1u; };
        uint32_t ndiff = 0, stops = 0;
        for (uint32_t p = lane; p + 1 < L; p += 32) {
            const uint32_t c0 = code(p), m0 = is_n(p), c1 = code(p + 1), m1 = is_n(p + 1);
            ndiff += (m0 != m1) || (!m0 && c0 != c1);
            if (p + 2 < L) {
                const uint32_t c2 = code(p + 2), m2 = is_n(p + 2);
                if (!(m0 | m1 | m2)) {
                    // codes A0 C1 T2 G3: forward stops TAA TAG TGA; triples whose reverse complement is one: TTA CTA TCA
                    if (c0 == 2 && ((c1 == 0 && (c2 == 0 || c2 == 3)) || (c1 == 3 && c2 == 0))) stops |= 1u << (p % 3);
                    if (c2 == 0 && ((c1 == 2 && (c0 == 2 || c0 == 1)) || (c1 == 1 && c0 == 2))) stops |= 8u << ((L - p) % 3);
                }
            }
        }
        ndiff = __reduce_add_sync(0xffffffffu, ndiff);
        stops = __reduce_or_sync(0xffffffffu, stops);
        if (lane == 0) {
            __align__(16) orph_read_result rec;
#pragma unroll
            for (int f = 0; f < 6; f++) { rec.n_kmers[f] = 0; rec.n_hits[f] = 0; rec.category[f] = ORPH_CAT_INVALID; }
            rec.best_frame = 0;
            rec.flags = 0;
            if (10ull * ndiff < 3ull * L) {  // (see prefilter_kernel for the integer form of ndiff / L < 0.3)
                rec.flags = ORPH_FLAG_LOW_COMPLEXITY;
#pragma unroll
                for (int f = 0; f < 6; f++) rec.category[f] = ORPH_CAT_LOW_COMPLEXITY_NUCLEOTIDE;
            } else {
                uint32_t open = 0;
#pragma unroll
                for (int f = 0; f < 6; f++) {
                    const uint32_t phase = f % 3;
                    const uint32_t plen = L >= phase ? (L - phase) / 3 : 0;
                    if ((stops >> f) & 1u) rec.category[f] = ORPH_CAT_STOP_CODONS;
                    else if (plen <= k) rec.category[f] = ORPH_CAT_TOO_SHORT_PEPTIDE;
                    else open |= 1u << f;
                }
                rec.flags = (uint8_t)open;
                if (open) {
                    uint32_t at = atomicAdd(&counters->n_tasks_n, __popc(open));
                    for (uint32_t bits = open; bits; bits &= bits - 1) task_queue_n[at++] = ((uint32_t)r << 3) | (__ffs(bits) - 1);
                }
            }
            uint4 *dst = reinterpret_cast<uint4 *>(out + r);
            const uint4 *src = reinterpret_cast<const uint4 *>(&rec);
            dst[0] = src[0];
            dst[1] = src[1];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Counting sort of the thread-per-frame tasks by peptide length (longest first), so that the 32 frames a
// warp walks in lock step have (nearly) the same number of residues.  Skipped when a tier's lengths span
// at most three adjacent values (fixed-length reads), where the prefilter's read order is kept.
// The sort is LOCAL: every chunk of ORPH_SORT_CHUNK tasks keeps its place in the queue and is ordered inside
// it.  A global sort makes the warps just as uniform but hands each of them 32 reads from all over the
// packed stream -- offsets, strand words and result records of neighbouring lanes then share nothing, and
// the task kernel ran mixed-length batches 7 % slower (profiles/r02b_kernel_ab_task_sort_locality.log).
// ------------------------------------------------------------------------------------------------
#ifndef ORPH_SORT_LOCAL
#define ORPH_SORT_LOCAL 1   // 1: tasks are ordered by frame length inside chunks of ORPH_SORT_CHUNK instead of globally (A/B switch)
#endif
#ifndef ORPH_SORT_CHUNK
#define ORPH_SORT_CHUNK (ORPH_SORT_LOCAL ? 16384 : 2048)
#endif
constexpr int kScatterBlock = ORPH_SORT_LOCAL ? 1024 : 256;  // threads per block of task_scatter_kernel

__global__ void task_scan_kernel(ScoreCounters *__restrict__ counters)
{
    const int tier = threadIdx.x >> 5, lane = threadIdx.x & 31;  // 64 threads: one warp per tier
    if (tier >= 2) return;
    uint32_t lo = kLenBuckets, hi = 0;
    for (uint32_t b = lane; b < kLenBuckets; b += 32)
        if (counters->hist[tier][b]) { lo = min(lo, b); hi = max(hi, b); }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    if (lane == 0) {
        uint32_t run = 0;
        for (int b = (int)kLenBuckets - 1; b >= 0; b--) {  // descending: the longest frames go first
            counters->cursor[tier][b] = run;
            run += counters->hist[tier][b];
        }
        counters->sort_tier[tier] = (lo < kLenBuckets && hi - lo >= 3) ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(kScatterBlock)
task_scatter_kernel(const uint64_t *__restrict__ offsets, const uint32_t *__restrict__ tasks, uint32_t *__restrict__ sorted,
                    uint32_t task_capacity, ScoreCounters *__restrict__ counters)
{
    // Each block sorts chunks of kChunk tasks: shared-memory histogram, output ranges by length from a scan of it (local
    // order) or from ONE global atomic per (chunk, length) (global order), shared-memory cursors inside them.
    constexpr uint32_t kChunk = ORPH_SORT_CHUNK;
    __shared__ uint32_t s_count[kLenBuckets], s_base[kLenBuckets];
#pragma unroll
    for (int tier = 0; tier < 2; tier++) {
        if (!counters->sort_tier[tier]) continue;
        const uint32_t n = tier == 0 ? counters->n_tasks : counters->n_tasks_long;
        auto plen_of = [&](uint32_t entry) {
            const uint32_t r = entry >> 3, f = entry & 7u;
            const uint32_t L = (uint32_t)(offsets[r + 1] - offsets[r]);
            return min((L - (f >= 3 ? f - 3 : f)) / 3, kLenBuckets - 1);
        };
        for (uint32_t c0 = blockIdx.x * kChunk; c0 < n; c0 += gridDim.x * kChunk) {
            for (uint32_t b = threadIdx.x; b < kLenBuckets; b += blockDim.x) s_count[b] = 0;
            __syncthreads();
            const uint32_t c1 = min(n, c0 + kChunk);
            for (uint32_t i = c0 + threadIdx.x; i < c1; i += blockDim.x)
                atomicAdd(&s_count[plen_of(tasks[tier == 0 ? i : task_capacity - 1 - i])], 1u);
            __syncthreads();
#if ORPH_SORT_LOCAL
            // the chunk keeps its place in the queue and is ordered inside it (longest frames first): the 32 tasks of a warp
            // are still frames of similar length, and they are reads from one neighbourhood of the packed stream
            if (threadIdx.x < 32) {  // exclusive scan over the buckets, longest first: four buckets per lane of one warp
                static_assert(kLenBuckets == 128, "four buckets per lane");
                const uint32_t lane = threadIdx.x;
                uint32_t cnt[4], mine = 0;
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    cnt[j] = s_count[4 * lane + j];
                    mine += cnt[j];
                }
                uint32_t from_here = mine;  // tasks in the buckets of this lane and of the lanes above it
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t up = __shfl_down_sync(0xffffffffu, from_here, o);
                    if (lane + o < 32) from_here += up;
                }
                uint32_t run = c0 + from_here - mine;
#pragma unroll
                for (int j = 3; j >= 0; j--) {
                    s_base[4 * lane + j] = run;
                    run += cnt[j];
                    s_count[4 * lane + j] = 0;
                }
            }
#else
            for (uint32_t b = threadIdx.x; b < kLenBuckets; b += blockDim.x) {
                const uint32_t cnt = s_count[b];
                s_base[b] = cnt ? atomicAdd(&counters->cursor[tier][b], cnt) : 0u;
                s_count[b] = 0;
            }
#endif
            __syncthreads();
            for (uint32_t i = c0 + threadIdx.x; i < c1; i += blockDim.x) {
                const uint32_t entry = tasks[tier == 0 ? i : task_capacity - 1 - i];
                const uint32_t b = plen_of(entry);
                const uint32_t pos = s_base[b] + atomicAdd(&s_count[b], 1u);
                sorted[tier == 0 ? pos : task_capacity - 1 - pos] = entry;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// score_frames_kernel: persistent warps, dynamic fetch from the queue written by prefilter_kernel
// ------------------------------------------------------------------------------------------------
constexpr int kFetchBatch = 16;          // queue entries a warp claims per atomic
constexpr uint32_t kFastMaxKmers = kFastMaxLen / 3;  // 512: bound of the per-block min-hits table

template <int K>
__global__ void __launch_bounds__(256)
score_frames_kernel(const ScoreParams P, const uint32_t *__restrict__ packed32, const uint64_t *__restrict__ offsets,
                    uint64_t base0, const uint32_t *__restrict__ queue, ScoreCounters *__restrict__ counters,
                    orph_read_result *__restrict__ out)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ FilterDev sF;
    __shared__ uint8_t lut_fwd[64], lut_rev[64];
    __shared__ uint16_t s_min_hits[kFastMaxKmers + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    stage_filter(&sF, P.F);
    if (threadIdx.x < 64) {
        const int t = threadIdx.x, b0 = t & 3, b1 = (t >> 2) & 3, b2 = t >> 4;
        lut_fwd[t] = P.lut[translate_codes(b0, b1, b2)];
        lut_rev[t] = P.lut[translate_codes(b2 ^ 2, b1 ^ 2, b0 ^ 2)];  // complement = code ^ 2, order reversed
    }
    const double thr = P.jaccard_threshold;
    for (uint32_t n = threadIdx.x; n <= P.fast_max_len / 3; n += blockDim.x)
        s_min_hits[n] = n ? (uint16_t)min_hits_for(n, thr) : (uint16_t)1;
    __syncthreads();
    // per-warp carve-out: [dedup slots | read words | peptide]
    const uint32_t per_warp = P.slots * 4 + P.read_words * 4 + P.pep_bytes;
    uint8_t *mine = smem + (size_t)warp * ((per_warp + 15) & ~15u);
    uint32_t *tab = reinterpret_cast<uint32_t *>(mine);
    uint32_t *rd = reinterpret_cast<uint32_t *>(mine + P.slots * 4);
    uint8_t *pep = mine + P.slots * 4 + P.read_words * 4;
    const uint32_t k = K > 0 ? (uint32_t)K : P.k;
    const bool force_exact = P.force_exact != 0;
    const uint32_t n_queue = counters->n_queue;

    while (true) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&counters->next_queue, (uint32_t)kFetchBatch);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_queue) break;
        const uint32_t cnt = min((uint32_t)kFetchBatch, n_queue - base);
        // lanes < cnt fetch one entry each together with its read's extent, so the latencies overlap
        uint32_t my_entry = 0, my_len = 0;
        uint64_t my_o0 = 0;
        if (lane < cnt) {
            my_entry = queue[base + lane];
            const uint32_t r = my_entry >> 6;
            my_o0 = offsets[r];
            my_len = (uint32_t)(offsets[r + 1] - my_o0);
        }
        for (uint32_t it = 0; it < cnt; it++) {
            const uint32_t entry = __shfl_sync(0xffffffffu, my_entry, it);
            const uint64_t o0 = __shfl_sync(0xffffffffu, my_o0, it);
            const uint32_t L = __shfl_sync(0xffffffffu, my_len, it);
            const uint32_t r = entry >> 6;
            uint32_t open = entry & 63u;
            const uint64_t bitpos = 2 * (o0 - base0);
            const uint64_t w0 = bitpos >> 5;
            const uint32_t bit0 = (uint32_t)(bitpos & 31);
            const uint32_t n_words = (bit0 + 2 * L + 31) / 32;
            __syncwarp();
            for (uint32_t t = lane; t <= n_words; t += kWarp) rd[t] = t < n_words ? __ldg(packed32 + w0 + t) : 0u;
            __syncwarp();
            BestFrame best;
            while (open) {
                const int f = __ffs(open) - 1;
                open &= open - 1;
                const uint32_t phase = f >= 3 ? f - 3 : f;
                const uint32_t plen = (L - phase) / 3;
                // translate + re-encode this frame only
                const uint8_t *lut = f < 3 ? lut_fwd : lut_rev;
                const int p0 = f < 3 ? (int)phase : (int)(L - 3 - phase), dp = f < 3 ? 3 : -3;
                for (uint32_t i = lane; i < plen; i += kWarp) {
                    const uint32_t bit = bit0 + 2 * (uint32_t)(p0 + dp * (int)i);
                    const uint32_t c = __funnelshift_r(rd[bit >> 5], rd[(bit >> 5) + 1], bit & 31) & 63u;
                    pep[i] = lut[c];
                }
                __syncwarp();
                uint32_t nd, nh;
                score_frame<K>(pep, plen, k, sF, tab, P.slots, lane, force_exact, nd, nh);
                // evaluate_is_kmer_low_complexity: n_kmers <= (len - k + 1) / 2  (translate.py:98-100), in integers
                const uint8_t cat = 2 * nd <= plen - k + 1 ? ORPH_CAT_LOW_COMPLEXITY_PEPTIDE
                                    : (nh >= s_min_hits[nd] ? ORPH_CAT_CODING : ORPH_CAT_NON_CODING);
                if (lane == 0) {
                    out[r].n_kmers[f] = (uint16_t)nd;
                    out[r].n_hits[f] = (uint16_t)nh;
                    out[r].category[f] = cat;
                }
                best.offer(nh, nd, cat, frame_key(f));
                __syncwarp();
            }
            if (lane == 0) out[r].best_frame = (int8_t)best.frame;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// score_bytes_kernel: byte-exact path, one warp per read taken from `list` (or all reads if list is null)
// ------------------------------------------------------------------------------------------------
// nmask (2-bit source only, may be NULL): the N plane of the stream -- a set bit reads back as 'N'
template <bool PACKED_SRC>
__device__ __forceinline__ uint8_t read_base(const void *src, uint64_t base0, uint64_t g, const uint32_t *nmask = nullptr)
{
    if (PACKED_SRC) {
        const uint64_t *w = static_cast<const uint64_t *>(src);
        const uint64_t rel = g - base0;
        if (nmask && ((nmask[rel >> 5] >> (rel & 31)) & 1u)) return (uint8_t)'N';
        return (uint8_t)"ACTG"[(w[rel >> 5] >> (2 * (rel & 31))) & 3];
    }
    return static_cast<const uint8_t *>(src)[g - base0];
}

template <bool PACKED_SRC>
__global__ void __launch_bounds__(256)
score_bytes_kernel(const ScoreParams P, const void *__restrict__ src, const uint64_t *__restrict__ offsets,
                   uint64_t base0, uint64_t n_reads, const uint32_t *__restrict__ list,
                   ScoreCounters *__restrict__ counters, orph_read_result *__restrict__ out, const uint32_t *__restrict__ nmask)
{
    extern __shared__ __align__(16) uint8_t smem[];
    __shared__ FilterDev sF;
    __shared__ uint8_t s_lut[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    stage_filter(&sF, P.F);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = P.lut[i];
    __syncthreads();
    // per-warp carve-out: [dedup slots | peptide | read bytes]
    const uint32_t rb_bytes = (P.max_len + 7) & ~3u;
    const uint32_t per_warp = P.slots * 4 + P.pep_bytes + rb_bytes;
    uint8_t *mine = smem + (size_t)warp * ((per_warp + 15) & ~15u);
    uint32_t *tab = reinterpret_cast<uint32_t *>(mine);
    uint8_t *pep = mine + P.slots * 4;
    uint8_t *rb = pep + P.pep_bytes;  // the read's bytes, staged once: every pass below works from shared memory
    const uint32_t k = P.k;

    while (true) {
        uint32_t item = 0;
        if (lane == 0) item = atomicAdd(&counters->next_bytes, 1u);
        item = __shfl_sync(0xffffffffu, item, 0);
        const uint64_t limit = list ? (uint64_t)counters->n_bytes : n_reads;
        if (item >= limit) break;
        const uint64_t r = list ? list[item] : item;
        const uint64_t o0 = offsets[r];
        const uint64_t L64 = offsets[r + 1] - o0;
        uint8_t cat[6];
        uint32_t nk[6] = {0, 0, 0, 0, 0, 0}, nh[6] = {0, 0, 0, 0, 0, 0};
        uint8_t flags = ORPH_FLAG_BYTE_PATH;
        BestFrame best;
        if (L64 == 0 || L64 > P.max_len) {
            if (lane == 0) atomicOr(&counters->status, L64 == 0 ? kStatusEmptyRead : kStatusTooLong);
#pragma unroll
            for (int f = 0; f < 6; f++) cat[f] = ORPH_CAT_INVALID;
        } else {
            const uint32_t L = (uint32_t)L64;
            __syncwarp();
            for (uint32_t i = lane; i < L; i += kWarp) rb[i] = read_base<PACKED_SRC>(src, base0, o0 + i, nmask);
            __syncwarp();
            // compute_fastp_complexity (translate.py:79-84): case-sensitive byte compare
            uint32_t ndiff = 0;
            for (uint32_t i = lane; i + 1 < L; i += kWarp)
                ndiff += rb[i] != rb[i + 1];
            ndiff = __reduce_add_sync(0xffffffffu, ndiff);
            if ((double)ndiff / (double)L < 0.3) {
                flags |= ORPH_FLAG_LOW_COMPLEXITY;
#pragma unroll
                for (int f = 0; f < 6; f++) cat[f] = ORPH_CAT_LOW_COMPLEXITY_NUCLEOTIDE;
            } else {
                // which frames contain a stop codon
                uint32_t stops = 0;
                for (uint32_t p = lane; p + 3 <= L; p += kWarp) {
                    const uint8_t a = rb[p], b = rb[p + 1],
                                  c = rb[p + 2];
                    if (translate_bytes(a, b, c) == '*') stops |= 1u << (p % 3);
                    if (translate_bytes(complement_byte(c), complement_byte(b), complement_byte(a)) == '*')
                        stops |= 8u << ((L - p) % 3);
                }
                stops = __reduce_or_sync(0xffffffffu, stops);
                uint32_t open = 0;
                for (int f = 0; f < 6; f++) {
                    const uint32_t phase = f % 3;
                    const uint32_t plen = L >= phase ? (L - phase) / 3 : 0;
                    if ((stops >> f) & 1u) { cat[f] = ORPH_CAT_STOP_CODONS; continue; }
                    if (plen <= k) { cat[f] = ORPH_CAT_TOO_SHORT_PEPTIDE; continue; }
                    open |= 1u << f;
                    __syncwarp();
                    for (uint32_t i = lane; i < plen; i += kWarp) {
                        uint8_t aa;
                        if (f < 3) {
                            const uint32_t g = phase + 3 * i;
                            aa = translate_bytes(rb[g], rb[g + 1], rb[g + 2]);
                        } else {
                            const uint32_t g = L - 3 - phase - 3 * i;  // forward position of the codon's last base
                            aa = translate_bytes(complement_byte(rb[g + 2]), complement_byte(rb[g + 1]), complement_byte(rb[g]));
                        }
                        pep[i] = s_lut[aa];
                    }
                    __syncwarp();
                    score_frame<0>(pep, plen, k, sF, tab, P.slots, lane, P.force_exact != 0, nk[f], nh[f]);
                    cat[f] = scored_category(nk[f], nh[f], plen, k, P.jaccard_threshold);
                    best.offer(nh[f], nk[f], cat[f], frame_key(f));
                }
                flags |= (uint8_t)open;
            }
        }
        if (lane == 0) {
            orph_read_result rec;
#pragma unroll
            for (int f = 0; f < 6; f++) { rec.n_kmers[f] = (uint16_t)nk[f]; rec.n_hits[f] = (uint16_t)nh[f]; rec.category[f] = cat[f]; }
            rec.best_frame = (int8_t)best.frame;
            rec.flags = flags;
            out[r] = rec;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// translate_frames_kernel: the peptides the reference EMITS (stdout FASTA of coding frames, the
// low-complexity-peptide FASTA; translate.py:244-249, 271-273): un-encoded amino acids of one frame per
// item, byte-exact codon semantics (translate_single_seq.py:13-42).  One warp per (read, frame) item.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
translate_frames_kernel(const uint8_t *__restrict__ ascii, uint64_t abase0, const uint64_t *__restrict__ offsets,
                        const uint64_t *__restrict__ item_read, const int8_t *__restrict__ item_frame, uint64_t n_items,
                        const uint64_t *__restrict__ out_offsets, uint8_t *__restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t it = warp0; it < n_items; it += n_warps) {
        const uint64_t r = item_read[it];
        const int key = item_frame[it];  // 1,2,3,-1,-2,-3
        const uint64_t o0 = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - o0);
        const uint32_t phase = (uint32_t)(key > 0 ? key - 1 : -key - 1);
        const uint32_t plen = L >= phase ? (L - phase) / 3 : 0;
        const uint8_t *seq = ascii + (o0 - abase0);
        uint8_t *dst = out + out_offsets[it];
        for (uint32_t i = lane; i < plen; i += kWarp) {
            uint8_t aa;
            if (key > 0) {
                const uint32_t p = phase + 3 * i;
                aa = translate_bytes(seq[p], seq[p + 1], seq[p + 2]);
            } else {
                const uint32_t p = L - 3 - phase - 3 * i;  // forward position of the reverse codon's last base
                aa = translate_bytes(complement_byte(seq[p + 2]), complement_byte(seq[p + 1]), complement_byte(seq[p]));
            }
            dst[i] = aa;
        }
    }
}

}  // namespace orph
