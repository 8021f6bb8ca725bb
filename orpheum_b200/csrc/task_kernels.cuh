// score_tasks_kernel: one THREAD per open frame ("task"), for reads up to kThreadMaxLen bases.
//
// This is the hot kernel for short reads.  A thread stages its strand (the read or its reverse complement, 2-bit words)
// in its own shared-memory column, turns it in place into the frame's encoded residues (exactly the bytes sourmash's
// hash_murmur sees after encode_peptide, sequence_encodings.py:457-468), then walks them: the current peptide k-mer is a
// rolling byte window in registers, hashed (translate.py:187) and probed in table 0 (translate.py:190).  Every lane of a
// warp is a different frame, so the k-mer loop runs at full lane occupancy; the probe of k-mer s is consumed one
// iteration later so its L2 latency overlaps the next hash, and a k-mer that passes table 0 has its table 1..3 probes
// issued together and consumed one iteration after that.
//
// Distinct k-mers (kmerize returns a set, compare_kmer_content.py:116-118): each thread keeps a private 1024-bit (reads
// <= 192 nt) or 2048-bit (<= 320 nt) blocked three-hash "seen" filter in shared memory.  A k-mer whose three bits are not
// all set is certainly new; otherwise the frame is walked backwards from the current k-mer and compared byte-exactly (a
// tandem repeat answers after one period, a false alarm walks the whole prefix, rarely), so the de-duplication is exact.
//
// What bounds it (round 2, profiles/README.md): with the table bins from mod_prime_tiny the kernel moves as many 32-byte
// sectors through L2 per second as the random-gather micro-benchmark does (114 per read, 100.5 of them required probes);
// fewer instructions (residue pre-pass, split seen bits) or fewer L1 look-ups (coalesced staging, table 1 first) no longer
// change its time (profiles/r02_kernel_ab_*.log).
#pragma once
#include "device_common.cuh"
#include "score_kernels.cuh"

namespace orph {

constexpr uint32_t kThreadMaxLenShort = 192;  // short tier: 1024-bit seen filter per thread
constexpr uint32_t kThreadMaxLenLong = 320;   // long tier: 2048-bit seen filter per thread (fewer resident blocks)
constexpr uint32_t kThreadMaxLen = kThreadMaxLenLong;
constexpr int kTaskBlock = 256;
constexpr uint32_t kSeenWordsShort = 16, kSeenWordsLong = 32;  // 64-bit words; word w of thread t at [w * kTaskBlock + t]

// rolling window over the last K encoded residues, little-endian bytes as MurmurHash reads them
template <int K>
struct KmerWindow {
    static constexpr int NW = (K + 3) / 4;
    static constexpr uint32_t kTopMask = 0xFFFFFFFFu >> (8 * (3 - ((K - 1) & 3)));
    uint32_t w[NW];
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int i = 0; i < NW; i++) w[i] = 0;
    }
    // append the next residue (drops the oldest)
    __device__ __forceinline__ void push(uint32_t byte)
    {
#pragma unroll
        for (int i = 0; i + 1 < NW; i++) w[i] = __funnelshift_r(w[i], w[i + 1], 8);
        w[NW - 1] = (w[NW - 1] >> 8) | (byte << (8 * ((K - 1) & 3)));
    }
    // prepend the previous residue (drops the newest): the window one position earlier
    __device__ __forceinline__ void push_front(uint32_t byte)
    {
#pragma unroll
        for (int i = NW - 1; i > 0; i--) w[i] = __funnelshift_l(w[i - 1], w[i], 8);
        w[0] = (w[0] << 8) | byte;
        w[NW - 1] &= kTopMask;
    }
    __device__ __forceinline__ bool equals(const KmerWindow &o) const
    {
        uint32_t d = 0;
#pragma unroll
        for (int i = 0; i < NW; i++) d |= w[i] ^ o.w[i];
        return d == 0;
    }
    // MurmurHash3_x64_128(bytes, K, seed 42).h1 with everything resolved at compile time
    __device__ __forceinline__ uint64_t hash() const
    {
        uint64_t h1 = kMurmurSeed, h2 = kMurmurSeed;
        constexpr int NB = K / 16;
#pragma unroll
        for (int b = 0; b < NB; b++) {
            const uint64_t k1 = (uint64_t)w[4 * b] | ((uint64_t)w[4 * b + 1] << 32);
            const uint64_t k2 = (uint64_t)w[4 * b + 2] | ((uint64_t)w[4 * b + 3] << 32);
            h1 ^= mix_k1(k1);
            h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
            h2 ^= mix_k2(k2);
            h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
        }
        constexpr int tail = K & 15, t0 = NB * 4;
        if constexpr (tail > 8) {
            uint64_t k2 = (uint64_t)w[t0 + 2];
            if constexpr (tail > 12) k2 |= (uint64_t)w[t0 + 3] << 32;
            h2 ^= mix_k2(k2);
        }
        if constexpr (tail > 0) {
            uint64_t k1 = (uint64_t)w[t0];
            if constexpr (tail > 4) k1 |= (uint64_t)w[t0 + 1] << 32;
            h1 ^= mix_k1(k1);
        }
        h1 ^= (uint64_t)K; h2 ^= (uint64_t)K;
        h1 += h2; h2 += h1;
        h1 = fmix64(h1); h2 = fmix64(h2);
        return h1 + h2;
    }
};

// One thread's strand, staged in shared memory: the bases of the strand the frame lives on (the read itself
// or its reverse complement), re-aligned so that base 0 sits at bit 0 of word 0.  Word t at col[t * kTaskBlock].
// With the strand materialised every frame is a forward walk: residue i of frame phase ph is the codon at
// bits [2*ph + 6*i, +6), looked up in ONE table (forward codon x alphabet).
// NM: the strand also has an N plane (one bit per base, column `mcol`, same alignment): reads that hold N.  A codon with an N
// is X -- encoded like every other residue, `x_res` -- unless the N is its third base and the first two are one of the seven
// four-fold families the reference's table lists (ACN CTN CGN GTN GCN GGN TCN, constants_translate.py:59-131; CCN is not
// in it), which translate like any member of the family.
constexpr uint32_t kFourFoldWithN = (1u << 4) | (1u << 9) | (1u << 13) | (1u << 11) | (1u << 7) | (1u << 15) | (1u << 6);  // index b0 | b1 << 2

__device__ __forceinline__ uint32_t residue_with_n(const uint8_t *lut, uint32_t codon, uint32_t n3, uint32_t x_res)
{
    if (n3 == 0) return lut[codon];
    if (n3 == 4u && ((kFourFoldWithN >> (codon & 15u)) & 1u)) return lut[codon & 15u];
    return x_res;
}

template <bool NM>
struct StrandColumn {
    const uint32_t *col;
    const uint32_t *mcol;
    uint32_t x_res;
    // 32 bits of the strand starting at `bit`
    __device__ __forceinline__ uint32_t window(uint32_t bit) const
    {
        const uint32_t wi = bit >> 5;
        return __funnelshift_r(col[wi * kTaskBlock], col[(wi + 1) * kTaskBlock], bit & 31u);
    }
    // the N bits of the three bases starting at base `b`
    __device__ __forceinline__ uint32_t n_bits(uint32_t b) const
    {
        const uint32_t wi = b >> 5;
        return __funnelshift_r(mcol[wi * kTaskBlock], mcol[(wi + 1) * kTaskBlock], b & 31u) & 7u;
    }
    // the N bits of the twelve bases (four codons) starting at base `b`
    __device__ __forceinline__ uint32_t n_bits12(uint32_t b) const
    {
        const uint32_t wi = b >> 5;
        return __funnelshift_r(mcol[wi * kTaskBlock], mcol[(wi + 1) * kTaskBlock], b & 31u) & 0xFFFu;
    }
    __device__ __forceinline__ uint32_t residue(const uint8_t *lut, uint32_t phase, uint32_t i) const
    {
        return (col[(i >> 2) * kTaskBlock] >> (8u * (i & 3u))) & 0xFFu;  // the column holds residues by now
    }
    // Turns the staged strand into the frame's encoded residues, in place: word j of the column = residues 4j .. 4j+3
    // (what encode_peptide makes of the translated frame, sequence_encodings.py:457-468).  Residues need 8 bits where
    // their codons need 6, so the walk goes from the last word down: word j is cut from strand words
    // floor((2 phase + 24 j) / 32) and the next one, both <= j for j >= 1 and word 0 alone for j = 0.
    __device__ __forceinline__ void to_residues(uint32_t *column, const uint8_t *lut, uint32_t phase, uint32_t plen, uint32_t max_plen) const
    {
        const uint32_t nw = (plen + 3u) >> 2;
        for (int j = (int)((max_plen + 3u) >> 2) - 1; j >= 0; j--) {
            if ((uint32_t)j < nw) {
                const uint32_t w = window(2u * phase + 24u * (uint32_t)j);
                uint32_t r0, r1, r2, r3;
                if constexpr (NM) {
                    const uint32_t nb = n_bits12(phase + 12u * (uint32_t)j);
                    r0 = residue_with_n(lut, w & 63u, nb & 7u, x_res);
                    r1 = residue_with_n(lut, (w >> 6) & 63u, (nb >> 3) & 7u, x_res);
                    r2 = residue_with_n(lut, (w >> 12) & 63u, (nb >> 6) & 7u, x_res);
                    r3 = residue_with_n(lut, (w >> 18) & 63u, nb >> 9, x_res);
                } else {
                    r0 = lut[w & 63u];
                    r1 = lut[(w >> 6) & 63u];
                    r2 = lut[(w >> 12) & 63u];
                    r3 = lut[(w >> 18) & 63u];
                }
                column[j * kTaskBlock] = r0 | (r1 << 8) | (r2 << 16) | (r3 << 24);
            }
        }
    }
};

// reverse complement of 16 packed bases: reverse the 2-bit fields, complement = code ^ 2
__device__ __forceinline__ uint32_t revcomp_word(uint32_t x)
{
    uint32_t y = __brev(x);
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    return y ^ 0xAAAAAAAAu;
}

// Exact answer behind a seen-filter alarm: does the k-mer ending at residue s (window `cur`) also end
// at an earlier residue of the same frame?  Walks backwards, so a tandem repeat answers after one period.
template <int K, bool NM>
__device__ __forceinline__ bool kmer_seen_before(const StrandColumn<NM> &sc, const uint8_t *lut, uint32_t phase, int s,
                                                 const KmerWindow<K> &cur)
{
    KmerWindow<K> win = cur;
    for (int e = s - 1; e >= K - 1; e--) {
        win.push_front(sc.residue(lut, phase, (uint32_t)(e - K + 1)));
        if (win.equals(cur)) return true;
    }
    return false;
}

// NM = true: the instantiation for reads that hold N (tasks_unsorted = their queue, `nmask` = the N plane of the 2-bit
// stream, long_tier = 2: the long tier's shared-memory geometry plus one N-plane column per thread).
// T4 = true: the filter has four tables of at most kTinyPrimeMax bits each (khmer's default geometry up to --tablesize 1.4e9):
// the table count is a compile-time fact and the bins come from mod_prime_tiny.
template <int K, bool NM, bool T4>
__global__ void __launch_bounds__(kTaskBlock)
score_tasks_kernel(const ScoreParams P, const uint32_t *__restrict__ packed32, const uint64_t *__restrict__ offsets,
                   uint64_t base0, const uint32_t *__restrict__ tasks_unsorted, const uint32_t *__restrict__ tasks_sorted,
                   int long_tier, ScoreCounters *__restrict__ counters, orph_read_result *__restrict__ out,
                   const uint32_t *__restrict__ nmask)
{
    extern __shared__ __align__(16) uint8_t dyn_smem[];  // [seen: P.seen_words u64 x 256][strand words: P.read_words u32 x 256][NM: N-plane words]
    __shared__ FilterDev sF;
    __shared__ uint8_t lut_fwd[64];
    __shared__ uint16_t s_min_hits[kThreadMaxLen / 3 + 2];
    const int lane = threadIdx.x & 31;
    stage_filter(&sF, P.F);
    if (threadIdx.x < 64) {
        const int t = threadIdx.x;
        lut_fwd[t] = P.lut[translate_codes(t & 3, (t >> 2) & 3, t >> 4)];
    }
    const double thr = P.jaccard_threshold;
    for (uint32_t n = threadIdx.x; n <= kThreadMaxLen / 3 + 1; n += blockDim.x)
        s_min_hits[n] = n ? (uint16_t)min_hits_for(n, thr) : (uint16_t)1;
    __syncthreads();
    const bool force_exact = P.force_exact != 0;
    const uint32_t n_tables = P.F.n_tables;
    const bool four_tables = T4 || n_tables == 4;
    const bool small = P.F.small != 0;
    // bin of hash h in table T (static index: the geometry comes straight from the constant bank)
#define ORPH_BIN(T, h, out)                                                                   \
    if constexpr (T4) {                                                                       \
        (out) = mod_prime_tiny((h), (uint32_t)P.F.size[T], P.F.magic[T]);                     \
    } else {                                                                                  \
        const uint64_t q__ = __umul64hi((h), P.F.magic[T]);                                   \
        if (small) {                                                                          \
            const uint32_t p__ = (uint32_t)P.F.size[T];                                       \
            uint32_t r__ = (uint32_t)(h) - (uint32_t)q__ * p__;                               \
            if (r__ >= p__) r__ -= p__;                                                       \
            (out) = r__;                                                                      \
        } else {                                                                              \
            uint64_t r__ = (h) - q__ * P.F.size[T];                                           \
            if (r__ >= P.F.size[T]) r__ -= P.F.size[T];                                       \
            (out) = r__;                                                                      \
        }                                                                                     \
    }
    const uint32_t n_tasks = NM ? counters->n_tasks_n : long_tier ? counters->n_tasks_long : counters->n_tasks;
    const uint32_t *tasks = NM ? tasks_unsorted : counters->sort_tier[long_tier ? 1 : 0] ? tasks_sorted : tasks_unsorted;
    uint32_t *next_task = NM ? &counters->next_task_n : long_tier ? &counters->next_task_long : &counters->next_task;
    const uint32_t seen_words = P.seen_words;
    unsigned long long *seen = reinterpret_cast<unsigned long long *>(dyn_smem) + threadIdx.x;
    uint32_t *strand = reinterpret_cast<uint32_t *>(dyn_smem + (size_t)seen_words * 8 * kTaskBlock) + threadIdx.x;
    const uint32_t strand_words = P.read_words;  // height of a thread's column
    const uint32_t stage_words = P.stage_words;  // of those, the words the 2-bit strand takes
    const uint32_t mask_words = (stage_words + 1) / 2 + 1;  // one bit per base against two (+ the word a 3-bit fetch may touch)
    uint32_t *mstrand = strand + (size_t)strand_words * kTaskBlock;
    const StrandColumn<NM> sc{strand, mstrand, (uint32_t)P.lut['X']};

    while (true) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(next_task, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_tasks) break;
        const bool valid = base + lane < n_tasks;
        uint32_t r = 0, f = 0, L = 0, plen = 0;
        uint64_t o0 = 0;
        if (valid) {
            const uint32_t entry = (long_tier && !NM) ? tasks[P.task_capacity - 1 - (base + lane)] : tasks[base + lane];
            r = entry >> 3;
            f = entry & 7u;
            o0 = offsets[r];
            L = (uint32_t)(offsets[r + 1] - o0);
        }
        const uint32_t phase = f >= 3 ? f - 3 : f;
        if (valid) plen = (L - phase) / 3;
        // ---- stage the strand: aligned copy of the read (frames 1..3) or of its reverse complement (-1..-3)
        {
            const uint64_t bitpos = 2 * (o0 - base0);
            const uint32_t bit0 = (uint32_t)(bitpos & 31);
            const uint32_t *src = packed32 + (bitpos >> 5);
            const uint32_t n_src = valid ? (bit0 + 2 * L + 31) / 32 : 0;  // source words that hold bases of this read
            auto src_word = [&](uint32_t i) { return i < n_src ? __ldg(src + i) : 0u; };
            // 32 bits of the read starting at (possibly negative) read-relative bit `start`
            auto read_bits = [&](int start) {
                const uint32_t a = bit0 + (uint32_t)(start > 0 ? start : 0);
                const uint32_t v = __funnelshift_r(src_word(a >> 5), src_word((a >> 5) + 1), a & 31u);
                return start >= 0 ? v : v << (uint32_t)(-start);
            };
            const uint32_t n_need = valid ? (2 * L + 31) / 32 : 0;
            for (uint32_t t = 0; t < stage_words; t++) {
                uint32_t w = 0;
                if (t < n_need) w = f < 3 ? read_bits((int)(32 * t)) : revcomp_word(read_bits(2 * (int)L - 32 * (int)(t + 1)));
                strand[t * kTaskBlock] = w;
            }
            if constexpr (NM) {
                // the N plane of the strand: bit i = base i of the read, or of its reverse complement (N is its own complement)
                const uint64_t g0 = o0 - base0;
                const uint32_t mb0 = (uint32_t)(g0 & 31);
                const uint32_t *msrc = nmask + (g0 >> 5);
                const uint32_t n_msrc = valid ? (mb0 + L + 31) / 32 : 0;
                auto msrc_word = [&](uint32_t i) { return i < n_msrc ? __ldg(msrc + i) : 0u; };
                auto mask_bits = [&](int start) {  // 32 N bits from (possibly negative) read-relative base `start`; beyond the read: 0
                    const uint32_t a = mb0 + (uint32_t)(start > 0 ? start : 0);
                    uint32_t v = __funnelshift_r(msrc_word(a >> 5), msrc_word((a >> 5) + 1), a & 31u);
                    const int first = start > 0 ? start : 0;
                    const int left = (int)L - first;  // bases of the read from `first` on
                    if (left < 32) v &= left > 0 ? ((1u << left) - 1u) : 0u;
                    return start >= 0 ? v : v << (uint32_t)(-start);
                };
                const uint32_t n_mneed = valid ? (L + 31) / 32 : 0;
                for (uint32_t t = 0; t < mask_words; t++) {
                    uint32_t w = 0;
                    if (t < n_mneed) w = f < 3 ? mask_bits((int)(32 * t)) : __brev(mask_bits((int)L - 32 * (int)(t + 1)));
                    mstrand[t * kTaskBlock] = w;
                }
            }
        }
        KmerWindow<K> win;
        win.clear();
        for (uint32_t w = 0; w < seen_words; w++) seen[w * kTaskBlock] = 0ull;
        uint32_t nd = 0, nh = 0;
        const uint32_t max_plen = __reduce_max_sync(0xffffffffu, plen);
        sc.to_residues(strand, lut_fwd, phase, plen, max_plen);
        __syncwarp();

        // Software pipeline over residues: no result of a shared- or global-memory load is consumed in the
        // iteration that issues it (shared-memory loads queue behind the 32-cycle gathers in the same LSU pipe,
        // so their latency is hundreds of cycles here).  Iteration s, in program order:
        //   (1) seen filter of k-mer s-1: its word has arrived -> store word|bits, decide new / duplicate
        //       (exact backward walk on an alarm; `win` is still the window of k-mer s-1);
        //   (2) consume the table 1..3 probes of k-mer s-2;
        //   (3) table-0 probe of k-mer s-1 has arrived: if it is new and hit, issue its table 1..3 probes;
        //   (4) push residue s (looked up during iteration s-1), hash k-mer s, issue its table-0 probe and the
        //       load of its seen-filter word;
        //   (5) residue s+1: residues come four to a word from the column, the next word is fetched four iterations
        //       before it is needed.
        bool t0_pending = false, s2_pending = false, seen_pending = false;
        uint64_t t0_h = 0;
        uint32_t t0_v = 0, t0_sh = 0;
        uint32_t s2_v1 = 0, s2_v2 = 0, s2_v3 = 0, s2_sh = 0;  // s2_sh: the wanted bits as a mask over v1 | v2 << 8 | v3 << 16
        unsigned long long *seen_slot = seen;
        unsigned long long seen_bits = 0, seen_old = 0;
        // residues come four to a word from the column; the next word is fetched four iterations before it is needed
        uint32_t res_word = strand[0], res_next = strand[kTaskBlock];
        uint32_t res_cur = res_word & 0xFFu;  // residue 0
        res_word >>= 8;
        for (uint32_t s = 0; s < max_plen + 2; s++) {
            // (1)
            bool kmer_new = false;
            if (seen_pending) {
                *seen_slot = seen_old | seen_bits;
                kmer_new = true;
                if ((seen_old & seen_bits) == seen_bits || force_exact)
                    kmer_new = !kmer_seen_before<K, NM>(sc, lut_fwd, phase, (int)s - 1, win);
                nd += kmer_new ? 1u : 0u;
                seen_pending = false;
            }
            // (2)
            if (s2_pending) nh += ((s2_v1 | (s2_v2 << 8) | (s2_v3 << 16)) & s2_sh) == s2_sh ? 1u : 0u;
            s2_pending = false;
            // (3)
            if (t0_pending && kmer_new && ((t0_v >> t0_sh) & 1u)) {
                if (!T4 && n_tables == 1) {
                    nh += 1;
                } else if (four_tables) {
                    uint64_t b1, b2, b3;
                    ORPH_BIN(1, t0_h, b1);
                    ORPH_BIN(2, t0_h, b2);
                    ORPH_BIN(3, t0_h, b3);
                    s2_v1 = ld_table_byte(P.F.table[1] + (b1 >> 3));
                    s2_v2 = ld_table_byte(P.F.table[2] + (b2 >> 3));
                    s2_v3 = ld_table_byte(P.F.table[3] + (b3 >> 3));
                    // the three wanted bits as one mask over the bytes side by side: one compare when they arrive
                    s2_sh = (1u << ((uint32_t)b1 & 7u)) | (0x100u << ((uint32_t)b2 & 7u)) | (0x10000u << ((uint32_t)b3 & 7u));
                    s2_pending = true;
                } else {
                    nh += filter_contains_rest(sF, t0_h) ? 1u : 0u;
                }
            }
            t0_pending = false;
            // (4)
            if (s < plen) {
                win.push(res_cur);
                if (s + 1 >= (uint32_t)K) {
                    const uint64_t h = win.hash();
                    uint64_t bin0;
                    ORPH_BIN(0, h, bin0);
                    t0_v = ld_table_byte(P.F.table[0] + (bin0 >> 3));
                    t0_sh = (uint32_t)bin0 & 7u;
                    t0_h = h;
                    t0_pending = true;
                    const uint32_t hw = (uint32_t)(h >> 32);
                    // seen filter: word chosen by 4 hash bits; two bits in its low half, one in its high half (15 more hash
                    // bits, 32-bit shifts only; the false-alarm rate is that of three bits anywhere in the word to within 20 %)
                    seen_bits = (unsigned long long)((1u << (hw & 31u)) | (1u << ((hw >> 5) & 31u))) | ((unsigned long long)(1u << ((hw >> 10) & 31u)) << 32);
                    seen_slot = seen + ((hw >> 15) & (seen_words - 1)) * kTaskBlock;
                    seen_old = *seen_slot;
                    seen_pending = true;
                }
            }
            // (5) residue s+1
            if (((s + 1u) & 3u) == 0) {
                res_word = res_next;
                res_next = strand[(((s + 1u) >> 2) + 1u) * kTaskBlock];
            }
            res_cur = res_word & 0xFFu;
            res_word >>= 8;
        }
#undef ORPH_BIN
        if (valid) {
            const uint32_t n = plen - (uint32_t)K + 1;
            // evaluate_is_kmer_low_complexity: n_kmers <= (len - k + 1) / 2  (translate.py:98-100), in integers
            const uint8_t cat = 2 * nd <= n ? ORPH_CAT_LOW_COMPLEXITY_PEPTIDE
                                : (nh >= s_min_hits[nd] ? ORPH_CAT_CODING : ORPH_CAT_NON_CODING);
            out[r].n_kmers[f] = (uint16_t)nd;
            out[r].n_hits[f] = (uint16_t)nh;
            out[r].category[f] = cat;
        }
        __syncwarp();
    }
}

// best_frame for every read that had frames scored by score_tasks_kernel (its frames are finished by
// different threads): exact rational argmax over the coding / non-coding frames of the record.
__global__ void __launch_bounds__(256)
finalize_best_kernel(orph_read_result *__restrict__ out, uint64_t n_reads)
{
    for (uint64_t r = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; r < n_reads; r += (uint64_t)gridDim.x * blockDim.x) {
        const uint4 *src = reinterpret_cast<const uint4 *>(out + r);
        const uint4 a = src[0], b = src[1];
        const uint32_t words[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
        const uint32_t flags = words[7] >> 24;
        if (!(flags & ORPH_FLAG_OPEN_MASK) || (flags & ORPH_FLAG_BYTE_PATH)) continue;
        BestFrame best;
#pragma unroll
        for (int f = 0; f < 6; f++) {
            const uint32_t nk = (words[f >> 1] >> (16 * (f & 1))) & 0xFFFFu;
            const uint32_t nh = (words[3 + (f >> 1)] >> (16 * (f & 1))) & 0xFFFFu;
            const uint32_t cat = f < 4 ? (words[6] >> (8 * f)) & 0xFFu : (words[7] >> (8 * (f - 4))) & 0xFFu;
            best.offer(nh, nk, (uint8_t)cat, frame_key(f));
        }
        out[r].best_frame = (int8_t)best.frame;
    }
}

}  // namespace orph
