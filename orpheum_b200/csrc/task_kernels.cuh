// score_tasks_kernel: one THREAD per open frame ("task"), for reads up to kThreadMaxLen bases.
//
// This is the hot kernel for short reads.  A thread walks its frame codon by codon over the read's
// 2-bit words (staged once into shared memory), keeps the current peptide k-mer as a rolling byte window
// in registers (exactly the bytes sourmash's hash_murmur sees after encode_peptide,
// sequence_encodings.py:457-468), hashes it (translate.py:187) and issues the table-0 Bloom probe
// (translate.py:190).  Every lane of a warp is a different frame, so the k-mer loop runs at full lane
// occupancy; the probe of k-mer s is consumed one iteration later so its L2 latency overlaps the next
// hash.  K-mers that survive table 0 are compacted into a per-warp shared-memory queue and finished 32
// at a time (tables 1..n-1), again at full occupancy.
//
// Distinct k-mers (kmerize returns a set, compare_kmer_content.py:116-118): each thread keeps a private
// 1024-bit blocked three-hash "seen" filter in shared memory.  A k-mer whose three bits are not all set
// is certainly new; otherwise the frame is walked backwards from the current k-mer and compared
// byte-exactly (a tandem repeat answers after one period, a false alarm walks the whole prefix, rarely),
// so the de-duplication is exact.
#pragma once
#include "device_common.cuh"
#include "score_kernels.cuh"

namespace orph {

constexpr uint32_t kThreadMaxLen = 192;  // longest read handled one-thread-per-frame
constexpr int kTaskBlock = 256;
constexpr uint32_t kSeenWords64 = 16;    // 16 x 64-bit words per thread, word w of thread t at [w * kTaskBlock + t]

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

// one thread's view of its read: 32-bit words of the 2-bit stream in a shared-memory column
struct ReadColumn {
    const uint32_t *col;  // word t at col[t * kTaskBlock]
    uint32_t bit0;        // bit of base 0 inside word 0
    // encoded residue i of the frame that starts at base p0 and advances by dp (= +3 forward, -3 reverse strand)
    __device__ __forceinline__ uint32_t residue(const uint8_t *lut, int p0, int dp, int i) const
    {
        const uint32_t bit = bit0 + 2u * (uint32_t)(p0 + dp * i);
        const uint32_t wi = bit >> 5;
        const uint32_t c = __funnelshift_r(col[wi * kTaskBlock], col[(wi + 1) * kTaskBlock], bit & 31u) & 63u;
        return lut[c];
    }
};

// Exact answer behind a seen-filter alarm: does the k-mer ending at residue s (window `cur`) also end
// at an earlier residue of the same frame?  Walks backwards, so a tandem repeat answers after one period.
template <int K>
__device__ __forceinline__ bool kmer_seen_before(const ReadColumn &rc, const uint8_t *lut, int p0, int dp, int s,
                                                 const KmerWindow<K> &cur)
{
    KmerWindow<K> win = cur;
    for (int e = s - 1; e >= K - 1; e--) {
        win.push_front(rc.residue(lut, p0, dp, e - K + 1));
        if (win.equals(cur)) return true;
    }
    return false;
}

template <int K>
__global__ void __launch_bounds__(kTaskBlock)
score_tasks_kernel(const ScoreParams P, const uint32_t *__restrict__ packed32, const uint64_t *__restrict__ offsets,
                   uint64_t base0, const uint32_t *__restrict__ tasks, ScoreCounters *__restrict__ counters,
                   orph_read_result *__restrict__ out)
{
    constexpr int kWarps = kTaskBlock / 32;
    extern __shared__ __align__(16) uint8_t dyn_smem[];  // [seen: 16 u64 x 256][read words: P.read_words u32 x 256]
    __shared__ FilterDev sF;
    __shared__ uint8_t lut_fwd[64], lut_rev[64];
    __shared__ uint16_t s_min_hits[kThreadMaxLen / 3 + 2];
    __shared__ unsigned long long q_hash[kWarps][64];
    __shared__ uint8_t q_lane[kWarps][64];
    __shared__ uint32_t s_hits[kWarps][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    stage_filter(&sF, P.F);
    if (threadIdx.x < 64) {
        const int t = threadIdx.x, b0 = t & 3, b1 = (t >> 2) & 3, b2 = t >> 4;
        lut_fwd[t] = P.lut[translate_codes(b0, b1, b2)];
        lut_rev[t] = P.lut[translate_codes(b2 ^ 2, b1 ^ 2, b0 ^ 2)];
    }
    const double thr = P.jaccard_threshold;
    for (uint32_t n = threadIdx.x; n <= kThreadMaxLen / 3 + 1; n += blockDim.x)
        s_min_hits[n] = n ? (uint16_t)min_hits_for(n, thr) : (uint16_t)1;
    __syncthreads();
    const bool force_exact = P.force_exact != 0;
    const uint32_t n_tables = sF.n_tables;
    const bool single_table = n_tables == 1;
    // table 0 is probed once per k-mer: keep its description in registers
    const uint64_t t0_size = sF.size[0], t0_magic = sF.magic[0];
    const uint8_t *t0_table = sF.table[0];
    const bool t0_small = sF.small != 0;
    const uint32_t n_tasks = counters->n_tasks;
    unsigned long long *seen = reinterpret_cast<unsigned long long *>(dyn_smem) + threadIdx.x;
    uint32_t *rd_col = reinterpret_cast<uint32_t *>(dyn_smem + (size_t)kSeenWords64 * 8 * kTaskBlock) + threadIdx.x;
    const uint32_t read_words = P.read_words;

    while (true) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&counters->next_task, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n_tasks) break;
        const bool valid = base + lane < n_tasks;
        uint32_t r = 0, f = 0, L = 0, plen = 0;
        uint64_t o0 = 0;
        if (valid) {
            const uint32_t entry = tasks[base + lane];
            r = entry >> 3;
            f = entry & 7u;
            o0 = offsets[r];
            L = (uint32_t)(offsets[r + 1] - o0);
        }
        const uint32_t phase = f >= 3 ? f - 3 : f;
        if (valid) plen = (L - phase) / 3;
        const int p0 = f < 3 ? (int)phase : (int)L - 3 - (int)phase, dp = f < 3 ? 3 : -3;
        const uint8_t *lut = f < 3 ? lut_fwd : lut_rev;
        // stage this read's words (plus one zero word) into the thread's shared-memory column
        const uint64_t bitpos = 2 * (o0 - base0);
        const uint32_t bit0 = (uint32_t)(bitpos & 31);
        {
            const uint32_t *src = packed32 + (bitpos >> 5);
            const uint32_t n_words = valid ? (bit0 + 2 * L + 31) / 32 : 0;
            for (uint32_t t = 0; t < read_words; t++) rd_col[t * kTaskBlock] = t < n_words ? __ldg(src + t) : 0u;
        }
        const ReadColumn rc{rd_col, bit0};
        KmerWindow<K> win;
        win.clear();
#pragma unroll
        for (uint32_t w = 0; w < kSeenWords64; w++) seen[w * kTaskBlock] = 0ull;
        s_hits[warp][lane] = 0;
        uint32_t nd = 0, nh_direct = 0, qlen = 0;
        const uint32_t max_plen = __reduce_max_sync(0xffffffffu, plen);
        __syncwarp();

        // Software pipeline.  Iteration s: (1) if 32 survivors are queued, issue their table 1..3 probes;
        // (2) roll in residue s (fetched one iteration ahead), hash k-mer s, issue its table-0 probe, update
        // the seen filter; (3) consume the probes of step 1; (4) consume the table-0 probe of k-mer s-1 and
        // queue it if it survived.  Every global load has a full hash between issue and use.
        bool prev_new = false;
        uint64_t prev_h = 0;
        uint32_t prev_v0 = 0, prev_sh = 0;
        uint32_t res_next = plen ? rc.residue(lut, p0, dp, 0) : 0u;
        for (uint32_t s = 0; s <= max_plen; s++) {
            // (1)
            const bool draining = !single_table && qlen >= 32;
            uint32_t d_hit = 0, d_owner = 0;
            uint64_t d_b1 = 0, d_b2 = 0, d_b3 = 0;
            uint32_t d_v1 = 0, d_v2 = 0, d_v3 = 0;
            bool d_generic = false;
            uint64_t d_h = 0;
            if (draining) {
                qlen -= 32;
                d_h = q_hash[warp][qlen + lane];
                d_owner = q_lane[warp][qlen + lane];
                if (n_tables == 4) {
                    d_b1 = filter_bin(sF, 1, d_h); d_b2 = filter_bin(sF, 2, d_h); d_b3 = filter_bin(sF, 3, d_h);
                    d_v1 = ld_table_byte(sF.table[1] + (d_b1 >> 3));
                    d_v2 = ld_table_byte(sF.table[2] + (d_b2 >> 3));
                    d_v3 = ld_table_byte(sF.table[3] + (d_b3 >> 3));
                } else {
                    d_generic = true;
                }
            }
            // (2)
            bool cur_new = false;
            uint64_t h = 0;
            uint32_t v0 = 0, sh = 0;
            if (s < plen) {
                win.push(res_next);
                if (s + 1 < plen) res_next = rc.residue(lut, p0, dp, (int)s + 1);
                if (s + 1 >= (uint32_t)K) {
                    h = win.hash();
                    uint64_t bin0;
                    {
                        const uint64_t q = __umul64hi(h, t0_magic);
                        if (t0_small) {
                            uint32_t rr = (uint32_t)h - (uint32_t)q * (uint32_t)t0_size;
                            if (rr >= (uint32_t)t0_size) rr -= (uint32_t)t0_size;
                            bin0 = rr;
                        } else {
                            uint64_t rr = h - q * t0_size;
                            if (rr >= t0_size) rr -= t0_size;
                            bin0 = rr;
                        }
                    }
                    v0 = ld_table_byte(t0_table + (bin0 >> 3));
                    sh = (uint32_t)bin0 & 7u;
                    // seen filter: word chosen by 4 hash bits, three bit positions inside it by 18 more
                    const uint32_t hw = (uint32_t)(h >> 32);
                    const unsigned long long bits = (1ull << (hw & 63u)) | (1ull << ((hw >> 6) & 63u)) | (1ull << ((hw >> 12) & 63u));
                    unsigned long long *slot = seen + ((hw >> 18) & (kSeenWords64 - 1)) * kTaskBlock;
                    const unsigned long long old = *slot;
                    *slot = old | bits;
                    cur_new = true;
                    if ((old & bits) == bits || force_exact) cur_new = !kmer_seen_before<K>(rc, lut, p0, dp, (int)s, win);
                    nd += cur_new ? 1u : 0u;
                }
            }
            // (3)
            if (draining) {
                if (d_generic) d_hit = filter_contains_rest(sF, d_h) ? 1u : 0u;
                else d_hit = (d_v1 >> (d_b1 & 7)) & (d_v2 >> (d_b2 & 7)) & (d_v3 >> (d_b3 & 7)) & 1u;
                if (d_hit) atomicAdd(&s_hits[warp][d_owner], 1u);
                __syncwarp();
            }
            // (4)
            const bool hit0 = prev_new && ((prev_v0 >> prev_sh) & 1u);
            const uint64_t hq = prev_h;
            prev_new = cur_new; prev_h = h; prev_v0 = v0; prev_sh = sh;
            if (single_table) {
                nh_direct += hit0 ? 1u : 0u;
                continue;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, hit0);
            if (m) {
                if (hit0) {
                    const uint32_t idx = qlen + __popc(m & ((1u << lane) - 1));
                    q_hash[warp][idx] = hq;
                    q_lane[warp][idx] = (uint8_t)lane;
                }
                qlen += __popc(m);
                __syncwarp();
            }
        }
        // at most 63 survivors are left: one full batch, then the remainder
        if (qlen >= 32) {
            qlen -= 32;
            const uint64_t qh = q_hash[warp][qlen + lane];
            const uint32_t ql = q_lane[warp][qlen + lane];
            if (filter_contains_rest(sF, qh)) atomicAdd(&s_hits[warp][ql], 1u);
        }
        if (qlen) {
            if ((uint32_t)lane < qlen) {
                const uint64_t qh = q_hash[warp][lane];
                const uint32_t ql = q_lane[warp][lane];
                if (filter_contains_rest(sF, qh)) atomicAdd(&s_hits[warp][ql], 1u);
            }
        }
        __syncwarp();
        if (valid) {
            const uint32_t nh = single_table ? nh_direct : s_hits[warp][lane];
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
