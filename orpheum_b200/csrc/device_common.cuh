// Device-side building blocks shared by the scoring and index kernels (sm_100a).
//
//   * MurmurHash3_x64_128 (seed 42, h1): what sourmash.minhash.hash_murmur computes
//     (reference call sites orpheum/translate.py:187, orpheum/index.py:115).
//   * khmer.Nodegraph bin arithmetic: bin = hash % prime_i, byte bin>>3, bit bin&7
//     (reference call sites orpheum/translate.py:190, orpheum/index.py:119); the modulo is an exact
//     Barrett reduction with a per-table precomputed floor(2^64 / p).
//   * the standard genetic code over 2-bit codes and over raw bytes
//     (orpheum/constants_translate.py:59-141, orpheum/translate_single_seq.py:13-32).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/orpheum_b200.h"

namespace orph {

constexpr uint32_t kMurmurSeed = 42u;
constexpr uint64_t kC1 = 0x87c37b91114253d5ULL;
constexpr uint64_t kC2 = 0x4cf5ad432745937fULL;
constexpr uint64_t kEmptyKey = 0xFFFFFFFFFFFFFFFFULL;  // dedup-table sentinel
constexpr int kMaxKsize = 64;                           // longest peptide k-mer the kernels hash
constexpr int kPepPad = kMaxKsize + 8;                  // over-read slack behind a peptide buffer

// Filter geometry as the kernels see it (passed by value in kernel parameters).
struct FilterDev {
    uint32_t ksize;
    uint32_t n_tables;
    uint64_t size[ORPH_MAX_TABLES];   // bits (primes)
    uint64_t magic[ORPH_MAX_TABLES];  // floor((2^64-1) / size)
    const uint8_t *table[ORPH_MAX_TABLES];
    uint32_t small;                   // every size < 2^31: remainders fit 32-bit arithmetic
    uint32_t tiny;                    // every size <= kTinyPrimeMax: mod_prime_tiny applies
};

__device__ __forceinline__ uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

__device__ __forceinline__ uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

__device__ __forceinline__ uint64_t mix_k1(uint64_t k1) { k1 *= kC1; k1 = rotl64(k1, 31); k1 *= kC2; return k1; }
__device__ __forceinline__ uint64_t mix_k2(uint64_t k2) { k2 *= kC2; k2 = rotl64(k2, 33); k2 *= kC1; return k2; }

// MurmurHash3_x64_128(key, len, seed 42).h1.  `word(t)` must return little-endian bytes [4t, 4t+4) of
// the key with every byte at index >= len forced to zero.
template <class WordFn>
__device__ __forceinline__ uint64_t murmur3_h1(WordFn word, uint32_t len)
{
    uint64_t h1 = kMurmurSeed, h2 = kMurmurSeed;
    const uint32_t nblocks = len >> 4;
    for (uint32_t b = 0; b < nblocks; b++) {
        uint64_t k1 = (uint64_t)word(4 * b) | ((uint64_t)word(4 * b + 1) << 32);
        uint64_t k2 = (uint64_t)word(4 * b + 2) | ((uint64_t)word(4 * b + 3) << 32);
        h1 ^= mix_k1(k1);
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        h2 ^= mix_k2(k2);
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint32_t tail = len & 15u;
    const uint32_t t0 = nblocks * 4;
    if (tail > 8) {
        uint64_t k2 = (uint64_t)word(t0 + 2);
        if (tail > 12) k2 |= (uint64_t)word(t0 + 3) << 32;
        h2 ^= mix_k2(k2);
    }
    if (tail > 0) {
        uint64_t k1 = (uint64_t)word(t0);
        if (tail > 4) k1 |= (uint64_t)word(t0 + 1) << 32;
        h1 ^= mix_k1(k1);
    }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    return h1 + h2;
}

// Word getter over a 4-byte-aligned shared-memory byte buffer: key starts at byte `j`.
// Reads one word past the key's last word, hence the kPepPad slack behind every buffer.
struct SmemKeyWords {
    const uint32_t *w;  // buffer as words
    uint32_t j;         // key start (bytes)
    uint32_t len;
    __device__ __forceinline__ uint32_t operator()(uint32_t t) const
    {
        const uint32_t byte = j + 4 * t;
        const uint32_t a = byte >> 2;
        const uint32_t v = __funnelshift_r(w[a], w[a + 1], (byte & 3u) * 8u);
        const int valid = (int)len - (int)(4 * t);  // bytes of this word that belong to the key
        if (valid >= 4) return v;
        if (valid <= 0) return 0u;
        return v & (0xFFFFFFFFu >> (32 - 8 * valid));
    }
};

// exact h % p.  With magic = floor((2^64-1)/p), q = umulhi(h, magic) is floor(h/p) or one less for every
// p >= 2, so a single conditional subtraction finishes the reduction.
__device__ __forceinline__ uint64_t mod_prime(uint64_t h, uint64_t p, uint64_t magic)
{
    const uint64_t q = __umul64hi(h, magic);
    uint64_t r = h - q * p;
    if (r >= p) r -= p;
    return r;
}

// h % p for p <= kTinyPrimeMax = (2^32 - 1) / 3, from 32-bit pieces only (six instructions against seventeen for
// mod_prime).  With m = magic = floor((2^64-1)/p), Q = floor(h m / 2^64) is floor(h/p) or one less.  Dropping the product of
// the two low words (lo mlo < 2^64) lowers the quotient estimate by at most one more:
//     Q' = hi mhi + floor((hi mlo + lo mhi) / 2^32)  in  {floor(h/p) - 2, ..., floor(h/p)},
// and only Q' mod 2^32 is needed (carries out of the 64-bit sum are multiples of 2^32 there), because
// h - Q' p < 3p < 2^32 is exact in 32-bit arithmetic.  Two conditional subtractions (min with the wrapped difference) finish.
constexpr uint64_t kTinyPrimeMax = 0xFFFFFFFFull / 3;
__device__ __forceinline__ uint32_t mod_prime_tiny(uint64_t h, uint32_t p, uint64_t magic)
{
    const uint32_t hi = (uint32_t)(h >> 32), lo = (uint32_t)h, mhi = (uint32_t)(magic >> 32), mlo = (uint32_t)magic;
    const uint64_t mid = (uint64_t)hi * mlo + (uint64_t)lo * mhi;  // mod 2^64
    const uint32_t q = hi * mhi + (uint32_t)(mid >> 32);
    const uint32_t neg_p = 0u - p;
    uint32_t r = lo + q * neg_p;
    r = __viaddmin_u32(r, neg_p, r);
    r = __viaddmin_u32(r, neg_p, r);
    return r;
}

// bin of hash h in table t.  F lives in shared memory (uniform, conflict-free broadcast reads).
__device__ __forceinline__ uint64_t filter_bin(const FilterDev &F, uint32_t t, uint64_t h)
{
    if (F.tiny) return mod_prime_tiny(h, (uint32_t)F.size[t], F.magic[t]);
    const uint64_t q = __umul64hi(h, F.magic[t]);
    if (F.small) {  // the remainder before correction is < 2p < 2^32: low words suffice
        const uint32_t p = (uint32_t)F.size[t];
        uint32_t r = (uint32_t)h - (uint32_t)q * p;
        if (r >= p) r -= p;
        return r;
    }
    const uint64_t p = F.size[t];
    uint64_t r = h - q * p;
    if (r >= p) r -= p;
    return r;
}

// Bloom-table probe: read-only path, no L1 allocation (random 1-byte gathers over >= 50 MB would only
// evict the lines the kernel actually re-uses); the tables live in L2.
__device__ __forceinline__ uint32_t ld_table_byte(const uint8_t *p)
{
    uint32_t v;
    asm("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Nodegraph.get: AND over tables with early exit.
__device__ __forceinline__ bool filter_contains(const FilterDev &F, uint64_t h)
{
    bool present = true;
    for (uint32_t t = 0; t < F.n_tables && present; t++) {
        const uint64_t bin = filter_bin(F, t, h);
        present = (ld_table_byte(F.table[t] + (bin >> 3)) >> (bin & 7)) & 1u;
    }
    return present;
}

// Tables 1.. of Nodegraph.get once table 0 is known to hit.  For the usual 4-table filter the three
// remaining probes are issued back to back so their L2 latencies overlap.
__device__ __forceinline__ bool filter_contains_rest(const FilterDev &F, uint64_t h)
{
    if (F.n_tables == 4) {
        const uint64_t b1 = filter_bin(F, 1, h), b2 = filter_bin(F, 2, h), b3 = filter_bin(F, 3, h);
        const uint32_t v1 = ld_table_byte(F.table[1] + (b1 >> 3));
        const uint32_t v2 = ld_table_byte(F.table[2] + (b2 >> 3));
        const uint32_t v3 = ld_table_byte(F.table[3] + (b3 >> 3));
        return ((v1 >> (b1 & 7)) & (v2 >> (b2 & 7)) & (v3 >> (b3 & 7)) & 1u) != 0;
    }
    bool present = true;
    for (uint32_t t = 1; t < F.n_tables && present; t++) {
        const uint64_t bin = filter_bin(F, t, h);
        present = (ld_table_byte(F.table[t] + (bin >> 3)) >> (bin & 7)) & 1u;
    }
    return present;
}

// block-wide copy of the kernel-parameter filter description into shared memory
__device__ __forceinline__ void stage_filter(FilterDev *dst, const FilterDev &src)
{
    if (threadIdx.x == 0) *dst = src;
}

// ---- genetic code --------------------------------------------------------------------------
// index = 16*b0 + 4*b1 + b2 with T,C,A,G = 0,1,2,3 (the classic NCBI table-1 ordering)
__device__ __constant__ char kAminoTCAG[65] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";

// 2-bit read code (ascii >> 1) & 3: A=0 C=1 T=2 G=3  ->  TCAG index
__device__ __forceinline__ int code_to_tcag(int c) { return (0xC6 >> (2 * c)) & 3; }  // {2,1,0,3} packed: 0b11'00'01'10

// amino acid of the codon with 2-bit codes (b0,b1,b2)
__device__ __forceinline__ uint8_t translate_codes(int b0, int b1, int b2)
{
    return (uint8_t)kAminoTCAG[16 * code_to_tcag(b0) + 4 * code_to_tcag(b1) + code_to_tcag(b2)];
}

// STANDARD_CODON_TABLE_MAPPING.get(codon, "X") on raw bytes: 64 upper-case ACGT codons plus
// ACN CTN CGN GTN GCN GGN TCN; everything else (lower case, other IUPAC codes, CCN, N first/second) -> X.
__device__ __forceinline__ int byte_to_tcag(uint8_t b)
{
    return b == 'T' ? 0 : b == 'C' ? 1 : b == 'A' ? 2 : b == 'G' ? 3 : -1;
}

__device__ __forceinline__ uint8_t translate_bytes(uint8_t a, uint8_t b, uint8_t c)
{
    const int ia = byte_to_tcag(a), ib = byte_to_tcag(b), ic = byte_to_tcag(c);
    if (ia < 0 || ib < 0) return 'X';
    if (ic >= 0) return (uint8_t)kAminoTCAG[16 * ia + 4 * ib + ic];
    if (c == 'N') {
        const int two = 4 * ia + ib;  // TCAG order
        switch (two) {
        case 4 * 2 + 1: return 'T';  // ACN
        case 4 * 1 + 0: return 'L';  // CTN
        case 4 * 1 + 3: return 'R';  // CGN
        case 4 * 3 + 0: return 'V';  // GTN
        case 4 * 3 + 1: return 'A';  // GCN
        case 4 * 3 + 3: return 'G';  // GGN
        case 4 * 0 + 1: return 'S';  // TCN
        default: break;
        }
    }
    return 'X';
}

// REVERSE_COMPLEMENT_TABLE: only upper-case ACGT are complemented
__device__ __forceinline__ uint8_t complement_byte(uint8_t c)
{
    return c == 'A' ? 'T' : c == 'T' ? 'A' : c == 'C' ? 'G' : c == 'G' ? 'C' : c;
}

}  // namespace orph
