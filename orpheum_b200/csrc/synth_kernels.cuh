// Measurement helper (bench.py --workload cfg4): counter-based synthetic reads generated straight
// into HBM.  1e9 reads are too many to stage from the host, so read i is a pure function of (seed, i): even i
// is back-translated from the proteome (random synonymous codons, 1 % substitutions, random phase and strand),
// odd i is uniform ACGT.  orpheum_b200/synth.py:counter_reads is the numpy twin used by the parity checks.
// Not on the scoring path.
#pragma once
#include "device_common.cuh"

namespace orph {

constexpr uint64_t kSynthGold = 0x9E3779B97F4A7C15ULL, kSynthStep = 0xD1B54A32D192ED03ULL;
constexpr uint32_t kSynthMaxLen = 320;

__device__ __forceinline__ uint64_t synth_mix(uint64_t z)
{
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t synth_draw(uint64_t seed, uint64_t i, uint32_t d)
{
    return synth_mix(seed + i * kSynthGold + (uint64_t)(d + 1) * kSynthStep);
}

// syn[aa]: bits 0..2 = number of sense codons of amino-acid byte aa, codon c at bits [4 + 6c, +6) as three 2-bit
// read codes (first base lowest).  out: the 2-bit stream, pre-zeroed; read i (global index first_index + t) occupies
// bases [t * L, (t + 1) * L).
__global__ void __launch_bounds__(256)
synth_reads_kernel(uint64_t seed, uint64_t first_index, uint64_t n_reads, uint32_t L, const uint8_t *__restrict__ residues,
                   const uint64_t *__restrict__ prot_offsets, uint64_t n_proteins, const uint64_t *__restrict__ syn,
                   uint32_t *__restrict__ out)
{
    const uint32_t n_res = L / 3 + 2;
    for (uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; t < n_reads; t += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t i = first_index + t;
        uint32_t w[kSynthMaxLen / 16 + 1];  // 16 bases per word
        const uint32_t n_words = (L + 15) / 16;
        if (i & 1) {
            for (uint32_t q = 0; q < n_words; q += 2) {
                const uint64_t d = synth_draw(seed, i, q / 2);
                w[q] = (uint32_t)d;
                if (q + 1 < n_words) w[q + 1] = (uint32_t)(d >> 32);
            }
        } else {
            const uint64_t d0 = synth_draw(seed, i, 0), d1 = synth_draw(seed, i, 1);
            uint64_t p = d0 % n_proteins;
            while (prot_offsets[p + 1] - prot_offsets[p] < n_res) p = p + 1 == n_proteins ? 0 : p + 1;
            const uint64_t plen = prot_offsets[p + 1] - prot_offsets[p];
            const uint64_t start = prot_offsets[p] + (d0 >> 32) % (plen - n_res + 1);
            const uint32_t phase = (uint32_t)(d1 % 3), flip = (uint32_t)(d1 >> 8) & 1u;
            for (uint32_t q = 0; q < n_words; q++) w[q] = 0;
            uint64_t picks = 0;
            for (uint32_t j = 0; j < n_res; j++) {
                if ((j & 7u) == 0) picks = synth_draw(seed, i, 2 + j / 8);
                const uint64_t s = syn[residues[start + j]];
                const uint32_t choice = (uint32_t)((picks >> (8 * (j & 7u))) & 0xFFu) % (uint32_t)(s & 7u);
                const uint32_t codon = (uint32_t)(s >> (4 + 6 * choice)) & 63u;
                for (uint32_t c = 0; c < 3; c++) {
                    const int b = (int)(3 * j + c) - (int)phase;  // position in the read
                    if (b >= 0 && b < (int)L) w[b >> 4] |= ((codon >> (2 * c)) & 3u) << (2 * (b & 15));
                }
            }
            uint64_t subs = 0;
            for (uint32_t b = 0; b < L; b++) {
                if ((b & 3u) == 0) subs = synth_draw(seed, i, 16 + b / 4);
                const uint32_t v = (uint32_t)(subs >> (16 * (b & 3u))) & 0xFFFFu;
                if (v < 655u) {
                    const uint32_t sh = 2 * (b & 15), code = (w[b >> 4] >> sh) & 3u;
                    w[b >> 4] = (w[b >> 4] & ~(3u << sh)) | (((code + 1 + v % 3) & 3u) << sh);
                }
            }
            if (flip) {  // reverse complement: base b <- base L-1-b, complement = code ^ 2
                uint32_t r[kSynthMaxLen / 16 + 1];
                for (uint32_t q = 0; q < n_words; q++) r[q] = 0;
                for (uint32_t b = 0; b < L; b++) {
                    const uint32_t src = L - 1 - b;
                    r[b >> 4] |= (((w[src >> 4] >> (2 * (src & 15))) & 3u) ^ 2u) << (2 * (b & 15));
                }
                for (uint32_t q = 0; q < n_words; q++) w[q] = r[q];
            }
        }
        // mask the tail of the last word, then OR the words into the stream at bit 2 * t * L
        if (L & 15u) w[n_words - 1] &= (1u << (2 * (L & 15u))) - 1u;
        const uint64_t bit = 2 * t * (uint64_t)L;
        uint32_t *dst = out + (bit >> 5);
        const uint32_t sh = (uint32_t)(bit & 31);
        uint32_t carry = 0;
        for (uint32_t q = 0; q < n_words; q++) {
            const uint32_t v = sh ? (w[q] << sh) | carry : w[q];
            carry = sh ? w[q] >> (32 - sh) : 0u;
            if (v) atomicOr(dst + q, v);
        }
        if (carry) atomicOr(dst + n_words, carry);
    }
}

}  // namespace orph
