// score_long_kernel: one BLOCK per read, for reads longer than the warp-per-read byte path holds in shared
// memory (kBytesMaxLen < L <= ORPH_MAX_READ_LEN; the reference has no length limit, translate.py:323-331).
//
// Same byte-exact semantics as score_bytes_kernel (xxN codons, everything else X, case-sensitive
// complexity, reverse complement that leaves non-ACGT bytes alone), but the frame's encoded peptide and the
// de-duplication table live in a per-block slice of a global scratch buffer (L2-resident), and the k-mers
// of a frame are spread over all threads of the block.  Distinct k-mers are exact: a slot holds
// [16-bit hash tag | 16-bit k-mer position] and a tag match is confirmed on the k-mer bytes.
#pragma once
#include "device_common.cuh"
#include "score_kernels.cuh"

namespace orph {

constexpr int kLongBlock = 512;
constexpr uint32_t kLongPosBits = 16;  // peptide positions < 65535 (ORPH_MAX_READ_LEN / 3 = 65535 residues)

// bytes of global scratch one block needs for reads up to `bound` bases
__host__ __device__ inline uint32_t long_slots_for(uint32_t bound)
{
    uint32_t n = bound / 3 + 1, p = 1024;
    while (p < 2 * n) p <<= 1;
    return p;
}
__host__ __device__ inline size_t long_scratch_bytes(uint32_t bound)
{
    return (size_t)long_slots_for(bound) * 4 + (((size_t)bound / 3 + kPepPad + 15) & ~(size_t)15);
}

template <bool PACKED_SRC>
__global__ void __launch_bounds__(kLongBlock)
score_long_kernel(const ScoreParams P, const void *__restrict__ src, const uint64_t *__restrict__ offsets, uint64_t base0,
                  const uint32_t *__restrict__ list_back, uint32_t list_capacity, ScoreCounters *__restrict__ counters,
                  orph_read_result *__restrict__ out, uint8_t *__restrict__ scratch, const uint32_t *__restrict__ nmask)
{
    __shared__ FilterDev sF;
    __shared__ uint8_t s_lut[256];
    __shared__ uint32_t s_item, s_ndiff, s_stops, s_nd, s_nh;
    const int lane = threadIdx.x & 31;
    stage_filter(&sF, P.F);
    for (int i = threadIdx.x; i < 256; i += blockDim.x) s_lut[i] = P.lut[i];
    const uint32_t slots = long_slots_for(P.max_len);
    uint8_t *mine = scratch + (size_t)blockIdx.x * long_scratch_bytes(P.max_len);
    uint32_t *tab = reinterpret_cast<uint32_t *>(mine);
    uint8_t *pep = mine + (size_t)slots * 4;
    const uint32_t *pepw = reinterpret_cast<const uint32_t *>(pep);
    const uint32_t k = P.k;
    const uint32_t n_long = counters->n_long;
    __syncthreads();

    while (true) {
        if (threadIdx.x == 0) {
            s_item = atomicAdd(&counters->next_long, 1u);
            s_ndiff = 0;
            s_stops = 0;
        }
        __syncthreads();
        const uint32_t item = s_item;
        if (item >= n_long) break;
        const uint64_t r = list_back[list_capacity - 1 - item];  // long reads are queued from the back of the byte queue
        const uint64_t o0 = offsets[r];
        const uint32_t L = (uint32_t)(offsets[r + 1] - o0);  // kBytesMaxLen < L (prefilter_kernel)
        // A read beyond ORPH_MAX_READ_LEN (the reference has no limit, translate.py:323-331) is classified exactly as long as
        // no frame of it needs scoring: every frame with a stop codon, or the whole read low-complexity -- which is what a
        // contig or a transcript of that length looks like.  A stop-free frame of more than 65535 residues cannot be
        // represented in the record's 16-bit counts: ORPH_E_TOO_LONG, the read's record stays ORPH_CAT_INVALID.
        const bool over_long = L > P.max_len;
        // compute_fastp_complexity (translate.py:79-84) and the stop codons of all six frames in one pass
        {
            uint32_t ndiff = 0, stops = 0;
            for (uint32_t p = threadIdx.x; p + 1 < L; p += blockDim.x) {
                const uint8_t a = read_base<PACKED_SRC>(src, base0, o0 + p, nmask), b = read_base<PACKED_SRC>(src, base0, o0 + p + 1, nmask);
                ndiff += a != b;
                if (p + 2 < L) {
                    const uint8_t c = read_base<PACKED_SRC>(src, base0, o0 + p + 2, nmask);
                    if (translate_bytes(a, b, c) == '*') stops |= 1u << (p % 3);
                    if (translate_bytes(complement_byte(c), complement_byte(b), complement_byte(a)) == '*')
                        stops |= 8u << ((L - p) % 3);
                }
            }
            ndiff = __reduce_add_sync(0xffffffffu, ndiff);
            stops = __reduce_or_sync(0xffffffffu, stops);
            if (lane == 0) {
                atomicAdd(&s_ndiff, ndiff);
                atomicOr(&s_stops, stops);
            }
        }
        __syncthreads();
        uint8_t cat[6];
        uint32_t nk[6] = {0, 0, 0, 0, 0, 0}, nh[6] = {0, 0, 0, 0, 0, 0};
        uint8_t flags = ORPH_FLAG_BYTE_PATH;
        BestFrame best;
        if ((double)s_ndiff / (double)L < 0.3) {
            flags |= ORPH_FLAG_LOW_COMPLEXITY;
#pragma unroll
            for (int f = 0; f < 6; f++) cat[f] = ORPH_CAT_LOW_COMPLEXITY_NUCLEOTIDE;
        } else {
            const uint32_t stops = s_stops;
            uint32_t open = 0;
            for (int f = 0; f < 6; f++) {
                const uint32_t phase = f % 3;
                const uint32_t plen = (L - phase) / 3;
                if ((stops >> f) & 1u) { cat[f] = ORPH_CAT_STOP_CODONS; continue; }
                if (plen <= k) { cat[f] = ORPH_CAT_TOO_SHORT_PEPTIDE; continue; }
                if (over_long) {
                    cat[f] = ORPH_CAT_INVALID;
                    if (threadIdx.x == 0) atomicOr(&counters->status, kStatusTooLong);
                    continue;
                }
                open |= 1u << f;
                const uint32_t n = plen - k + 1;
                uint32_t mask = 1023;
                while (mask + 1 < 4 * n && mask + 1 < slots) mask = 2 * mask + 1;
                __syncthreads();  // the previous frame is done with pep / tab / s_nd / s_nh
                for (uint32_t i = threadIdx.x; i < plen; i += blockDim.x) {
                    uint8_t aa;
                    if (f < 3) {
                        const uint64_t g = o0 + phase + 3 * i;
                        aa = translate_bytes(read_base<PACKED_SRC>(src, base0, g, nmask), read_base<PACKED_SRC>(src, base0, g + 1, nmask),
                                             read_base<PACKED_SRC>(src, base0, g + 2, nmask));
                    } else {
                        const uint64_t g = o0 + (L - 3 - phase - 3 * i);  // forward position of the codon's last base
                        aa = translate_bytes(complement_byte(read_base<PACKED_SRC>(src, base0, g + 2, nmask)),
                                             complement_byte(read_base<PACKED_SRC>(src, base0, g + 1, nmask)),
                                             complement_byte(read_base<PACKED_SRC>(src, base0, g, nmask)));
                    }
                    pep[i] = s_lut[aa];
                }
                for (uint32_t i = plen + threadIdx.x; i < plen + kPepPad; i += blockDim.x) pep[i] = 0;
                {
                    uint4 *t4 = reinterpret_cast<uint4 *>(tab);
                    for (uint32_t s = threadIdx.x; s <= (mask >> 2); s += blockDim.x) t4[s] = make_uint4(kEmptySlot, kEmptySlot, kEmptySlot, kEmptySlot);
                }
                if (threadIdx.x == 0) { s_nd = 0; s_nh = 0; }
                __syncthreads();
                uint32_t my_nd = 0, my_nh = 0;
                for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
                    const uint64_t h = murmur3_h1(SmemKeyWords{pepw, j, k}, k);
                    bool winner = false;
                    if (P.force_exact) {  // test hook: all-pairs restatement, no hashing shortcuts
                        winner = true;
                        for (uint32_t e = 0; winner && e < j; e++)
                            if (kmer_bytes_equal(pep, e, j, k)) winner = false;
                    } else {
                        const uint32_t tag = (uint32_t)(h >> 40) & 0xFFFFu;
                        const uint32_t entry = (tag << kLongPosBits) | j;
                        uint32_t slot = (uint32_t)h & mask;
                        const uint32_t step = ((uint32_t)(h >> 20) & mask) | 1u;  // odd: visits every slot
                        while (true) {
                            const uint32_t old = atomicCAS(&tab[slot], kEmptySlot, entry);
                            if (old == kEmptySlot) { winner = true; break; }
                            if ((old >> kLongPosBits) == tag && kmer_bytes_equal(pep, old & 0xFFFFu, j, k)) break;
                            slot = (slot + step) & mask;
                        }
                    }
                    if (winner) {
                        my_nd++;
                        my_nh += filter_contains(sF, h) ? 1u : 0u;
                    }
                }
                my_nd = __reduce_add_sync(0xffffffffu, my_nd);
                my_nh = __reduce_add_sync(0xffffffffu, my_nh);
                if (lane == 0 && my_nd) {
                    atomicAdd(&s_nd, my_nd);
                    atomicAdd(&s_nh, my_nh);
                }
                __syncthreads();
                nk[f] = s_nd;
                nh[f] = s_nh;
                cat[f] = scored_category(nk[f], nh[f], plen, k, P.jaccard_threshold);
                best.offer(nh[f], nk[f], cat[f], frame_key(f));
            }
            flags |= (uint8_t)open;
            if (over_long && open == 0) {
                bool invalid = false;
#pragma unroll
                for (int f = 0; f < 6; f++) invalid |= cat[f] == ORPH_CAT_INVALID;
                if (invalid) {  // the record of a read that could not be scored carries no partial answer
#pragma unroll
                    for (int f = 0; f < 6; f++) cat[f] = ORPH_CAT_INVALID;
                }
            }
        }
        if (threadIdx.x == 0) {
            orph_read_result rec;
#pragma unroll
            for (int f = 0; f < 6; f++) { rec.n_kmers[f] = (uint16_t)nk[f]; rec.n_hits[f] = (uint16_t)nh[f]; rec.category[f] = cat[f]; }
            rec.best_frame = (int8_t)best.frame;
            rec.flags = flags;
            out[r] = rec;
        }
        __syncthreads();  // s_item / s_ndiff / s_stops are rewritten at the top
    }
}

}  // namespace orph
