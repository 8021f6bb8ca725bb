// Index-side kernels: khmer.Nodegraph.add / get / occupied_bins on device-resident bit tables, and
// the inner loop of index.make_peptide_bloom_filter (orpheum/index.py:107-122).
#pragma once
#include "device_common.cuh"

namespace orph {

struct BuildParams {
    FilterDev F;       // table pointers are written through (cast away const in the kernel)
    uint8_t lut[256];  // encode_peptide as a byte table
};

// Nodegraph.add(h): OR bit (h % size_i) into every table; returns true if any bit was newly set.
__device__ __forceinline__ bool filter_add(const FilterDev &F, uint64_t h)
{
    bool is_new = false;
    for (uint32_t t = 0; t < F.n_tables; t++) {
        const uint64_t bin = filter_bin(F, t, h);
        // byte bin>>3, bit bin&7 of a little-endian table == bit (bin & 31) of u32 word bin>>5
        uint32_t *word = reinterpret_cast<uint32_t *>(const_cast<uint8_t *>(F.table[t])) + (bin >> 5);
        const uint32_t bit = 1u << (bin & 31);
        if (!(*reinterpret_cast<volatile uint32_t *>(word) & bit)) {
            const uint32_t old = atomicOr(word, bit);
            is_new |= !(old & bit);
        }
    }
    return is_new;
}

// word getter over global residues re-encoded through a shared-memory LUT
struct GlobalLutKeyWords {
    const uint8_t *p;    // first residue of the k-mer
    const uint8_t *lut;  // shared memory, 256 bytes
    uint32_t len;
    __device__ __forceinline__ uint32_t operator()(uint32_t t) const
    {
        uint32_t v = 0;
#pragma unroll
        for (uint32_t i = 0; i < 4; i++) {
            const uint32_t b = 4 * t + i;
            if (b < len) v |= (uint32_t)lut[__ldg(p + b)] << (8 * i);
        }
        return v;
    }
};

// One warp per peptide record (dynamic fetch).  A record containing '*' is skipped as a whole
// (index.py:108-109); records shorter than k contribute nothing (index.py:120-122).
//
// `seen` (optional, `seen_mask`+1 slots initialised to kEmptyKey) is a global hash set of the k-mer
// hashes of this call: only the first occurrence of a hash goes on to the bit tables.  The bits are
// the same either way; what it buys is khmer's n_unique_kmers (the number of adds that set a new bit),
// which concurrent duplicate k-mers -- isoforms of one gene in neighbouring records -- would otherwise
// over-count, because two in-flight copies can each win a different table's bit.
__global__ void __launch_bounds__(256)
bloom_add_peptides_kernel(const BuildParams P, const uint8_t *__restrict__ residues,
                          const uint64_t *__restrict__ offsets, uint64_t n_records,
                          unsigned long long *__restrict__ next_record, unsigned long long *__restrict__ n_unique,
                          unsigned long long *__restrict__ seen, uint64_t seen_mask)
{
    __shared__ uint8_t lut[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) lut[i] = P.lut[i];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const uint32_t k = P.F.ksize;
    uint32_t my_new = 0;
    while (true) {
        unsigned long long rec = 0;
        if (lane == 0) rec = atomicAdd(next_record, 1ull);
        rec = __shfl_sync(0xffffffffu, rec, 0);
        if (rec >= n_records) break;
        const uint64_t o0 = offsets[rec], len = offsets[rec + 1] - o0;
        if (len < k) continue;
        const uint8_t *seq = residues + o0;
        bool star = false;
        for (uint64_t i = lane; i < len; i += 32) star |= __ldg(seq + i) == '*';
        if (__any_sync(0xffffffffu, star)) continue;
        const uint64_t n = len - k + 1;
        for (uint64_t j = lane; j < n; j += 32) {
            const uint64_t h = murmur3_h1(GlobalLutKeyWords{seq + j, lut, k}, k);
            bool first = true;
            if (seen && h != kEmptyKey) {
                uint64_t slot = h & seen_mask;
                while (true) {
                    const unsigned long long old = atomicCAS(&seen[slot], (unsigned long long)kEmptyKey, (unsigned long long)h);
                    if (old == kEmptyKey) break;
                    if (old == h) { first = false; break; }
                    slot = (slot + 1) & seen_mask;
                }
            }
            if (first) my_new += filter_add(P.F, h) ? 1u : 0u;
        }
    }
    my_new = __reduce_add_sync(0xffffffffu, my_new);
    if (lane == 0 && my_new) atomicAdd(n_unique, (unsigned long long)my_new);
}

__global__ void bloom_add_hashes_kernel(const FilterDev F, const uint64_t *__restrict__ hashes, uint64_t n,
                                        uint8_t *__restrict__ is_new, unsigned long long *__restrict__ n_unique)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const bool fresh = filter_add(F, hashes[i]);
        if (is_new) is_new[i] = fresh;
        if (fresh) atomicAdd(n_unique, 1ull);
    }
}

__global__ void bloom_query_hashes_kernel(const FilterDev F, const uint64_t *__restrict__ hashes, uint64_t n,
                                          uint8_t *__restrict__ out01)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out01[i] = filter_contains(F, hashes[i]);
}

// popcount over `n_words` u32 words (tables are zero padded to a multiple of 256 bytes)
__global__ void popcount_kernel(const uint32_t *__restrict__ words, uint64_t n_words, unsigned long long *__restrict__ out)
{
    unsigned long long acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n_words; i += (uint64_t)gridDim.x * blockDim.x)
        acc += __popc(__ldg(words + i));
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(out, acc);
}

struct GlobalKeyWords {
    const uint8_t *p;
    uint32_t len;
    __device__ __forceinline__ uint32_t operator()(uint32_t t) const
    {
        uint32_t v = 0;
#pragma unroll
        for (uint32_t i = 0; i < 4; i++) {
            const uint32_t b = 4 * t + i;
            if (b < len) v |= (uint32_t)p[b] << (8 * i);
        }
        return v;
    }
};

__global__ void hash_kmers_kernel(const uint8_t *__restrict__ keys, uint32_t len, uint64_t n, uint64_t *__restrict__ out)
{
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = murmur3_h1(GlobalKeyWords{keys + i * len, len}, len);
}

// Random 32-byte-sector gather micro-benchmark (the measured ceiling for the Bloom probes).
__global__ void __launch_bounds__(256)
gather_peak_kernel(const uint8_t *__restrict__ buf, uint64_t n_bytes, uint32_t loads_per_thread, uint32_t *__restrict__ sink)
{
    uint64_t x = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * 0x9E3779B97F4A7C15ULL + 0x1234567ULL;
    uint32_t acc = 0;
    for (uint32_t i = 0; i < loads_per_thread; i += 4) {
        uint64_t a[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            x ^= x >> 12; x ^= x << 25; x ^= x >> 27;
            a[u] = __umul64hi(x * 0x2545F4914F6CDD1DULL, n_bytes);
        }
#pragma unroll
        for (int u = 0; u < 4; u++) acc += __ldg(buf + a[u]);
    }
    if (acc == 0xFFFFFFFFu) *sink = acc;
}

}  // namespace orph
