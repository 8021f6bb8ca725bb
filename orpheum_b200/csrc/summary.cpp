// Host-side shares of the JSON summary (create_save_summary.py:136-240), computed from the result records without
// per-row objects in the caller's language: one call per batch for the read classes, the coding-frame histogram and the
// jaccard column, one call for the mean / standard deviation of the whole column.
//
// The reference takes mean and std with pandas / numpy, whose float64 sums are PAIRWISE sums with a fixed shape (numpy
// umath, pairwise_sum: blocks of at most 128 elements summed in eight interleaved accumulators, larger ranges split at
// n/2 rounded down to a multiple of 8).  The value of a floating-point sum depends on that shape, so it is restated here
// exactly -- and because the shape is a tree, its subtrees can be summed on different threads without changing a bit.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/orpheum_b200.h"
#include "hostpack.h"

extern int orph_set_error(int code, const char *fmt, ...);  // defined in orpheum_b200.cu

#pragma GCC optimize("fp-contract=off")  // (x - m) * (x - m) + acc must stay a multiply and an add, as numpy computes it

namespace {

constexpr uint64_t kBlock = 128;          // numpy's PW_BLOCKSIZE
constexpr uint64_t kTaskElems = 1u << 17;  // subtrees of at most this many elements are one unit of parallel work

// element i of the summed column: x[i] (mode 0) or (x[i] - shift)^2 (mode 1); 0 where ok[i] == 0
struct Column {
    const double *x;
    const uint8_t *ok;
    double shift;
    int square;
    inline double at(uint64_t i) const
    {
        if (ok && !ok[i]) return 0.0;
        if (!square) return x[i];
        const double d = x[i] - shift;
        return d * d;
    }
};

// numpy/_core/src/umath/loops_utils.h.src, @TYPE@_pairwise_sum, for a contiguous range
double pairwise(const Column &c, uint64_t a, uint64_t n)
{
    if (n < 8) {
        double res = 0.0;
        for (uint64_t i = 0; i < n; i++) res += c.at(a + i);
        return res;
    }
    if (n <= kBlock) {
        double r[8];
        for (int j = 0; j < 8; j++) r[j] = c.at(a + j);
        uint64_t i;
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += c.at(a + i + j);
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += c.at(a + i);
        return res;
    }
    uint64_t n2 = n / 2;
    n2 -= n2 % 8;
    return pairwise(c, a, n2) + pairwise(c, a + n2, n - n2);
}

struct Task {
    uint64_t a, n;
    double sum;
};

// the subtrees at which the recursion is cut, in the order the recursion visits them
void collect(uint64_t a, uint64_t n, std::vector<Task> &tasks)
{
    if (n <= kTaskElems) {
        tasks.push_back({a, n, 0.0});
        return;
    }
    uint64_t n2 = n / 2;
    n2 -= n2 % 8;
    collect(a, n2, tasks);
    collect(a + n2, n - n2, tasks);
}

double combine(uint64_t n, const std::vector<Task> &tasks, size_t &next)
{
    if (n <= kTaskElems) return tasks[next++].sum;
    uint64_t n2 = n / 2;
    n2 -= n2 % 8;
    const double left = combine(n2, tasks, next);
    return left + combine(n - n2, tasks, next);
}

double pairwise_parallel(const Column &c, uint64_t n, int n_threads)
{
    if (n <= kTaskElems) return pairwise(c, 0, n);
    std::vector<Task> tasks;
    collect(0, n, tasks);
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    orph_host::Pool pool(std::min<int>(n_threads, (int)tasks.size()));
    pool.run(tasks.size(), [&](uint64_t t) { tasks[t].sum = pairwise(c, tasks[t].a, tasks[t].n); });
    size_t next = 0;
    return combine(n, tasks, next);
}

}  // namespace

extern "C" int orph_f64_pairwise_sum(const double *x, const uint8_t *ok, uint64_t n, double shift, int square, int n_threads, double *out)
{
    if (!out || (n && !x)) return orph_set_error(ORPH_E_INVALID, "orph_f64_pairwise_sum: NULL argument");
    const Column c{x, ok, shift, square};
    *out = pairwise_parallel(c, n, n_threads);
    return ORPH_OK;
}

extern "C" int orph_summary_partial(const orph_read_result *records, uint64_t n_reads, double *jaccard, uint8_t *scored, uint64_t classes[10],
                                    uint64_t coding_hist[7], uint64_t first_with[7], int n_threads)
{
    if (!classes || !coding_hist || !first_with || (n_reads && (!records || !jaccard || !scored)))
        return orph_set_error(ORPH_E_INVALID, "orph_summary_partial: NULL argument");
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    constexpr uint64_t kChunk = 1u << 15;
    const uint64_t n_chunks = (n_reads + kChunk - 1) / kChunk;
    struct Part {
        uint64_t classes[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        uint64_t hist[7] = {0, 0, 0, 0, 0, 0, 0};
        uint64_t first[7] = {~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull, ~0ull};
    };
    std::vector<Part> parts(n_chunks);
    orph_host::Pool pool((int)std::min<uint64_t>((uint64_t)n_threads, std::max<uint64_t>(1, n_chunks)));
    pool.run(n_chunks, [&](uint64_t ck) {
        Part &p = parts[ck];
        const uint64_t r1 = std::min(n_reads, (ck + 1) * kChunk);
        for (uint64_t r = ck * kChunk; r < r1; r++) {
            const orph_read_result &rec = records[r];
            uint32_t present = 0, n_coding = 0;
            for (int f = 0; f < 6; f++) {
                const uint32_t cat = rec.category[f];
                present |= 1u << (cat & 31u);
                const bool sc = cat == ORPH_CAT_CODING || cat == ORPH_CAT_NON_CODING;
                n_coding += cat == ORPH_CAT_CODING;
                scored[6 * r + f] = sc ? 1 : 0;
                // the one IEEE division of translate.py:204
                jaccard[6 * r + f] = sc ? (double)rec.n_hits[f] / (double)rec.n_kmers[f] : std::nan("");
            }
            present &= 0x3Fu;
            // one class per read: Coding > too-short peptide > too-short read > the only category > Non-coding
            if (present & (1u << ORPH_CAT_CODING)) p.classes[0]++;
            else if (present & (1u << ORPH_CAT_TOO_SHORT_PEPTIDE)) p.classes[1]++;
            else if (present & (1u << ORPH_CAT_LOW_COMPLEXITY_NUCLEOTIDE)) p.classes[2]++;
            else if (present && (present & (present - 1)) == 0) p.classes[3 + __builtin_ctz(present)]++;
            else p.classes[9]++;
            if (n_coding) {
                p.hist[n_coding]++;
                if (p.first[n_coding] == ~0ull) p.first[n_coding] = r;
            }
        }
    });
    for (int i = 0; i < 10; i++) classes[i] = 0;
    for (int i = 0; i < 7; i++) {
        coding_hist[i] = 0;
        first_with[i] = ~0ull;
    }
    for (const Part &p : parts) {
        for (int i = 0; i < 10; i++) classes[i] += p.classes[i];
        for (int i = 0; i < 7; i++) {
            coding_hist[i] += p.hist[i];
            first_with[i] = std::min(first_with[i], p.first[i]);
        }
    }
    return ORPH_OK;
}
