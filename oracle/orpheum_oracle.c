/*
 * orpheum_oracle.c -- CPU restatement of the orpheum `translate` scoring path
 * and the `index` Bloom build.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/ (incl. the checker
 * scripts under tests/scripts/), the smoke check in __graft_entry__.py and
 * bench.py's cpu_baseline / --impl reference legs may load it.  The product (orpheum_b200/) never calls into this file; it fails
 * loudly when its CUDA library is missing.
 *
 * Parity status: PINNED.  tests/test_oracle.py check this file against
 *   - the reference's golden .nodegraph files (byte-identical rebuild),
 *   - the reference's golden score CSVs / coding FASTA for the 23 test reads,
 *   - the inline known-answer tests of the reference's unit tests,
 *   - vectors produced by running the UNMODIFIED reference python
 *     (oracle/gen_golden.py, run in the build container; tests/golden/).
 *
 * The arithmetic of two third-party native dependencies that are not vendored
 * in the reference tree is restated from their published algorithms:
 *   - sourmash==3.2.2  sourmash.minhash.hash_murmur  = MurmurHash3_x64_128
 *     (Austin Appleby, public domain), seed 42, low 64 bits (h1).
 *     reference call sites: orpheum/translate.py:187, orpheum/index.py:115
 *   - khmer==3.0.0a3 / 2.1.1  khmer.Nodegraph (oxli Hashgraph over BitStorage):
 *     n_tables bit arrays sized by the largest odd primes below `tablesize`,
 *     bin = hash % size, byte bin>>3, bit bin&7; get = AND, add = OR.
 *     reference call sites: orpheum/index.py:70-77,101,119,202-218,
 *     orpheum/translate.py:145,190,198
 *
 * Every function names the reference lines it follows (paths relative to the
 * reference checkout).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>

#define ORC_MAX_TABLES 16

/* category codes shared with include/orpheum_b200.h */
enum {
    ORC_CAT_STOP = 0,          /* "Translation frame has stop codon(s)"                   translate.py:218-226 */
    ORC_CAT_SHORT_PEPTIDE = 1, /* "Translation is shorter than peptide k-mer size + 1"    translate.py:227-237 */
    ORC_CAT_LOWCPLX_PEPTIDE = 2, /* "Low complexity peptide in {alphabet} alphabet"       translate.py:244-259 */
    ORC_CAT_CODING = 3,        /* "Coding"                                                translate.py:271-284 */
    ORC_CAT_NONCODING = 4,     /* "Non-coding"                                            translate.py:285-298 */
    ORC_CAT_LOWCPLX_NT = 5     /* "Read length was shorter than 3 * peptide k-mer size"   translate.py:301-331 (quirk) */
};

/* ------------------------------------------------------------------ */
/* MurmurHash3_x64_128, seed 42, h1  (sourmash hash_murmur)            */
/* ------------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

uint64_t orc_hash_murmur_seed(const uint8_t *data, uint64_t len, uint32_t seed)
{
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = seed, h2 = seed;
    uint64_t nblocks = len / 16;
    for (uint64_t i = 0; i < nblocks; i++) {
        uint64_t k1, k2;
        memcpy(&k1, data + 16 * i, 8);
        memcpy(&k2, data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    switch (len & 15) {
    case 15: k2 ^= (uint64_t)tail[14] << 48; /* fallthrough */
    case 14: k2 ^= (uint64_t)tail[13] << 40; /* fallthrough */
    case 13: k2 ^= (uint64_t)tail[12] << 32; /* fallthrough */
    case 12: k2 ^= (uint64_t)tail[11] << 24; /* fallthrough */
    case 11: k2 ^= (uint64_t)tail[10] << 16; /* fallthrough */
    case 10: k2 ^= (uint64_t)tail[9] << 8;   /* fallthrough */
    case 9:  k2 ^= (uint64_t)tail[8];
             k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; /* fallthrough */
    case 8:  k1 ^= (uint64_t)tail[7] << 56; /* fallthrough */
    case 7:  k1 ^= (uint64_t)tail[6] << 48; /* fallthrough */
    case 6:  k1 ^= (uint64_t)tail[5] << 40; /* fallthrough */
    case 5:  k1 ^= (uint64_t)tail[4] << 32; /* fallthrough */
    case 4:  k1 ^= (uint64_t)tail[3] << 24; /* fallthrough */
    case 3:  k1 ^= (uint64_t)tail[2] << 16; /* fallthrough */
    case 2:  k1 ^= (uint64_t)tail[1] << 8;  /* fallthrough */
    case 1:  k1 ^= (uint64_t)tail[0];
             k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= len; h2 ^= len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

/* hash_murmur(kmer) with sourmash's default seed (translate.py:187, index.py:115) */
uint64_t orc_hash_murmur(const uint8_t *data, uint64_t len)
{
    return orc_hash_murmur_seed(data, len, 42u);
}

/* ------------------------------------------------------------------ */
/* khmer.Nodegraph                                                     */
/* ------------------------------------------------------------------ */
typedef struct orc_nodegraph {
    uint32_t ksize;
    uint32_t n_tables;
    uint64_t sizes[ORC_MAX_TABLES];   /* in bits, primes, descending */
    uint8_t *tables[ORC_MAX_TABLES];  /* sizes[i]/8 + 1 bytes each */
    uint64_t n_unique_kmers;
} orc_nodegraph;

static int is_prime_u64(uint64_t n)
{
    if (n < 2) return 0;
    if (n < 4) return 1;
    if ((n & 1) == 0) return 0;
    for (uint64_t d = 3; d * d <= n; d += 2)
        if (n % d == 0) return 0;
    return 1;
}

/* khmer get_n_primes_near_x: the n largest odd primes strictly below x, descending. */
int orc_table_sizes(uint64_t tablesize, uint32_t n_tables, uint64_t *out)
{
    if (tablesize < 3) return -1;
    uint64_t i = tablesize - 1;
    if ((i & 1) == 0) i--;
    uint32_t found = 0;
    while (found < n_tables && i > 0) {
        if (is_prime_u64(i)) out[found++] = i;
        if (i < 2) break;
        i -= 2;
    }
    return found == n_tables ? 0 : -1;
}

void orc_ng_free(orc_nodegraph *g)
{
    if (!g) return;
    for (uint32_t i = 0; i < g->n_tables; i++) free(g->tables[i]);
    free(g);
}

orc_nodegraph *orc_ng_new_sizes(uint32_t ksize, uint32_t n_tables, const uint64_t *sizes)
{
    if (n_tables == 0 || n_tables > ORC_MAX_TABLES) return NULL;
    orc_nodegraph *g = (orc_nodegraph *)calloc(1, sizeof(*g));
    if (!g) return NULL;
    g->ksize = ksize;
    g->n_tables = n_tables;
    for (uint32_t i = 0; i < n_tables; i++) {
        g->sizes[i] = sizes[i];
        g->tables[i] = (uint8_t *)calloc(sizes[i] / 8 + 1, 1);
        if (!g->tables[i]) { orc_ng_free(g); return NULL; }
    }
    return g;
}

/* khmer.Nodegraph(ksize, tablesize, n_tables)   index.py:101 */
orc_nodegraph *orc_ng_new(uint32_t ksize, uint64_t tablesize, uint32_t n_tables)
{
    uint64_t sizes[ORC_MAX_TABLES];
    if (n_tables == 0 || n_tables > ORC_MAX_TABLES) return NULL;
    if (orc_table_sizes(tablesize, n_tables, sizes) != 0) return NULL;
    return orc_ng_new_sizes(ksize, n_tables, sizes);
}

uint32_t orc_ng_ksize(const orc_nodegraph *g) { return g->ksize; }
uint32_t orc_ng_n_tables(const orc_nodegraph *g) { return g->n_tables; }
uint64_t orc_ng_table_size(const orc_nodegraph *g, uint32_t i) { return g->sizes[i]; }
uint8_t *orc_ng_table(orc_nodegraph *g, uint32_t i) { return g->tables[i]; }
uint64_t orc_ng_n_unique_kmers(const orc_nodegraph *g) { return g->n_unique_kmers; }

/* Nodegraph.add(hash): OR the bit into every table; "new" if any bit was unset.  index.py:119 */
int orc_ng_add(orc_nodegraph *g, uint64_t h)
{
    int is_new = 0;
    for (uint32_t i = 0; i < g->n_tables; i++) {
        uint64_t bin = h % g->sizes[i];
        uint8_t bit = (uint8_t)(1u << (bin & 7));
        if (!(g->tables[i][bin >> 3] & bit)) {
            g->tables[i][bin >> 3] |= bit;
            is_new = 1;
        }
    }
    if (is_new) g->n_unique_kmers++;
    return is_new;
}

/* Nodegraph.get(hash): AND over tables, early exit.  translate.py:190.
 * *n_probes (optional) receives the number of tables looked at. */
int orc_ng_get_counted(const orc_nodegraph *g, uint64_t h, uint32_t *n_probes)
{
    uint32_t probes = 0;
    int present = 1;
    for (uint32_t i = 0; i < g->n_tables; i++) {
        uint64_t bin = h % g->sizes[i];
        probes++;
        if (!(g->tables[i][bin >> 3] & (1u << (bin & 7)))) { present = 0; break; }
    }
    if (n_probes) *n_probes = probes;
    return present;
}

int orc_ng_get(const orc_nodegraph *g, uint64_t h) { return orc_ng_get_counted(g, h, NULL); }

/* popcount of table i; table 0 is the OXLI header's occupied_bins */
uint64_t orc_ng_popcount(const orc_nodegraph *g, uint32_t i)
{
    uint64_t n = 0, nbytes = g->sizes[i] / 8 + 1;
    for (uint64_t b = 0; b < nbytes; b++) n += (uint64_t)__builtin_popcount(g->tables[i][b]);
    return n;
}

/* OXLI v4 nodegraph file (khmer save/load).  index.py:206,218 */
int orc_ng_save(const orc_nodegraph *g, const char *path)
{
    FILE *f = fopen(path, "wb");
    if (!f) return -1;
    uint8_t hdr[19];
    memcpy(hdr, "OXLI", 4);
    hdr[4] = 4; hdr[5] = 2;
    uint32_t k = g->ksize;
    memcpy(hdr + 6, &k, 4);
    hdr[10] = (uint8_t)g->n_tables;
    uint64_t occ = orc_ng_popcount(g, 0);
    memcpy(hdr + 11, &occ, 8);
    int ok = fwrite(hdr, 1, 19, f) == 19;
    for (uint32_t i = 0; ok && i < g->n_tables; i++) {
        uint64_t sz = g->sizes[i], nbytes = sz / 8 + 1;
        ok = fwrite(&sz, 8, 1, f) == 1 && fwrite(g->tables[i], 1, nbytes, f) == nbytes;
    }
    fclose(f);
    return ok ? 0 : -1;
}

/* khmer.Nodegraph.load / khmer.load_nodegraph   index.py:70-77 */
orc_nodegraph *orc_ng_load(const char *path)
{
    FILE *f = fopen(path, "rb");
    if (!f) return NULL;
    uint8_t hdr[19];
    orc_nodegraph *g = NULL;
    if (fread(hdr, 1, 19, f) != 19 || memcmp(hdr, "OXLI", 4) != 0 || hdr[4] != 4 || hdr[5] != 2) goto fail;
    uint32_t k; memcpy(&k, hdr + 6, 4);
    uint32_t nt = hdr[10];
    if (nt == 0 || nt > ORC_MAX_TABLES) goto fail;
    g = (orc_nodegraph *)calloc(1, sizeof(*g));
    if (!g) goto fail;
    g->ksize = k; g->n_tables = nt;
    for (uint32_t i = 0; i < nt; i++) {
        uint64_t sz;
        if (fread(&sz, 8, 1, f) != 1) goto fail;
        uint64_t nbytes = sz / 8 + 1;
        g->sizes[i] = sz;
        g->tables[i] = (uint8_t *)malloc(nbytes);
        if (!g->tables[i] || fread(g->tables[i], 1, nbytes, f) != nbytes) goto fail;
    }
    fclose(f);
    return g;
fail:
    fclose(f);
    orc_ng_free(g);
    return NULL;
}

/* ------------------------------------------------------------------ */
/* distinct k-mers of a byte string: kmerize()  compare_kmer_content.py:116-118 */
/* ------------------------------------------------------------------ */
typedef struct { const uint8_t *p; } kmer_ref;
static __thread uint32_t g_cmp_k;
static int cmp_kmer(const void *a, const void *b)
{
    return memcmp(((const kmer_ref *)a)->p, ((const kmer_ref *)b)->p, g_cmp_k);
}

/* Fills refs[] (capacity len-k+1) with one pointer per DISTINCT k-mer; returns the count. */
static uint64_t distinct_kmers(const uint8_t *seq, uint64_t len, uint32_t k, kmer_ref *refs)
{
    if (len < k) return 0;
    uint64_t n = len - k + 1;
    for (uint64_t i = 0; i < n; i++) refs[i].p = seq + i;
    g_cmp_k = k;
    qsort(refs, n, sizeof(kmer_ref), cmp_kmer);
    uint64_t m = 0;
    for (uint64_t i = 0; i < n; i++)
        if (i == 0 || memcmp(refs[i].p, refs[m - 1].p, k) != 0) refs[m++] = refs[i];
    return m;
}

uint64_t orc_count_distinct_kmers(const uint8_t *seq, uint64_t len, uint32_t k)
{
    if (len < k) return 0;
    kmer_ref *refs = (kmer_ref *)malloc((len - k + 1) * sizeof(kmer_ref));
    uint64_t m = distinct_kmers(seq, len, k, refs);
    free(refs);
    return m;
}

/* ------------------------------------------------------------------ */
/* index build: make_peptide_bloom_filter inner loop  index.py:107-122 */
/* ------------------------------------------------------------------ */
/* One peptide record: skipped entirely if it contains '*' (index.py:108-109);
 * re-encoded through lut (encode_peptide, sequence_encodings.py:457-468); every
 * distinct k-mer hashed and added (index.py:112-119).  Returns #k-mers added. */
uint64_t orc_ng_add_peptide(orc_nodegraph *g, const uint8_t *seq, uint64_t len, const uint8_t *lut)
{
    uint32_t k = g->ksize;
    if (memchr(seq, '*', len)) return 0;
    if (len < k) return 0;
    uint8_t *enc = (uint8_t *)malloc(len ? len : 1);
    for (uint64_t i = 0; i < len; i++) enc[i] = lut[seq[i]];
    kmer_ref *refs = (kmer_ref *)malloc((len - k + 1) * sizeof(kmer_ref));
    uint64_t m = distinct_kmers(enc, len, k, refs);
    /* python iterates the set in arbitrary order; the bits are order independent, and
     * n_unique_kmers depends on order only through Bloom collisions (the reference's own test
     * allows rtol 1e-3).  Sorted order keeps this deterministic. */
    for (uint64_t i = 0; i < m; i++) orc_ng_add(g, orc_hash_murmur(refs[i].p, k));
    free(refs);
    free(enc);
    return m;
}

void orc_ng_add_peptides(orc_nodegraph *g, const uint8_t *residues, const uint64_t *offsets,
                         uint64_t n_records, const uint8_t *lut)
{
    for (uint64_t r = 0; r < n_records; r++)
        orc_ng_add_peptide(g, residues + offsets[r], offsets[r + 1] - offsets[r], lut);
}

/* ------------------------------------------------------------------ */
/* translation: translate_single_seq.py:13-78, constants_translate.py:59-141 */
/* ------------------------------------------------------------------ */
/* STANDARD_CODON_TABLE_MAPPING.get(codon, "X"): 64 ACGT codons plus the seven xxN entries. */
static uint8_t translate_codon(uint8_t a, uint8_t b, uint8_t c)
{
    /* index order TCAG per position, the classic NCBI table-1 string */
    static const char AAS[] = "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
    int ia = a == 'T' ? 0 : a == 'C' ? 1 : a == 'A' ? 2 : a == 'G' ? 3 : -1;
    int ib = b == 'T' ? 0 : b == 'C' ? 1 : b == 'A' ? 2 : b == 'G' ? 3 : -1;
    int ic = c == 'T' ? 0 : c == 'C' ? 1 : c == 'A' ? 2 : c == 'G' ? 3 : -1;
    if (ia < 0 || ib < 0) return 'X';
    if (ic >= 0) return (uint8_t)AAS[ia * 16 + ib * 4 + ic];
    if (c == 'N') {
        /* ACN T, CTN L, CGN R, GTN V, GCN A, GGN G, TCN S -- and notably no CCN */
        if (a == 'A' && b == 'C') return 'T';
        if (a == 'C' && b == 'T') return 'L';
        if (a == 'C' && b == 'G') return 'R';
        if (a == 'G' && b == 'T') return 'V';
        if (a == 'G' && b == 'C') return 'A';
        if (a == 'G' && b == 'G') return 'G';
        if (a == 'T' && b == 'C') return 'S';
    }
    return 'X';
}

uint8_t orc_translate_codon(const uint8_t *codon) { return translate_codon(codon[0], codon[1], codon[2]); }

/* REVERSE_COMPLEMENT_TABLE: only ACGT are mapped, everything else unchanged. */
static uint8_t complement(uint8_t c)
{
    switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A'; default: return c; }
}

void orc_reverse_complement(const uint8_t *seq, uint64_t len, uint8_t *out)
{
    for (uint64_t i = 0; i < len; i++) out[i] = complement(seq[len - 1 - i]);
}

/* _single_seq_translation(seq, frame): i in range(int(len/3)), 3-char codons only.  Returns peptide length. */
static uint64_t translate_frame(const uint8_t *seq, uint64_t len, uint32_t frame, uint8_t *out)
{
    uint64_t n = 0;
    for (uint64_t i = 0; i < len / 3; i++) {
        uint64_t start = frame + 3 * i;
        if (start + 3 <= len) out[n++] = translate_codon(seq[start], seq[start + 1], seq[start + 2]);
    }
    return n;
}

/* six_frame_translation(): frames in key order 1,2,3,-1,-2,-3.  out must hold 6*(len/3+1) bytes;
 * frame f's peptide starts at out + f*(len/3+1); lens[f] receives its length. */
void orc_six_frame_translation(const uint8_t *seq, uint64_t len, uint8_t *out, uint64_t *lens)
{
    uint64_t stride = len / 3 + 1;
    uint8_t *rc = (uint8_t *)malloc(len ? len : 1);
    orc_reverse_complement(seq, len, rc);
    for (uint32_t f = 0; f < 3; f++) {
        lens[f] = translate_frame(seq, len, f, out + f * stride);
        lens[3 + f] = translate_frame(rc, len, f, out + (3 + f) * stride);
    }
    free(rc);
}

/* ------------------------------------------------------------------ */
/* scoring                                                             */
/* ------------------------------------------------------------------ */
typedef struct orc_frame_score {
    double jaccard;     /* NaN unless category is coding / non-coding */
    int64_t n_kmers;    /* -1 where the reference reports NaN */
    int64_t n_hits;     /* -1 where not scored */
    int32_t category;
    int32_t frame;      /* 1,2,3,-1,-2,-3 */
    uint64_t probes_min; /* table look-ups under early-exit AND semantics (SURVEY 8d P_min) */
    uint64_t probes_max; /* n_tables * distinct k-mers */
} orc_frame_score;

/* compute_fastp_complexity: #(seq[i] != seq[i+1]) -- the caller divides.  translate.py:79-84 */
uint64_t orc_fastp_ndiff(const uint8_t *seq, uint64_t len)
{
    uint64_t n = 0;
    for (uint64_t i = 0; i + 1 < len; i++) n += seq[i] != seq[i + 1];
    return n;
}

/* Translate.maybe_score_single_read -> check_nucleotide_content / check_peptide_content
 * (translate.py:208-331).  Returns 0, or -1 for an empty read (the reference raises
 * ZeroDivisionError at translate.py:83). */
int orc_score_read(const orc_nodegraph *g, const uint8_t *seq, uint64_t len, const uint8_t *lut,
                   double jaccard_threshold, orc_frame_score *out)
{
    static const int32_t FRAMES[6] = {1, 2, 3, -1, -2, -3};
    uint32_t k = g->ksize;
    if (len == 0) return -1;
    for (int f = 0; f < 6; f++) {
        out[f].jaccard = NAN; out[f].n_kmers = -1; out[f].n_hits = -1;
        out[f].frame = FRAMES[f]; out[f].probes_min = 0; out[f].probes_max = 0;
    }
    /* evaluate_is_fastp_low_complexity: ndiff / len < 0.3   (translate.py:56-84) */
    double complexity = (double)orc_fastp_ndiff(seq, len) / (double)len;
    if (complexity < 0.3) {
        /* check_nucleotide_content(description, nan, sequence): nan > 0 is False (translate.py:308-321) */
        for (int f = 0; f < 6; f++) out[f].category = ORC_CAT_LOWCPLX_NT;
        return 0;
    }
    uint64_t stride = len / 3 + 1;
    uint8_t *pep = (uint8_t *)malloc(6 * stride);
    uint8_t *enc = (uint8_t *)malloc(stride);
    kmer_ref *refs = (kmer_ref *)malloc(stride * sizeof(kmer_ref));
    uint64_t lens[6];
    orc_six_frame_translation(seq, len, pep, lens);
    for (int f = 0; f < 6; f++) {
        const uint8_t *t = pep + f * stride;
        uint64_t tl = lens[f];
        if (memchr(t, '*', tl)) { out[f].category = ORC_CAT_STOP; continue; }          /* translate.py:218 */
        if (tl <= k) { out[f].category = ORC_CAT_SHORT_PEPTIDE; continue; }             /* translate.py:227 */
        for (uint64_t i = 0; i < tl; i++) enc[i] = lut[t[i]];                           /* translate.py:239 */
        uint64_t n = distinct_kmers(enc, tl, k, refs);                                  /* translate.py:186 */
        uint64_t hits = 0;
        for (uint64_t i = 0; i < n; i++) {                                              /* translate.py:187-191 */
            uint32_t probes;
            hits += (uint64_t)orc_ng_get_counted(g, orc_hash_murmur(refs[i].p, k), &probes);
            out[f].probes_min += probes;
        }
        out[f].probes_max = n * g->n_tables;
        double jaccard = (double)hits / (double)n;                                      /* translate.py:204 */
        out[f].n_kmers = (int64_t)n;
        out[f].n_hits = (int64_t)hits;
        /* evaluate_is_kmer_low_complexity: n_kmers <= (len - k + 1) / 2  (translate.py:87-107) */
        if ((double)n <= (double)(tl - k + 1) / 2.0) {
            out[f].category = ORC_CAT_LOWCPLX_PEPTIDE;                                  /* jaccard stays NaN, n_kmers kept */
        } else if (jaccard > jaccard_threshold) {                                       /* translate.py:271 */
            out[f].category = ORC_CAT_CODING; out[f].jaccard = jaccard;
        } else {
            out[f].category = ORC_CAT_NONCODING; out[f].jaccard = jaccard;
        }
    }
    free(refs); free(enc); free(pep);
    return 0;
}

/* Batch form over concatenated reads, pthreads over dynamic chunks of reads (the reference
 * itself is single threaded, translate.py:367-379; threads only make the CPU baseline generous).
 * Returns the number of empty reads encountered (their rows get category -1). */
typedef struct {
    const orc_nodegraph *g; const uint8_t *reads; const uint64_t *offsets; uint64_t n_reads;
    const uint8_t *lut; double thr; orc_frame_score *out;
    uint64_t next; int64_t n_empty; pthread_mutex_t mu;
} batch_job;

static void *batch_worker(void *arg)
{
    batch_job *j = (batch_job *)arg;
    const uint64_t CHUNK = 256;
    int64_t empty = 0;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        uint64_t lo = j->next;
        j->next = lo + CHUNK;
        pthread_mutex_unlock(&j->mu);
        if (lo >= j->n_reads) break;
        uint64_t hi = lo + CHUNK < j->n_reads ? lo + CHUNK : j->n_reads;
        for (uint64_t r = lo; r < hi; r++) {
            int rc = orc_score_read(j->g, j->reads + j->offsets[r], j->offsets[r + 1] - j->offsets[r], j->lut,
                                    j->thr, j->out + 6 * r);
            if (rc != 0) {
                empty++;
                for (int f = 0; f < 6; f++) j->out[6 * r + f].category = -1;
            }
        }
    }
    pthread_mutex_lock(&j->mu);
    j->n_empty += empty;
    pthread_mutex_unlock(&j->mu);
    return NULL;
}

int orc_max_threads(void)
{
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int64_t orc_score_reads(const orc_nodegraph *g, const uint8_t *reads, const uint64_t *offsets, uint64_t n_reads,
                        const uint8_t *lut, double jaccard_threshold, orc_frame_score *out, int n_threads)
{
    batch_job job = {g, reads, offsets, n_reads, lut, jaccard_threshold, out, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    if (n_threads <= 0) n_threads = orc_max_threads();
    if (n_threads > 256) n_threads = 256;
    if (n_threads == 1) { batch_worker(&job); return job.n_empty; }
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < n_threads; t++)
        if (pthread_create(&th[started], NULL, batch_worker, &job) == 0) started++;
    if (started == 0) batch_worker(&job);
    for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
    return job.n_empty;
}
