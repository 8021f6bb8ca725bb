/*
 * orpheum_b200.h -- C ABI of the B200-native orpheum `translate` / `index` hot path.
 *
 * The reference (czbiohub-sf/orpheum) is pure Python and has no FFI of its own; its per-read
 * arithmetic lives behind two third-party native calls (sourmash `hash_murmur`, khmer
 * `Nodegraph.add/get`).  This library replaces the whole per-read body -- everything under
 * `Translate.score_reads_per_file` (orpheum/translate.py:367-372) and the inner loop of
 * `make_peptide_bloom_filter` (orpheum/index.py:105-122) -- with sm_100a CUDA kernels.  Each entry
 * point below names the reference interface it stands in for.  INTEGRATION.md shows the ctypes
 * binding a maintainer of the reference would add.
 *
 * Conventions
 *   - Plain C, no C++/torch types.  Every function returns 0 (ORPH_OK) or a negative orph_status;
 *     orph_last_error() returns a thread-local message for the last failure on this thread.
 *   - The caller owns every host array it passes in; the library owns device memory behind the
 *     opaque handles.  Nothing returned needs freeing except handles (`*_destroy`).
 *   - One orph_ctx per GPU per host thread.  Filters are immutable while scoring and may be shared
 *     by several contexts on the same device.
 *   - "host" entry points take host pointers and include the H2D/D2H copies; "_device" entry points
 *     take device pointers (e.g. torch tensors' data_ptr()) and only enqueue work on the context's
 *     stream -- call orph_ctx_synchronize (or synchronize the stream yourself) before reading.
 *   - Frame order in every [6] array is 1, 2, 3, -1, -2, -3 (translate_single_seq.py:70-78,
 *     translate.py:320).
 */
#ifndef ORPHEUM_B200_H
#define ORPHEUM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORPH_ABI_VERSION 2
#define ORPH_MAX_TABLES 16

typedef enum orph_status {
    ORPH_OK = 0,
    ORPH_E_INVALID = -1,    /* bad argument */
    ORPH_E_CUDA = -2,       /* CUDA runtime failure (message has the CUDA error string) */
    ORPH_E_NOMEM = -3,      /* host or device allocation failed */
    ORPH_E_TOO_LONG = -4,   /* a read longer than ORPH_MAX_READ_LEN has a stop-free frame (or a read exceeds the declared bound) */
    ORPH_E_EMPTY_READ = -5, /* a zero-length read: the reference raises ZeroDivisionError (translate.py:83) */
    ORPH_E_IO = -6
} orph_status;

/* Per-frame category codes.  Text forms: constants_translate.py:27-44. */
typedef enum orph_category {
    ORPH_CAT_STOP_CODONS = 0,       /* "Translation frame has stop codon(s)"                 translate.py:218-226 */
    ORPH_CAT_TOO_SHORT_PEPTIDE = 1, /* "Translation is shorter than peptide k-mer size + 1"  translate.py:227-237 */
    ORPH_CAT_LOW_COMPLEXITY_PEPTIDE = 2, /* "Low complexity peptide in {alphabet} alphabet"  translate.py:244-259 */
    ORPH_CAT_CODING = 3,            /* "Coding"                                              translate.py:271-284 */
    ORPH_CAT_NON_CODING = 4,        /* "Non-coding"                                          translate.py:285-298 */
    ORPH_CAT_LOW_COMPLEXITY_NUCLEOTIDE = 5, /* fastp-low-complexity read; the reference labels it "Read length was
                                       shorter than 3 * peptide k-mer size" (translate.py:301-331 quirk) */
    ORPH_CAT_INVALID = 255          /* empty read (no reference result exists) */
} orph_category;

/* One record per read, 32 bytes.  For categories 0, 1 and 5 the reference reports NaN for both
 * jaccard and n_kmers (counts are 0 here); for category 2 jaccard is NaN but n_kmers is kept;
 * for 3 and 4 jaccard = n_hits / n_kmers (one IEEE double division, done by the caller or by
 * orph_jaccard()).  Replaces the 6 SingleReadScore tuples of Translate.maybe_score_single_read
 * (translate.py:323-331, constants_translate.py:46-54). */
typedef struct orph_read_result {
    uint16_t n_kmers[6];  /* distinct peptide k-mers of the frame (translate.py:186,188) */
    uint16_t n_hits[6];   /* of those, how many are in the Bloom filter (translate.py:189-191) */
    uint8_t category[6];  /* orph_category */
    int8_t best_frame;    /* frame (1,2,3,-1,-2,-3) with the maximum jaccard among categories 3/4, exact rational
                             comparison, first in frame order on ties; 0 if no frame was scored (README.md:49-54) */
    uint8_t flags;        /* ORPH_FLAG_* */
} orph_read_result;

#define ORPH_FLAG_OPEN_MASK 0x3f      /* bit f: frame f had no stop codon and was longer than k (was scored) */
#define ORPH_FLAG_LOW_COMPLEXITY 0x40 /* fastp nucleotide complexity < 0.3 (translate.py:56-84) */
#define ORPH_FLAG_BYTE_PATH 0x80      /* read went through a byte-exact kernel (lower case / IUPAC bytes, N in a read longer than
                                         320 nt, reads beyond the packed kernels' length); informational: the other fields do
                                         not depend on the route */

/* Input encodings of read batches. */
typedef enum orph_reads_format {
    ORPH_READS_ASCII = 0,  /* concatenated raw sequence bytes exactly as screed yields them (any byte, any case) */
    ORPH_READS_PACKED2 = 1 /* 2 bits per base, base g of the concatenated stream at bits [2*(g%32), +2) of
                              little-endian u64 word g/32; code = (ascii >> 1) & 3, i.e. A=0 C=1 T=2 G=3.
                              Only upper-case ACGT reads can be represented: the caller routes other reads
                              through an ASCII batch (orph_pack_reads reports which). */
} orph_reads_format;

#define ORPH_MAX_READ_LEN 196607u /* longest read whose frames can be SCORED: a frame of 65535 residues is the most whose
                                     distinct k-mer count fits the record's 16-bit fields.  Longer reads (contigs, transcripts;
                                     the reference has no limit, translate.py:323-331) are still classified exactly whenever
                                     none of their frames needs scoring -- the read is low-complexity or every frame holds a
                                     stop codon, i.e. practically always.  Only a longer read with a stop-free frame gives
                                     ORPH_E_TOO_LONG; its record is ORPH_CAT_INVALID, every other read's record is filled. */

typedef struct orph_ctx orph_ctx;
typedef struct orph_filter orph_filter;

/* ---- devices -------------------------------------------------------------------------------- */
/* Number of CUDA devices this process can use (0 when there is none or the driver cannot be reached; never an error):
 * what Translate asks before it spreads the lanes of a scoring stream over the GPUs of the box (the reference walks its
 * files and reads serially, orpheum/translate.py:367-379). */
int orph_device_count(void);

/* ---- context ------------------------------------------------------------------------------- */
/* `stream`: a cudaStream_t to enqueue on (e.g. torch.cuda.current_stream().cuda_stream), or NULL to
 * let the context create its own non-blocking stream. */
int orph_ctx_create(int device, void *stream, orph_ctx **out);
void orph_ctx_destroy(orph_ctx *ctx);
int orph_ctx_synchronize(orph_ctx *ctx);
/* Returns (and clears) the sticky device-side status of work enqueued so far: ORPH_OK,
 * ORPH_E_EMPTY_READ or ORPH_E_TOO_LONG.  Synchronizes the stream. */
int orph_ctx_check(orph_ctx *ctx);
/* Test hook: force the exact all-pairs de-duplication path in the scoring kernels (normally only
 * taken on a 64-bit MurmurHash collision between two different k-mers of one frame). */
int orph_ctx_set_force_exact_dedup(orph_ctx *ctx, int on);
/* Test hook: on = 0 routes short reads through the warp-per-read kernel instead of the thread-per-frame
 * kernel (both must give identical results). */
int orph_ctx_set_thread_path(orph_ctx *ctx, int on);
/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
uint64_t orph_ctx_launch_count(const orph_ctx *ctx);
/* Per-kernel device timing (CUDA events on the context's stream around every scoring kernel), for
 * bench.py's roofline.  Kernel ids: 0 pack_ascii, 1 prefilter, 2 score_frames (warp per read),
 * 3 score_bytes, 4 score_tasks (thread per frame), 5 finalize_best, 6 task sort (scan + scatter),
 * 7 score_long (block per read). */
#define ORPH_N_TIMED_KERNELS 8
int orph_ctx_set_kernel_timing(orph_ctx *ctx, int on);
/* Adds up (and clears) the intervals recorded so far: ms[i] total milliseconds and launches[i] count of
 * kernel i.  Synchronizes the stream. */
int orph_ctx_kernel_times(orph_ctx *ctx, double *ms, uint64_t *launches);
/* Host ingest of orph_score_reads (ASCII batches): the call is PCIe-bound at 1 byte per base, so the library packs part
 * of each large batch to 2 bits per base on `n_threads` host threads while the copy engine ships the rest as ASCII
 * (a pinned input buffer is needed for that overlap; a pageable one is always packed).  n_threads < 0 (default): the
 * cores this process may run on, divided by LOCAL_WORLD_SIZE when that is set; 0: never pack on the host.  Results do
 * not depend on the setting. */
int orph_ctx_set_host_threads(orph_ctx *ctx, int n_threads);
/* Counters of the ingest pipeline since the context was created (any output may be NULL): chunks packed on the host,
 * chunks sent as ASCII, size of the host pool, measured host packing rate in input bytes/s, measured H2D rate. */
int orph_ctx_ingest_stats(orph_ctx *ctx, uint64_t *chunks_packed_on_host, uint64_t *chunks_sent_ascii, int *host_threads,
                          double *pack_bytes_per_s, double *h2d_bytes_per_s);
const char *orph_last_error(void);
int orph_abi_version(void);

/* ---- Bloom filter: khmer.Nodegraph ---------------------------------------------------------- */
/* khmer's table sizing: the n_tables largest odd primes below `tablesize`, descending
 * (khmer.Nodegraph(k, tablesize, n_tables), index.py:101). */
int orph_table_sizes(uint64_t tablesize, uint32_t n_tables, uint64_t *out_sizes);

/* Empty filter with explicit table sizes (bits).  khmer.Nodegraph ctor, index.py:101. */
int orph_filter_create(orph_ctx *ctx, uint32_t ksize, uint32_t n_tables, const uint64_t *table_sizes,
                       orph_filter **out);
/* Filter from host tables laid out exactly as in a .nodegraph file: table i is table_sizes[i]/8+1
 * bytes, bit b at byte b>>3, mask 1<<(b&7).  khmer load path, index.py:70-77. */
int orph_filter_from_host(orph_ctx *ctx, uint32_t ksize, uint32_t n_tables, const uint64_t *table_sizes,
                          const uint8_t *const *tables, orph_filter **out);
/* Copies the tables back (same layout) and returns popcount(table 0), the OXLI header's
 * occupied_bins field.  Either output may be NULL.  Nodegraph.save, index.py:206,218. */
int orph_filter_to_host(const orph_filter *f, uint8_t *const *tables, uint64_t *occupied_bins_table0);
void orph_filter_destroy(orph_filter *f);
/* Geometry; Nodegraph.ksize() (translate.py:145, index.py:164,202).  Outputs may be NULL. */
int orph_filter_info(const orph_filter *f, uint32_t *ksize, uint32_t *n_tables, uint64_t *table_sizes);
/* Nodegraph.n_unique_kmers() (tests/test_index.py:54,82): number of add operations that set at least
 * one new bit.  Order dependent in khmer, therefore only approximately reproducible under parallel
 * insertion (the reference checks it to rtol 1e-3); the bits themselves are exact. */
uint64_t orph_filter_n_unique_kmers(const orph_filter *f);
/* popcount of one table (table 0 = occupied_bins). */
int orph_filter_popcount(const orph_filter *f, uint32_t table, uint64_t *out);
/* The single contiguous device allocation that holds all tables (table i at byte offset
 * table_offsets[i], 256-byte aligned), for replication with ncclBroadcast / cudaMemcpyPeer by the
 * caller's plumbing.  A filter created with the same geometry on another device has the same layout. */
int orph_filter_device_buffer(const orph_filter *f, void **device_ptr, uint64_t *n_bytes, uint64_t *table_offsets);
/* Single-process replication onto another context's device over NVLink (cudaMemcpyPeerAsync). */
int orph_filter_replicate(const orph_filter *src, orph_ctx *dst_ctx, orph_filter **out);

/* The inner loop of index.make_peptide_bloom_filter (index.py:107-122): for each record, skip it if it
 * contains '*', re-encode through lut (encode_peptide), hash every k-mer (MurmurHash3_x64_128 seed 42,
 * low 64 bits) and OR bit (hash % size_i) into every table.  residues/offsets are HOST arrays
 * (offsets has n_records+1 entries).  n_unique_kmers_inc (optional) receives the increment of
 * Nodegraph.n_unique_kmers(); passing NULL skips the de-duplication set that exact count needs (the bits are
 * the same) and orph_filter_n_unique_kmers() then over-counts k-mers that were inserted concurrently. */
int orph_filter_add_peptides(orph_filter *f, const uint8_t *residues, const uint64_t *offsets, uint64_t n_records,
                             const uint8_t lut[256], uint64_t *n_unique_kmers_inc);
/* Nodegraph.add(hash) for a batch of precomputed hashes (index.py:119); is_new (optional) gets 0/1.  The bits are exact;
 * is_new / n_unique_kmers are what khmer reports only for batches without duplicates and without k-mers racing for the same
 * new bit (concurrent insertion: two such hashes can both see themselves as new). */
int orph_filter_add_hashes(orph_filter *f, const uint64_t *hashes, uint64_t n, uint8_t *is_new);
/* Nodegraph.get(hash) for a batch of hashes (translate.py:190); out01[i] = 0/1.  Host arrays. */
int orph_filter_query_hashes(const orph_filter *f, const uint64_t *hashes, uint64_t n, uint8_t *out01);
/* sourmash.minhash.hash_murmur for a batch of equal-length keys, computed on the GPU (test hook for the
 * device hash; translate.py:187).  keys is n*len bytes. */
int orph_hash_kmers(orph_ctx *ctx, const uint8_t *keys, uint32_t len, uint64_t n, uint64_t *out_hashes);

/* ---- scoring: Translate.score_reads_per_file's per-read body -------------------------------- */
/* Host-buffer form (the drop-in seam, translate.py:323-372).  reads/offsets/out are HOST arrays;
 * offsets has n_reads+1 entries, in bases, offsets[0] may be non-zero.  For ORPH_READS_ASCII every read
 * gets the reference's byte-exact semantics (N, lower case, ... included).  Includes H2D, all kernels,
 * D2H and a stream synchronize.  Returns ORPH_E_EMPTY_READ / ORPH_E_TOO_LONG after filling every other
 * read's record if such reads were present. */
int orph_score_reads(orph_ctx *ctx, const orph_filter *f, const void *reads, int reads_format,
                     const uint64_t *offsets, uint64_t n_reads, const uint8_t lut[256], double jaccard_threshold,
                     orph_read_result *out);
/* Device-buffer form: same contract, but reads/offsets/out are DEVICE pointers (reads 16-byte aligned)
 * and the call only enqueues on the context's stream.  max_read_len is an upper bound on the read
 * length in this batch: it sizes the per-read scratch of the kernels and decides which of them are launched
 * (a read longer than the declared bound is reported through ORPH_E_TOO_LONG); pass 0 if unknown (assumes
 * ORPH_MAX_READ_LEN, which costs scratch memory and idle launches).  Use orph_ctx_check() for the deferred
 * status. */
int orph_score_reads_device(orph_ctx *ctx, const orph_filter *f, const void *d_reads, int reads_format,
                            const uint64_t *d_offsets, uint64_t n_reads, uint32_t max_read_len,
                            const uint8_t lut[256], double jaccard_threshold, orph_read_result *d_out);
/* 2-bit reads WITH N: ORPH_READS_PACKED2 plus an N plane -- bit g % 32 of 32-bit word g / 32 of d_nmask is set where base g
 * of the stream is an upper-case 'N' (its 2-bit code is then ignored) -- and d_has_n[i] = 1 for every read that holds one.
 * Reads with N translate as in the reference (translate_single_seq.py:13-32 over constants_translate.py:59-131: a codon with
 * an N is X unless it is one of ACN CTN CGN GTN GCN GGN TCN) on the same kernels as plain reads; only bytes other than ACGTN
 * (lower case, IUPAC codes) still need the ASCII form.  Same contract as orph_score_reads_device otherwise. */
int orph_score_reads_device_n(orph_ctx *ctx, const orph_filter *f, const void *d_packed, const uint32_t *d_nmask,
                              const uint8_t *d_has_n, const uint64_t *d_offsets, uint64_t n_reads, uint32_t max_read_len,
                              const uint8_t lut[256], double jaccard_threshold, orph_read_result *d_out);
/* Builds that form on the device from ASCII reads (all DEVICE pointers; d_offsets[0] must be 0, total_bases =
 * d_offsets[n_reads]): d_packed_out ((total_bases + 31) / 32 u64 words), d_nmask_out (as many u32 words), d_has_n_out and
 * d_needs_bytes_out (n_reads bytes each; needs_bytes[i] = 1: read i holds a byte other than ACGTN and must be scored from
 * its ASCII form).  Only enqueues on the context's stream. */
int orph_pack_reads_device(orph_ctx *ctx, const uint8_t *d_ascii, const uint64_t *d_offsets, uint64_t n_reads,
                           uint64_t total_bases, uint64_t *d_packed_out, uint32_t *d_nmask_out, uint8_t *d_has_n_out,
                           uint8_t *d_needs_bytes_out);
/* Host helper for callers that want to ship 2-bit reads over PCIe: packs the concatenated ASCII stream
 * [offsets[0], offsets[n_reads]) into out_packed (ceil(total/32) u64 words) and sets needs_ascii[i]=1
 * for every read that contains a byte other than upper-case ACGT (its packed bits are unspecified).
 * Multi-threaded on the host; n_threads <= 0 picks the hardware concurrency. */
int orph_pack_reads(const uint8_t *reads, const uint64_t *offsets, uint64_t n_reads, uint64_t *out_packed,
                    uint8_t *needs_ascii, int n_threads);
/* The peptides the reference emits (stdout FASTA of coding frames, low-complexity-peptide FASTA;
 * translate.py:244-249,271-273): the UN-encoded translation of frame item_frame[i] (1,2,3,-1,-2,-3) of read
 * item_read[i], with the reference's codon semantics (translate_single_seq.py:13-42: 64 ACGT codons, the seven
 * xxN codons, everything else 'X', '*' for stops).  Host arrays; the caller pre-computes out_offsets (n_items+1,
 * item i gets (len(read) - (|frame|-1)) / 3 bytes) and provides out_peptides of out_offsets[n_items] bytes. */
int orph_translate_frames(orph_ctx *ctx, const uint8_t *reads, const uint64_t *offsets, uint64_t n_reads,
                          const uint64_t *item_read, const int8_t *item_frame, uint64_t n_items,
                          const uint64_t *out_offsets, uint8_t *out_peptides);
/* jaccard of frame i of a record exactly as the reference computes it: n_hits / n_kmers in IEEE double
 * for categories 3/4, NaN otherwise (translate.py:204). */
double orph_jaccard(const orph_read_result *r, int frame_index);

/* ---- host ingest: FASTA / FASTQ (plain or gzip) with screed's record semantics ------------------- */
/* Stands in for `screed.open(path)` (translate.py:370, index.py:83): name = the whole header line, sequence
 * untouched, empty file = no records.  ORPH_E_INVALID if the file is neither FASTA nor FASTQ (the reference
 * sees screed's ValueError), ORPH_E_IO if it cannot be read or a FASTQ record is malformed. */
typedef struct orph_fastx orph_fastx;
int orph_fastx_open(const char *path, orph_fastx **out);
/* Parses up to max_records records / about max_bases bases and exposes them in the scoring ABI's layout:
 * concatenated sequence bytes + offsets[n+1], concatenated names + offsets[n+1].  The pointers stay valid
 * until the next call on the same handle.  *n_records == 0 means end of file. */
int orph_fastx_next_batch(orph_fastx *fx, uint64_t max_records, uint64_t max_bases, uint64_t *n_records,
                          const uint8_t **seq_bytes, const uint64_t **seq_offsets, const uint8_t **name_bytes,
                          const uint64_t **name_offsets);
/* Copies the sequence bytes and offsets of the batch returned by the last orph_fastx_next_batch into caller-owned
 * arrays (seq_offsets[n] bytes and n + 1 offsets), using the parser's threads. */
int orph_fastx_copy_batch(orph_fastx *fx, uint8_t *seq_dst, uint64_t *seq_offsets_dst);
void orph_fastx_close(orph_fastx *fx);

/* ---- fused host ingest: sequence file -> 2-bit packed batches ------------------------------- */
/* The same records as orph_fastx_* (screed's semantics, translate.py:370), delivered the way the GPU takes them: the
 * reads of a batch as ONE 2-bit stream (ORPH_READS_PACKED2, base 0 of the batch at bit 0) with 32-bit base offsets,
 * the record names, and for every read where its bytes sit in the text.  Strict four-line FASTQ is parsed and packed in a
 * single parallel pass over the text (the file is mapped; gzip input is inflated block by block); any other layout
 * (FASTA, wrapped FASTQ, ...) goes through the general parser, same outputs.  Host only: no GPU is touched. */
typedef struct orph_ingest orph_ingest;
typedef struct orph_ingest_batch {
    uint64_t n_reads;              /* 0: end of file */
    const uint64_t *packed;        /* (total_bases + 31) / 32 words */
    const uint32_t *offsets32;     /* [n_reads + 1]: read i = bases [offsets32[i], offsets32[i + 1]) of the stream */
    uint64_t total_bases;
    uint32_t max_read_len;
    uint32_t reserved;
    const uint8_t *names;          /* record names, concatenated */
    const uint64_t *name_offsets;  /* [n_reads + 1] */
    const uint8_t *text;           /* read i's bytes as the file has them: text[seq_start[i] .. + seq_len[i]) */
    const uint64_t *seq_start;     /* [n_reads] */
    const uint32_t *seq_len;       /* [n_reads] */
    const uint32_t *flagged;       /* ascending indices of the reads that hold a byte other than upper-case ACGT: */
    uint64_t n_flagged;            /*   their packed bits are unspecified, score them from their bytes */
} orph_ingest_batch;
/* n_threads <= 0: every core of the host.  ORPH_E_INVALID if the file is neither FASTA nor FASTQ. */
int orph_ingest_open(const char *path, int n_threads, orph_ingest **out);
/* The next batch of about max_reads reads (batches end at a segment boundary of the text; 0 = 2^20).  The pointers stay
 * valid until the call after the next one on the same handle. */
int orph_ingest_next(orph_ingest *ing, uint64_t max_reads, orph_ingest_batch *out);
void orph_ingest_close(orph_ingest *ing);
/* The bytes of the reads index[0 .. n_index) of a batch (text / seq_start / seq_len as delivered by orph_ingest_next or
 * orph_stream_next), concatenated: out_offsets[n_index + 1] always, out_bytes (out_offsets[n_index] bytes) unless NULL. */
int orph_gather_reads(const uint8_t *text, const uint64_t *seq_start, const uint32_t *seq_len, const uint64_t *index, uint64_t n_index,
                      uint8_t *out_bytes, uint64_t *out_offsets);

/* ---- scoring stream: file in, records out ------------------------------------------------------------------------ */
/* orph_ingest feeding the scoring kernels: a background thread parses and packs batch k + 1 into pinned memory while
 * batch k is copied to the GPU and scored, so a caller looping over orph_stream_next sees one finished batch per call
 * and never touches the bases.  Replaces the loop `for record in screed.open(reads): score_reads_per_record(...)` of
 * Translate.score_reads_per_file (translate.py:367-372).  The stream owns a private context on ctx's device: the
 * caller's ctx stays free for other calls (orph_translate_frames, ...) while the stream runs. */
typedef struct orph_stream orph_stream;
typedef struct orph_stream_batch {
    uint64_t n_reads;                /* 0: end of file */
    const orph_read_result *results; /* [n_reads] */
    const uint8_t *names;
    const uint64_t *name_offsets;    /* [n_reads + 1] */
    const uint8_t *text;             /* as in orph_ingest_batch */
    const uint64_t *seq_start;
    const uint32_t *seq_len;
    uint64_t n_byte_path;            /* reads that were scored from their bytes (N, lower case, ...) */
    int status;                      /* ORPH_OK, or ORPH_E_EMPTY_READ / ORPH_E_TOO_LONG if such reads were in the batch */
} orph_stream_batch;
int orph_stream_open(orph_ctx *ctx, const orph_filter *f, const char *path, const uint8_t lut[256], double jaccard_threshold,
                     uint64_t batch_reads, int n_threads, orph_stream **out);
/* The same stream over several GPUs of the box (or several private contexts of one GPU): ONE parser feeds n_lanes device
 * lanes, batch k is copied to and scored on lane k % n_lanes, and orph_stream_next still delivers the batches in file order.
 * filters[i] must be a replica of filters[0] living on ctxs[i]'s device (orph_filter_replicate).  This is what replaces the
 * reference's serial loop over files and reads (translate.py:367-379) on a multi-GPU box; the parser, not a GPU, is the
 * limiting stage, so lanes beyond the first only add head-room. */
int orph_stream_open_multi(orph_ctx *const *ctxs, const orph_filter *const *filters, int n_lanes, const char *path,
                           const uint8_t lut[256], double jaccard_threshold, uint64_t batch_reads, int n_threads,
                           orph_stream **out);
/* Blocks until the next batch is scored.  The pointers stay valid until the next call on the same stream. */
int orph_stream_next(orph_stream *st, orph_stream_batch *out);
/* The FASTA records Translate emits for the batch last returned by orph_stream_next (translate.py:244-298), formatted by
 * the library: bit k of kinds_mask asks for kind k of orph_format_fasta (0 coding peptides = the stdout stream, 1 / 2 coding /
 * non-coding nucleotides, 3 low-complexity peptides).  The emitting frames are found in the records, their reads gathered
 * from the text, their peptides translated by one orph_translate_frames call on `ctx` (a context of the caller's, not the
 * stream's), the text formatted block-parallel.  out[k] / out_len[k]: the bytes of kind k (NULL / 0 if not requested or
 * empty), valid until the next call on the stream. */
int orph_stream_emit(orph_stream *st, orph_ctx *ctx, uint32_t kinds_mask, const uint8_t *out[4], uint64_t out_len[4]);
void orph_stream_close(orph_stream *st);

/* ---- score CSV straight from the result records -------------------------------------------------- */
/* The rows of Translate.score_reads_per_record (translate.py:342-365) in the exact text csv.writer produces for
 * them (create_save_summary.py:52-64): `read_id,jaccard,n_kmers,category,frame,filename`, 6 rows per read,
 * jaccard as Python's repr(float) of n_hits / n_kmers or `nan`, n_kmers as an int or `nan`.  name_fields holds
 * the read ids ALREADY rendered as CSV fields (the caller quotes the rare ids that need it), category_text[c]
 * the CSV field of category code c (NULL where the alphabet has none), filename_field likewise.
 * Does not write the header line.  Blocks of reads are formatted on n_threads host threads (<= 0: every core) and copied to
 * their places in a shared mapping of the file's new tail by all threads (pwrite() where the file cannot be mapped, where the
 * file system is short of space -- a full disk must be an error code, not a SIGBUS -- or with ORPHEUM_B200_CSV_PWRITE set).  The block buffers are kept between calls (one caller at a time reuses them; a concurrent
 * call allocates its own). */
int orph_csv_write_scores(const char *path, int append, const uint8_t *name_fields, const uint64_t *name_offsets,
                          uint64_t n_reads, const orph_read_result *results, const char *const *category_text,
                          const char *filename_field, int n_threads);
/* Helpers of the JSON summary (create_save_summary.py:180-222 groups rows by read id): 64-bit hashes of a batch's read
 * ids, and whether a set of such hashes is duplicate-free.  all_distinct == 1 proves the ids pairwise different;
 * 0 means two ids are equal or, very rarely, collide: the caller then compares the ids themselves. */
int orph_hash_names(const uint8_t *names, const uint64_t *name_offsets, uint64_t n, uint64_t *out_hashes, int n_threads);
int orph_u64_all_distinct(const uint64_t *values, uint64_t n, int n_threads, int *all_distinct);
/* One batch's share of the JSON summary (create_save_summary.py:136-240), straight from its records (read ids proven
 * distinct by the two calls above): per row (6 per read, frame order) jaccard[6r+f] = n_hits / n_kmers for Coding and
 * Non-coding frames, NaN otherwise, and scored[6r+f] = 1 / 0; classes = reads by their one category {Coding, too-short
 * peptide, too-short read, reads whose only category is code c (6 entries), Non-coding}; coding_hist[c] = reads with c
 * coding frames (c = 1..6), first_with[c] = index of the first such read (~0 if none: the reference's histogram keeps
 * first-appearance order). */
int orph_summary_partial(const orph_read_result *records, uint64_t n_reads, double *jaccard, uint8_t *scored, uint64_t classes[10],
                         uint64_t coding_hist[7], uint64_t first_with[7], int n_threads);
/* numpy's float64 sum of a contiguous column, to the last bit (pairwise summation with numpy's tree shape: the mean and
 * std of the jaccard column in the reference's summary are such sums, create_save_summary.py:143-151), subtrees summed on
 * n_threads threads.  Element i is x[i] (square == 0) or (x[i] - shift)^2 (square != 0), and 0 where ok[i] == 0 (ok may be
 * NULL: every element counts) -- np.nanmean / np.nanstd's arithmetic without their temporaries. */
int orph_f64_pairwise_sum(const double *x, const uint8_t *ok, uint64_t n, double shift, int square, int n_threads, double *out);
/* The FASTA records Translate emits for one batch (translate.py:244-298), in read order then frame order:
 *   kind 0  coding peptides (the stdout stream): ">{name}__translation_frame:{f}__jaccard:{j}" with every space
 *           replaced by "__", then the peptide;
 *   kind 1 / 2  coding / non-coding nucleotides (--coding-nucleotide-fasta, --noncoding-nucleotide-fasta): the same
 *           header, then the read;
 *   kind 3  low-complexity peptides: ">{name} translation_frame: {}.format(frame)" verbatim (translate.py:247).
 * peptides / pep_offsets: the un-encoded translations (orph_translate_frames) of every frame whose category is CODING
 * -- and LOW_COMPLEXITY_PEPTIDE when peptides_include_low_complexity -- in read order then frame order.  Writes at
 * most out_cap bytes and always returns the full length in *out_len (call again with a larger buffer if needed). */
int orph_format_fasta(int kind, const uint8_t *names, const uint64_t *name_offsets, uint64_t n_reads,
                      const orph_read_result *results, const uint8_t *reads, const uint64_t *read_offsets,
                      const uint8_t *peptides, const uint64_t *pep_offsets, int peptides_include_low_complexity,
                      uint8_t *out, uint64_t out_cap, uint64_t *out_len);
/* Python's repr(float) (the formatting the CSV uses), exposed for tests.  out32 must hold 32 bytes. */
int orph_format_float_repr(double v, char *out32);

/* ---- measurement helpers (bench.py) ---------------------------------------------------------- */
/* Random 32-byte-sector gather micro-benchmark over `footprint_bytes` of device memory: every thread
 * issues `loads_per_thread` dependent-free byte loads at pseudo-random addresses.  Returns sectors/s
 * measured with CUDA events on the context's stream.  This is the ceiling the Bloom-probe rate is
 * graded against. */
int orph_bench_gather_peak(orph_ctx *ctx, uint64_t footprint_bytes, uint32_t loads_per_thread, int persist_l2,
                           double *sectors_per_second);

/* Counter-based synthetic reads generated in HBM (BASELINE configs[3]: 1e9 reads are too many to stage from the host).
 * Read i = first_index + t is a pure function of (seed, i): even i is back-translated from the proteome (random
 * synonymous codons from d_syn_table, 1 % substitutions, random phase and strand), odd i is uniform ACGT;
 * orpheum_b200/synth.py:counter_reads is the host twin.  All pointers are DEVICE pointers; d_packed_out receives the
 * ORPH_READS_PACKED2 stream of n_reads reads of read_len bases ((n_reads * read_len + 31) / 32 + 1 u64 words).
 * d_syn_table[aa]: bits 0..2 = number of sense codons of amino-acid byte aa, codon c at bits [4 + 6c, +6). */
int orph_synth_reads_device(orph_ctx *ctx, uint64_t seed, uint64_t first_index, uint64_t n_reads, uint32_t read_len,
                            const uint8_t *d_residues, const uint64_t *d_prot_offsets, uint64_t n_proteins,
                            const uint64_t *d_syn_table, void *d_packed_out);

#ifdef __cplusplus
}
#endif
#endif /* ORPHEUM_B200_H */
