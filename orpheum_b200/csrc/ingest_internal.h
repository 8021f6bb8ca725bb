// Internals of orph_ingest shared between ingest.cpp and the scoring stream in orpheum_b200.cu.
#pragma once
#include <zlib.h>

#include <string>
#include <vector>

#include "../../include/orpheum_b200.h"
#include "fused_ingest.h"

// where the arrays the DMA engine reads come from: plain memory for host-only use, pinned memory under a stream
struct IngestAlloc {
    void *(*alloc)(size_t bytes);
    void (*release)(void *p);
};

struct IngestSlot {
    orph_host::FusedBatch b;
    IngestAlloc alloc{nullptr, nullptr};
    uint64_t packed_cap = 0, reads_cap = 0, names_cap = 0;
    const uint8_t *text = nullptr;  // the bytes seq_start points into: the mapped file or own_text
    uint64_t text_len = 0;
    uint8_t *own_text = nullptr;    // inflated block (gzip) or the general parser's concatenated sequences
    uint64_t own_text_cap = 0;
    IngestSlot() = default;
    IngestSlot(const IngestSlot &) = delete;
    IngestSlot &operator=(const IngestSlot &) = delete;
    IngestSlot(IngestSlot &&o) noexcept { *this = std::move(o); }
    IngestSlot &operator=(IngestSlot &&o) noexcept
    {
        std::swap(b, o.b);
        std::swap(alloc, o.alloc);
        std::swap(packed_cap, o.packed_cap);
        std::swap(reads_cap, o.reads_cap);
        std::swap(names_cap, o.names_cap);
        std::swap(text, o.text);
        std::swap(text_len, o.text_len);
        std::swap(own_text, o.own_text);
        std::swap(own_text_cap, o.own_text_cap);
        return *this;
    }
    ~IngestSlot();
};

struct orph_ingest {
    std::string path;
    char kind = 0;                 // '>' FASTA, '@' FASTQ, 0 empty
    const uint8_t *map = nullptr;  // plain file, mapped
    uint64_t map_size = 0, pos = 0;
    gzFile gz = nullptr;           // compressed FASTQ: blocks inflated into the slots
    bool gz_eof = false;
    const uint8_t *zmap = nullptr; // BGZF (blocked gzip: bgzip, samtools, many sequencing pipelines): the compressed file, mapped;
    uint64_t zsize = 0, zpos = 0;  //   its independent <= 64 KiB members are inflated in parallel on the pool
    std::vector<uint8_t> carry;    // text of a record cut by the end of the previous block
    bool general = false;          // the general sequential parser has taken over
    orph_fastx *fx = nullptr;
    uint64_t records_done = 0;
    double bytes_per_record = 0;
    orph_host::Pool *pool = nullptr;
    std::vector<IngestSlot> slots;
    int cur = -1;
    IngestAlloc alloc{nullptr, nullptr};
};

int orph_ingest_open_with(const char *path, int n_threads, int n_slots, IngestAlloc alloc, orph_ingest **out);
// the next batch, in the slot ring's next slot (valid until the ring wraps around)
int orph_ingest_next_slot(orph_ingest *ing, uint64_t max_reads, IngestSlot **out);

// the inputs of orph_format_fasta; the reads either contiguous (reads + read_offsets) or scattered in a text (stream batches)
struct OrphFastaSource {
    const uint8_t *names;
    const uint64_t *name_offsets;
    const orph_read_result *results;
    const uint8_t *reads;
    const uint64_t *read_offsets;
    const uint8_t *text;
    const uint64_t *seq_start;
    const uint32_t *seq_len;
    const uint8_t *peptides;
    const uint64_t *pep_offsets;
    int peptides_include_low_complexity;
};
int64_t orph_format_fasta_range(int kind, const OrphFastaSource &src, uint64_t r0, uint64_t r1, uint64_t item0, uint8_t *out, uint64_t out_cap);
