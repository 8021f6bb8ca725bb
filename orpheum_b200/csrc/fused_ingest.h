// Fused host ingest: FASTQ text -> 2-bit packed reads + offsets + names in ONE pass over the text, parallel over the
// host cores (stands in for the serial `screed.open` loop of orpheum/translate.py:367-372).
//
// A window of text is cut into fixed-size segments.  Every worker finds the first record of its segment, parses the
// strict four-line records that start inside it (phase A: line ends, validation, lengths), waits for the running totals
// of the segments before it (reads, bases, name bytes), and then writes its reads at their final positions (phase B:
// SIMD 2-bit packing straight into the batch's stream, 32-bit offsets, names), while the text is still in its cache.
// The guessed segment starts are verified against the end of the preceding segment, so the result is exactly what a
// sequential parse yields; anything that is not strict four-line FASTQ ends the batch in front of the offending
// record and is left to the general parser (fastx.cpp).
#pragma once
#include <cstdint>
#include <vector>

#include "hostpack.h"

namespace orph_host {

// Arrays of one batch, owned by the caller.  Capacities bound what a window may deliver; a window that holds more is
// cut at a segment boundary (the rest is parsed again into the next batch).
struct FusedBatch {
    uint64_t *packed = nullptr;       // ORPH_READS_PACKED2 stream, base 0 of the batch at bit 0 of word 0
    uint64_t packed_cap_words = 0;
    uint32_t *off32 = nullptr;        // [cap_reads + 1] base offsets into the stream
    uint64_t *seq_start = nullptr;    // [cap_reads] byte offset of the read's sequence in the text
    uint32_t *seq_len = nullptr;      // [cap_reads]
    uint64_t *name_off = nullptr;     // [cap_reads + 1] byte offsets into names
    uint8_t *names = nullptr;         // record names (header line without '@', stripped), concatenated
    uint64_t names_cap = 0;
    uint64_t cap_reads = 0;
    std::vector<uint32_t> flagged;    // reads holding a byte other than upper-case ACGT (their packed bits are unspecified), ascending
    // results of parse_fastq_window
    uint64_t n_reads = 0, total_bases = 0, names_bytes = 0;
    uint32_t max_len = 0;
};

struct FusedResult {
    uint64_t consumed_to = 0;   // text position behind the last record delivered (== lo when nothing was)
    bool not_strict = false;    // the record starting at consumed_to is not strict four-line FASTQ
    bool out_of_capacity = false;  // the batch was cut because an array was full
};

// Parses records of text[lo, hi) (lo is a record start).  at_eof: text[hi] is the end of the file, so a last line
// without a terminator is complete; otherwise a record that runs past hi is left for the next window.  text_end bounds
// what may be read (>= hi; SIMD loads read up to 31 bytes past a sequence).
FusedResult parse_fastq_window(Pool *pool, const uint8_t *text, uint64_t lo, uint64_t hi, uint64_t text_end, bool at_eof,
                               FusedBatch &batch);

}  // namespace orph_host
