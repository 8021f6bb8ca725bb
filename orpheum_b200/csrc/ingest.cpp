// orph_ingest: FASTA / FASTQ (plain or gzip) file -> batches of 2-bit packed reads, host only (include/orpheum_b200.h).
//
// Strict four-line FASTQ -- what sequencers write -- goes through the fused parallel parser (fused_ingest.cpp): the text
// (the mapped file, or blocks inflated from a gzip stream) is parsed and packed in one pass.  Everything else (FASTA,
// wrapped or blank-line-separated FASTQ, a malformed tail) is handed to the general sequential parser of fastx.cpp from
// the record where strictness ends, and packed with the SIMD packer of hostpack.cpp; the batches look the same.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "ingest_internal.h"

extern int orph_set_error(int code, const char *fmt, ...);  // orpheum_b200.cu
int orph_fastx_open_general(const char *path, uint64_t byte_pos, uint64_t skip_records, orph_fastx **out);  // fastx.cpp

using namespace orph_host;

static uint64_t bgzf_block_size(const uint8_t *z, uint64_t zsize, uint64_t at);

namespace {

void *plain_alloc(size_t n)
{
    void *p = nullptr;
    return posix_memalign(&p, 4096, n ? n : 1) == 0 ? p : nullptr;
}
void plain_release(void *p) { free(p); }

template <class T>
bool grow(const IngestAlloc &a, T *&p, uint64_t &cap, uint64_t need)
{
    if (need <= cap && p) return true;
    if (p) a.release(p);
    p = nullptr;
    cap = 0;
    const uint64_t want = need + need / 8 + 64;
    p = static_cast<T *>(a.alloc(want * sizeof(T)));
    if (!p) return false;
    cap = want;
    return true;
}

}  // namespace

IngestSlot::~IngestSlot()
{
    if (alloc.release) {
        if (b.packed) alloc.release(b.packed);
        if (b.off32) alloc.release(b.off32);
    }
    free(b.seq_start);
    free(b.seq_len);
    free(b.name_off);
    free(b.names);
    free(own_text);
}

// arrays for up to cap_reads reads out of `text_bytes` bytes of text
static bool ensure_slot(IngestSlot &s, uint64_t cap_reads, uint64_t text_bytes)
{
    FusedBatch &b = s.b;
    // a strict record holds its bases twice (sequence and quality): at most text/2 bases
    const uint64_t words = text_bytes / 64 + cap_reads / 32 + 8;
    if (!grow(s.alloc, b.packed, s.packed_cap, words)) return false;
    b.packed_cap_words = s.packed_cap;
    if (cap_reads + 1 > s.reads_cap) {
        if (b.off32) s.alloc.release(b.off32);
        b.off32 = nullptr;
        const uint64_t want = cap_reads + cap_reads / 8 + 64;
        b.off32 = static_cast<uint32_t *>(s.alloc.alloc((want + 1) * sizeof(uint32_t)));
        b.seq_start = static_cast<uint64_t *>(realloc(b.seq_start, want * sizeof(uint64_t)));
        b.seq_len = static_cast<uint32_t *>(realloc(b.seq_len, want * sizeof(uint32_t)));
        b.name_off = static_cast<uint64_t *>(realloc(b.name_off, (want + 1) * sizeof(uint64_t)));
        if (!b.off32 || !b.seq_start || !b.seq_len || !b.name_off) return false;
        s.reads_cap = want;
    }
    b.cap_reads = cap_reads;
    if (text_bytes > s.names_cap) {
        free(b.names);
        b.names = static_cast<uint8_t *>(malloc(text_bytes + 64));  // untouched pages cost nothing
        if (!b.names) return false;
        s.names_cap = text_bytes;
    }
    b.names_cap = s.names_cap;
    return true;
}

int orph_ingest_open_with(const char *path, int n_threads, int n_slots, IngestAlloc alloc, orph_ingest **out)
{
    if (!path || !out) return orph_set_error(ORPH_E_INVALID, "orph_ingest_open: NULL argument");
    *out = nullptr;
    // format sniffing and the error behaviour are the general reader's (screed's): unknown format -> ORPH_E_INVALID
    orph_fastx *probe = nullptr;
    int rc = orph_fastx_open(path, &probe);
    if (rc != ORPH_OK) return rc;
    orph_ingest *ing = new (std::nothrow) orph_ingest();
    if (!ing) {
        orph_fastx_close(probe);
        return orph_set_error(ORPH_E_NOMEM, "out of host memory");
    }
    ing->path = path;
    ing->alloc = alloc.alloc ? alloc : IngestAlloc{plain_alloc, plain_release};
    if (n_threads <= 0) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    ing->pool = new (std::nothrow) Pool(std::min(n_threads, 64));
    ing->slots.resize((size_t)std::max(2, n_slots));
    for (auto &s : ing->slots) s.alloc = ing->alloc;
    // what kind of file is it?  (the probe has read the first block)
    gzFile gz = gzopen(path, "rb");
    const bool compressed = gz && !gzdirect(gz);
    struct stat st;
    const int fd = open(path, O_RDONLY);
    const bool regular = fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode);
    char first = 0;
    if (gz) {
        gzbuffer(gz, 1 << 20);
        const int c = gzgetc(gz);
        if (c >= 0) { first = (char)c; gzungetc(c, gz); }
    }
    ing->kind = first;
    if (first == '@' && !compressed && regular && st.st_size > 0) {
        void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) {
            madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
            ing->map = static_cast<const uint8_t *>(m);
            ing->map_size = (uint64_t)st.st_size;
        }
    }
    if (fd >= 0) close(fd);
    if (first == '@' && compressed) {
        ing->gz = gz;  // fused parser over inflated blocks
        gz = nullptr;
        // blocked gzip?  then the members are inflated in parallel straight from the mapped file
        const int zfd = regular ? open(path, O_RDONLY) : -1;
        if (zfd >= 0) {
            void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, zfd, 0);
            close(zfd);
            if (m != MAP_FAILED) {
                if (bgzf_block_size(static_cast<const uint8_t *>(m), (uint64_t)st.st_size, 0) > 0) {
                    ing->zmap = static_cast<const uint8_t *>(m);
                    ing->zsize = (uint64_t)st.st_size;
                    madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
                } else {
                    munmap(m, (size_t)st.st_size);
                }
            }
        }
    }
    if (gz) gzclose(gz);
    if (ing->map || ing->gz) {
        orph_fastx_close(probe);
    } else {
        ing->general = true;  // FASTA, empty file, unmappable input: the general parser from the start
        ing->fx = probe;
    }
    *out = ing;
    return ORPH_OK;
}

extern "C" int orph_ingest_open(const char *path, int n_threads, orph_ingest **out)
{
    return orph_ingest_open_with(path, n_threads, 2, IngestAlloc{nullptr, nullptr}, out);
}

extern "C" void orph_ingest_close(orph_ingest *ing)
{
    if (!ing) return;
    if (ing->map) munmap(const_cast<uint8_t *>(ing->map), ing->map_size);
    if (ing->zmap) munmap(const_cast<uint8_t *>(ing->zmap), ing->zsize);
    if (ing->gz) gzclose(ing->gz);
    if (ing->fx) orph_fastx_close(ing->fx);
    ing->slots.clear();
    delete ing->pool;
    delete ing;
}

// the general parser's next batch, copied into the slot and packed
static int next_general(orph_ingest *ing, IngestSlot &s, uint64_t max_reads)
{
    uint64_t n = 0;
    const uint8_t *seq, *names;
    const uint64_t *seq_off, *name_off;
    const int rc = orph_fastx_next_batch(ing->fx, max_reads, 1ull << 31, &n, &seq, &seq_off, &names, &name_off);
    if (rc != ORPH_OK) return rc;
    FusedBatch &b = s.b;
    b.n_reads = 0;
    b.total_bases = b.names_bytes = 0;
    b.max_len = 0;
    b.flagged.clear();
    if (n == 0) return ORPH_OK;
    const uint64_t total = seq_off[n];
    if (total > 0xFFFFFFFFull) return orph_set_error(ORPH_E_INVALID, "orph_ingest: a batch of more than 2^32 bases");
    if (!ensure_slot(s, n, std::max<uint64_t>(2 * total, name_off[n]) + 64)) return orph_set_error(ORPH_E_NOMEM, "out of host memory");
    if (total + 64 > s.own_text_cap) {
        free(s.own_text);
        s.own_text = static_cast<uint8_t *>(malloc(total + total / 8 + 64));
        if (!s.own_text) return orph_set_error(ORPH_E_NOMEM, "out of host memory");
        s.own_text_cap = total + total / 8 + 64;
    }
    memcpy(s.own_text, seq, total);
    s.text = s.own_text;
    memcpy(b.names, names, name_off[n]);
    memcpy(b.name_off, name_off, (n + 1) * sizeof(uint64_t));
    std::vector<uint8_t> needs(n, 0);
    pack_stream(ing->pool, s.own_text, 0, total, seq_off, n, b.packed, needs.data());
    uint32_t max_len = 0;
    for (uint64_t i = 0; i < n; i++) {
        const uint64_t len = seq_off[i + 1] - seq_off[i];
        b.seq_start[i] = seq_off[i];
        b.seq_len[i] = (uint32_t)len;
        b.off32[i] = (uint32_t)seq_off[i];
        max_len = std::max(max_len, (uint32_t)std::min<uint64_t>(len, 0xFFFFFFFFull));
        if (needs[i]) b.flagged.push_back((uint32_t)i);
    }
    b.off32[n] = (uint32_t)total;
    b.n_reads = n;
    b.total_bases = total;
    b.names_bytes = name_off[n];
    b.max_len = max_len;
    return ORPH_OK;
}

// BGZF: gzip members of at most 64 KiB, each with an extra field "BC" holding (member size - 1) [SAM spec 4.1].  Returns the
// size of the member that starts at `at`, 0 if there is no well-formed BGZF member header there.
static uint64_t bgzf_block_size(const uint8_t *z, uint64_t zsize, uint64_t at)
{
    if (at + 18 > zsize) return 0;
    const uint8_t *h = z + at;
    if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return 0;
    const uint32_t xlen = h[10] | (h[11] << 8);
    if (at + 12 + xlen > zsize) return 0;
    for (uint32_t x = 0; x + 4 <= xlen;) {
        const uint8_t *f = h + 12 + x;
        const uint32_t slen = f[2] | (f[3] << 8);
        if (f[0] == 'B' && f[1] == 'C' && slen == 2 && x + 6 <= xlen) {
            const uint64_t bsize = (uint64_t)(f[4] | (f[5] << 8)) + 1;
            if (bsize < 12 + xlen + 8 || at + bsize > zsize || (h[3] & ~4)) return 0;  // no name / comment / hcrc fields in BGZF
            return bsize;
        }
        x += 4 + slen;
    }
    return 0;
}

// BGZF: the members that follow zpos, about `want` bytes of text, inflated in parallel behind the carried-over tail.
// false: a member is damaged (error set) -- or *not_bgzf: the file stops being BGZF here (the caller falls back).
static bool fill_bgzf_text(orph_ingest *ing, IngestSlot &s, uint64_t want, bool *not_bgzf)
{
    struct Member { uint64_t in, in_len, out, out_len; uint32_t crc; };
    std::vector<Member> members;
    uint64_t at = ing->zpos, out_total = 0;
    *not_bgzf = false;
    while (at < ing->zsize && out_total < want) {
        const uint64_t bsize = bgzf_block_size(ing->zmap, ing->zsize, at);
        if (bsize == 0) {
            *not_bgzf = true;
            break;
        }
        const uint8_t *h = ing->zmap + at;
        const uint32_t xlen = h[10] | (h[11] << 8);
        const uint8_t *tail = h + bsize - 8;
        const uint32_t crc = tail[0] | (tail[1] << 8) | (tail[2] << 16) | ((uint32_t)tail[3] << 24);
        const uint32_t isize = tail[4] | (tail[5] << 8) | (tail[6] << 16) | ((uint32_t)tail[7] << 24);
        if (isize > (1u << 16)) {  // a BGZF member holds at most 64 KiB of text
            orph_set_error(ORPH_E_IO, "%s: a damaged BGZF block (size field)", ing->path.c_str());
            return false;
        }
        members.push_back({at + 12 + xlen, bsize - 12 - xlen - 8, out_total, isize, crc});
        out_total += isize;
        at += bsize;
    }
    if (*not_bgzf && members.empty()) return false;
    const uint64_t carry = ing->carry.size();
    const uint64_t need = carry + out_total + 64;
    if (need > s.own_text_cap) {
        uint8_t *q = static_cast<uint8_t *>(realloc(s.own_text, need + need / 8));
        if (!q) {
            orph_set_error(ORPH_E_NOMEM, "out of host memory");
            return false;
        }
        s.own_text = q;
        s.own_text_cap = need + need / 8;
    }
    if (carry) memcpy(s.own_text, ing->carry.data(), carry);
    ing->carry.clear();
    std::atomic<int> bad{0};
    uint8_t *dst = s.own_text + carry;
    const uint8_t *zmap = ing->zmap;
    auto inflate_some = [&](uint64_t blk) {  // 16 members per work item
        z_stream zs;
        memset(&zs, 0, sizeof(zs));
        if (inflateInit2(&zs, -15) != Z_OK) {
            bad = 1;
            return;
        }
        const uint64_t m1 = std::min<uint64_t>(members.size(), (blk + 1) * 16);
        for (uint64_t m = blk * 16; m < m1; m++) {
            const Member &mb = members[m];
            if (mb.out_len == 0) continue;
            inflateReset(&zs);
            zs.next_in = const_cast<Bytef *>(zmap + mb.in);
            zs.avail_in = (uInt)mb.in_len;
            zs.next_out = dst + mb.out;
            zs.avail_out = (uInt)mb.out_len;
            const int rc = inflate(&zs, Z_FINISH);
            if (rc != Z_STREAM_END || zs.avail_out != 0 || (uint32_t)crc32(0L, dst + mb.out, (uInt)mb.out_len) != mb.crc) bad = 1;
        }
        inflateEnd(&zs);
    };
    const uint64_t n_items = (members.size() + 15) / 16;
    if (ing->pool) ing->pool->run(n_items, inflate_some);
    else
        for (uint64_t b = 0; b < n_items; b++) inflate_some(b);
    if (bad) {
        orph_set_error(ORPH_E_IO, "%s: a damaged BGZF block (inflate / CRC)", ing->path.c_str());
        return false;
    }
    ing->zpos = at;
    if (at >= ing->zsize) ing->gz_eof = true;
    s.text = s.own_text;
    s.text_len = carry + out_total;
    return true;
}

// gzip: the slot's text = what the previous window left over + freshly inflated bytes
static bool fill_gz_text(orph_ingest *ing, IngestSlot &s, uint64_t want)
{
    const uint64_t need = ing->carry.size() + want + 64;
    if (need > s.own_text_cap) {
        uint8_t *q = static_cast<uint8_t *>(realloc(s.own_text, need));
        if (!q) return false;
        s.own_text = q;
        s.own_text_cap = need;
    }
    if (!ing->carry.empty()) memcpy(s.own_text, ing->carry.data(), ing->carry.size());
    uint64_t at = ing->carry.size();
    ing->carry.clear();
    while (!ing->gz_eof && at < need - 64) {
        const int n = gzread(ing->gz, s.own_text + at, (unsigned)std::min<uint64_t>(need - 64 - at, 1u << 30));
        if (n <= 0) {
            ing->gz_eof = true;
            break;
        }
        at += (uint64_t)n;
    }
    s.text = s.own_text;
    s.text_len = at;
    return true;
}

int orph_ingest_next_slot(orph_ingest *ing, uint64_t max_reads, IngestSlot **out)
{
    if (!ing || !out) return orph_set_error(ORPH_E_INVALID, "orph_ingest_next: NULL argument");
    if (max_reads == 0) max_reads = 1u << 20;
    max_reads = std::max<uint64_t>(4096, std::min<uint64_t>(max_reads, 1u << 26));
    ing->cur = (ing->cur + 1) % (int)ing->slots.size();
    IngestSlot &s = ing->slots[(size_t)ing->cur];
    *out = &s;
    s.b.n_reads = 0;
    s.b.flagged.clear();
    if (!ing->general) {
        uint64_t want = ing->bytes_per_record > 0 ? (uint64_t)(0.97 * ing->bytes_per_record * (double)max_reads) + (1u << 16) : 64u << 20;
        for (;;) {
            uint64_t lo, hi, text_end;
            bool at_eof;
            const uint8_t *text;
            if (ing->map) {
                text = ing->map;
                lo = ing->pos;
                hi = std::min(ing->map_size, lo + want);
                at_eof = hi == ing->map_size;
                text_end = ing->map_size;
                s.text = text;
            } else if (ing->zmap) {
                bool not_bgzf = false;
                if (!fill_bgzf_text(ing, s, want, &not_bgzf)) {
                    if (!not_bgzf) return ORPH_E_IO;
                    // blocked gzip followed by something else: only a file that is BGZF from its first to its last member is
                    // read in parallel
                    return orph_set_error(ORPH_E_IO, "%s: not a BGZF member at byte %llu (mixed gzip / BGZF files are not supported)", ing->path.c_str(),
                                          (unsigned long long)ing->zpos);
                }
                text = s.own_text;
                lo = 0;
                hi = text_end = s.text_len;
                at_eof = ing->gz_eof;
            } else {
                if (!fill_gz_text(ing, s, want)) return orph_set_error(ORPH_E_NOMEM, "out of host memory");
                text = s.own_text;
                lo = 0;
                hi = text_end = s.text_len;
                at_eof = ing->gz_eof;
            }
            if (hi <= lo) return ORPH_OK;  // end of file
            if (!ensure_slot(s, max_reads, hi - lo)) return orph_set_error(ORPH_E_NOMEM, "out of host memory");
            const FusedResult r = parse_fastq_window(ing->pool, text, lo, hi, text_end, at_eof, s.b);
            if (ing->map) ing->pos = r.consumed_to;
            else ing->carry.assign(text + r.consumed_to, text + hi);
            ing->records_done += s.b.n_reads;
            if (s.b.n_reads) ing->bytes_per_record = (double)(r.consumed_to - lo) / (double)s.b.n_reads;
            if (r.not_strict) {
                // the rest of the file belongs to the general parser
                ing->general = true;
                const int rc = orph_fastx_open_general(ing->path.c_str(), ing->map ? ing->pos : 0, ing->map ? 0 : ing->records_done, &ing->fx);
                if (rc != ORPH_OK) return rc;
                if (s.b.n_reads) return ORPH_OK;
                break;
            }
            if (s.b.n_reads) return ORPH_OK;
            if (r.out_of_capacity) {      // not even the first segment's records fit: a larger batch
                max_reads *= 2;
                continue;
            }
            if (at_eof) return ORPH_OK;  // nothing but an empty tail
            want *= 4;                    // not one whole record in the window (long reads): widen it
            if (want > (1ull << 36)) return orph_set_error(ORPH_E_IO, "bad FASTQ: a record longer than 64 GiB");
        }
    }
    return next_general(ing, s, max_reads);
}

extern "C" int orph_ingest_next(orph_ingest *ing, uint64_t max_reads, orph_ingest_batch *out)
{
    if (!out) return orph_set_error(ORPH_E_INVALID, "orph_ingest_next: NULL argument");
    memset(out, 0, sizeof(*out));
    IngestSlot *s = nullptr;
    const int rc = orph_ingest_next_slot(ing, max_reads, &s);
    if (rc != ORPH_OK) return rc;
    const FusedBatch &b = s->b;
    out->n_reads = b.n_reads;
    out->packed = b.packed;
    out->offsets32 = b.off32;
    out->total_bases = b.total_bases;
    out->max_read_len = b.max_len;
    out->names = b.names;
    out->name_offsets = b.name_off;
    out->text = s->text;
    out->seq_start = b.seq_start;
    out->seq_len = b.seq_len;
    out->flagged = b.flagged.data();
    out->n_flagged = b.flagged.size();
    return ORPH_OK;
}

// Bytes of selected reads of a batch, concatenated (the emit path needs the sequences of the coding reads only).
extern "C" int orph_gather_reads(const uint8_t *text, const uint64_t *seq_start, const uint32_t *seq_len, const uint64_t *index, uint64_t n_index,
                                 uint8_t *out_bytes, uint64_t *out_offsets)
{
    if (!out_offsets || (n_index && (!text || !seq_start || !seq_len || !index))) return orph_set_error(ORPH_E_INVALID, "orph_gather_reads: NULL argument");
    out_offsets[0] = 0;
    for (uint64_t i = 0; i < n_index; i++) out_offsets[i + 1] = out_offsets[i] + seq_len[index[i]];
    if (out_bytes)
        for (uint64_t i = 0; i < n_index; i++) memcpy(out_bytes + out_offsets[i], text + seq_start[index[i]], seq_len[index[i]]);
    return ORPH_OK;
}
