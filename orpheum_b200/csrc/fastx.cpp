// Host ingest: FASTA / FASTQ (plain or gzip) -> the batch layout the scoring ABI takes (concatenated
// sequence bytes + u64 offsets) plus the record names.  Record semantics are screed's, which is what the
// reference reads its inputs with (orpheum/translate.py:370, orpheum/index.py:83): the name is the whole
// header line without the leading '>' / '@' (no split at white space), sequence lines are concatenated
// as they are (no upper-casing), FASTQ may wrap its sequence over several lines, an empty file has no
// records, and a file that starts with anything else is not a sequence file.
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/orpheum_b200.h"

extern int orph_set_error(int code, const char *fmt, ...);  // defined in orpheum_b200.cu

struct orph_fastx {
    gzFile gz = nullptr;
    char kind = 0;  // '>' FASTA, '@' FASTQ, 0 empty file
    std::vector<char> buf;
    size_t pos = 0, end = 0;
    bool eof = false;
    std::string pending;  // FASTA: header line of the next record, already consumed
    bool have_pending = false;
    // current batch
    std::vector<uint8_t> seq, names;
    std::vector<uint64_t> seq_off, name_off;
    std::string line;  // scratch
};

static bool fill(orph_fastx *fx)
{
    if (fx->eof) return false;
    if (fx->pos > 0 && fx->pos < fx->end) memmove(fx->buf.data(), fx->buf.data() + fx->pos, fx->end - fx->pos);
    fx->end -= fx->pos;
    fx->pos = 0;
    if (fx->end == fx->buf.size()) fx->buf.resize(fx->buf.size() * 2);
    const int n = gzread(fx->gz, fx->buf.data() + fx->end, (unsigned)(fx->buf.size() - fx->end));
    if (n <= 0) {
        fx->eof = true;
        return false;
    }
    fx->end += (size_t)n;
    return true;
}

// next line without its terminator; false at end of file
static bool next_line(orph_fastx *fx, const char **s, size_t *len)
{
    for (;;) {
        const char *base = fx->buf.data() + fx->pos;
        const char *nl = (const char *)memchr(base, '\n', fx->end - fx->pos);
        if (nl) {
            *s = base;
            *len = (size_t)(nl - base);
            fx->pos += *len + 1;
            return true;
        }
        if (!fill(fx)) {
            if (fx->pos < fx->end) {  // last line without a newline
                *s = fx->buf.data() + fx->pos;
                *len = fx->end - fx->pos;
                fx->pos = fx->end;
                return true;
            }
            return false;
        }
    }
}

static void strip(const char **s, size_t *len)
{
    while (*len && ((*s)[*len - 1] == '\r' || (*s)[*len - 1] == ' ' || (*s)[*len - 1] == '\t' || (*s)[*len - 1] == '\n')) (*len)--;
    while (*len && (**s == ' ' || **s == '\t' || **s == '\r')) {
        (*s)++;
        (*len)--;
    }
}

extern "C" int orph_fastx_open(const char *path, orph_fastx **out)
{
    if (!path || !out) return orph_set_error(ORPH_E_INVALID, "orph_fastx_open: NULL argument");
    *out = nullptr;
    gzFile gz = gzopen(path, "rb");
    if (!gz) return orph_set_error(ORPH_E_IO, "cannot open '%s'", path);
    gzbuffer(gz, 1 << 20);
    orph_fastx *fx = new orph_fastx();
    fx->gz = gz;
    fx->buf.resize(4 << 20);
    fill(fx);
    if (fx->end == 0) {
        fx->kind = 0;  // empty file: no records
    } else if (fx->buf[0] == '>' || fx->buf[0] == '@') {
        fx->kind = fx->buf[0];
    } else {
        gzclose(gz);
        delete fx;
        return orph_set_error(ORPH_E_INVALID, "unknown file format for '%s'", path);
    }
    *out = fx;
    return ORPH_OK;
}

extern "C" void orph_fastx_close(orph_fastx *fx)
{
    if (!fx) return;
    if (fx->gz) gzclose(fx->gz);
    delete fx;
}

static void begin_record(orph_fastx *fx, const char *name, size_t len)
{
    strip(&name, &len);
    fx->names.insert(fx->names.end(), (const uint8_t *)name, (const uint8_t *)name + len);
    fx->name_off.push_back(fx->names.size());
}

extern "C" int orph_fastx_next_batch(orph_fastx *fx, uint64_t max_records, uint64_t max_bases, uint64_t *n_records,
                                     const uint8_t **seq_bytes, const uint64_t **seq_offsets, const uint8_t **name_bytes,
                                     const uint64_t **name_offsets)
{
    if (!fx || !n_records) return orph_set_error(ORPH_E_INVALID, "orph_fastx_next_batch: NULL argument");
    fx->seq.clear();
    fx->names.clear();
    fx->seq_off.assign(1, 0);
    fx->name_off.assign(1, 0);
    uint64_t n = 0;
    const char *s;
    size_t len;
    if (fx->kind == '>') {
        while (n < max_records && fx->seq.size() < max_bases) {
            if (!fx->have_pending) {
                if (!next_line(fx, &s, &len)) break;
                if (len == 0 || s[0] != '>') continue;  // stray line before the first header
                fx->pending.assign(s + 1, len - 1);
                fx->have_pending = true;
            }
            begin_record(fx, fx->pending.data(), fx->pending.size());
            fx->have_pending = false;
            while (next_line(fx, &s, &len)) {
                if (len && s[0] == '>') {
                    fx->pending.assign(s + 1, len - 1);
                    fx->have_pending = true;
                    break;
                }
                strip(&s, &len);
                fx->seq.insert(fx->seq.end(), (const uint8_t *)s, (const uint8_t *)s + len);
            }
            fx->seq_off.push_back(fx->seq.size());
            n++;
        }
    } else if (fx->kind == '@') {
        while (n < max_records && fx->seq.size() < max_bases) {
            if (!next_line(fx, &s, &len)) break;
            if (len == 0) continue;
            if (s[0] != '@') return orph_set_error(ORPH_E_IO, "bad FASTQ: no '@' at the beginning of a record");
            begin_record(fx, s + 1, len - 1);
            const size_t seq_start = fx->seq.size();
            bool plus = false;
            while (next_line(fx, &s, &len)) {
                if (len && (s[0] == '+' || s[0] == '#')) {
                    plus = true;
                    break;
                }
                strip(&s, &len);
                fx->seq.insert(fx->seq.end(), (const uint8_t *)s, (const uint8_t *)s + len);
            }
            const size_t seq_len = fx->seq.size() - seq_start;
            size_t qual_len = 0;
            while (plus && qual_len < seq_len && next_line(fx, &s, &len)) {
                strip(&s, &len);
                qual_len += len;
            }
            if (qual_len != seq_len) return orph_set_error(ORPH_E_IO, "bad FASTQ: quality and sequence lengths differ");
            fx->seq_off.push_back(fx->seq.size());
            n++;
        }
    }
    *n_records = n;
    if (seq_bytes) *seq_bytes = fx->seq.data();
    if (seq_offsets) *seq_offsets = fx->seq_off.data();
    if (name_bytes) *name_bytes = fx->names.data();
    if (name_offsets) *name_offsets = fx->name_off.data();
    return ORPH_OK;
}

// ------------------------------------------------------------------------------------------------
// Score CSV writer: the rows Translate.score_reads_per_record builds (translate.py:342-365) and
// CreateSaveSummary.maybe_write_csv writes with csv.writer(lineterminator="\n")
// (create_save_summary.py:52-64), produced straight from the result records.
// ------------------------------------------------------------------------------------------------
#include <charconv>

// Python's repr(float) for a finite double: shortest round-trip digits; fixed notation when the decimal
// exponent e satisfies -4 <= e < 16, otherwise d.ddde[+-]XX; integers get a trailing ".0".
static size_t py_repr_double(double v, char *out)
{
    char sci[40];
    auto res = std::to_chars(sci, sci + sizeof(sci), v, std::chars_format::scientific);  // d[.ddd]e[+-]XX, shortest
    *res.ptr = 0;
    char *e = strchr(sci, 'e');
    const int exp10 = atoi(e + 1);
    char digits[32];
    size_t nd = 0;
    for (char *p = sci; p < e; p++)
        if (*p >= '0' && *p <= '9') digits[nd++] = *p;
    char *o = out;
    if (v < 0) *o++ = '-';
    if (exp10 >= -4 && exp10 < 16) {
        if (exp10 < 0) {
            *o++ = '0';
            *o++ = '.';
            for (int i = 0; i < -exp10 - 1; i++) *o++ = '0';
            for (size_t i = 0; i < nd; i++) *o++ = digits[i];
        } else {
            for (size_t i = 0; i < nd || (int)i <= exp10; i++) {
                if ((int)i == exp10 + 1) *o++ = '.';
                *o++ = i < nd ? digits[i] : '0';
            }
            if ((int)nd <= exp10 + 1) {
                *o++ = '.';
                *o++ = '0';
            }
        }
    } else {
        *o++ = digits[0];
        if (nd > 1) {
            *o++ = '.';
            for (size_t i = 1; i < nd; i++) *o++ = digits[i];
        }
        *o++ = 'e';
        *o++ = exp10 < 0 ? '-' : '+';
        const int a = exp10 < 0 ? -exp10 : exp10;
        if (a < 10) *o++ = '0';
        o += snprintf(o, 8, "%d", a);
    }
    *o = 0;
    return (size_t)(o - out);
}

extern "C" int orph_format_float_repr(double v, char *out32)
{
    if (!out32) return orph_set_error(ORPH_E_INVALID, "orph_format_float_repr: NULL buffer");
    if (v != v) {
        strcpy(out32, "nan");
        return ORPH_OK;
    }
    py_repr_double(v, out32);
    return ORPH_OK;
}

extern "C" int orph_csv_write_scores(const char *path, int append, const uint8_t *name_fields, const uint64_t *name_offsets,
                                     uint64_t n_reads, const orph_read_result *results, const char *const *category_text,
                                     const char *filename_field)
{
    if (!path || !category_text || !filename_field || (n_reads && (!name_fields || !name_offsets || !results)))
        return orph_set_error(ORPH_E_INVALID, "orph_csv_write_scores: NULL argument");
    FILE *f = fopen(path, append ? "ab" : "wb");
    if (!f) return orph_set_error(ORPH_E_IO, "cannot open '%s' for writing", path);
    static const char *kFrames[6] = {"1", "2", "3", "-1", "-2", "-3"};
    std::vector<char> buf;
    buf.reserve(1 << 22);
    const size_t fn_len = strlen(filename_field);
    size_t cat_len[6];
    for (int c = 0; c < 6; c++) cat_len[c] = category_text[c] ? strlen(category_text[c]) : 0;
    char num[48];
    bool ok = true;
    for (uint64_t r = 0; r < n_reads && ok; r++) {
        const uint8_t *name = name_fields + name_offsets[r];
        const size_t name_len = (size_t)(name_offsets[r + 1] - name_offsets[r]);
        const orph_read_result &res = results[r];
        for (int fr = 0; fr < 6; fr++) {
            const unsigned cat = res.category[fr];
            if (cat > 5 || !category_text[cat]) {
                fclose(f);
                return orph_set_error(ORPH_E_INVALID, "orph_csv_write_scores: read %llu frame %d has category %u without a text",
                                      (unsigned long long)r, fr, cat);
            }
            buf.insert(buf.end(), name, name + name_len);
            buf.push_back(',');
            const bool scored = cat == ORPH_CAT_CODING || cat == ORPH_CAT_NON_CODING;
            if (scored) {
                const size_t n = py_repr_double((double)res.n_hits[fr] / (double)res.n_kmers[fr], num);
                buf.insert(buf.end(), num, num + n);
            } else {
                buf.insert(buf.end(), {'n', 'a', 'n'});
            }
            buf.push_back(',');
            if (scored || cat == ORPH_CAT_LOW_COMPLEXITY_PEPTIDE) {
                const int n = snprintf(num, sizeof(num), "%u", (unsigned)res.n_kmers[fr]);
                buf.insert(buf.end(), num, num + n);
            } else {
                buf.insert(buf.end(), {'n', 'a', 'n'});
            }
            buf.push_back(',');
            buf.insert(buf.end(), category_text[cat], category_text[cat] + cat_len[cat]);
            buf.push_back(',');
            buf.insert(buf.end(), kFrames[fr], kFrames[fr] + strlen(kFrames[fr]));
            buf.push_back(',');
            buf.insert(buf.end(), filename_field, filename_field + fn_len);
            buf.push_back('\n');
        }
        if (buf.size() >= (1 << 22)) {
            ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
            buf.clear();
        }
    }
    if (ok && !buf.empty()) ok = fwrite(buf.data(), 1, buf.size(), f) == buf.size();
    ok = (fclose(f) == 0) && ok;
    return ok ? ORPH_OK : orph_set_error(ORPH_E_IO, "write to '%s' failed", path);
}
