// ba_seqio.inl -- FASTA / FASTQ text -> packed batch, host side.
//
// The step before the alignment path: instead of one Go object and several allocations per
// record (io/seqio/fasta/fasta.go:60-113, io/seqio/fastq/fastq.go:62-141, seq/linear), the whole
// file is parsed once into ONE pinned letters buffer (directly uploadable by ba_align_batch) plus
// offset/length arrays and one text buffer for names and descriptions.  The record grammar is
// the reference readers': lines are trimmed of ASCII white space, empty lines are skipped, a
// header's name ends at the first blank or tab and the rest is the description, white space
// inside sequence lines is dropped, FASTQ is the single-line four-line form with the
// reference's tolerance for a '+' line directly after the header.
// Included by ba_api.cu (same translation unit; no device code here).

namespace {

inline bool sq_space(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

struct SqBuilder {
    std::vector<int64_t> seq_off, name_off, desc_off;
    std::vector<int32_t> seq_len, name_len, desc_len;
    std::vector<uint8_t> letters;
    std::string text;
    int stride = 1;
    void header(const uint8_t *line, int64_t len)
    {
        // fasta.go:115-121 / fastq.go:151-157: name = line[1:first blank], description = the rest
        int64_t mark = -1;
        for (int64_t x = 0; x < len; x++)
            if (line[x] == ' ' || line[x] == '\t') {
                mark = x;
                break;
            }
        const int64_t nb = 1, ne = mark < 0 ? len : std::max<int64_t>(mark, 1);
        name_off.push_back((int64_t)text.size());
        name_len.push_back((int32_t)(ne - nb));
        text.append((const char *)line + nb, (size_t)(ne - nb));
        desc_off.push_back((int64_t)text.size());
        if (mark >= 0 && mark + 1 < len) {
            desc_len.push_back((int32_t)(len - mark - 1));
            text.append((const char *)line + mark + 1, (size_t)(len - mark - 1));
        } else {
            desc_len.push_back(0);
        }
        seq_off.push_back((int64_t)letters.size());
        seq_len.push_back(0);
    }
};

// next line [b, e) of data without its terminator; false at the end
inline bool sq_line(const uint8_t *d, int64_t len, int64_t &pos, int64_t &b, int64_t &e)
{
    if (pos >= len) return false;
    b = pos;
    while (pos < len && d[pos] != '\n') pos++;
    e = pos;
    if (pos < len) pos++;
    while (b < e && sq_space(d[b])) b++;       // bytes.TrimSpace
    while (e > b && sq_space(d[e - 1])) e--;
    return true;
}

int sq_finish(SqBuilder &B, ba_seqfile *out)
{
    const int64_t n = (int64_t)B.seq_off.size();
    out->n = n;
    out->stride = B.stride;
    out->letters_bytes = (int64_t)B.letters.size();
    out->text_bytes = (int64_t)B.text.size();
    out->letters = (uint8_t *)ba_alloc_host((size_t)B.letters.size() + 16);
    out->text = (char *)malloc(B.text.size() + 1);
    out->seq_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    out->name_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    out->desc_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n + 1));
    out->seq_len = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    out->name_len = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    out->desc_len = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n + 1));
    if (!out->letters || !out->text || !out->seq_off || !out->name_off || !out->desc_off || !out->seq_len ||
        !out->name_len || !out->desc_len) {
        ba_seqfile_free(out);
        return BA_ERR_NOMEM;
    }
    if (!B.letters.empty()) memcpy(out->letters, B.letters.data(), B.letters.size());
    if (!B.text.empty()) memcpy(out->text, B.text.data(), B.text.size());
    for (int64_t k = 0; k < n; k++) {
        out->seq_off[k] = B.seq_off[k];
        out->seq_len[k] = B.seq_len[k];
        out->name_off[k] = B.name_off[k];
        out->name_len[k] = B.name_len[k];
        out->desc_off[k] = B.desc_off[k];
        out->desc_len[k] = B.desc_len[k];
    }
    return BA_OK;
}

int sq_fail(ba_seqfile *out, int64_t line_no, const char *msg, const uint8_t *line, int64_t len)
{
    out->err_line = line_no;
    if (line)
        snprintf(out->err, sizeof out->err, "%s \"%.*s\"", msg, (int)std::min<int64_t>(len, 80), (const char *)line);
    else
        snprintf(out->err, sizeof out->err, "%s", msg);
    return BA_ERR_FORMAT;
}

}  // namespace

extern "C" {

int ba_fasta_parse(const uint8_t *data, int64_t len, int stride, ba_seqfile *out)
{
    if (!out || len < 0 || (len > 0 && !data) || (stride != 1 && stride != 2)) return BA_ERR_ARG;
    memset(out, 0, sizeof *out);
    SqBuilder B;
    B.stride = stride;
    int64_t pos = 0, b, e, line_no = 0;
    while (sq_line(data, len, pos, b, e)) {
        line_no++;
        if (e == b) continue;                                  // fasta.go:87-89
        if (data[b] == '>') {                                  // fasta.go:91-99
            B.header(data + b, e - b);
        } else {
            if (B.seq_off.empty())                             // fasta.go:101-103
                return sq_fail(out, line_no, "fasta: badly formed line", data + b, e - b);
            for (int64_t x = b; x < e; x++) {                  // fasta.go:104-105: white space inside the line is dropped
                if (sq_space(data[x])) continue;
                B.letters.push_back(data[x]);
                if (stride == 2) B.letters.push_back(0);       // QLetter{L, Q: 0}
                B.seq_len.back()++;
            }
        }
    }
    return sq_finish(B, out);
}

int ba_fastq_parse(const uint8_t *data, int64_t len, int quality_offset, ba_seqfile *out)
{
    if (!out || len < 0 || (len > 0 && !data)) return BA_ERR_ARG;
    memset(out, 0, sizeof *out);
    SqBuilder B;
    B.stride = 2;
    enum { ID1, LETTERS, ID2, QUALITY };
    int64_t pos = 0, b, e, line_no = 0;
    for (;;) {
        // one record (fastq.go:62-141)
        int state = ID1;
        int64_t lb = 0, le = 0;        // the header line ("label")
        bool have = false, got_seq = false;
        int64_t first = 0, count = 0;  // letters of this record
        int64_t qb = 0, qe = 0;
        bool eof = false;
        for (;;) {
            if (!sq_line(data, len, pos, b, e)) {
                eof = true;
                break;
            }
            line_no++;
            const int64_t n = e - b;
            if (state == ID1 && n > 0 && data[b] == '@') {
                state = LETTERS;
                B.header(data + b, n);
                have = true;
                lb = b;
                le = e;
                first = (int64_t)B.letters.size();
            } else if (state == ID2 && n > 0 && data[b] == '+') {
                state = QUALITY;
                if (n != 1 && (le - lb != n || memcmp(data + lb + 1, data + b + 1, (size_t)(n - 1)) != 0))
                    return sq_fail(out, line_no, "fastq: quality header does not match sequence header", nullptr, 0);
            } else if (state == LETTERS && n > 0) {
                if (data[b] == '+' && (n == 1 || (le - lb == n && memcmp(data + lb + 1, data + b + 1, (size_t)(n - 1)) == 0))) {
                    state = QUALITY;   // fastq.go:111-114: a '+' line straight after the header
                    continue;
                }
                state = ID2;
                got_seq = true;
                for (int64_t x = b; x < e; x++) {
                    const uint8_t c = data[x];
                    if (sq_space(c) || c == 0x85 || c == 0xA0) continue;   // fastq.go:143-149
                    B.letters.push_back(c);
                    B.letters.push_back(0);
                    count++;
                }
            } else if (state == QUALITY) {
                if (n == 0 && count != 0) continue;            // fastq.go:127-129
                qb = b;
                qe = e;
                break;
            }
            // any other line in any other state is skipped, like the reference's switch
        }
        if (eof) {
            if (have && state == QUALITY) {
                qb = qe = 0;  // io.EOF in the quality state: an empty quality line
            } else if (!have && state == ID1) {
                break;        // clean end of file
            } else {
                return sq_fail(out, line_no, "fastq: unexpected end of file inside a record", nullptr, 0);
            }
        }
        (void)got_seq;
        // qualities: white space dropped, one per letter (fastq.go:132-138)
        int64_t nq = 0;
        for (int64_t x = qb; x < qe; x++) {
            if (sq_space(data[x])) continue;
            if (nq < count) B.letters[(size_t)(first + 2 * nq + 1)] = (uint8_t)(data[x] - quality_offset);
            nq++;
        }
        if (nq != count) return sq_fail(out, line_no, "fastq: sequence/quality length mismatch", nullptr, 0);
        B.seq_len.back() = (int32_t)count;
        if (eof) break;
    }
    return sq_finish(B, out);
}

void ba_seqfile_free(ba_seqfile *f)
{
    if (!f) return;
    if (f->letters) ba_free_host(f->letters);
    free(f->text);
    free(f->seq_off);
    free(f->name_off);
    free(f->desc_off);
    free(f->seq_len);
    free(f->name_len);
    free(f->desc_len);
    const int64_t el = f->err_line;
    char err[sizeof f->err];
    memcpy(err, f->err, sizeof err);
    memset(f, 0, sizeof *f);
    f->err_line = el;           // the diagnosis survives the release
    memcpy(f->err, err, sizeof err);
}

}  // extern "C"

// ---- one batch over several devices: the host deals pairs to devices and
// gathers results by pair index; this helper packs a device's share contiguously so that it
// uploads only its own letters.
extern "C" int ba_gather_pairs(const uint8_t *letters, int stride, const int64_t *ref_off, const int32_t *ref_len,
                               const int64_t *qry_off, const int32_t *qry_len, const int64_t *ids, int64_t k,
                               uint8_t *dst, int64_t dst_bytes, int64_t *new_ref_off, int64_t *new_qry_off)
{
    if (k < 0 || (stride != 1 && stride != 2) || (k > 0 && (!ids || !ref_off || !ref_len || !qry_off || !qry_len ||
                                                            !new_ref_off || !new_qry_off)))
        return BA_ERR_ARG;
    int64_t o = 0;
    for (int64_t x = 0; x < k; x++) {
        const int64_t p = ids[x];
        const int64_t rb = (int64_t)ref_len[p] * stride, qb = (int64_t)qry_len[p] * stride;
        if (rb < 0 || qb < 0 || o + rb + qb > dst_bytes) return BA_ERR_ARG;
        if (rb) memcpy(dst + o, letters + ref_off[p], (size_t)rb);
        new_ref_off[x] = o;
        o += rb;
        if (qb) memcpy(dst + o, letters + qry_off[p], (size_t)qb);
        new_qry_off[x] = o;
        o += qb;
    }
    return BA_OK;
}
