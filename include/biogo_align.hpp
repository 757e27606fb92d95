// biogo_align.hpp -- header-only C++ host mirror of bíogo's `align` surface over the C ABI of
// libbiogoalign.so (include/biogo_align.h).  Same names, argument meaning and error behaviour as
// the Go package (align/align.go:19-73 and the six wrappers align/sw.go:22-49 ...), written in
// C++ because this image has no Go toolchain; the cgo equivalent is go/align/cuda_bridge.go.
//
//   using namespace biogo;
//   align::SWAffine smith{{{0,-1,-1,-1,-1},{-1,1,-1,-1,-1},...}, -5};
//   seq::Seq a{&alphabet::DNAgapped(), "ATAGGAAG"}, b{&alphabet::DNAgapped(), "ATTGGCAATG"};
//   std::vector<feat::Pair> aln;
//   align::Error err = smith.Align(a, b, aln);          // err == align::nil on success
//
// The DP always runs in the CUDA library; a missing GPU is reported as align::ErrCUDA (no CPU
// fallback).  Go panics surface as align::GoPanic exceptions.
#pragma once

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "biogo_align.h"

namespace biogo {

namespace alphabet {
// the subset of alphabet.Alphabet that align uses (alphabet/alphabet.go:183-255)
class Alphabet {
  public:
    Alphabet(const std::string &letters, char gap, bool case_sensitive = false) : len_((int)letters.size()), gap_(gap)
    {
        index_.fill(-1);
        for (int i = 0; i < len_; i++) {
            const unsigned char c = (unsigned char)letters[i];
            if (case_sensitive) {
                index_[c] = i;
            } else {
                index_[(unsigned char)std::tolower(c)] = i;
                index_[(unsigned char)std::toupper(c)] = i;
            }
        }
    }
    int Len() const { return len_; }
    char Gap() const { return gap_; }
    int IndexOf(char l) const { return index_[(unsigned char)l]; }
    const std::array<int32_t, 256> &LetterIndex() const { return index_; }

  private:
    int len_;
    char gap_;
    std::array<int32_t, 256> index_;
};
// alphabet/alphabet.go:26-79
inline const Alphabet &DNA() { static const Alphabet a("acgt", '-'); return a; }
inline const Alphabet &DNAgapped() { static const Alphabet a("-acgt", '-'); return a; }
inline const Alphabet &DNAredundant() { static const Alphabet a("-acmgrsvtwyhkdbn", '-'); return a; }
inline const Alphabet &RNAgapped() { static const Alphabet a("-acgu", '-'); return a; }
inline const Alphabet &Protein() { static const Alphabet a("-abcdefghijklmnpqrstvwxyz*", '-'); return a; }

struct QLetter {
    uint8_t L, Q;  // alphabet/letters.go:182-185
};
}  // namespace alphabet

namespace seq {
// align.AlphabetSlicer (align/align.go:19-22): an alphabet plus Letters or QLetters
enum class SliceKind { Letters, QLetters, Other };
struct Slicer {
    const alphabet::Alphabet *Alpha = nullptr;
    SliceKind kind = SliceKind::Letters;
    std::string letters;                     // alphabet.Letters
    std::vector<alphabet::QLetter> qletters; // alphabet.QLetters
    const alphabet::Alphabet *Alphabet() const { return Alpha; }
};
inline Slicer Seq(const alphabet::Alphabet *a, const std::string &s)
{
    Slicer x;
    x.Alpha = a;
    x.letters = s;
    return x;
}
inline Slicer QSeq(const alphabet::Alphabet *a, const std::string &s, uint8_t q = 40)
{
    Slicer x;
    x.Alpha = a;
    x.kind = SliceKind::QLetters;
    for (char c : s) x.qletters.push_back({(uint8_t)c, q});
    return x;
}
}  // namespace seq

namespace feat {
// featPair (align/align.go:121-144)
struct Pair {
    int a_start, a_end, b_start, b_end, score;
    int Score() const { return score; }
    std::string String() const
    {
        auto r = [](int s, int e) { return "[" + std::to_string(s) + "," + std::to_string(e) + ")"; };
        if (a_start == a_end) return "-/" + r(b_start, b_end) + "=" + std::to_string(score);
        if (b_start == b_end) return r(a_start, a_end) + "/-=" + std::to_string(score);
        return r(a_start, a_end) + "/" + r(b_start, b_end) + "=" + std::to_string(score);
    }
};
inline std::string Sprint(const std::vector<Pair> &aln)
{
    std::string s = "[";
    for (size_t k = 0; k < aln.size(); k++) s += (k ? " " : "") + aln[k].String();
    return s + "]";
}
}  // namespace feat

namespace align {

// align/align.go:57-73.  Sentinels compare by code; ErrMatrixWrongSize carries Size/Len.
enum class Code {
    Nil, MismatchedTypes, MismatchedAlphabets, NoAlphabet, NotGappedAlphabet, TypeNotHandled, MatrixNotSquare,
    MatrixWrongSize, IllegalLetter, CUDA
};
struct Error {
    Code code = Code::Nil;
    int Size = 0, Len = 0;   // ErrMatrixWrongSize
    std::string message;
    bool operator==(const Error &o) const { return code == o.code && Size == o.Size && Len == o.Len; }
    bool operator!=(const Error &o) const { return !(*this == o); }
    explicit operator bool() const { return code != Code::Nil; }
    const std::string &What() const { return message; }
};
inline const Error nil{};
inline const Error ErrMismatchedTypes{Code::MismatchedTypes, 0, 0, "align: mismatched sequence types"};
inline const Error ErrMismatchedAlphabets{Code::MismatchedAlphabets, 0, 0, "align: mismatched alphabets"};
inline const Error ErrNoAlphabet{Code::NoAlphabet, 0, 0, "align: no alphabet"};
inline const Error ErrNotGappedAlphabet{Code::NotGappedAlphabet, 0, 0, "align: alphabet does not have gap at position 0"};
inline const Error ErrTypeNotHandled{Code::TypeNotHandled, 0, 0, "align: sequence type not handled"};
inline const Error ErrMatrixNotSquare{Code::MatrixNotSquare, 0, 0, "align: scoring matrix is not square"};
inline Error ErrMatrixWrongSize(int size, int len)
{
    return Error{Code::MatrixWrongSize, size, len,
                 "align: scoring matrix size " + std::to_string(size) + " does not match alphabet length " +
                     std::to_string(len)};
}
struct GoPanic : std::runtime_error {
    using std::runtime_error::runtime_error;
};

using Linear = std::vector<std::vector<int>>;  // align.Linear (align/align.go:32-34)

namespace detail {
struct Ctx {
    ba_ctx *h = nullptr;
    Ctx(int device = 0)
    {
        if (ba_create(device, &h) != BA_OK) {
            std::string msg = ba_last_error(h);
            if (h) ba_destroy(h);
            h = nullptr;
            throw std::runtime_error("libbiogoalign: " + msg);
        }
    }
    ~Ctx()
    {
        if (h) ba_destroy(h);
    }
};
inline Ctx &ctx()
{
    static thread_local Ctx c;  // one context per thread: Align stays reentrant
    return c;
}

// the wrapper's validation, in the reference's order (align/sw.go:23-48)
inline Error validate(const seq::Slicer &r, const seq::Slicer &q)
{
    const alphabet::Alphabet *alpha = r.Alphabet();
    if (!alpha) return ErrNoAlphabet;
    if (alpha != q.Alphabet()) return ErrMismatchedAlphabets;
    if (alpha->IndexOf(alpha->Gap()) != 0) return ErrNotGappedAlphabet;
    if (r.kind == seq::SliceKind::Other) return ErrTypeNotHandled;
    if (r.kind != q.kind) return ErrMismatchedTypes;
    return nil;
}

inline Error run(int kind, bool affine, const Linear &m, int gap_open, bool check_size, const seq::Slicer &r,
                 const seq::Slicer &q, std::vector<feat::Pair> &out)
{
    out.clear();
    if (Error e = validate(r, q)) return e;
    const alphabet::Alphabet &alpha = *r.Alphabet();
    // align/sw_letters.go:69-79
    const int let = (int)m.size();
    if (check_size && let < alpha.Len()) return ErrMatrixWrongSize(let, alpha.Len());
    std::vector<int64_t> la;
    for (const auto &row : m) {
        if ((int)row.size() != let) return ErrMatrixNotSquare;
        for (int v : row) la.push_back(v);
    }
    std::vector<uint8_t> buf;
    int stride = 1;
    int64_t roff = 0, qoff;
    int32_t rlen, qlen;
    if (r.kind == seq::SliceKind::Letters) {
        buf.assign(r.letters.begin(), r.letters.end());
        qoff = (int64_t)buf.size();
        buf.insert(buf.end(), q.letters.begin(), q.letters.end());
        rlen = (int32_t)r.letters.size();
        qlen = (int32_t)q.letters.size();
    } else {
        stride = 2;
        for (auto l : r.qletters) { buf.push_back(l.L); buf.push_back(l.Q); }
        qoff = (int64_t)buf.size();
        for (auto l : q.qletters) { buf.push_back(l.L); buf.push_back(l.Q); }
        rlen = (int32_t)r.qletters.size();
        qlen = (int32_t)q.qletters.size();
    }
    Ctx &c = ctx();
    int rc = ba_set_scoring(c.h, kind, affine ? 1 : 0, la.data(), let, alpha.Len(), gap_open, alpha.LetterIndex().data());
    if (rc != BA_OK) return Error{Code::CUDA, 0, 0, ba_last_error(c.h)};
    ba_result res;
    rc = ba_align_batch(c.h, buf.data(), stride, &roff, &rlen, &qoff, &qlen, 1, &res);
    if (rc != BA_OK) return Error{Code::CUDA, 0, 0, ba_last_error(c.h)};
    Error err = nil;
    const int st = res.status[0];
    if (st == BA_PAIR_OK) {
        for (int k = 0; k < res.seg_count[0]; k++) {
            const int32_t *s = res.segs + 5 * (res.seg_start[0] + k);
            out.push_back({s[0], s[1], s[2], s[3], s[4]});
        }
    } else if (st == BA_PAIR_ILLEGAL_REF || st == BA_PAIR_ILLEGAL_QRY) {
        const bool ref = st == BA_PAIR_ILLEGAL_REF;
        const int64_t pos = res.err_pos[0];
        const uint8_t ch = buf[(size_t)((ref ? roff : qoff) + pos * stride)];
        err = Error{Code::IllegalLetter, 0, 0,
                    std::string("align: illegal letter '") + (char)ch + "' at position " + std::to_string(pos) +
                        " in " + (ref ? "rSeq" : "qSeq")};
    }
    ba_free_result(c.h, &res);
    if (st == BA_PAIR_PANIC) throw GoPanic("runtime error: index out of range");
    if (st == BA_PAIR_NO_PATH) throw GoPanic("align: internal error: no path");
    return err;
}
}  // namespace detail

// The six aligner types (align/sw.go:18, nw.go:16, fitted.go:16, sw_affine.go:16, nw_affine.go:16,
// fitted_affine.go:16).  Align(reference, query) -> ([]feat.Pair, error) becomes
// Error Align(reference, query, out).
struct SW {
    Linear Matrix;
    Error Align(const seq::Slicer &r, const seq::Slicer &q, std::vector<feat::Pair> &out) const
    { return detail::run(BA_SW, false, Matrix, 0, true, r, q, out); }
};
struct NW {
    Linear Matrix;
    Error Align(const seq::Slicer &r, const seq::Slicer &q, std::vector<feat::Pair> &out) const
    { return detail::run(BA_NW, false, Matrix, 0, true, r, q, out); }
};
struct Fitted {
    Linear Matrix;
    Error Align(const seq::Slicer &r, const seq::Slicer &q, std::vector<feat::Pair> &out) const
    { return detail::run(BA_FITTED, false, Matrix, 0, false, r, q, out); }
};
struct SWAffine {
    Linear Matrix;
    int GapOpen = 0;
    Error Align(const seq::Slicer &r, const seq::Slicer &q, std::vector<feat::Pair> &out) const
    { return detail::run(BA_SW, true, Matrix, GapOpen, true, r, q, out); }
};
struct NWAffine {
    Linear Matrix;
    int GapOpen = 0;
    Error Align(const seq::Slicer &r, const seq::Slicer &q, std::vector<feat::Pair> &out) const
    { return detail::run(BA_NW, true, Matrix, GapOpen, true, r, q, out); }
};
struct FittedAffine {
    Linear Matrix;
    int GapOpen = 0;
    Error Align(const seq::Slicer &r, const seq::Slicer &q, std::vector<feat::Pair> &out) const
    { return detail::run(BA_FITTED, true, Matrix, GapOpen, false, r, q, out); }
};

// align.Format (align/align.go:148-170) for Letters / QLetters (the L component)
inline std::array<std::string, 2> Format(const seq::Slicer &a, const seq::Slicer &b, const std::vector<feat::Pair> &f,
                                         char gap)
{
    auto sub = [](const seq::Slicer &s, int from, int to) {
        if (s.kind == seq::SliceKind::Letters) return s.letters.substr((size_t)from, (size_t)(to - from));
        std::string o;
        for (int k = from; k < to; k++) o.push_back((char)s.qletters[(size_t)k].L);
        return o;
    };
    std::array<std::string, 2> out;
    for (const auto &p : f) {
        const int alen = p.a_end - p.a_start, blen = p.b_end - p.b_start;
        out[0] += alen == 0 ? std::string((size_t)blen, gap) : sub(a, p.a_start, p.a_end);
        out[1] += blen == 0 ? std::string((size_t)alen, gap) : sub(b, p.b_start, p.b_end);
    }
    return out;
}

}  // namespace align
}  // namespace biogo
