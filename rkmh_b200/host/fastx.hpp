// fastx.hpp -- FASTA/FASTQ record reader for the rkmh host front-end.
//
// Same record rules as the reference's parser (kseq_read as used by parse_fastas, reference
// kseq_reader/kseq.hpp:170-208 and src/rkmh.cpp:238-292), written as a plain state machine over a zlib
// stream (gzip input is transparent, as with gzopen in the reference):
//   * a record starts at '>' or '@'; the name is the header up to the first white-space byte, the rest of the
//     header line is a comment and is dropped;
//   * the sequence is every isgraph() byte up to the next '>', '@' or '+' byte (anywhere, not only at a line start);
//   * after '+' the rest of that line is skipped and quality bytes in [33,127] are collected until there are as
//     many as sequence bytes; one more byte is consumed after the last quality byte (the reference does the same);
//   * a truncated quality string ends the file (kseq returns -2 and the reference's loop stops).
// Sequences are stored exactly as read: upper-casing (mkmh::to_upper) happens on the device.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace rkhost {

class ByteStream {
  public:
    // "-" reads standard input (gzip or plain, like any other path)
    explicit ByteStream(const char* path) : buf_(1 << 22) {
        if (const char* e = getenv("RKMH_PARSE_BUF")) buf_.resize((size_t)std::max<long>(atol(e), 16));   // testing knob
        fp_ = (path[0] == '-' && path[1] == 0) ? gzdopen(0, "r") : gzopen(path, "r");
        if (fp_) gzbuffer(fp_, 1 << 20);
    }
    ~ByteStream() { if (fp_) gzclose(fp_); }
    bool ok() const { return fp_ != nullptr; }
    int get() {
        if (pos_ >= len_ && fill(1) == 0) return -1;
        return (unsigned char)buf_[pos_++];
    }
    // Makes at least `need` unread bytes available (fewer only at the end of the input); returns how many there are.
    // Pointers obtained from ptr() before the call are invalid afterwards.
    size_t fill(size_t need) {
        if (len_ - pos_ >= need || eof_) return len_ - pos_;
        if (pos_ > 0) { memmove(buf_.data(), buf_.data() + pos_, len_ - pos_); len_ -= pos_; pos_ = 0; }
        if (need > buf_.size()) buf_.resize(need + (need >> 1));
        while (len_ < need && !eof_) {
            const int got = gzread(fp_, buf_.data() + len_, (unsigned)std::min<size_t>(buf_.size() - len_, 1u << 30));
            if (got <= 0) eof_ = true; else len_ += (size_t)got;
        }
        return len_ - pos_;
    }
    const char* ptr() const { return buf_.data() + pos_; }
    void skip(size_t n) { pos_ += n; }
    bool at_eof() const { return eof_; }
  private:
    gzFile fp_ = nullptr;
    std::vector<char> buf_;
    size_t pos_ = 0, len_ = 0;
    bool eof_ = false;
};

// A growing set of records in the layout the C ABI takes: concatenated bases + offsets, plus names (and
// qualities when asked for).
struct SeqSet {
    std::vector<char> bases;
    std::vector<int64_t> offs{0};
    std::vector<std::string> names;
    std::vector<std::string> quals;
    size_t size() const { return names.size(); }
};

// A record as views into the reader's buffers: valid until the next call of FastxReader::next.
struct Record { const char* name; size_t name_len; const char* seq; size_t seq_len; const char* qual; size_t qual_len; };

// One record at a time.  The general path is the byte-wise state machine of the header comment above.  Well-formed
// four-line FASTQ records ("@name...\n" + one line of letters + "+...\n" + as many quality bytes + "\n", followed by
// the next header or the end of the input) take a fast path that finds the four line ends with memchr and hands out
// views into the input buffer; on such records both paths give the same name, sequence and quality by construction,
// and everything else -- FASTA, wrapped lines, CRLF, '>' '+' '@' inside a sequence, truncated files -- is left to
// the state machine.  `next` returns false at the end of the file (or at a truncated quality string, where the
// reference stops too).
class FastxReader {
  public:
    explicit FastxReader(const char* path) : in_(path) {
        if (!in_.ok()) return;
        c_ = in_.get();
        while (c_ != -1 && c_ != '>' && c_ != '@') c_ = in_.get();
    }
    bool ok() const { return in_.ok(); }
    bool next(Record& r) {
        if (c_ == -1) return false;                      // c_ is the header byte of the current record
        if (c_ == '@' && fast(r)) return true;
        if (!slow()) return false;
        r = Record{name_.data(), name_.size(), seq_.data(), seq_.size(), qual_.data(), qual_.size()};
        return true;
    }
    bool next(std::string& name, std::string& seq, std::string& qual) {
        Record r;
        if (!next(r)) return false;
        name.assign(r.name, r.name_len); seq.assign(r.seq, r.seq_len); qual.assign(r.qual, r.qual_len);
        return true;
    }
  private:
    static bool all_letters(const char* p, size_t n) {
        unsigned bad = 0;
        for (size_t i = 0; i < n; ++i) bad |= (unsigned)((unsigned char)((p[i] | 0x20) - 'a') > 25);
        return bad == 0;
    }
    static bool all_quality(const char* p, size_t n) {
        unsigned bad = 0;
        for (size_t i = 0; i < n; ++i) bad |= (unsigned)((unsigned char)(p[i] - 33) > 94);
        return bad == 0;
    }
    bool fast(Record& r) {                               // nothing is consumed unless the record is accepted
        size_t need = 1024;
        for (;;) {
            const size_t avail = in_.fill(need);
            const char* p = in_.ptr();
            const char* end = p + avail;
            const bool eof = in_.at_eof();
            const char* e1 = (const char*)memchr(p, '\n', avail);
            const char* e2 = e1 ? (const char*)memchr(e1 + 1, '\n', (size_t)(end - (e1 + 1))) : nullptr;
            const char* e3 = nullptr;
            if (e2) {
                if (e2 + 1 >= end) { if (eof) return false; } else if (e2[1] != '+') return false;
                else e3 = (const char*)memchr(e2 + 1, '\n', (size_t)(end - (e2 + 1)));
            }
            if (e3) {
                const char* sq = e1 + 1;
                const size_t L = (size_t)(e2 - sq);
                const char* q = e3 + 1;
                if (L == 0) return false;
                if (q + L + 2 <= end || (eof && q + L + 1 == end)) {
                    if (q[L] != '\n') return false;
                    const bool last = q + L + 1 == end;
                    if (!last && q[L + 1] != '@' && q[L + 1] != '>') return false;
                    if (!all_letters(sq, L) || !all_quality(q, L)) return false;
                    size_t nl = 0;
                    while (p + nl < e1 && !isspace((unsigned char)p[nl])) ++nl;
                    r = Record{p, nl, sq, L, q, L};
                    if (last) { in_.skip((size_t)(q + L + 1 - p)); c_ = -1; }
                    else { c_ = (unsigned char)q[L + 1]; in_.skip((size_t)(q + L + 2 - p)); }
                    return true;
                }
            }
            if (eof) return false;                       // the tail of the input: the state machine decides
            need = std::max<size_t>(2 * need, 2 * avail);
        }
    }
    bool slow() {
        name_.clear(); seq_.clear(); qual_.clear();
        int c = in_.get();
        while (c != -1 && !isspace(c)) { name_.push_back((char)c); c = in_.get(); }
        while (c != -1 && c != '\n') c = in_.get();      // comment
        c = in_.get();
        while (c != -1 && c != '>' && c != '+' && c != '@') {
            if (isgraph(c)) seq_.push_back((char)c);
            c = in_.get();
        }
        if (c == '+') {
            c = in_.get();
            while (c != -1 && c != '\n') c = in_.get();  // rest of the '+' line
            if (c == -1) { c_ = -1; return false; }
            c = in_.get();
            while (c != -1 && qual_.size() < seq_.size()) {
                if (c >= 33 && c <= 127) qual_.push_back((char)c);
                c = in_.get();                           // note: one byte past the last quality byte is consumed
            }
            if (qual_.size() != seq_.size()) { c_ = -1; return false; }
            c = in_.get();                               // the byte after the quality string is dropped unseen
            while (c != -1 && c != '>' && c != '@') c = in_.get();
        }
        c_ = c;
        return true;
    }
    ByteStream in_;
    int c_ = -1;
    std::string name_, seq_, qual_;
};

// Appends every record of `path` to `out`.  Returns false when the file cannot be opened.
inline bool read_fastx(const char* path, SeqSet& out, bool keep_qual) {
    FastxReader rd(path);
    if (!rd.ok()) return false;
    std::string name, seq, qual;
    while (rd.next(name, seq, qual)) {
        out.bases.insert(out.bases.end(), seq.begin(), seq.end());
        out.offs.push_back((int64_t)out.bases.size());
        out.names.push_back(name);
        if (keep_qual) out.quals.push_back(qual);
    }
    return true;
}

}  // namespace rkhost
