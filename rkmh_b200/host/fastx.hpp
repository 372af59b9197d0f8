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

#include <cctype>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

namespace rkhost {

class ByteStream {
  public:
    // "-" reads standard input (gzip or plain, like any other path)
    explicit ByteStream(const char* path) : buf_(1 << 20) {
        fp_ = (path[0] == '-' && path[1] == 0) ? gzdopen(0, "r") : gzopen(path, "r");
        if (fp_) gzbuffer(fp_, 1 << 20);
    }
    ~ByteStream() { if (fp_) gzclose(fp_); }
    bool ok() const { return fp_ != nullptr; }
    int get() {
        if (pos_ >= len_) {
            if (eof_) return -1;
            len_ = gzread(fp_, buf_.data(), (unsigned)buf_.size());
            pos_ = 0;
            if (len_ <= 0) { eof_ = true; len_ = 0; return -1; }
        }
        return (unsigned char)buf_[pos_++];
    }
  private:
    gzFile fp_ = nullptr;
    std::vector<char> buf_;
    int pos_ = 0, len_ = 0;
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

// One record at a time; the state machine of the header comment above.  `next` returns false at the end of the file
// (or at a truncated quality string, where the reference stops too).
class FastxReader {
  public:
    explicit FastxReader(const char* path) : in_(path) {
        if (!in_.ok()) return;
        c_ = in_.get();
        while (c_ != -1 && c_ != '>' && c_ != '@') c_ = in_.get();
    }
    bool ok() const { return in_.ok(); }
    bool next(std::string& name, std::string& seq, std::string& qual) {
        if (c_ == -1) return false;                      // c_ is the header byte of the current record
        name.clear(); seq.clear(); qual.clear();
        int c = in_.get();
        while (c != -1 && !isspace(c)) { name.push_back((char)c); c = in_.get(); }
        while (c != -1 && c != '\n') c = in_.get();      // comment
        c = in_.get();
        while (c != -1 && c != '>' && c != '+' && c != '@') {
            if (isgraph(c)) seq.push_back((char)c);
            c = in_.get();
        }
        if (c == '+') {
            c = in_.get();
            while (c != -1 && c != '\n') c = in_.get();  // rest of the '+' line
            if (c == -1) { c_ = -1; return false; }
            c = in_.get();
            while (c != -1 && qual.size() < seq.size()) {
                if (c >= 33 && c <= 127) qual.push_back((char)c);
                c = in_.get();                           // note: one byte past the last quality byte is consumed
            }
            if (qual.size() != seq.size()) { c_ = -1; return false; }
            c = in_.get();                               // the byte after the quality string is dropped unseen
            while (c != -1 && c != '>' && c != '@') c = in_.get();
        }
        c_ = c;
        return true;
    }
  private:
    ByteStream in_;
    int c_ = -1;
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
