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
    explicit ByteStream(const char* path) : buf_(1 << 20) { fp_ = gzopen(path, "r"); if (fp_) gzbuffer(fp_, 1 << 20); }
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

// Appends every record of `path` to `out`.  Returns false when the file cannot be opened.
inline bool read_fastx(const char* path, SeqSet& out, bool keep_qual) {
    ByteStream in(path);
    if (!in.ok()) return false;
    int c = in.get();
    while (c != -1 && c != '>' && c != '@') c = in.get();
    std::string name, qual;
    while (c != -1) {                                   // c is the header byte of the current record
        name.clear();
        c = in.get();
        while (c != -1 && !isspace(c)) { name.push_back((char)c); c = in.get(); }
        while (c != -1 && c != '\n') c = in.get();       // comment
        const size_t start = out.bases.size();
        c = in.get();
        while (c != -1 && c != '>' && c != '+' && c != '@') {
            if (isgraph(c)) out.bases.push_back((char)c);
            c = in.get();
        }
        const size_t len = out.bases.size() - start;
        if (c == '+') {
            c = in.get();
            while (c != -1 && c != '\n') c = in.get();   // rest of the '+' line
            if (c == -1) { out.bases.resize(start); return true; }
            qual.clear();
            c = in.get();
            while (c != -1 && qual.size() < len) {
                if (c >= 33 && c <= 127) qual.push_back((char)c);
                c = in.get();                            // note: one byte past the last quality byte is consumed
            }
            if (qual.size() != len) { out.bases.resize(start); return true; }
            out.names.push_back(name);
            out.offs.push_back((int64_t)out.bases.size());
            if (keep_qual) out.quals.push_back(qual);
            c = in.get();                                // the byte after the quality string is dropped unseen
            while (c != -1 && c != '>' && c != '@') c = in.get();
        } else {
            out.names.push_back(name);
            out.offs.push_back((int64_t)out.bases.size());
            if (keep_qual) out.quals.emplace_back();
        }
    }
    return true;
}

}  // namespace rkhost
