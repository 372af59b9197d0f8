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
#include <sys/stat.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cctype>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <cstdio>
#include <string>
#include <thread>
#include <vector>

namespace rkhost {

// Blocked gzip (BGZF: bgzip, samtools, htslib): a series of gzip members of at most 64 KB, each of which states its
// compressed size in an extra field ('B' 'C', RFC 1952 FEXTRA), so the members can be found without inflating anything and
// inflated by several threads at once.  zlib alone reads such a file as one stream on one thread (about 0.3 GB/s of FASTQ);
// this reader takes batches of members and inflates them in parallel, in order.  Anything that is not BGZF from its first
// byte on is not touched (ByteStream then uses gzopen as before).
class BgzfReader {
  public:
    ~BgzfReader() {
        if (!workers_.empty()) {
            { std::lock_guard<std::mutex> lk(m_); stop_ = true; ++gen_; }
            cv_.notify_all();
            for (auto& t : workers_) t.join();
        }
        if (fp_) fclose(fp_);
    }
    // true when `path` is a regular file that starts with a BGZF member
    bool open(const char* path, int threads) {
        struct stat st;
        if (stat(path, &st) != 0 || !S_ISREG(st.st_mode)) return false;      // a FIFO or a device must not lose the bytes a probe reads
        fp_ = fopen(path, "rb");
        if (!fp_) return false;
        unsigned char h[18];
        const bool ok = fread(h, 1, 18, fp_) == 18 && member_size(h) > 0;
        if (!ok || fseek(fp_, 0, SEEK_SET) != 0) { fclose(fp_); fp_ = nullptr; return false; }
        threads_ = std::max(1, std::min(threads, 32));
        return true;
    }
    // the next batch of inflated bytes appended to out (up to ~256 members); false at the end of the file or on a damaged member
    bool next(std::vector<char>& out, size_t& out_len) {
        if (!fp_ || done_) return false;
        struct Member { size_t in_at, in_len, out_at, out_len; };
        std::vector<Member> ms;
        in_.clear();
        size_t total_out = 0;
        while (ms.size() < 256) {
            unsigned char h[18];
            const size_t got = fread(h, 1, 18, fp_);
            if (got == 0) { done_ = true; break; }
            const long bsize = got == 18 ? member_size(h) : -1;
            if (bsize < 26) { done_ = true; bad_ = true; break; }             // not a BGZF member: stop here (truncated / foreign data)
            const size_t at = in_.size();
            in_.resize(at + (size_t)bsize);
            memcpy(in_.data() + at, h, 18);
            if (fread(in_.data() + at + 18, 1, (size_t)bsize - 18, fp_) != (size_t)bsize - 18) { in_.resize(at); done_ = true; bad_ = true; break; }
            const unsigned char* tail = reinterpret_cast<const unsigned char*>(in_.data()) + at + bsize - 4;
            const size_t isize = (size_t)tail[0] | ((size_t)tail[1] << 8) | ((size_t)tail[2] << 16) | ((size_t)tail[3] << 24);
            ms.push_back(Member{at, (size_t)bsize, total_out, isize});
            total_out += isize;
        }
        if (ms.empty()) return false;
        if (out.size() < out_len + total_out) out.resize(out_len + total_out);
        char* dst = out.data() + out_len;
        std::atomic<bool> fail{false};
        auto work = [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i) {
                const unsigned char* src = reinterpret_cast<const unsigned char*>(in_.data()) + ms[i].in_at;
                const size_t xlen = (size_t)src[10] | ((size_t)src[11] << 8);
                z_stream z{};
                if (inflateInit2(&z, -15) != Z_OK) { fail = true; return; }
                z.next_in = const_cast<unsigned char*>(src) + 12 + xlen;
                z.avail_in = (unsigned)(ms[i].in_len - 12 - xlen - 8);
                z.next_out = reinterpret_cast<unsigned char*>(dst + ms[i].out_at);
                z.avail_out = (unsigned)ms[i].out_len;
                const int rc = inflate(&z, Z_FINISH);
                if (!(rc == Z_STREAM_END || (rc == Z_OK && ms[i].out_len == 0) || (rc == Z_BUF_ERROR && ms[i].out_len == 0)) || z.total_out != ms[i].out_len) fail = true;
                inflateEnd(&z);
            }
        };
        // The helpers persist for the life of the reader: creating threads maps and unmaps their stacks, and while the CUDA
        // driver starts up on its own thread those calls queue on the process's mmap lock (a batch then took as long as
        // zlib's single stream).
        const size_t nt = std::min<size_t>((size_t)threads_, ms.size());
        if (nt <= 1) work(0, ms.size());
        else {
            while (workers_.size() + 1 < (size_t)threads_) {
                const size_t idx = workers_.size() + 1;
                workers_.emplace_back([this, idx] { helper(idx); });
            }
            {
                std::lock_guard<std::mutex> lk(m_);
                job_ = [&, nt](size_t t) { if (t < nt) work(ms.size() * t / nt, ms.size() * (t + 1) / nt); };
                pending_ = workers_.size();
                ++gen_;
            }
            cv_.notify_all();
            work(0, ms.size() / nt);                         // the caller is helper 0
            std::unique_lock<std::mutex> lk(m_);
            done_cv_.wait(lk, [&] { return pending_ == 0; });
        }
        if (fail) { done_ = true; bad_ = true; return false; }
        out_len += total_out;
        return true;
    }
    bool damaged() const { return bad_; }

  private:
    // total size of the BGZF member whose first 18 bytes are h, or -1
    static long member_size(const unsigned char* h) {
        if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) return -1;
        const unsigned xlen = h[10] | (h[11] << 8);
        if (xlen < 6 || h[12] != 'B' || h[13] != 'C' || h[14] != 2 || h[15] != 0) return -1;      // bgzip writes BC first
        return (long)(h[16] | (h[17] << 8)) + 1;
    }
    void helper(size_t idx) {
        unsigned long seen = 0;
        for (;;) {
            std::function<void(size_t)> job;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                if (stop_) return;
                job = job_;
            }
            job(idx);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--pending_ == 0) done_cv_.notify_one();
            }
        }
    }
    FILE* fp_ = nullptr;
    int threads_ = 1;
    bool done_ = false, bad_ = false;
    std::vector<char> in_;
    std::vector<std::thread> workers_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    std::function<void(size_t)> job_;
    size_t pending_ = 0;
    unsigned long gen_ = 0;
    bool stop_ = false;
};

class ByteStream {
  public:
    // "-" reads standard input (gzip or plain, like any other path)
    explicit ByteStream(const char* path, long long offset = 0) : buf_(1 << 22) {
        if (const char* e = getenv("RKMH_PARSE_BUF")) buf_.resize((size_t)std::max<long>(atol(e), 16));   // testing knob
        const bool is_stdin = path[0] == '-' && path[1] == 0;
        if (!is_stdin && offset == 0 && !getenv("RKMH_NO_BGZF")) {
            int threads = (int)std::max(1u, std::min(8u, std::thread::hardware_concurrency() / 2));
            if (const char* e = getenv("RKMH_PARSE_THREADS")) threads = std::max(1, atoi(e));
            if (bgzf_.open(path, threads)) { use_bgzf_ = true; return; }
        }
        fp_ = is_stdin ? gzdopen(0, "r") : gzopen(path, "r");
        if (fp_) gzbuffer(fp_, 1 << 20);
        if (fp_ && offset > 0 && gzseek(fp_, (z_off_t)offset, SEEK_SET) < 0) { gzclose(fp_); fp_ = nullptr; }
    }
    ~ByteStream() { if (fp_) gzclose(fp_); }
    bool ok() const { return fp_ != nullptr || use_bgzf_; }
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
            if (use_bgzf_) {
                if (!bgzf_.next(buf_, len_)) eof_ = true;       // (grows buf_ by a batch of inflated members)
                continue;
            }
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
    BgzfReader bgzf_;
    bool use_bgzf_ = false;
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

// The fast path on a memory range: p points just AFTER the '@' of a header, [p, end) is what is buffered, `eof` says
// that `end` is the end of the input.  Returns 1 and fills r / consumed / next_header when [p, ...) starts a
// well-formed four-line record (name line, one line of letters only, '+' line, as many quality bytes in [33, 127],
// newline, and then the next header byte '@' or '>' or the end of the input); 0 when the record is not of that form
// (the state machine must parse it); -1 when more bytes are needed to decide.  `consumed` counts from p up to and
// including the next record's header byte (or to the end of the input, next_header = -1).
inline bool fastx_all_letters(const char* p, size_t n) {
    unsigned bad = 0;
    for (size_t i = 0; i < n; ++i) bad |= (unsigned)((unsigned char)((p[i] | 0x20) - 'a') > 25);
    return bad == 0;
}
inline bool fastx_all_quality(const char* p, size_t n) {
    unsigned bad = 0;
    for (size_t i = 0; i < n; ++i) bad |= (unsigned)((unsigned char)(p[i] - 33) > 94);
    return bad == 0;
}
inline int fastx_fast_record(const char* p, const char* end, bool eof, Record& r, size_t& consumed, int& next_header) {
    const size_t avail = (size_t)(end - p);
    const char* e1 = (const char*)memchr(p, '\n', avail);
    const char* e2 = e1 ? (const char*)memchr(e1 + 1, '\n', (size_t)(end - (e1 + 1))) : nullptr;
    if (!e2) return eof ? 0 : -1;
    if (e2 + 1 >= end) return eof ? 0 : -1;
    if (e2[1] != '+') return 0;
    const char* e3 = (const char*)memchr(e2 + 1, '\n', (size_t)(end - (e2 + 1)));
    if (!e3) return eof ? 0 : -1;
    const char* sq = e1 + 1;
    const size_t L = (size_t)(e2 - sq);
    const char* q = e3 + 1;
    if (L == 0) return 0;
    const bool last = eof && q + L + 1 == end;
    if (!(q + L + 2 <= end || last)) return eof ? 0 : -1;
    if (q[L] != '\n') return 0;
    if (!last && q[L + 1] != '@' && q[L + 1] != '>') return 0;
    if (!fastx_all_letters(sq, L) || !fastx_all_quality(q, L)) return 0;
    size_t nl = 0;
    while (p + nl < e1 && !isspace((unsigned char)p[nl])) ++nl;
    r = Record{p, nl, sq, L, q, L};
    if (last) { consumed = (size_t)(q + L + 1 - p); next_header = -1; }
    else { consumed = (size_t)(q + L + 2 - p); next_header = (unsigned char)q[L + 1]; }
    return 1;
}

// One record at a time.  The general path is the byte-wise state machine of the header comment above.  Well-formed
// four-line FASTQ records ("@name...\n" + one line of letters + "+...\n" + as many quality bytes + "\n", followed by
// the next header or the end of the input) take a fast path that finds the four line ends with memchr and hands out
// views into the input buffer; on such records both paths give the same name, sequence and quality by construction,
// and everything else -- FASTA, wrapped lines, CRLF, '>' '+' '@' inside a sequence, truncated files -- is left to
// the state machine.  `next` returns false at the end of the file (or at a truncated quality string, where the
// reference stops too).
class FastxReader {
  public:
    explicit FastxReader(const char* path, long long offset = 0) : in_(path, offset) {
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
    bool fast(Record& r) {                               // nothing is consumed unless the record is accepted
        size_t need = 1024;
        for (;;) {
            const size_t avail = in_.fill(need);
            const char* p = in_.ptr();
            size_t consumed = 0;
            int nh = -1;
            const int rc = fastx_fast_record(p, p + avail, in_.at_eof(), r, consumed, nh);
            if (rc == 0) return false;
            if (rc == 1) { in_.skip(consumed); c_ = nh; return true; }
            need = std::max<size_t>(2 * need, 2 * avail);
        }
    }
    bool slow() {
        name_.clear(); seq_.clear(); qual_.clear();
        int c = in_.get();
        // a header byte that is the last byte of the input starts no record: ks_getuntil returns -1 before it reads a
        // name when the stream is exhausted (kseq.hpp:97, `ks->begin >= ks->end && ks->is_eof`) and kseq_read ends the file
        if (c == -1) { c_ = -1; return false; }
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
