// sketch_io.hpp -- persistence of the host front-end: reference sketches as JSON in the schema the
// reference sketches out in its dead code (dump_hash_json, src/rkmh.cpp:489-525: a Mash-like object per sequence), the
// Vowpal Wabbit line of print_wabbit (src/rkmh.cpp:463-487), and a count table as a sparse binary file (the reference only
// has stubs: HASHTCounter::write_to_binary, mkmh/HASHTCounter.cpp:29-38, writes its numbers without separators and is
// never called; docs/todo.md:1).
//
// JSON written here (one array, one object per sequence, keys in the alphabetical order nlohmann::json would give them):
//   [{"alphabet":"ATGC","canonical":"true","hashBits":64,"hashSeed":42,"hashType":"MurmurHash3_x64_128","kmer":"12",
//     "name":"...","preserveCase":"false","seqLen":7906,
//     "sketches":{"comment":"","hashes":[...],"length":1000,"name":"..."}}, ...]
// "kmer" is the space-joined k list, "length" the REQUESTED sketch size (the array may be shorter).  The reader accepts
// any key order and ignores keys it does not know.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace rkhost {

struct SketchRecord { std::string name; int64_t seq_len = 0; std::vector<uint64_t> hashes; };
struct SketchFile { std::vector<int> ks; int sketch_size = 0; std::vector<SketchRecord> recs; };

inline void json_escape(std::string& out, const std::string& s) {
    for (unsigned char c : s) {
        if (c == '"' || c == '\\') { out += '\\'; out += (char)c; }
        else if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); out += b; }
        else out += (char)c;
    }
}

inline std::string kmer_string(const std::vector<int>& ks) {
    std::string s;
    for (size_t i = 0; i < ks.size(); ++i) { if (i) s += ' '; s += std::to_string(ks[i]); }
    return s;
}

inline bool write_sketch_json(const char* path, const SketchFile& f) {
    FILE* fo = fopen(path, "w");
    if (!fo) return false;
    std::string out = "[";
    const std::string kmer = kmer_string(f.ks);
    for (size_t i = 0; i < f.recs.size(); ++i) {
        const SketchRecord& r = f.recs[i];
        std::string name; json_escape(name, r.name);
        if (i) out += ",";
        out += "\n{\"alphabet\":\"ATGC\",\"canonical\":\"true\",\"hashBits\":64,\"hashSeed\":42,\"hashType\":\"MurmurHash3_x64_128\",\"kmer\":\"";
        out += kmer; out += "\",\"name\":\""; out += name; out += "\",\"preserveCase\":\"false\",\"seqLen\":"; out += std::to_string(r.seq_len);
        out += ",\"sketches\":{\"comment\":\"\",\"hashes\":[";
        for (size_t j = 0; j < r.hashes.size(); ++j) { if (j) out += ','; out += std::to_string((unsigned long long)r.hashes[j]); }
        out += "],\"length\":"; out += std::to_string(f.sketch_size); out += ",\"name\":\""; out += name; out += "\"}}";
        if (out.size() > (1u << 22)) { fwrite(out.data(), 1, out.size(), fo); out.clear(); }
    }
    out += "\n]\n";
    fwrite(out.data(), 1, out.size(), fo);
    return fclose(fo) == 0;
}

// A pull parser over the whole file in memory: just enough JSON for the schema above.
class JsonIn {
  public:
    explicit JsonIn(std::string text) : t_(std::move(text)) {}
    std::string err;
    void ws() { while (p_ < t_.size() && (t_[p_] == ' ' || t_[p_] == '\n' || t_[p_] == '\t' || t_[p_] == '\r')) ++p_; }
    bool eat(char c) { ws(); if (p_ < t_.size() && t_[p_] == c) { ++p_; return true; } return false; }
    char peek() { ws(); return p_ < t_.size() ? t_[p_] : '\0'; }
    bool fail(const char* what) { if (err.empty()) err = std::string(what) + " at byte " + std::to_string(p_); return false; }
    bool string(std::string& out) {
        out.clear();
        if (!eat('"')) return fail("expected a string");
        while (p_ < t_.size() && t_[p_] != '"') {
            char c = t_[p_++];
            if (c == '\\' && p_ < t_.size()) {
                const char e = t_[p_++];
                if (e == 'n') c = '\n'; else if (e == 't') c = '\t'; else if (e == 'r') c = '\r'; else if (e == 'b') c = '\b'; else if (e == 'f') c = '\f';
                else if (e == 'u' && p_ + 4 <= t_.size()) { c = (char)strtol(t_.substr(p_, 4).c_str(), nullptr, 16); p_ += 4; }
                else c = e;
            }
            out += c;
        }
        if (p_ >= t_.size()) return fail("unterminated string");
        ++p_;
        return true;
    }
    bool u64(uint64_t& v) {
        ws();
        if (p_ >= t_.size() || t_[p_] < '0' || t_[p_] > '9') return fail("expected an unsigned integer");
        v = 0;
        while (p_ < t_.size() && t_[p_] >= '0' && t_[p_] <= '9') { v = v * 10 + (uint64_t)(t_[p_] - '0'); ++p_; }
        return true;
    }
    bool skip() {                                        // any value
        const char c = peek();
        if (c == '"') { std::string s; return string(s); }
        if (c == '{' || c == '[') {
            const char close = c == '{' ? '}' : ']';
            ++p_;
            if (eat(close)) return true;
            do {
                if (c == '{') { std::string k; if (!string(k) || !eat(':')) return fail("bad object"); }
                if (!skip()) return false;
            } while (eat(','));
            return eat(close) ? true : fail("unterminated container");
        }
        const size_t b = p_;
        while (p_ < t_.size() && strchr(",]} \n\t\r", t_[p_]) == nullptr) ++p_;
        return p_ > b ? true : fail("expected a value");
    }
  private:
    std::string t_;
    size_t p_ = 0;
};

inline bool read_sketch_json(const char* path, SketchFile& f, std::string& err) {
    FILE* fi = fopen(path, "rb");
    if (!fi) { err = std::string("cannot open ") + path; return false; }
    std::string text;
    char buf[1 << 16];
    size_t got;
    while ((got = fread(buf, 1, sizeof buf, fi)) > 0) text.append(buf, got);
    fclose(fi);
    JsonIn in(std::move(text));
    auto bad = [&](const char* what) { err = std::string(path) + ": " + (in.err.empty() ? what : in.err.c_str()); return false; };
    if (!in.eat('[')) return bad("expected an array of sketches");
    std::string kmer_seen;
    if (in.peek() != ']') {
        do {
            if (!in.eat('{')) return bad("expected an object");
            SketchRecord rec;
            std::string kmer;
            uint64_t length = 0;
            bool have_hashes = false;
            if (in.peek() != '}') {
                do {
                    std::string key;
                    if (!in.string(key) || !in.eat(':')) return bad("bad object");
                    if (key == "name") { if (!in.string(rec.name)) return bad("name"); }
                    else if (key == "kmer") { if (!in.string(kmer)) return bad("kmer"); }
                    else if (key == "seqLen") { uint64_t v = 0; if (!in.u64(v)) return bad("seqLen"); rec.seq_len = (int64_t)v; }
                    else if (key == "sketches") {
                        if (!in.eat('{')) return bad("sketches");
                        if (in.peek() != '}') {
                            do {
                                std::string k2;
                                if (!in.string(k2) || !in.eat(':')) return bad("bad sketches object");
                                if (k2 == "hashes") {
                                    if (!in.eat('[')) return bad("hashes");
                                    have_hashes = true;
                                    if (in.peek() != ']') {
                                        do { uint64_t v = 0; if (!in.u64(v)) return bad("hash value"); rec.hashes.push_back(v); } while (in.eat(','));
                                    }
                                    if (!in.eat(']')) return bad("hashes");
                                } else if (k2 == "length") { if (!in.u64(length)) return bad("length"); }
                                else if (!in.skip()) return bad("value");
                            } while (in.eat(','));
                        }
                        if (!in.eat('}')) return bad("sketches");
                    } else if (!in.skip()) return bad("value");
                } while (in.eat(','));
            }
            if (!in.eat('}')) return bad("unterminated object");
            if (!have_hashes) return bad("an entry without sketches.hashes");
            if (f.recs.empty()) {
                kmer_seen = kmer; f.sketch_size = (int)length;
                f.ks.clear();
                const char* q = kmer.c_str();
                while (*q) { char* e; const long v = strtol(q, &e, 10); if (e == q) break; f.ks.push_back((int)v); q = e; }
            } else if (kmer != kmer_seen || (int)length != f.sketch_size) return bad("entries with different k-mer sizes or sketch sizes");
            f.recs.push_back(std::move(rec));
        } while (in.eat(','));
    }
    if (!in.eat(']')) return bad("unterminated array");
    return true;
}

// print_wabbit (src/rkmh.cpp:463-487): label 1.0 `key|namespace hash:1 ... |sketch k:K s:S
inline void wabbit_line(std::string& out, const std::string& key, const uint64_t* mins, int n, int k0, int sketch_size) {
    out += "XYX 1.0 `";
    // join(split(key, '|'), "_"): empty fields between consecutive '|' survive, a trailing one does not (getline semantics)
    std::string joined;
    size_t b = 0;
    bool first = true;
    while (b <= key.size()) {
        const size_t e = key.find('|', b);
        if (e == std::string::npos) { if (b < key.size()) { if (!first) joined += '_'; joined += key.substr(b); } break; }
        if (!first) joined += '_';
        joined += key.substr(b, e - b);
        first = false;
        b = e + 1;
    }
    out += joined; out += "|vir";
    for (int i = 0; i < n; ++i) { out += ' '; out += std::to_string((unsigned long long)mins[i]); out += ":1"; }
    out += " |sketch k:"; out += std::to_string(k0); out += " s:"; out += std::to_string(sketch_size); out += '\n';
}

// ---- count table ------------------------------------------------------------------------------------------------
// "RKMHCNT1", u32 nk, u32 ks[8], u32 count_zeros, u32 reserved, u64 nslots, u64 n, then n x {u32 slot, i32 count} (little endian)
struct CountFile { std::vector<int> ks; int count_zeros = 0; uint64_t nslots = 0; std::vector<uint32_t> slots; std::vector<int32_t> counts; };

inline bool write_count_file(const char* path, const CountFile& c) {
    FILE* fo = fopen(path, "wb");
    if (!fo) return false;
    uint32_t head[12] = {0};
    head[0] = (uint32_t)c.ks.size();
    for (size_t i = 0; i < c.ks.size() && i < 8; ++i) head[1 + i] = (uint32_t)c.ks[i];
    head[9] = (uint32_t)c.count_zeros;
    const uint64_t n = c.slots.size();
    bool ok = fwrite("RKMHCNT1", 1, 8, fo) == 8 && fwrite(head, 4, 12, fo) == 12 && fwrite(&c.nslots, 8, 1, fo) == 1 && fwrite(&n, 8, 1, fo) == 1;
    std::vector<uint32_t> pairs(2 * (size_t)n);
    for (uint64_t i = 0; i < n; ++i) { pairs[2 * i] = c.slots[i]; pairs[2 * i + 1] = (uint32_t)c.counts[i]; }
    ok = ok && (n == 0 || fwrite(pairs.data(), 8, (size_t)n, fo) == (size_t)n);
    return (fclose(fo) == 0) && ok;
}

inline bool read_count_file(const char* path, CountFile& c, std::string& err) {
    FILE* fi = fopen(path, "rb");
    if (!fi) { err = std::string("cannot open ") + path; return false; }
    char magic[8];
    uint32_t head[12];
    uint64_t n = 0;
    bool ok = fread(magic, 1, 8, fi) == 8 && memcmp(magic, "RKMHCNT1", 8) == 0 && fread(head, 4, 12, fi) == 12 &&
              fread(&c.nslots, 8, 1, fi) == 1 && fread(&n, 8, 1, fi) == 1 && head[0] <= 8 && n <= c.nslots;
    if (ok) {
        c.ks.assign(head + 1, head + 1 + head[0]);
        c.count_zeros = (int)head[9];
        std::vector<uint32_t> pairs(2 * (size_t)n);
        ok = n == 0 || fread(pairs.data(), 8, (size_t)n, fi) == (size_t)n;
        c.slots.resize((size_t)n); c.counts.resize((size_t)n);
        for (uint64_t i = 0; ok && i < n; ++i) { c.slots[i] = pairs[2 * i]; c.counts[i] = (int32_t)pairs[2 * i + 1]; }
    }
    fclose(fi);
    if (!ok) err = std::string(path) + ": not a count table file of this program";
    return ok;
}

}  // namespace rkhost
