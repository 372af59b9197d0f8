// rkmh_main.cpp -- the `rkmh` command line (classify / stream / filter / hash) on top of the C ABI
// (include/rkmh_gpu.h).  Host side only: option parsing with the reference's letters and defaults
// (reference src/rkmh.cpp:626-720, 1037-1123, 1963-2037), FASTA/FASTQ parsing into pinned batch buffers,
// sharding of read chunks over the GPUs of the box, and formatting of the reference's exact output lines.
// Every arithmetic step of the hot path runs in the CUDA library; there is no CPU fallback.
//
// Reads are STREAMED: a parser thread fills chunks of pinned host memory (RKMH_CHUNK_MB, default 32 MB of bases each, pinned on demand),
// one worker thread per GPU pushes chunks through rkg_count_batch / rkg_classify_batch, four threads format the result
// lines, and the main thread writes them in input order, so host memory stays bounded whatever the input size (the reference loads every read
// first, src/rkmh.cpp:238-292).  `-f -` reads standard input.  With -M the read files are streamed twice (pass 1
// counts, pass 2 classifies); a read file that is not a regular file (standard input, a FIFO,
// /dev/fd/N) cannot be read twice: the reads are buffered instead.
//
// Extra, non-reference option: -g/--gpus N (default: RKMH_GPUS or 1) shards reads over N GPUs.  -t is accepted
// and ignored (the reference's OpenMP thread count).
#include <fcntl.h>
#include <getopt.h>
#include <sys/mman.h>
#include <sys/prctl.h>
#include <sys/wait.h>
#include <csignal>
#include <cerrno>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <deque>
#include <functional>
#include <list>
#include <map>
#include <set>
#include <mutex>
#include <cstdlib>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/rkmh_gpu.h"
#include "fastx.hpp"
#include "sketch_io.hpp"

#ifdef RKMH_WITH_NCCL
#include <cuda_runtime.h>
#include <nccl.h>
#endif

using rkhost::SeqSet;

namespace {

constexpr uint64_t STREAM_TABLE = 200000000ULL;   // src/rkmh.cpp:739,742
constexpr uint64_t FILTER_TABLE = 10000000ULL;    // src/rkmh.cpp:1187-1188

struct Options {
    std::vector<const char*> refs, reads;
    std::vector<int> ks;
    int sketch = 1000, threads = 1, min_occ = -1, min_matches = -1, min_diff = 0, max_samples = 100000;
    bool do_read_depth = false, do_ref_depth = false, hash_sketch_given = false, in_stream = false;
    int gpus = 1;
    std::vector<const char*> pre_refs;       // -R: pre-hashed references (JSON written by `rkmh hash -o`)
    const char* read_map_file = nullptr;      // -p: the read k-mer count table, loaded when the file exists, written otherwise
    const char* outname = nullptr;            // hash -o: JSON sketches
    bool wabbit = false;                      // hash -w: Vowpal Wabbit lines
};

void usage(const char* argv0) {
    fprintf(stderr,
            "Usage: %s { classify | stream | filter | hash | hpv16 | call } [options]\n"
            "    classify: match each read to the reference it most closely resembles using MinHash sketches.\n"
            "    stream:   the same (classify forwards to stream, as in the reference).\n"
            "    filter:   spit out reads that meet thresholds for match to ref, uniqueness, etc.\n"
            "    hash:     compute the MinHash sketches of a set of reads and/or references.\n"
            "    call:     alignment-free variant calls (VCF) of a read set against a reference, from k-mer depth (-w window).\n"
            "    hpv16:    HPV type, HPV16 lineage and sublineage of each read (-R <dir with all_pave_ref.fa and new_refs.fa>).\n"
            "Options:\n"
            "  --reference/-r <REF>      reference FASTA/FASTQ (repeatable)\n"
            "  --fasta/-f <FASTAFILE>    reads (repeatable)\n"
            "  --kmer/-k <KMERSIZE>      k-mer size, 1..32 (repeatable; default 16)\n"
            "  --sketch-size/-s <N>      sketch size (default 1000)\n"
            "  --threads/-t <N>          accepted for compatibility, ignored\n"
            "  --min-kmer-occurence/-M <N>  only use read k-mers seen at least N times in the read set\n"
            "  --min-matches/-N <N>      flag/filter reads with fewer than N matches\n"
            "  --min-diff/-D <N>         flag/filter reads whose best match is not N better than the runner-up\n"
            "  --max-samples/-I <N>      drop reference k-mers seen more than N times in the reference set\n"
            "  --pre-references/-R <F>   stream / filter: reference sketches from a JSON file written by `hash -s N -o F`\n"
            "  --read-kmer-map-file/-p <F>  with -M: load the read k-mer count table from F if it exists, else write it there\n"
            "  -o <F> / -w               hash -s N: sketches as JSON into F / as Vowpal Wabbit lines\n"
            "  --in-stream/-i            filter: then classify and report every record of standard input\n"
            "  --gpus/-g <N>             shard reads over N B200s (default $RKMH_GPUS or 1)\n",
            argv0);
}

[[noreturn]] void die(rkg_ctx* ctx, const char* what) {
    fprintf(stderr, "rkmh: %s: %s\n", what, rkg_last_error(ctx));
    exit(2);
}
#define RK(ctx, call) do { if ((call) != RKG_OK) die(ctx, #call); } while (0)

// RKMH_TIMING=1: wall-clock marks of the start-up and streaming stages on stderr
const std::chrono::steady_clock::time_point g_t0 = std::chrono::steady_clock::now();
void tmark(const char* what) {
    static const bool on = getenv("RKMH_TIMING") != nullptr;
    if (on) fprintf(stderr, "rkmh: %8.3f s  %s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - g_t0).count(), what);
}

// getopt loop shared by stream/classify and filter (same letters in the reference; `d` and unknown letters
// print the usage and abort(), 'h' / '?' print it and exit(1)).
void parse_common(int argc, char** argv, Options& o, const char* optstring) {
    static struct option long_options[] = {
        {"help", no_argument, 0, 'h'},           {"kmer", no_argument, 0, 'k'},   // sic: the reference declares --kmer no_argument
        {"fasta", required_argument, 0, 'f'},    {"reference", required_argument, 0, 'r'},
        {"sketch-size", required_argument, 0, 's'}, {"ref-sketch", required_argument, 0, 'S'},
        {"threads", required_argument, 0, 't'},  {"min-kmer-occurence", required_argument, 0, 'M'},
        {"min-matches", required_argument, 0, 'N'}, {"min-diff", required_argument, 0, 'D'},
        {"max-samples", required_argument, 0, 'I'}, {"pre-reads", required_argument, 0, 'F'},
        {"pre-references", required_argument, 0, 'R'}, {"read-kmer-map-file", required_argument, 0, 'p'},
        {"ref-kmer-map-file", required_argument, 0, 'q'}, {"in-stream", no_argument, 0, 'i'},
        {"output-reads", no_argument, 0, 'z'},   {"merge-sketch", no_argument, 0, 'm'},
        {"gpus", required_argument, 0, 'g'},     {0, 0, 0, 0}};
    optind = 2;
    for (;;) {
        int idx = 0;
        int c = getopt_long(argc, argv, optstring, long_options, &idx);
        if (c == -1) break;
        switch (c) {
            case 'i': o.in_stream = true; break;                       // filter: classify the records of standard input (below)
            case 'm': case 'z': break;                                 // parsed, no effect (reference: same)
            case 'R': o.pre_refs.push_back(optarg); break;            // the reference parses these and stops there (src/rkmh.cpp:659-670)
            case 'p': o.read_map_file = optarg; break;
            case 'F': case 'q': case 'S': break;                       // parsed, no effect
            case 'o': o.outname = optarg; break;
            case 'w': o.wabbit = true; break;
            case 'c': case 'K': case 'T': break;                       // hash: parsed, no effect
            case 't': o.threads = atoi(optarg); break;
            case 'r': o.refs.push_back(optarg); break;
            case 'f': o.reads.push_back(optarg); break;
            case 'k': if (optarg) o.ks.push_back(atoi(optarg)); break;
            case 'N': o.min_matches = atoi(optarg); break;
            case 'D': o.min_diff = atoi(optarg); break;
            case 's': o.sketch = atoi(optarg); o.hash_sketch_given = true; break;
            case 'M': o.min_occ = atoi(optarg); o.do_read_depth = true; break;
            case 'I': o.max_samples = atoi(optarg); o.do_ref_depth = true; break;
            case 'g': o.gpus = atoi(optarg); break;
            case '?': case 'h': usage(argv[0]); exit(1);
            default: usage(argv[0]); abort();
        }
    }
    if (o.ks.empty()) {
        fprintf(stderr, "No kmer size(s) provided. Will use a default kmer size of 16.\n");
        o.ks.push_back(16);
    }
}

void load_all(const std::vector<const char*>& files, SeqSet& set, bool keep_qual) {
    for (const char* f : files)
        if (!rkhost::read_fastx(f, set, keep_qual)) { fprintf(stderr, "rkmh: cannot open %s\n", f); exit(2); }
}

// One context per GPU, parameters and references replicated on each.
struct Fleet {
    std::vector<rkg_ctx*> ctx;
    std::thread bg; std::mutex m; std::condition_variable cv; bool ready = true;
    void create_all() {                                  // contexts are created concurrently: ~1.3 s each on a B200 box
        std::vector<std::thread> th;
        for (size_t d = 0; d < ctx.size(); ++d)
            th.emplace_back([this, d] {
                if (rkg_create((int)d, &ctx[d]) != RKG_OK) { fprintf(stderr, "rkmh: %s\n", rkg_last_error(nullptr)); exit(2); }
            });
        for (auto& t : th) t.join();
        tmark("CUDA contexts created");
    }
    void open(int n) { ctx.assign((size_t)n, nullptr); create_all(); }
    // The same on a background thread, followed by `after` (parameters, references): CUDA start-up takes over a second,
    // during which the read files are already being parsed into chunks.  Users call wait() before touching a context.
    void open_async(int n, std::function<void()> after) {
        ctx.assign((size_t)n, nullptr);
        ready = false;
        bg = std::thread([this, after] {
            create_all();
            after();
            tmark("references loaded");
            { std::lock_guard<std::mutex> g(m); ready = true; }
            cv.notify_all();
        });
    }
    void wait() { std::unique_lock<std::mutex> g(m); cv.wait(g, [&] { return ready; }); }
    bool is_ready() { std::lock_guard<std::mutex> g(m); return ready; }
    ~Fleet() { if (bg.joinable()) bg.join(); for (auto c : ctx) rkg_destroy(c); }
    template <typename F> void each(F&& f) {
        std::vector<std::thread> th;
        for (size_t d = 0; d < ctx.size(); ++d) th.emplace_back([&, d] { f((int)d, ctx[d]); });
        for (auto& t : th) t.join();
    }
};

// The only global point of the path: sum the per-GPU read count tables (src/rkmh.cpp:910, the barrier between
// pass 1 and pass 2).
void allreduce_counts(Fleet& fl) {
    if (fl.ctx.size() < 2) return;
#ifdef RKMH_WITH_NCCL
    const int n = (int)fl.ctx.size();
    std::vector<ncclComm_t> comms(n);
    std::vector<int> devs(n);
    for (int d = 0; d < n; ++d) devs[d] = d;
    auto nccl_ok = [](ncclResult_t r, const char* what) {
        if (r != ncclSuccess) { fprintf(stderr, "rkmh: %s: %s\n", what, ncclGetErrorString(r)); exit(2); }
    };
    // NCCL may print its version banner on stdout (NCCL_DEBUG=VERSION); stdout is the result stream, so send
    // anything the library prints during initialisation to stderr instead.
    fflush(stdout);
    const int saved_stdout = dup(1);
    dup2(2, 1);
    nccl_ok(ncclCommInitAll(comms.data(), n, devs.data()), "ncclCommInitAll");
    fflush(stdout);
    dup2(saved_stdout, 1);
    close(saved_stdout);
    std::vector<cudaStream_t> streams(n);
    for (int d = 0; d < n; ++d) { cudaSetDevice(d); cudaDeviceSynchronize(); cudaStreamCreate(&streams[d]); }
    nccl_ok(ncclGroupStart(), "ncclGroupStart");
    for (int d = 0; d < n; ++d) {
        void* p = nullptr; uint64_t slots = 0;
        RK(fl.ctx[d], rkg_count_table(fl.ctx[d], &p, &slots));
        cudaSetDevice(d);
        nccl_ok(ncclAllReduce(p, p, (size_t)slots, ncclInt32, ncclSum, comms[d], streams[d]), "ncclAllReduce");
    }
    nccl_ok(ncclGroupEnd(), "ncclGroupEnd");
    for (int d = 0; d < n; ++d) {
        cudaSetDevice(d);
        cudaError_t e = cudaStreamSynchronize(streams[d]);
        if (e != cudaSuccess) { fprintf(stderr, "rkmh: all-reduce failed: %s\n", cudaGetErrorString(e)); exit(2); }
        cudaStreamDestroy(streams[d]);
        ncclCommDestroy(comms[d]);
    }
#else
    fprintf(stderr, "rkmh: -M with more than one GPU needs the NCCL build (make WITH_NCCL=1)\n");
    exit(2);
#endif
}

// Reads are sharded over the GPUs in contiguous ranges; every range goes through rkg_classify_batch.
void shard(size_t n, int parts, int d, size_t& lo, size_t& hi) { lo = n * d / parts; hi = n * (d + 1) / parts; }

void append_int(std::string& s, long long v) { char b[24]; int n = snprintf(b, sizeof b, "%lld", v); s.append(b, n); }

// Pre-hashed references (-R file.json, written by `rkmh hash -o`): names into refs.names (zero-length sequences), sketches
// kept for setup_refs.  The file's k list and sketch size must be the command line's.
struct PreRefs { std::vector<uint64_t> sketches; std::vector<int32_t> lens; };
void load_pre_refs(const Options& o, SeqSet& refs, PreRefs& pre) {
    for (const char* path : o.pre_refs) {
        rkhost::SketchFile f; std::string err;
        if (!rkhost::read_sketch_json(path, f, err)) { fprintf(stderr, "rkmh: %s\n", err.c_str()); exit(2); }
        if (f.recs.empty()) continue;
        if (f.ks != o.ks || f.sketch_size != o.sketch) {
            fprintf(stderr, "rkmh: %s holds sketches for -k %s -s %d, the command line asks for -k %s -s %d\n", path,
                    rkhost::kmer_string(f.ks).c_str(), f.sketch_size, rkhost::kmer_string(o.ks).c_str(), o.sketch);
            exit(2);
        }
        for (auto& r : f.recs) {
            if ((int)r.hashes.size() > o.sketch) { fprintf(stderr, "rkmh: %s: a sketch longer than its sketch size\n", path); exit(2); }
            refs.names.push_back(r.name);
            refs.offs.push_back(refs.offs.back());
            pre.lens.push_back((int32_t)r.hashes.size());
            r.hashes.resize((size_t)o.sketch, 0);
            pre.sketches.insert(pre.sketches.end(), r.hashes.begin(), r.hashes.end());
        }
    }
}

void setup_refs(Fleet& fl, const Options& o, const SeqSet& refs, uint64_t table, bool filter_mode, const PreRefs* pre = nullptr) {
    if (pre && !pre->lens.empty()) {
        // references that come as sequences are sketched on the GPU first, then all sketches make one table
        const bool use_i_pre = filter_mode ? (o.max_samples < 100000) : o.do_ref_depth;
        if (use_i_pre) { fprintf(stderr, "rkmh: -I needs the reference sequences (it filters by k-mer counts over them); not with -R\n"); exit(2); }
        const size_t nseq = refs.size() - pre->lens.size();
        fl.each([&](int, rkg_ctx* c) {
            RK(c, rkg_set_params(c, o.ks.data(), (int)o.ks.size(), o.sketch, o.do_read_depth ? table : 0, 0));
            std::vector<uint64_t> sk((size_t)refs.size() * (size_t)o.sketch, 0);
            std::vector<int32_t> ln(refs.size(), 0);
            if (nseq) {
                RK(c, rkg_load_refs(c, refs.bases.data(), refs.offs.data(), (int)nseq, -1, filter_mode ? 1 : 0));
                RK(c, rkg_get_ref_sketches(c, sk.data(), ln.data()));
            }
            std::copy(pre->sketches.begin(), pre->sketches.end(), sk.begin() + nseq * (size_t)o.sketch);
            std::copy(pre->lens.begin(), pre->lens.end(), ln.begin() + nseq);
            RK(c, rkg_load_ref_sketches(c, sk.data(), ln.data(), (int)refs.size()));
        });
        return;
    }
    const bool use_i = filter_mode ? (o.max_samples < 100000) : o.do_ref_depth;
    fl.each([&](int, rkg_ctx* c) {
        RK(c, rkg_set_params(c, o.ks.data(), (int)o.ks.size(), o.sketch, o.do_read_depth ? table : 0, use_i ? table : 0));
        RK(c, rkg_load_refs(c, refs.bases.data(), refs.offs.data(), (int)refs.size(), use_i ? o.max_samples : -1,
                            filter_mode ? 1 : 0));
    });
}

// -p FILE (the reference's --read-kmer-map-file, parsed and ignored there): the read k-mer count table of pass 1 is loaded
// from FILE when it exists -- pass 1 is then skipped --, and written to FILE after pass 1 otherwise.
bool load_count_table(Fleet& fl, const Options& o, bool count_zeros) {
    if (!o.read_map_file) return false;
    struct stat st;
    if (stat(o.read_map_file, &st) != 0) return false;
    rkhost::CountFile cf; std::string err;
    if (!rkhost::read_count_file(o.read_map_file, cf, err)) { fprintf(stderr, "rkmh: %s\n", err.c_str()); exit(2); }
    void* p = nullptr; uint64_t nslots = 0;
    RK(fl.ctx[0], rkg_count_table(fl.ctx[0], &p, &nslots));
    if (cf.ks != o.ks || cf.nslots != nslots || cf.count_zeros != (count_zeros ? 1 : 0)) {
        fprintf(stderr, "rkmh: %s was counted with other parameters (k list, table size or command)\n", o.read_map_file);
        exit(2);
    }
    fl.each([&](int, rkg_ctx* c) { RK(c, rkg_count_import(c, 0, cf.slots.data(), cf.counts.data(), cf.slots.size())); });
    return true;
}

void save_count_table(Fleet& fl, const Options& o, bool count_zeros) {
    if (!o.read_map_file) return;
    rkg_ctx* c = fl.ctx[0];
    rkhost::CountFile cf;
    cf.ks = o.ks; cf.count_zeros = count_zeros ? 1 : 0;
    void* p = nullptr;
    RK(c, rkg_count_table(c, &p, &cf.nslots));
    uint64_t n = 0;
    RK(c, rkg_count_export(c, 0, nullptr, nullptr, 0, &n));
    cf.slots.resize((size_t)n); cf.counts.resize((size_t)n);
    if (n) RK(c, rkg_count_export(c, 0, cf.slots.data(), cf.counts.data(), n, &n));
    if (!rkhost::write_count_file(o.read_map_file, cf)) { fprintf(stderr, "rkmh: cannot write %s\n", o.read_map_file); exit(2); }
}

std::vector<rkg_hit> classify_all(Fleet& fl, const Options& o, const SeqSet& reads, bool count_zeros) {
    fl.wait();
    const int G = (int)fl.ctx.size();
    const size_t n = reads.size();
    std::vector<rkg_hit> hits(n);
    if (o.do_read_depth && !load_count_table(fl, o, count_zeros)) {   // pass 1: global k-mer count table
        fl.each([&](int d, rkg_ctx* c) {
            size_t lo, hi; shard(n, G, d, lo, hi);
            RK(c, rkg_count_reset(c));
            if (hi > lo) RK(c, rkg_count_batch(c, reads.bases.data(), reads.offs.data() + lo, (int64_t)(hi - lo), count_zeros));
        });
        allreduce_counts(fl);
        save_count_table(fl, o, count_zeros);
    }
    fl.each([&](int d, rkg_ctx* c) {
        size_t lo, hi; shard(n, G, d, lo, hi);
        if (hi > lo)
            RK(c, rkg_classify_batch(c, reads.bases.data(), reads.offs.data() + lo, (int64_t)(hi - lo),
                                     o.do_read_depth ? o.min_occ : -1, hits.data() + lo));
    });
    return hits;
}

// ---- streaming pipeline -----------------------------------------------------------------------------------
// One chunk of reads in pinned host memory, in the layout the C ABI takes.
struct Chunk {
    char* bases = nullptr; int64_t* offs = nullptr; rkg_hit* hits = nullptr;
    size_t cap_bases = 0, cap_reads = 0, nreads = 0;
    std::vector<char> big;                 // a single record larger than cap_bases goes here (pageable)
    std::vector<char> text;                // names (and quality strings) of the chunk's reads, back to back
    std::vector<uint64_t> name_at, qual_at; // name i = text[name_at[i], name_at[i + 1]); qualities likewise when kept
    uint64_t seq = 0;                      // position of the chunk in the input order
    bool pinned = false;                   // bases / offs / hits are cudaHostAlloc'ed (else malloc'ed: host-only modes, or made before CUDA was up)
    std::string out;                       // formatted result lines
    std::vector<int32_t> extra;            // hpv16: per read, the counts of the lineage / sublineage tables
    const char* data() const { return big.empty() ? bases : big.data(); }
    void reset(uint64_t s) {
        nreads = 0; offs[0] = 0; big.clear(); big.shrink_to_fit(); text.clear(); name_at.clear(); qual_at.clear(); out.clear(); seq = s;
    }
    void add_text(const rkhost::Record& r, bool keep_qual) {
        name_at.push_back(text.size());
        text.insert(text.end(), r.name, r.name + r.name_len);
        if (keep_qual) { qual_at.push_back(text.size()); text.insert(text.end(), r.qual, r.qual + r.qual_len); }
    }
    // (pointer, length) of read i's name / quality string; valid once the chunk is complete
    std::pair<const char*, size_t> name(size_t i) const {
        const uint64_t e = !qual_at.empty() ? qual_at[i] : (i + 1 < name_at.size() ? name_at[i + 1] : text.size());
        return {text.data() + name_at[i], (size_t)(e - name_at[i])};
    }
    std::pair<const char*, size_t> qual(size_t i) const {
        const uint64_t e = i + 1 < name_at.size() ? name_at[i + 1] : text.size();
        return {text.data() + qual_at[i], (size_t)(e - qual_at[i])};
    }
};

template <typename T> class Channel {      // unbounded MPMC queue; the chunk pool bounds what is in flight
  public:
    void push(T v) { { std::lock_guard<std::mutex> g(m_); q_.push_back(v); } cv_.notify_one(); }
    T pop() { std::unique_lock<std::mutex> g(m_); cv_.wait(g, [&] { return !q_.empty(); }); T v = q_.front(); q_.pop_front(); return v; }
  private:
    std::mutex m_; std::condition_variable cv_; std::deque<T> q_;
};

// Chunks are pinned on demand (page-locking costs ~0.5 s per GB: a small input should not pay for the whole pool) up to
// `max_chunks`; after that the parser waits for the writer to hand one back.
struct ChunkPool {
    Fleet* fleet = nullptr;                              // null: host-only modes (ordinary memory)
    std::vector<Chunk*> all;
    std::deque<Chunk*> free_list;
    std::mutex m; std::condition_variable cv;
    int max_chunks = 0;
    size_t cap_bases = 0;
    void open(Fleet* fl, int n, size_t cap) { fleet = fl; max_chunks = n; cap_bases = cap; }
    Chunk* make() {
        const size_t cap_reads = std::max<size_t>(cap_bases / 32, 1024);
        Chunk* c = new Chunk();
        void *pb = nullptr, *po = nullptr, *ph = nullptr;
        // pinned once CUDA is up; the chunks made while the contexts are still being created are pageable (their copies
        // are staged by the driver: slower, but they exist a second earlier)
        rkg_ctx* owner = (fleet && fleet->is_ready()) ? fleet->ctx[0] : nullptr;
        if (owner) {
            RK(owner, rkg_host_alloc(owner, cap_bases, &pb));
            RK(owner, rkg_host_alloc(owner, sizeof(int64_t) * (cap_reads + 1), &po));
            RK(owner, rkg_host_alloc(owner, sizeof(rkg_hit) * cap_reads, &ph));
            c->pinned = true;
        } else {
            // huge pages: 4 KB first-touch faults of eight parser threads (60 k of them for the pool) queue on the process's
            // mmap lock together with the CUDA driver's start-up on its own thread and were seen to triple its time
            auto big = [](size_t bytes) -> void* {
                void* p = nullptr;
                const size_t huge = (size_t)2 << 20, sz = (bytes + huge - 1) / huge * huge;
                if (posix_memalign(&p, huge, sz) != 0) return nullptr;
                madvise(p, sz, MADV_HUGEPAGE);
                return p;
            };
            pb = big(cap_bases); po = big(sizeof(int64_t) * (cap_reads + 1)); ph = big(sizeof(rkg_hit) * cap_reads);
            if (!pb || !po || !ph) { fprintf(stderr, "rkmh: out of memory\n"); exit(2); }
        }
        c->bases = (char*)pb; c->offs = (int64_t*)po; c->hits = (rkg_hit*)ph;
        c->cap_bases = cap_bases; c->cap_reads = cap_reads;
        return c;
    }
    Chunk* acquire() {
        std::unique_lock<std::mutex> g(m);
        for (;;) {
            if (!free_list.empty()) { Chunk* c = free_list.front(); free_list.pop_front(); return c; }
            if ((int)all.size() < max_chunks) {
                all.push_back(nullptr);
                const size_t slot = all.size() - 1;
                g.unlock();
                Chunk* c = make();
                g.lock();
                all[slot] = c;
                return c;
            }
            cv.wait(g);
        }
    }
    void release(Chunk* c) { { std::lock_guard<std::mutex> g(m); free_list.push_back(c); } cv.notify_one(); }
    ~ChunkPool() {
        for (Chunk* c : all) if (c) {
            if (c->pinned) { rkg_ctx* o = fleet->ctx[0]; rkg_host_free(o, c->bases); rkg_host_free(o, c->offs); rkg_host_free(o, c->hits); }
            else { free(c->bases); free(c->offs); free(c->hits); }
            delete c;
        }
    }
};

size_t chunk_bytes() {
    if (const char* b = getenv("RKMH_CHUNK_BYTES")) return (size_t)std::max<long>(atol(b), 64);   // testing knob
    const char* e = getenv("RKMH_CHUNK_MB");
    // 8 MB of bases per chunk (about 55 k short reads): small enough that the pool of pinned chunks (a few per parser
    // thread) page-locks in tens of milliseconds, large enough that a chunk's kernel launches do not show
    const long mb = e ? atol(e) : 8;
    return (size_t)std::max<long>(mb, 1) << 20;
}

// Threads of the parallel segment parser (plain four-line FASTQ files of at least two segments; everything else is read
// by the sequential state machine).  RKMH_PARSE_THREADS overrides; 1 turns the parallel parser off.  Round 1 measured the
// parallel parser SLOWER than the sequential one on the B200 box (2.9 s against 1.95 s for 8 M reads): its pool of 32 MB
// chunks page-locked 1.4 GB before the first read was classified.  With 8 MB chunks and two chunks of text per segment
// the pool is 4x smaller and the parser scales.
int parse_threads() {
    if (const char* e = getenv("RKMH_PARSE_THREADS")) return std::min(std::max(atoi(e), 1), 16);
    const unsigned hw = std::thread::hardware_concurrency();
    return (int)std::min(8u, std::max(1u, hw / 2));      // half of the cores, at most 8: the rest format, write and feed the GPUs
}

size_t parse_segment_bytes(size_t cap_bases) {          // about one chunk of bases per segment of FASTQ text
    if (const char* e = getenv("RKMH_PARSE_SEG")) return (size_t)std::max<long>(atol(e), 256);   // testing knob
    return 2 * cap_bases;
}

// A plain (uncompressed, regular) FASTQ file for the parallel parser.  Segments are read with pread() into per-thread
// buffers, not mapped: eight threads page-faulting a mapping hold the process's mmap lock against the CUDA driver, whose
// start-up (itself full of mmap calls) ran 2.3 s instead of 1.2 s beside them on the B200 box.
struct PlainFile {
    size_t size = 0; int fd = -1;
    bool open(const char* path) {
        if (path[0] == '-' && path[1] == 0) return false;
        fd = ::open(path, O_RDONLY);
        if (fd < 0) return false;
        struct stat st;
        if (fstat(fd, &st) != 0 || !S_ISREG(st.st_mode) || st.st_size < 4) return false;
        size = (size_t)st.st_size;
        unsigned char magic[2] = {0, 0};
        if (pread(fd, magic, 2, 0) != 2 || (magic[0] == 0x1f && magic[1] == 0x8b)) return false;   // gzip: sequential reader
        posix_fadvise(fd, 0, 0, POSIX_FADV_SEQUENTIAL);
        return true;
    }
    // bytes [lo, hi) into dst; false on a short read
    bool read(size_t lo, size_t hi, char* dst) const {
        while (lo < hi) {
            const ssize_t got = pread(fd, dst, hi - lo, (off_t)lo);
            if (got <= 0) return false;
            lo += (size_t)got; dst += got;
        }
        return true;
    }
    ~PlainFile() { if (fd >= 0) ::close(fd); }
};

// Appends a record to the chunk being filled, closing it (-> `full`) and taking a new one from the pool when needed.
struct ChunkFiller {
    ChunkPool& pool; bool keep_qual; std::function<void(Chunk*)> full;
    Chunk* cur = nullptr;
    void add(const rkhost::Record& rec) {
        if (cur && (cur->nreads >= cur->cap_reads || !cur->big.empty() ||
                    (size_t)cur->offs[cur->nreads] + rec.seq_len > cur->cap_bases)) flush();
        if (!cur) { cur = pool.acquire(); cur->reset(0); }
        const size_t at = (size_t)cur->offs[cur->nreads];
        if (rec.seq_len > cur->cap_bases) cur->big.assign(rec.seq, rec.seq + rec.seq_len);      // alone in its chunk
        else memcpy(cur->bases + at, rec.seq, rec.seq_len);
        cur->offs[++cur->nreads] = (int64_t)(at + rec.seq_len);
        cur->add_text(rec, keep_qual);
    }
    void flush() { if (cur) { full(cur); cur = nullptr; } }
};

// Parallel parse of the well-formed prefix of a plain four-line FASTQ file.  The file is cut into segments; a team of
// threads parses them with the fast path only (fastx.hpp), each starting at the first line-initial '@' of its segment
// from which two well-formed records follow.  The calling thread then walks the segments IN ORDER and accepts one only
// if it starts exactly where the previous one ended: by induction every accepted record is what the sequential
// state machine would have produced, because on well-formed records both paths agree and the chain starts at the
// file's first header.  At the first segment that does not chain, or the first record that is not well-formed, the
// function stops and returns the byte offset at which the sequential reader must take over (the file size when
// everything was parsed here, 0 when nothing was).
size_t parse_parallel_prefix(const char* path, ChunkPool& pool, bool keep_qual, int threads, size_t seg_bytes,
                             const std::function<void(Chunk*)>& emit_chunk, uint64_t& nrecords) {
    PlainFile pf;
    if (threads < 2 || !pf.open(path) || pf.size < 2 * seg_bytes) return 0;
    const size_t size = pf.size, npos = (size_t)-1;
    // what a thread reads beyond its segment: the record that straddles the end, and the next one (which the search for a
    // segment's first record wants to see whole); a record that needs more ends the parallel prefix
    const size_t slack = std::min<size_t>(std::max<size_t>(seg_bytes / 2, 1 << 16), 4 << 20);
    // RKMH_PARSE_MMAP=1: map the file instead (A/B timing of the two ways to read it)
    const char* mapped = nullptr;
    if (getenv("RKMH_PARSE_MMAP")) {
        void* mp = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, pf.fd, 0);
        if (mp != MAP_FAILED) { mapped = (const char*)mp; madvise(mp, size, MADV_SEQUENTIAL); }
    }
    struct Unmap { const char* p; size_t n; ~Unmap() { if (p) munmap((void*)p, n); } } unmap{mapped, size};
    size_t start0 = 0;
    {
        std::vector<char> head(std::min<size_t>(size, 1 << 16));
        if (!pf.read(0, head.size(), head.data())) return 0;
        while (start0 < head.size() && head[start0] != '>' && head[start0] != '@') ++start0;
        if (start0 >= head.size() || head[start0] != '@') return 0;
    }
    struct Seg { size_t first = (size_t)-1, end = (size_t)-1; bool anomaly = false, done = false; std::vector<Chunk*> chunks; uint64_t nrec = 0; };
    const size_t nseg = (size + seg_bytes - 1) / seg_bytes;
    std::vector<Seg> segs(nseg);
    std::mutex m; std::condition_variable cv;
    size_t next_seg = 0, verified = 0;
    std::atomic<bool> abort_all{false};
    auto parse_segment = [&](size_t i, std::vector<char>& buf) {
        Seg& sg = segs[i];
        const size_t lo = i * seg_bytes, hi = std::min(size, lo + seg_bytes);
        // file bytes [rlo, rhi) are in buf; d[x] is file byte x for x in that range
        const size_t rlo = lo > 0 ? lo - 1 : 0, rhi = std::min(size, hi + slack);
        sg.first = npos;
        const char* d = mapped;
        if (!mapped) {
            if (buf.size() < rhi - rlo) buf.resize(rhi - rlo);
            if (!pf.read(rlo, rhi, buf.data())) { sg.anomaly = true; return; }
            d = buf.data() - rlo;
        }
        auto well_formed_at = [&](size_t pos, rkhost::Record& r, size_t& next) -> bool {     // d[pos] == '@'
            size_t consumed = 0; int nh = -1;
            if (rkhost::fastx_fast_record(d + pos + 1, d + rhi, rhi == size, r, consumed, nh) != 1) return false;
            next = nh == -1 ? size : pos + consumed;
            return true;
        };
        rkhost::Record r; size_t nx = 0, nx2 = 0;
        size_t pos = npos;
        if (i == 0) pos = start0;
        else
            for (size_t p = std::max<size_t>(lo, 1); p < hi; ++p) {
                const char* hit = (const char*)memchr(d + p, '@', hi - p);
                if (!hit) break;
                p = (size_t)(hit - d);
                if (d[p - 1] != '\n') continue;
                if (well_formed_at(p, r, nx) && (nx >= size || (nx < rhi && d[nx] == '@' && well_formed_at(nx, r, nx2)))) { pos = p; break; }
            }
        sg.first = pos;
        if (pos == npos) return;
        ChunkFiller filler{pool, keep_qual, [&](Chunk* c) { sg.chunks.push_back(c); }};
        while (pos < hi) {
            if (abort_all.load(std::memory_order_relaxed)) break;
            // anything the fast path does not take, a record larger than a chunk, or a fourth chunk for one segment
            // (the pool is sized for three per segment in flight) ends the parallel prefix here
            if (d[pos] != '@' || !well_formed_at(pos, r, nx) || r.seq_len > pool.cap_bases ||
                (sg.chunks.size() >= 2 && filler.cur && ((size_t)filler.cur->offs[filler.cur->nreads] + r.seq_len > filler.cur->cap_bases ||
                                                         filler.cur->nreads >= filler.cur->cap_reads))) { sg.anomaly = true; break; }
            filler.add(r);
            ++sg.nrec;
            pos = nx;
        }
        filler.flush();
        sg.end = pos;
    };
    std::vector<std::thread> team;
    for (int t = 0; t < threads; ++t)
        team.emplace_back([&] {
            std::vector<char> buf;                       // this thread's segment buffer, reused (touched once)
            for (;;) {
                size_t i;
                {
                    std::unique_lock<std::mutex> g(m);
                    cv.wait(g, [&] { return abort_all || next_seg >= nseg || next_seg < verified + (size_t)threads; });
                    if (abort_all || next_seg >= nseg) return;
                    i = next_seg++;
                }
                parse_segment(i, buf);
                { std::lock_guard<std::mutex> g(m); segs[i].done = true; }
                cv.notify_all();
            }
        });
    size_t expected = start0;
    for (size_t i = 0; i < nseg; ++i) {
        { std::unique_lock<std::mutex> g(m); cv.wait(g, [&] { return segs[i].done; }); }
        Seg& sg = segs[i];
        bool stop = false;
        if (sg.first != npos) {
            if (sg.first != expected) stop = true;                   // does not chain: the sequential reader decides
            else {
                for (Chunk* c : sg.chunks) emit_chunk(c);
                sg.chunks.clear();
                nrecords += sg.nrec;
                expected = sg.end;
                stop = sg.anomaly;
            }
        }
        if (stop) break;
        { std::lock_guard<std::mutex> g(m); verified = i + 1; }
        cv.notify_all();
    }
    // segments that were parsed ahead but not accepted give their chunks back (first the finished ones, so that a
    // thread still waiting for a chunk can see the abort flag and leave)
    std::vector<Chunk*> back;
    {
        std::lock_guard<std::mutex> g(m);
        abort_all = true;
        for (Seg& sg : segs) if (sg.done) { back.insert(back.end(), sg.chunks.begin(), sg.chunks.end()); sg.chunks.clear(); }
    }
    cv.notify_all();
    for (Chunk* c : back) pool.release(c);
    for (auto& t : team) t.join();
    for (Seg& sg : segs) for (Chunk* c : sg.chunks) pool.release(c);
    return std::min(expected, size);
}

// Every record of `files`, in order, as chunks handed to `emit_chunk` (called on this thread).
uint64_t parse_files(const std::vector<const char*>& files, ChunkPool& pool, bool keep_qual,
                     const std::function<void(Chunk*)>& emit_chunk) {
    uint64_t nrecords = 0;
    ChunkFiller filler{pool, keep_qual, emit_chunk};
    rkhost::Record rec;
    const int pthreads = parse_threads();
    for (const char* f : files) {
        filler.flush();                                  // chunks of the parallel parser must follow what came before
        const size_t resume = parse_parallel_prefix(f, pool, keep_qual, pthreads, parse_segment_bytes(pool.cap_bases),
                                                    emit_chunk, nrecords);
        if (getenv("RKMH_PARSE_DEBUG")) fprintf(stderr, "rkmh: %s: %llu records from the parallel parser, sequential from byte %zu\n", f, (unsigned long long)nrecords, resume);
        rkhost::FastxReader rd(f, (long long)resume);
        if (!rd.ok()) { fprintf(stderr, "rkmh: cannot open %s\n", f); exit(2); }
        while (rd.next(rec)) { ++nrecords; filler.add(rec); }
    }
    filler.flush();
    return nrecords;
}

// Streams every record of `files` through `work` (called on one worker thread per GPU, any order), `format` (a small
// thread pool) and then through `emit` (called on the calling thread, in input order).  Returns the number of records.
uint64_t stream_reads(Fleet& fl, ChunkPool& pool, const std::vector<const char*>& files, bool keep_qual,
                      const std::function<void(int, rkg_ctx*, Chunk&)>& work, const std::function<void(Chunk&)>& format,
                      const std::function<void(Chunk&)>& emit) {
    Channel<Chunk*> work_q, format_q;
    const int nformat = format ? 4 : 0;                  // result lines are formatted off the GPU workers' threads
    std::mutex dm; std::condition_variable dcv; std::map<uint64_t, Chunk*> done;
    std::atomic<int64_t> total_chunks{-1};
    uint64_t nrecords = 0;
    std::thread reader([&] {
        uint64_t seq = 0;
        nrecords = parse_files(files, pool, keep_qual, [&](Chunk* c) { c->seq = seq++; work_q.push(c); });
        tmark("input parsed");
        // stored under the writer's mutex: a notify between the writer's predicate check and its wait would otherwise
        // be lost, and with every chunk already emitted (or none at all: an empty input) nothing would wake it again
        { std::lock_guard<std::mutex> g(dm); total_chunks = (int64_t)seq; }
        dcv.notify_all();
        for (size_t d = 0; d < fl.ctx.size(); ++d) work_q.push(nullptr);
    });
    std::vector<std::thread> workers;
    for (size_t d = 0; d < fl.ctx.size(); ++d)
        workers.emplace_back([&, d] {
            fl.wait();                                   // contexts created, parameters and references in place
            for (;;) {
                Chunk* c = work_q.pop();
                if (!c) break;
                work((int)d, fl.ctx[d], *c);
                if (nformat) { format_q.push(c); continue; }
                { std::lock_guard<std::mutex> g(dm); done[c->seq] = c; }
                dcv.notify_all();
            }
        });
    std::vector<std::thread> formatters;
    for (int i = 0; i < nformat; ++i)
        formatters.emplace_back([&] {
            for (;;) {
                Chunk* c = format_q.pop();
                if (!c) break;
                format(*c);
                { std::lock_guard<std::mutex> g(dm); done[c->seq] = c; }
                dcv.notify_all();
            }
        });
    for (uint64_t next = 0;; ++next) {
        Chunk* c = nullptr;
        {
            std::unique_lock<std::mutex> g(dm);
            dcv.wait(g, [&] { return done.count(next) || (total_chunks >= 0 && (int64_t)next >= total_chunks); });
            auto it = done.find(next);
            if (it == done.end()) break;
            c = it->second; done.erase(it);
        }
        emit(*c);
        pool.release(c);
    }
    tmark("results written");
    reader.join();
    for (auto& w : workers) w.join();
    for (int i = 0; i < nformat; ++i) format_q.push(nullptr);
    for (auto& f : formatters) f.join();
    return nrecords;
}

// -M reads the input twice (pass 1 counts, pass 2 classifies): only regular files can be re-opened.  Standard input,
// FIFOs, /dev/stdin, /dev/fd/N (`<(zcat ...)`) and anything else that is not a regular file is buffered instead, as the
// reference does for every input (src/rkmh.cpp:786-788).
bool any_unrewindable(const std::vector<const char*>& files) {
    for (const char* f : files) {
        if (f[0] == '-' && f[1] == 0) return true;
        struct stat st;
        if (stat(f, &st) != 0 || !S_ISREG(st.st_mode)) return true;
    }
    return false;
}

// The classification of every read of `files`, streamed; `format` turns a classified chunk into output lines.
void run_streamed(Fleet& fl, const Options& o, bool keep_qual, bool count_zeros,
                  const std::function<void(Chunk&)>& format) {
    ChunkPool pool;
    pool.open(&fl, 2 * (int)fl.ctx.size() + 6 + 3 * parse_threads(), chunk_bytes());
    const auto write_out = [](Chunk& c) { fwrite(c.out.data(), 1, c.out.size(), stdout); };
    uint64_t pass1 = 0;
    bool counted = false;
    if (o.do_read_depth) {                              // pass 1: global k-mer count table (src/rkmh.cpp:904-910)
        fl.wait();
        if (!load_count_table(fl, o, count_zeros)) {
            fl.each([&](int, rkg_ctx* c) { RK(c, rkg_count_reset(c)); });
            pass1 = stream_reads(fl, pool, o.reads, false,
                         [&](int, rkg_ctx* c, Chunk& ch) { RK(c, rkg_count_batch(c, ch.data(), ch.offs, (int64_t)ch.nreads, count_zeros)); },
                         nullptr, [](Chunk&) {});
            allreduce_counts(fl);
            save_count_table(fl, o, count_zeros);
            counted = true;
        }
    }
    const uint64_t pass2 = stream_reads(fl, pool, o.reads, keep_qual,
                 [&](int, rkg_ctx* c, Chunk& ch) {
                     RK(c, rkg_classify_batch(c, ch.data(), ch.offs, (int64_t)ch.nreads, o.do_read_depth ? o.min_occ : -1, ch.hits));
                 },
                 format, write_out);
    if (counted && pass1 != pass2) {                    // an input that changed (or drained) between the two passes
        fflush(stdout);
        fprintf(stderr, "rkmh: -M pass 1 saw %llu records, pass 2 %llu: the read files changed between the passes\n",
                (unsigned long long)pass1, (unsigned long long)pass2);
        exit(2);
    }
}

// result line of stream / classify: src/rkmh.cpp:884-893
void format_stream_line(std::string& out, const Options& o, const SeqSet& refs, const char* read_name, size_t read_name_len,
                        const rkg_hit& h) {
    const int diff = h.best - h.prev;
    const bool depth_fail = h.sketch_len <= o.min_matches, match_fail = h.best < o.min_matches, diff_ok = diff > o.min_diff;
    out += refs.names[h.ref_id]; out += '\t'; out.append(read_name, read_name_len); out += '\t';
    append_int(out, h.best); out += '\t'; append_int(out, o.sketch);
    if (depth_fail) out += "FAIL:DEPTH";
    out += '\t'; if (match_fail) out += "FAIL:MATCHES";
    out += '\t'; if (!diff_ok) out += "FAIL:DIFF";
    out += '\n';
}

// record of filter: src/rkmh.cpp:1289-1309, src/equiv.hpp:328-352
void format_filter_record(std::string& out, const Options& o, const char* name, size_t name_len, const char* s, int64_t L,
                          const char* qual, size_t qual_len, const rkg_hit& h) {
    const int best = std::max(h.best, 0), prev = std::max(h.prev, 0);
    if (!(h.sketch_len > 0) || best < o.min_matches || !(best - prev > o.min_diff)) return;
    out += '>'; out.append(name, name_len); out += '\n';
    for (int64_t j = 0; j < L; ++j) {
        signed char ch = (signed char)s[j];
        out += (char)((int)ch - 91 > 0 ? ch - 32 : ch);          // to_upper, mkmh.hpp:231-236
    }
    out += "\n+\n"; out.append(qual, qual_len); out += '\n';
}

// The end of a command: everything is written, so the process leaves without tearing down the CUDA contexts, the pinned
// chunk pool and the device tables one by one (measured on the B200 box: more than a second for an 8 M-read run -- the
// operating system reclaims all of it at once).  RKMH_CLEAN_EXIT=1 returns through the destructors instead (sanitizers).
//
// Even so the kernel needs ~0.6 s after _exit to unpin the pool and release the GPU contexts, and whoever waits for the
// process waits for that too.  So the GPU commands run in a WORKER process (forked before anything touches CUDA); the
// process the caller started only waits for the worker's one-byte "everything is written" message and leaves at once,
// while the worker's teardown finishes on its own.  A worker that ends any other way (error exit, signal) is waited for and
// its status passed on.  RKMH_NO_DETACH=1 runs everything in one process.
int g_done_fd = -1;

int leave(int rc) {
    tmark("done");
    if (getenv("RKMH_CLEAN_EXIT")) return rc;
    fflush(stdout); fflush(stderr);
    if (g_done_fd >= 0) {
        prctl(PR_SET_PDEATHSIG, 0);                  // the front process leaves first now: that is not a reason to die
        close(1); close(2);                          // readers of our output see end-of-file now, not after the teardown
        const unsigned char b = (unsigned char)rc;
        if (write(g_done_fd, &b, 1) != 1) {}
    }
    _exit(rc);
}

// Returns in the worker (and in the only process when detaching is off); the front process never returns.
void detach_teardown() {
    if (getenv("RKMH_NO_DETACH") || getenv("RKMH_CLEAN_EXIT")) return;
    int fds[2];
    if (pipe(fds) != 0) return;
    fflush(stdout); fflush(stderr);
    const pid_t pid = fork();
    if (pid < 0) { close(fds[0]); close(fds[1]); return; }
    if (pid == 0) {
        close(fds[0]);
        g_done_fd = fds[1];
        prctl(PR_SET_PDEATHSIG, SIGTERM);            // a front process killed half-way takes the worker with it
        return;
    }
    close(fds[1]);
    unsigned char b = 0;
    ssize_t n;
    do { n = read(fds[0], &b, 1); } while (n < 0 && errno == EINTR);
    if (n == 1) _exit((int)b);
    int st = 0;
    while (waitpid(pid, &st, 0) < 0 && errno == EINTR) {}
    if (WIFSIGNALED(st)) {                           // e.g. the abort() of an unknown option: end the same way
        signal(WTERMSIG(st), SIG_DFL);
        raise(WTERMSIG(st));
    }
    _exit(WIFEXITED(st) ? WEXITSTATUS(st) : 1);
}

// ---- stream / classify: src/rkmh.cpp:584-989 ------------------------------------------------------------
int main_stream(int argc, char** argv) {
    if (argc <= 2) { usage(argv[0]); exit(1); }
    Options o;
    const char* env = getenv("RKMH_GPUS");
    if (env) o.gpus = atoi(env);
    parse_common(argc, argv, o, "zmhdk:f:r:s:S:t:M:N:I:R:F:p:q:iD:g:");
    SeqSet refs;
    PreRefs pre;
    load_all(o.refs, refs, false);
    load_pre_refs(o, refs, pre);                         // after the sequences: reference indices follow this order
    if (refs.size() == 0) { fprintf(stderr, "rkmh: no reference sequences (the reference binary is undefined here)\n"); return 2; }
    Fleet fl;
    fl.open_async(std::max(1, o.gpus), [&] { setup_refs(fl, o, refs, STREAM_TABLE, false, &pre); });
    if (!(o.do_read_depth && any_unrewindable(o.reads))) {
        run_streamed(fl, o, false, false, [&](Chunk& c) {
            c.out.reserve(c.nreads * 64);
            for (size_t i = 0; i < c.nreads; ++i) { const auto nm = c.name(i); format_stream_line(c.out, o, refs, nm.first, nm.second, c.hits[i]); }
        });
        return leave(0);
    }
    SeqSet reads;                                        // -M on standard input: two passes need the reads twice
    load_all(o.reads, reads, false);
    std::vector<rkg_hit> hits = classify_all(fl, o, reads, false);
    std::string out;
    out.reserve(1 << 22);
    for (size_t i = 0; i < reads.size(); ++i) {
        format_stream_line(out, o, refs, reads.names[i].data(), reads.names[i].size(), hits[i]);
        if (out.size() > (1 << 22) - 4096) { fwrite(out.data(), 1, out.size(), stdout); out.clear(); }
    }
    fwrite(out.data(), 1, out.size(), stdout);
    return leave(0);
}

// ---- filter: src/rkmh.cpp:996-1424 (intended flow) ----------------------------------------------
int main_filter(int argc, char** argv) {
    if (argc <= 2) { usage(argv[0]); exit(1); }
    Options o;
    const char* env = getenv("RKMH_GPUS");
    if (env) o.gpus = atoi(env);
    parse_common(argc, argv, o, "hdk:f:r:s:S:t:M:N:I:R:F:p:q:iD:g:");
    SeqSet refs;
    PreRefs pre;
    load_all(o.refs, refs, false);
    load_pre_refs(o, refs, pre);
    if (refs.size() == 0) { fprintf(stderr, "rkmh: no reference sequences\n"); return 2; }
    Fleet fl;
    fl.open_async(std::max(1, o.gpus), [&] { setup_refs(fl, o, refs, FILTER_TABLE, true, &pre); });
    if (!(o.do_read_depth && any_unrewindable(o.reads))) {
        if (!o.reads.empty() || !o.in_stream)
            run_streamed(fl, o, true, true, [&](Chunk& c) {
                for (size_t i = 0; i < c.nreads; ++i) {
                    const auto nm = c.name(i), ql = c.qual(i);
                    format_filter_record(c.out, o, nm.first, nm.second, c.data() + c.offs[i], c.offs[i + 1] - c.offs[i], ql.first, ql.second, c.hits[i]);
                }
            });
    } else {
        SeqSet reads;
        load_all(o.reads, reads, true);
        std::vector<rkg_hit> hits = classify_all(fl, o, reads, true);
        std::string out;
        for (size_t i = 0; i < reads.size(); ++i) {
            format_filter_record(out, o, reads.names[i].data(), reads.names[i].size(), reads.bases.data() + reads.offs[i],
                                 reads.offs[i + 1] - reads.offs[i], reads.quals[i].data(), reads.quals[i].size(), hits[i]);
            if (out.size() > (1 << 22)) { fwrite(out.data(), 1, out.size(), stdout); out.clear(); }
        }
        fwrite(out.data(), 1, out.size(), stdout);
    }
    if (!o.in_stream) return leave(0);
    // -i / --in-stream (src/rkmh.cpp:1329-1414): after the read files, every record of STANDARD INPUT is sketched and classified
    // against the references and REPORTED (not filtered), one line each:
    //   Sample: <name> TAB Result: <reference or empty> TAB <shared> TAB <min(sketch length, reference sketch length)> TAB
    //   [FAIL:DEPTH] TAB [FAIL:MATCHES] TAB [FAIL:DIFF]
    // The reference's own code path cannot run (its OpenMP tasks read the kseq buffers after kseq_destroy; ASan: heap
    // use after free at src/rkmh.cpp:1353), so this follows its text: -M m > 0 keeps the hashes that occur at least m times
    // in the table counted from the read FILES (:1366-1376), best / previous best start from 0 (src/equiv.hpp:328-345).
    std::vector<int32_t> ref_lens(refs.size());
    fl.wait();
    {
        std::vector<uint64_t> sk((size_t)refs.size() * (size_t)o.sketch);
        RK(fl.ctx[0], rkg_get_ref_sketches(fl.ctx[0], sk.data(), ref_lens.data()));
    }
    const int occ = (o.do_read_depth && o.min_occ > 0) ? o.min_occ : -1;
    ChunkPool pool;
    pool.open(&fl, 2 * (int)fl.ctx.size() + 6, chunk_bytes());
    const std::vector<const char*> stdin_only{"-"};
    stream_reads(fl, pool, stdin_only, false,
                 [&](int, rkg_ctx* c, Chunk& ch) { RK(c, rkg_classify_batch(c, ch.data(), ch.offs, (int64_t)ch.nreads, occ, ch.hits)); },
                 [&](Chunk& c) {
                     for (size_t i = 0; i < c.nreads; ++i) {
                         const auto nm = c.name(i);
                         const rkg_hit& h = c.hits[i];
                         const int best = std::max(h.best, 0), prev = std::max(h.prev, 0);
                         const bool found = best > 0;
                         c.out += "Sample: "; c.out.append(nm.first, nm.second); c.out += "\tResult: ";
                         if (found) c.out += refs.names[h.ref_id];
                         c.out += '\t'; append_int(c.out, best);
                         c.out += '\t'; append_int(c.out, found ? std::min<int>(h.sketch_len, ref_lens[h.ref_id]) : 0);
                         c.out += '\t'; if (h.sketch_len <= 0) c.out += "FAIL:DEPTH";
                         c.out += '\t'; if (best < o.min_matches) c.out += "FAIL:MATCHES";
                         c.out += '\t'; if (!(best - prev > o.min_diff)) c.out += "FAIL:DIFF";
                         c.out += '\n';
                     }
                 },
                 [](Chunk& c) { fwrite(c.out.data(), 1, c.out.size(), stdout); });
    return leave(0);
}

// ---- hash: src/rkmh.cpp:1931-2116; with -s the intended behaviour (README.md:114-117), without it HEAD's ---------
int main_hash(int argc, char** argv) {
    if (argc <= 2) { usage(argv[0]); exit(1); }
    Options o;
    o.sketch = 0;
    const char* env = getenv("RKMH_GPUS");
    if (env) o.gpus = atoi(env);
    parse_common(argc, argv, o, "hdk:f:r:s:S:t:M:N:I:R:F:p:q:iD:g:mwcKTo:");
    fprintf(stderr, "Using a kmer size of %d\n", o.ks[0]);
    SeqSet seqs;
    load_all(o.refs, seqs, false);                        // -r is accepted here (HEAD aborts on it)
    load_all(o.reads, seqs, false);
    Fleet fl; fl.open(std::max(1, o.gpus));
    rkg_ctx* c = fl.ctx[0];
    const bool sketch_mode = o.sketch > 0;
    fl.each([&](int, rkg_ctx* cc) { RK(cc, rkg_set_params(cc, o.ks.data(), (int)o.ks.size(), sketch_mode ? o.sketch : 1, 0, 0)); });
    std::string out;
    const size_t n = seqs.size();
    if (sketch_mode) {
        // Records of at least RKMH_SHARD_MIN bases (default 16 Mi; single k only) are sketched as ONE device-resident
        // record by rkg_sketch_shard, their windows split over the GPUs of the fleet and the partial bottom-s sketches
        // merged; runs of shorter records go through rkg_sketch_batch on the first GPU.
        const char* sm = getenv("RKMH_SHARD_MIN");
        const int64_t shard_min = sm ? std::max<long long>(atoll(sm), 1) : (int64_t)1 << 24;
        const int G = (int)fl.ctx.size(), k0 = o.ks[0];
        std::vector<uint64_t> sk(n * (size_t)o.sketch);
        std::vector<int32_t> lens(n);
        auto is_big = [&](size_t i) { return o.ks.size() == 1 && seqs.offs[i + 1] - seqs.offs[i] >= shard_min; };
        size_t i = 0;
        while (i < n) {
            if (!is_big(i)) {
                size_t j = i;
                while (j < n && !is_big(j)) ++j;
                RK(c, rkg_sketch_batch(c, seqs.bases.data(), seqs.offs.data() + i, (int64_t)(j - i), -1, sk.data() + i * o.sketch, lens.data() + i));
                i = j;
                continue;
            }
            const int64_t L = seqs.offs[i + 1] - seqs.offs[i], nk = std::max<int64_t>(L - k0, 0);
            const char* rec = seqs.bases.data() + seqs.offs[i];
            std::vector<uint64_t> parts((size_t)G * o.sketch, 0);
            std::vector<int32_t> plens(G, 0);
            fl.each([&](int d, rkg_ctx* cc) {
                const int64_t b = nk * d / G, e = nk * (d + 1) / G;
                if (e > b) RK(cc, rkg_sketch_shard(cc, rec + b, std::min<int64_t>(e + k0, L) - b, b, e, L, parts.data() + (size_t)d * o.sketch, &plens[d]));
            });
            if (G == 1) { lens[i] = plens[0]; std::copy(parts.begin(), parts.begin() + plens[0], sk.begin() + i * o.sketch); }
            else RK(c, rkg_merge_sketches(c, parts.data(), plens.data(), G, sk.data() + i * o.sketch, &lens[i]));
            ++i;
        }
        if (o.outname) {                                  // -o FILE: the sketches as JSON (dump_hash_json's schema, sketch_io.hpp)
            rkhost::SketchFile jf;
            jf.ks = o.ks; jf.sketch_size = o.sketch;
            for (size_t i2 = 0; i2 < n; ++i2) {
                rkhost::SketchRecord r;
                r.name = seqs.names[i2]; r.seq_len = seqs.offs[i2 + 1] - seqs.offs[i2];
                r.hashes.assign(sk.begin() + i2 * (size_t)o.sketch, sk.begin() + i2 * (size_t)o.sketch + lens[i2]);
                jf.recs.push_back(std::move(r));
            }
            if (!rkhost::write_sketch_json(o.outname, jf)) { fprintf(stderr, "rkmh: cannot write %s\n", o.outname); return 2; }
            return leave(0);
        }
        for (size_t i2 = 0; i2 < n; ++i2) {               // print_hashes format, mkmh.hpp:1438-1449
            if (o.wabbit) {                               // -w: print_wabbit, src/rkmh.cpp:463-487
                rkhost::wabbit_line(out, seqs.names[i2], sk.data() + i2 * (size_t)o.sketch, lens[i2], o.ks[0], o.sketch);
                if (out.size() > (1 << 22)) { fwrite(out.data(), 1, out.size(), stdout); out.clear(); }
                continue;
            }
            out += seqs.names[i2]; out += '\t';
            for (int j = 0; j < lens[i2]; ++j) { if (j) out += '\t'; char b[24]; out.append(b, snprintf(b, sizeof b, "%llu", (unsigned long long)sk[i2 * o.sketch + j])); }
            out += '\n';
            if (out.size() > (1 << 22)) { fwrite(out.data(), 1, out.size(), stdout); out.clear(); }
        }
    } else {                                              // HEAD-literal: every hash, unsorted, zeros included
        size_t total = 0;
        for (size_t i = 0; i < n; ++i)
            for (int k : o.ks) total += (size_t)std::max<int64_t>(seqs.offs[i + 1] - seqs.offs[i] - k, 0);
        std::vector<uint64_t> h(std::max<size_t>(total, 1));
        std::vector<int64_t> ho(n + 1);
        if (n) RK(c, rkg_hash_batch(c, seqs.bases.data(), seqs.offs.data(), (int64_t)n, h.data(), ho.data()));
        for (size_t i = 0; i < n; ++i) {
            out += seqs.names[i]; out += '\t';
            for (int64_t j = ho[i]; j < ho[i + 1]; ++j) { if (j > ho[i]) out += '\t'; char b[24]; out.append(b, snprintf(b, sizeof b, "%llu", (unsigned long long)h[j])); }
            out += '\n';
            if (out.size() > (1 << 22)) { fwrite(out.data(), 1, out.size(), stdout); out.clear(); }
        }
    }
    fwrite(out.data(), 1, out.size(), stdout);
    return leave(0);
}

// ---- hpv16: src/rkmh.cpp:2366-2723 ----------------------------------------------------------------------
// The tiered HPV16 classifier: every hash of a read (no sketch) against the FULL hash sets of the 182 HPV type references
// (hash_set_intersection_size, mkmh/mkmh.hpp:842-863) and against the lineage / sublineage specific k-mer tables built
// from <refpath>/new_refs.fa (sort_by_similarity, mkmh/mkmh.hpp:1304-1333).  Hashing, the set table and the per-read
// counts run on the GPU (rkg_hash_batch, rkg_load_ref_sets, rkg_classify_sets); the lineage tables are small host-side
// set differences, done with the same containers as the reference so that their order (and the file the reference
// writes next to its output, lineage_specific_hashes.<k>.tst) comes out the same.
std::string find_ref_file(const std::string& dir, const char* name) {
    const std::string plain = dir + "/" + name;
    struct stat st;
    if (stat(plain.c_str(), &st) == 0) return plain;
    const std::string gz = plain + ".gz";                   // the fixtures of this repository are stored compressed
    if (stat(gz.c_str(), &st) == 0) return gz;
    return plain;
}

template <typename Key>
void specific_tables(const std::map<Key, std::set<uint64_t>>& groups, std::vector<std::string>& names,
                     std::vector<std::vector<uint64_t>>& tables, const std::function<std::string(const Key&)>& name_of) {
    for (const auto& x : groups) {                          // src/rkmh.cpp:2568-2596 / 2620-2644
        std::vector<uint64_t> diff(x.second.size() + 1000);
        std::vector<uint64_t> xdiff(x.second.begin(), x.second.end());
        for (const auto& y : groups) {
            if (x.first == y.first) continue;
            auto it = std::set_difference(xdiff.begin(), xdiff.end(), y.second.begin(), y.second.end(), diff.begin());
            diff.resize((size_t)(it - diff.begin()));
            xdiff = diff;
        }
        names.push_back(name_of(x.first));
        std::sort(xdiff.begin(), xdiff.end());
        tables.push_back(xdiff);
    }
}

void append_double(std::string& s, double v) { char b[40]; int n = snprintf(b, sizeof b, "%g", v); s.append(b, n); }   // ostream << double

int main_hpv16(int argc, char** argv) {
    if (argc <= 2) { usage(argv[0]); exit(1); }
    std::vector<const char*> read_files;
    std::string refpath = "data";
    std::vector<int> ks;
    int min_kmer_occ = 0, gpus = 1;
    bool do_read_depth = false;
    if (const char* env = getenv("RKMH_GPUS")) gpus = atoi(env);
    static struct option long_options[] = {
        {"help", no_argument, 0, 'h'}, {"kmer", no_argument, 0, 'k'}, {"fasta", required_argument, 0, 'f'},
        {"reference", required_argument, 0, 'r'}, {"sketch", required_argument, 0, 's'}, {"threads", required_argument, 0, 't'},
        {"min-kmer-occurence", required_argument, 0, 'M'}, {"min-matches", required_argument, 0, 'N'},
        {"min-diff", required_argument, 0, 'D'}, {"max-samples", required_argument, 0, 'I'}, {"gpus", required_argument, 0, 'g'},
        {0, 0, 0, 0}};
    optind = 2;
    for (;;) {
        int idx = 0;
        const int c = getopt_long(argc, argv, "hk:f:R:s:t:M:N:D:g:", long_options, &idx);      // src/rkmh.cpp:2408
        if (c == -1) break;
        switch (c) {
            case 't': case 's': case 'N': case 'D': break;          // parsed by the reference, without effect on the output
            case 'f': read_files.push_back(optarg); break;
            case 'R': refpath = optarg; break;
            case 'k': if (optarg) ks.push_back(atoi(optarg)); break;
            case 'M': min_kmer_occ = atoi(optarg); do_read_depth = true; break;
            case 'g': gpus = atoi(optarg); break;
            case '?': case 'h': usage(argv[0]); exit(1);
            default: usage(argv[0]); abort();
        }
    }
    if (ks.empty()) {
        fprintf(stderr, "NO KMER SIZE PROVIDED. USING A DEFAULT KMER SIZE OF 16\n");
        ks.push_back(16);
    }
    const std::string type_file = find_ref_file(refpath, "all_pave_ref.fa"), sub_file = find_ref_file(refpath, "new_refs.fa");
    SeqSet types, subs;
    load_all({type_file.c_str()}, types, false);
    load_all({sub_file.c_str()}, subs, false);
    if (types.size() == 0) { fprintf(stderr, "rkmh: no type references in %s\n", type_file.c_str()); return 2; }
    Fleet fl; fl.open(std::max(1, gpus));
    rkg_ctx* c0 = fl.ctx[0];

    // every hash of every reference with the FIRST k-mer size (src/rkmh.cpp:2546-2552)
    const int k0 = ks[0];
    RK(c0, rkg_set_params(c0, &k0, 1, 1, 0, 0));
    auto hash_all = [&](const SeqSet& set, std::vector<uint64_t>& h, std::vector<int64_t>& ho) {
        size_t total = 0;
        for (size_t i = 0; i < set.size(); ++i) total += (size_t)std::max<int64_t>(set.offs[i + 1] - set.offs[i] - k0, 0);
        h.assign(std::max<size_t>(total, 1), 0); ho.assign(set.size() + 1, 0);
        if (set.size()) RK(c0, rkg_hash_batch(c0, set.bases.data(), set.offs.data(), (int64_t)set.size(), h.data(), ho.data()));
    };
    std::vector<uint64_t> type_h, sub_h;
    std::vector<int64_t> type_ho, sub_ho;
    hash_all(types, type_h, type_ho);
    hash_all(subs, sub_h, sub_ho);

    // lineage = first letter of the name, sublineage = first two (src/rkmh.cpp:2560-2566, 2612-2618)
    std::map<char, std::set<uint64_t>> lin_to_hashes;
    std::map<std::string, std::set<uint64_t>> sublin_to_hashes;
    for (size_t i = 0; i < subs.size(); ++i) {
        const std::string& key = subs.names[i];
        const char lin = key.empty() ? '\0' : key[0];
        const std::string sublin = key.substr(0, 2);
        for (int64_t j = sub_ho[i]; j < sub_ho[i + 1]; ++j) { lin_to_hashes[lin].insert(sub_h[j]); sublin_to_hashes[sublin].insert(sub_h[j]); }
    }
    std::vector<std::string> lin_names, sublin_names;
    std::vector<std::vector<uint64_t>> lin_tables, sublin_tables;
    specific_tables<char>(lin_to_hashes, lin_names, lin_tables, [](const char& ch) { return std::string(1, ch); });
    specific_tables<std::string>(sublin_to_hashes, sublin_names, sublin_tables, [](const std::string& x) { return x; });
    {   // the side file and the messages of the reference (src/rkmh.cpp:2598-2610, 2646-2649)
        const std::string fn = "lineage_specific_hashes." + std::to_string(k0) + ".tst";
        if (FILE* fo = fopen(fn.c_str(), "w")) {
            for (size_t i = 0; i < lin_names.size(); ++i) {
                fprintf(fo, "%s\t", lin_names[i].c_str());
                for (uint64_t x : lin_tables[i]) fprintf(fo, "%llu\t", (unsigned long long)x);
                fputc('\n', fo);
            }
            fclose(fo);
        }
        fprintf(stderr, "Lineage specific kmer table created:\n");
        for (size_t i = 0; i < lin_names.size(); ++i) fprintf(stderr, "\t%s\t%zu\n", lin_names[i].c_str(), lin_tables[i].size());
        fprintf(stderr, "Sublineage specific kmer table created:\n");
        for (size_t i = 0; i < sublin_names.size(); ++i) fprintf(stderr, "\t%s\t%zu\n", sublin_names[i].c_str(), sublin_tables[i].size());
    }

    // one table: the type references' full hash arrays, then the lineage and the sublineage tables
    const int ntypes = (int)types.size(), nlin = (int)lin_names.size(), nsub = (int)sublin_names.size();
    const int nextra = nlin + nsub;
    std::vector<uint64_t> set_h(type_h.begin(), type_h.begin() + type_ho[ntypes]);
    std::vector<int64_t> set_o(type_ho.begin(), type_ho.end());
    for (const auto* tabs : {&lin_tables, &sublin_tables})
        for (const auto& t : *tabs) { set_h.insert(set_h.end(), t.begin(), t.end()); set_o.push_back((int64_t)set_h.size()); }
    if (set_h.empty()) set_h.push_back(0);
    fl.each([&](int, rkg_ctx* c) {
        RK(c, rkg_set_params(c, ks.data(), (int)ks.size(), 1, do_read_depth ? 800000000ULL : 0, 0));       // src/rkmh.cpp:2516
        RK(c, rkg_load_ref_sets(c, set_h.data(), set_o.data(), ntypes + nextra));
    });

    auto format = [&](Chunk& c) {                           // src/rkmh.cpp:2665-2704
        std::vector<std::pair<int, int>> helper;
        for (size_t i = 0; i < c.nreads; ++i) {
            const auto nm = c.name(i);
            const int64_t len = c.offs[i + 1] - c.offs[i];
            int64_t hashnum = 0;
            for (int k : ks) hashnum += std::max<int64_t>(len - k, 0);
            const rkg_hit& h = c.hits[i];
            const int32_t* ex = c.extra.data() + i * (size_t)nextra;
            c.out.append(nm.first, nm.second); c.out += '\t';
            c.out += types.names[h.ref_id]; c.out += '\t';
            append_int(c.out, h.best); c.out += '/'; append_int(c.out, hashnum); c.out += '\t';
            std::string sims[2], inters[2];
            for (int tier = 0; tier < 2; ++tier) {          // sort_by_similarity, mkmh/mkmh.hpp:1304-1333
                const int n = tier ? nsub : nlin;
                const int32_t* cnt = ex + (tier ? nlin : 0);
                const std::vector<std::string>& names = tier ? sublin_names : lin_names;
                helper.resize((size_t)n);
                for (int j = 0; j < n; ++j) helper[j] = std::make_pair(j, (int)cnt[j]);
                std::sort(helper.begin(), helper.end(),
                          [](const std::pair<int, int>& lhs, const std::pair<int, int>& rhs) { return lhs.second > rhs.second; });
                for (int j = 0; j < n; ++j) {
                    sims[tier] += names[helper[j].first]; sims[tier] += ':';
                    append_double(sims[tier], (double)helper[j].second / (double)hashnum); sims[tier] += ';';
                    append_int(inters[tier], helper[j].second); inters[tier] += ';';
                }
            }
            c.out += sims[0]; c.out += '\t'; c.out += sims[1]; c.out += '\t'; c.out += inters[0]; c.out += '\t'; c.out += inters[1];
            c.out += '\n';
        }
    };
    auto classify = [&](rkg_ctx* c, Chunk& ch) {
        ch.extra.resize(std::max<size_t>(ch.nreads * (size_t)std::max(nextra, 1), 1));
        RK(c, rkg_classify_sets(c, ch.data(), ch.offs, (int64_t)ch.nreads, do_read_depth ? min_kmer_occ : -1, ntypes, ch.hits, ch.extra.data()));
    };
    const auto write_out = [](Chunk& c) { fwrite(c.out.data(), 1, c.out.size(), stdout); };
    ChunkPool pool;
    pool.open(&fl, 2 * (int)fl.ctx.size() + 6 + 3 * parse_threads(), chunk_bytes());
    if (!(do_read_depth && any_unrewindable(read_files))) {
        if (do_read_depth) {                                // src/rkmh.cpp:2514-2530: valid windows only
            fl.each([&](int, rkg_ctx* c) { RK(c, rkg_count_reset(c)); });
            stream_reads(fl, pool, read_files, false,
                         [&](int, rkg_ctx* c, Chunk& ch) { RK(c, rkg_count_batch(c, ch.data(), ch.offs, (int64_t)ch.nreads, 0)); },
                         nullptr, [](Chunk&) {});
            allreduce_counts(fl);
        }
        stream_reads(fl, pool, read_files, false, [&](int, rkg_ctx* c, Chunk& ch) { classify(c, ch); }, format, write_out);
        return leave(0);
    }
    // -M on an input that cannot be read twice: buffer the reads, then the same two passes through one big chunk
    SeqSet reads;
    load_all(read_files, reads, false);
    RK(c0, rkg_count_reset(c0));
    if (reads.size()) RK(c0, rkg_count_batch(c0, reads.bases.data(), reads.offs.data(), (int64_t)reads.size(), 0));
    Chunk all;
    std::vector<int64_t> offs(reads.offs);
    std::vector<rkg_hit> hits(std::max<size_t>(reads.size(), 1));
    all.bases = reads.bases.data(); all.offs = offs.data(); all.hits = hits.data(); all.nreads = reads.size();
    for (size_t i = 0; i < reads.size(); ++i) { all.name_at.push_back(all.text.size()); all.text.insert(all.text.end(), reads.names[i].begin(), reads.names[i].end()); }
    if (reads.size()) classify(c0, all);
    format(all);
    write_out(all);
    all.bases = nullptr; all.offs = nullptr; all.hits = nullptr;
    return leave(0);
}

// ---- call: src/rkmh.cpp:1455-1904 -----------------------------------------------------------------------
// Alignment-free variant calling: the exact k-mer depth of the read set is kept on the GPU (rkg_depth_*); for every window
// of the reference whose depth is under half of the running average over the last `window_len` windows, every single-base
// substitution and every one-base deletion is looked up, and the ones that "rescue" the depth become VCF lines.  The walk
// along the reference, the thresholds and the bookkeeping of the calls are the reference's, statement for statement.
int main_call(int argc, char** argv) {
    if (argc <= 2) { usage(argv[0]); exit(1); }
    std::vector<const char*> ref_files, read_files;
    std::vector<int> ks;
    int window_len = 100, gpus = 1;
    bool show_depth = false, output_vcf = true;
    static struct option long_options[] = {
        {"help", no_argument, 0, 'h'}, {"kmer", no_argument, 0, 'k'}, {"fasta", required_argument, 0, 'f'},
        {"reference", required_argument, 0, 'r'}, {"sketch", required_argument, 0, 's'}, {"threads", required_argument, 0, 't'},
        {"show-depth", required_argument, 0, 'd'}, {"window-len", required_argument, 0, 'w'}, {"gpus", required_argument, 0, 'g'},
        {0, 0, 0, 0}};
    optind = 2;
    for (;;) {
        int idx = 0;
        const int c = getopt_long(argc, argv, "hdk:f:r:s:t:w:g:", long_options, &idx);     // src/rkmh.cpp:1494
        if (c == -1) break;
        switch (c) {
            case 't': case 's': break;                        // parsed, without effect on the output
            case 'r': ref_files.push_back(optarg); break;
            case 'f': read_files.push_back(optarg); break;
            case 'k': if (optarg) ks.push_back(atoi(optarg)); break;
            case 'w': window_len = atoi(optarg); break;
            case 'd': show_depth = true; output_vcf = false; break;
            case 'g': gpus = atoi(optarg); break;
            case '?': case 'h': usage(argv[0]); exit(1);
            default: usage(argv[0]); abort();
        }
    }
    (void)gpus;
    if (ks.empty()) {
        fprintf(stderr, "No kmer size(s) provided. Will use a default kmer size of 16.\n");
        ks.push_back(16);
    } else if (ks.size() > 1) {
        fprintf(stderr, "Only a single kmer size may be used for calling.\nSizes provided: ");
        for (int k : ks) fprintf(stderr, "%d ", k);
        fprintf(stderr, "\nPlease choose a single kmer size.\n");
        exit(1);
    }
    fprintf(stderr, "Parsing sequences...\n");
    if (ref_files.empty()) {
        fprintf(stderr, "No references were provided. Please provide at least one reference file in fasta/fastq format.\n");
        usage(argv[0]); exit(1);
    }
    if (read_files.empty()) {
        fprintf(stderr, "No reads were provided. Please provide at least one read file in fasta/fastq format.\n");
        usage(argv[0]); exit(1);
    }
    SeqSet refs, reads;
    load_all(ref_files, refs, false);
    load_all(read_files, reads, false);
    const int k = ks[0];
    Fleet fl; fl.open(1);
    rkg_ctx* c = fl.ctx[0];
    RK(c, rkg_set_params(c, &k, 1, 1, 0, 0));
    uint64_t windows = 0;
    for (size_t i = 0; i < reads.size(); ++i) windows += (uint64_t)std::max<int64_t>(reads.offs[i + 1] - reads.offs[i] - k, 0);
    RK(c, rkg_depth_begin(c, windows));
    if (reads.size()) RK(c, rkg_depth_add(c, reads.bases.data(), reads.offs.data(), (int64_t)reads.size()));

    if (refs.size() > 1) fprintf(stderr, "WARNING: more than one ref provided. VCF will not be correct\n");
    std::string out;
    if (output_vcf) {                                     // src/rkmh.cpp:1727-1733 (the RD and OD lines share one line there)
        out += "##fileformat=VCF4.2\n##source=rkmh\n##reference="; out += ref_files[0]; out += '\n';
        out += "##INFO=<ID=KD,Number=1,Type=Integer,Description=\"Number of times call for specific kmer appears\">\n";
        out += "##INFO=<ID=MD,Number=1,Type=Integer,Description=\"Maximum depth found for the rescue kmer.\">\n";
        out += "##INFO=<ID=RD,Number=1,Type=Integer,Description=\"Average depth in region\">";
        out += "##INFO=<ID=OD,Number=1,Type=Integer,Description=\"Depth of original kmer at site before modification.\">\n";
    }
    std::map<std::string, int> call_count, call_max_depth, call_avg_depth, call_orig_depth;
    std::list<int> d_window;                              // not reset between references (the reference's is per thread)
    long long d_sum = 0;
    struct Low { int64_t j; int depth, avg_d; };
    for (size_t i = 0; i < refs.size(); ++i) {
        const char* seq = refs.bases.data() + refs.offs[i];
        const int64_t len = refs.offs[i + 1] - refs.offs[i], nh = std::max<int64_t>(len - k, 0);
        std::vector<int32_t> depth((size_t)std::max<int64_t>(nh, 1));
        if (nh) RK(c, rkg_depth_of_sequence(c, seq, len, depth.data()));
        std::vector<Low> low;
        std::vector<int64_t> low_pos;
        for (int64_t j = 0; j < nh; ++j) {                // src/rkmh.cpp:1772-1792
            d_window.push_back(depth[j]); d_sum += depth[j];
            if ((int)d_window.size() > window_len) { d_sum -= d_window.front(); d_window.pop_front(); }
            const int avg_d = (int)((double)(int)d_sum / (double)d_window.size());
            if (depth[j] < .5 * avg_d) { low.push_back({j, depth[j], avg_d}); low_pos.push_back(j); }
        }
        if (low.empty() || show_depth) continue;          // -d: the reference builds depth lines and never prints them
        std::vector<int32_t> alt((size_t)low.size() * 4 * (size_t)k);
        RK(c, rkg_depth_rescue(c, seq, len, low_pos.data(), (int64_t)low_pos.size(), alt.data()));
        auto upper = [](char ch) { return (char)(((int)(signed char)ch - 91 > 0) ? ch - 32 : ch); };
        static const char rot[4][3] = {{'C', 'T', 'G'}, {'T', 'G', 'A'}, {'C', 'G', 'A'}, {'A', 'C', 'T'}};   // A C T G
        auto record = [&](const std::string& key, int avg_d, int dep, int alt_depth) {
            call_count[key] += 1;
            call_avg_depth[key] = std::max(avg_d, call_avg_depth[key]);
            call_orig_depth[key] = std::max(call_orig_depth[key], dep);
            if (alt_depth > call_max_depth[key]) call_max_depth[key] = alt_depth;
        };
        for (size_t w = 0; w < low.size(); ++w) {
            const Low& lw = low[w];
            const int32_t* a = alt.data() + w * 4 * (size_t)k;
            for (int p = 0; p < k; ++p) {                 // substitutions, src/rkmh.cpp:1796-1831
                const char orig = upper(seq[lw.j + p]);
                const int row = orig == 'A' ? 0 : orig == 'C' ? 1 : orig == 'T' ? 2 : orig == 'G' ? 3 : -1;
                if (row < 0) continue;
                for (int x = 0; x < 3; ++x) {
                    const int alt_depth = a[3 * p + x];
                    if (alt_depth < 0) continue;
                    if ((alt_depth >= .1 * lw.avg_d) & (alt_depth > lw.depth)) {
                        std::string key = refs.names[i]; key += '\t'; append_int(key, lw.j + p + 1); key += "\t.\t"; key += orig; key += '\t'; key += rot[row][x];
                        record(key, lw.avg_d, lw.depth, alt_depth);
                    }
                }
            }
            if (lw.j > 0)                                 // deletions, src/rkmh.cpp:1840-1858
                for (int d = 1; d <= k; ++d) {
                    const int alt_depth = a[3 * k + d - 1];
                    if (alt_depth < 0) continue;
                    if (output_vcf && alt_depth > 0.9 * lw.avg_d) {
                        std::string key = refs.names[i]; key += '\t'; append_int(key, lw.j + d + 1); key += "\t.\t"; key += upper(seq[lw.j - 1 + d]); key += "\t-";
                        record(key, lw.avg_d, lw.depth, alt_depth);
                    }
                }
        }
    }
    if (output_vcf)
        for (const auto& x : call_count) {                // src/rkmh.cpp:1876-1879
            out += x.first; out += "\t99\tPASS\tKC="; append_int(out, x.second); out += ";MD="; append_int(out, call_max_depth[x.first]);
            out += ";RD="; append_int(out, call_avg_depth[x.first]); out += ";OD="; append_int(out, call_orig_depth[x.first]); out += '\n';
        }
    fwrite(out.data(), 1, out.size(), stdout);
    return leave(0);
}

}  // namespace

int main(int argc, char** argv) {
    if (argc <= 1) { usage(argv[0]); exit(1); }
    const std::string cmd = argv[1];
    if (cmd == "classify" || cmd == "stream" || cmd == "filter" || cmd == "hash" || cmd == "hpv16" || cmd == "call") detach_teardown();
    if (cmd == "classify") {
        fprintf(stderr, "CLASSIFY COMMAND IS TEMPORARILY UNAVAILABLE: TRY rkmh stream INSTEAD.\n");   // src/rkmh.cpp:2746
        return main_stream(argc, argv);
    }
    if (cmd == "stream") return main_stream(argc, argv);
    if (cmd == "filter") return main_filter(argc, argv);
    if (cmd == "hash") return main_hash(argc, argv);
    if (cmd == "hpv16") return main_hpv16(argc, argv);
    if (cmd == "call") return main_call(argc, argv);
    if (cmd == "parse-stream") {     // host-only: the records as the streaming front-end chunks them (parallel parser included)
        std::vector<const char*> files(argv + 2, argv + argc);
        ChunkPool pool;
        pool.open(nullptr, 6 + 3 * parse_threads(), chunk_bytes());
        parse_files(files, pool, true, [&](Chunk* c) {
            for (size_t r = 0; r < c->nreads; ++r) {
                const auto nm = c->name(r), ql = c->qual(r);
                fwrite(nm.first, 1, nm.second, stdout); fputc('\t', stdout);
                fwrite(c->data() + c->offs[r], 1, (size_t)(c->offs[r + 1] - c->offs[r]), stdout); fputc('\t', stdout);
                fwrite(ql.first, 1, ql.second, stdout); fputc('\n', stdout);
            }
            pool.release(c);
        });
        return 0;
    }
    if (cmd == "parse-count") {      // host-only: time the streaming front-end's parser alone (records and bases seen)
        std::vector<const char*> files(argv + 2, argv + argc);
        ChunkPool pool;
        pool.open(nullptr, 6 + 3 * parse_threads(), chunk_bytes());
        uint64_t nb = 0, nchunks = 0;
        const uint64_t nrec = parse_files(files, pool, true, [&](Chunk* c) { nb += (uint64_t)c->offs[c->nreads]; ++nchunks; pool.release(c); });
        printf("records %llu bases %llu chunks %llu\n", (unsigned long long)nrec, (unsigned long long)nb, (unsigned long long)nchunks);
        return 0;
    }
    if (cmd == "parse") {            // host-only debugging aid: dump the records the parser sees (no GPU needed)
        for (int i = 2; i < argc; ++i) {
            SeqSet set;
            if (!rkhost::read_fastx(argv[i], set, true)) { fprintf(stderr, "rkmh: cannot open %s\n", argv[i]); return 2; }
            for (size_t r = 0; r < set.size(); ++r) {
                fwrite(set.names[r].data(), 1, set.names[r].size(), stdout); fputc('\t', stdout);
                fwrite(set.bases.data() + set.offs[r], 1, (size_t)(set.offs[r + 1] - set.offs[r]), stdout); fputc('\t', stdout);
                fwrite(set.quals[r].data(), 1, set.quals[r].size(), stdout); fputc('\n', stdout);
            }
        }
        return 0;
    }
    usage(argv[0]);
    return 1;
}
