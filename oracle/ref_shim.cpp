// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
//
// A thin extern "C" face over the REAL mkmh library, compiled from /root/reference/mkmh (see
// oracle/Makefile: _ref/libmkmh_ref.so).  It lets the Python tests call the reference's own
// calc_hashes / minhashes / hash_intersection_size / HASHTCounter and compare them with the C
// restatement (oracle/rkmh_oracle.c) and, through committed golden vectors, with the CUDA product.
// Nothing here restates an algorithm: every function forwards to the reference symbol it names.
#include <cstdint>
#include <cstring>
#include "mkmh.hpp"          // /root/reference/mkmh/mkmh.hpp (via -I)
#include "HASHTCounter.hpp"

using mkmh::hash_t;

extern "C" {

// MurmurHash3_x64_128 -- mkmh/murmur3/murmur3.cpp:234
void ref_murmur3_x64_128(const void* key, int len, uint32_t seed, void* out) {
    MurmurHash3_x64_128(key, len, seed, out);
}

// mkmh::to_upper -- mkmh/mkmh.hpp:231
void ref_to_upper(char* seq, int64_t len) { mkmh::to_upper(seq, (size_t)len); }

// mkmh::calc_hashes(const char*, const int&, const int&, hash_t*&, int&) -- mkmh/mkmh.hpp:505
int64_t ref_calc_hashes(const char* seq, int64_t len, int k, uint64_t* out) {
    hash_t* h = nullptr; int n = 0;
    mkmh::calc_hashes(seq, (int)len, k, h, n);
    std::memcpy(out, h, sizeof(uint64_t) * (size_t)n);
    delete[] h;
    return n;
}

// mkmh::calc_hashes(const char*, int, vector<int>, hash_t*&, int&) -- mkmh/mkmh.hpp:689
int64_t ref_calc_hashes_multi(const char* seq, int64_t len, const int* ks, int nk, uint64_t* out) {
    std::vector<int> kv(ks, ks + nk);
    hash_t* h = nullptr; int n = 0;
    mkmh::calc_hashes(seq, (int)len, kv, h, n);
    std::memcpy(out, h, sizeof(uint64_t) * (size_t)n);
    delete[] h;
    return n;
}

// mkmh::minhashes(hash_t*&, int, int, hash_t*&, int&, bool) -- mkmh/mkmh.hpp:931 (sorts `hashes`)
int ref_minhashes(uint64_t* hashes, int64_t n, int s, uint64_t* ret) {
    hash_t* in = reinterpret_cast<hash_t*>(hashes);
    hash_t* mins = nullptr; int len = 0;
    mkmh::minhashes(in, (int)n, s, mins, len);
    std::memcpy(ret, mins, sizeof(uint64_t) * (size_t)len);
    delete[] mins;
    return len;
}

// mkmh::hash_intersection_size -- mkmh/mkmh.hpp:822
int ref_intersection_size(const uint64_t* a, int na, const uint64_t* b, int nb) {
    int ret = 0;
    mkmh::hash_intersection_size(reinterpret_cast<const hash_t*>(a), na,
                                 reinterpret_cast<const hash_t*>(b), nb, ret);
    return ret;
}

// mkmh::HASHTCounter -- mkmh/HASHTCounter.cpp:12-15,46-58.  The reference leaves the array
// uninitialised (new int[]); the shim zeroes it, which is what fresh pages give the reference.
void* ref_counter_new(uint64_t size) {
    mkmh::HASHTCounter* c = new mkmh::HASHTCounter(size);
    for (uint64_t i = 0; i < size; ++i) { int& v = c->get((hash_t)i); v = 0; }
    return c;
}
void ref_counter_free(void* c) { delete static_cast<mkmh::HASHTCounter*>(c); }
void ref_counter_increment(void* c, uint64_t key) { static_cast<mkmh::HASHTCounter*>(c)->increment(key); }
int ref_counter_get(void* c, uint64_t key) { return static_cast<mkmh::HASHTCounter*>(c)->get(key); }

// mkmh::mask_by_frequency -- mkmh/mkmh.hpp:898
void ref_mask_by_frequency(uint64_t* h, int64_t n, void* c, int min_occ) {
    hash_t* p = reinterpret_cast<hash_t*>(h);
    mkmh::mask_by_frequency(p, (int)n, static_cast<mkmh::HASHTCounter*>(c), min_occ);
}

}  // extern "C"
