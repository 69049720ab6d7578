// TEST INFRASTRUCTURE ONLY: host build of agc_b200/csrc/lz_chunk_core.cuh (the chunk parser and the stitcher of the chunk-parallel
// LZ-diff encoder) so the CPU test-suite can compare it with the oracle without a GPU.  The product (libagcgpu.so) only contains
// the device build.
#include "../../agc_b200/csrc/lz_chunk_core.cuh"
#include <cstring>
#include <vector>
extern "C" {
// text / ref: 1 byte per symbol (0..3).  ht: the reference's hash table widened to u32 (0xFFFFFFFF = empty).  When is_rc the packed
// store holds the REVERSE COMPLEMENT of text and the view reads it back in place, as the device does for rc segments.
// lead = bases of other data placed before the segment in the store (tests unaligned starts).
// returns the delta length, -1 = "sequential kernel decides", -2 = overflow; *n_fallback_flags gets the chunk count
__attribute__((visibility("default"))) long lzc_host_encode(const unsigned char* text, unsigned n, const unsigned char* ref, unsigned m,
                                                            const unsigned* ht, unsigned ht_size, int is_short, unsigned mml, int is_rc, unsigned lead,
                                                            unsigned char* out, unsigned cap);
static unsigned g_chunk = LZC_CHUNK;
__attribute__((visibility("default"))) void lzc_host_set_chunk(unsigned c) { g_chunk = c; }
// same, also returning the chunk records (64 bytes each)
__attribute__((visibility("default"))) long lzc_host_encode_rec(const unsigned char* text, unsigned n, const unsigned char* ref, unsigned m,
                                                                const unsigned* ht, unsigned ht_size, int is_short, unsigned mml, int is_rc, unsigned lead,
                                                                unsigned char* out, unsigned cap, void* recs_out, unsigned recs_cap);
long lzc_host_encode(const unsigned char* text, unsigned n, const unsigned char* ref, unsigned m,
                     const unsigned* ht, unsigned ht_size, int is_short, unsigned mml, int is_rc, unsigned lead, unsigned char* out, unsigned cap)
{
    return lzc_host_encode_rec(text, n, ref, m, ht, ht_size, is_short, mml, is_rc, lead, out, cap, nullptr, 0);
}
long lzc_host_encode_rec(const unsigned char* text, unsigned n, const unsigned char* ref, unsigned m,
                         const unsigned* ht, unsigned ht_size, int is_short, unsigned mml, int is_rc, unsigned lead,
                         unsigned char* out, unsigned cap, void* recs_out, unsigned recs_cap)
{
    auto pack = [](const std::vector<unsigned char>& sym, std::vector<uint64_t>& w) {
        w.assign(sym.size() / 32 + 4, 0);
        unsigned char* b = (unsigned char*)w.data();
        for (size_t i = 0; i < sym.size(); ++i) b[i >> 2] |= (unsigned char)((sym[i] & 3u) << (6 - 2 * (i & 3)));
    };
    std::vector<unsigned char> store(lead, 1);                       // some leading bases (C's)
    if (!is_rc) store.insert(store.end(), text, text + n);
    else for (unsigned i = 0; i < n; ++i) store.push_back((unsigned char)(3 - text[n - 1 - i]));
    for (int i = 0; i < 40; ++i) store.push_back(2);                 // and trailing ones
    std::vector<unsigned char> rs(ref, ref + m);
    std::vector<uint64_t> T, R; pack(store, T); pack(rs, R);
    std::vector<uint16_t> h16; std::vector<uint32_t> h32;
    if (is_short) { h16.resize(ht_size); for (unsigned i = 0; i < ht_size; ++i) h16[i] = ht[i] == 0xffffffffu ? 0xffffu : (uint16_t)ht[i]; }
    else h32.assign(ht, ht + ht_size);
    LzcView<false> a; a.T = T.data(); a.gs = lead; a.n = n; a.rc = is_rc; a.R = R.data(); a.r_s = 0;
    a.ht = is_short ? (const void*)h16.data() : (const void*)h32.data(); a.ht_s = 0; a.mask = ht_size - 1; a.is_short = is_short; a.m = m;
    const unsigned nch = n ? (n + g_chunk - 1) / g_chunk : 1;
    std::vector<LzcRec> rec(nch);
    std::vector<unsigned char> cslab((size_t)nch * LZC_CSLAB + 64, 0xEE);
    for (unsigned k = 0; k < nch; ++k) {
        const unsigned c0 = k * g_chunk, c1 = lzc_min(n, c0 + g_chunk);
        lzc_parse_chunk<false>(a, c0, c1, mml, cslab.data() + (size_t)k * LZC_CSLAB, rec[k]);
        if (rec[k].bytes > LZC_CSLAB) return -3;
    }
    if (recs_out && recs_cap >= nch) memcpy(recs_out, rec.data(), (size_t)nch * sizeof(LzcRec));
    LzcReq q; memset(&q, 0, sizeof q);
    q.gstart = lead; q.n = n; q.is_rc = is_rc; q.nch = nch; q.out_cap = cap; q.chunk = g_chunk;
    return (long)lzc_stitch_segment(a, q, mml, rec.data(), cslab.data(), out, cap);
}
// GetCodingCostVector through the same chunk parse + stitch in cost mode; costs has n entries (zero-filled here).
// returns 0, or <= -10 when the stitcher hands the segment to the sequential kernel
__attribute__((visibility("default"))) long lzc_host_costs(const unsigned char* text, unsigned n, const unsigned char* ref, unsigned m,
                                                           const unsigned* ht, unsigned ht_size, int is_short, unsigned mml, int is_rc, unsigned lead,
                                                           int prefix, unsigned* costs)
{
    auto pack = [](const std::vector<unsigned char>& sym, std::vector<uint64_t>& w) {
        w.assign(sym.size() / 32 + 4, 0);
        unsigned char* b = (unsigned char*)w.data();
        for (size_t i = 0; i < sym.size(); ++i) b[i >> 2] |= (unsigned char)((sym[i] & 3u) << (6 - 2 * (i & 3)));
    };
    std::vector<unsigned char> store(lead, 1);
    if (!is_rc) store.insert(store.end(), text, text + n);
    else for (unsigned i = 0; i < n; ++i) store.push_back((unsigned char)(3 - text[n - 1 - i]));
    for (int i = 0; i < 40; ++i) store.push_back(2);
    std::vector<unsigned char> rs(ref, ref + m);
    std::vector<uint64_t> T, R; pack(store, T); pack(rs, R);
    std::vector<uint16_t> h16; std::vector<uint32_t> h32;
    if (is_short) { h16.resize(ht_size); for (unsigned i = 0; i < ht_size; ++i) h16[i] = ht[i] == 0xffffffffu ? 0xffffu : (uint16_t)ht[i]; }
    else h32.assign(ht, ht + ht_size);
    LzcView<false> a; a.T = T.data(); a.gs = lead; a.n = n; a.rc = is_rc; a.R = R.data(); a.r_s = 0;
    a.ht = is_short ? (const void*)h16.data() : (const void*)h32.data(); a.ht_s = 0; a.mask = ht_size - 1; a.is_short = is_short; a.m = m;
    const unsigned nch = n ? (n + g_chunk - 1) / g_chunk : 1;
    std::vector<LzcRec> rec(nch);
    for (unsigned i = 0; i < n; ++i) costs[i] = 0;
    for (unsigned k = 0; k < nch; ++k) {
        const unsigned c0 = k * g_chunk, c1 = lzc_min(n, c0 + g_chunk);
        lzc_parse_chunk<false, true>(a, c0, c1, mml, nullptr, rec[k], costs, (unsigned)prefix);
    }
    LzcReq q; memset(&q, 0, sizeof q);
    q.gstart = lead; q.n = n; q.is_rc = is_rc; q.nch = nch; q.out_cap = prefix; q.chunk = g_chunk;
    return (long)lzc_stitch_segment<LzcView<false>, true>(a, q, mml, rec.data(), nullptr, nullptr, 0, costs, (unsigned)prefix);
}
}
