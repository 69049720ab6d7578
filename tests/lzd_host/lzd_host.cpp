// TEST INFRASTRUCTURE ONLY: host build of agc_b200/csrc/lz_diag_core.cuh (the warp-per-segment LZ-diff encoder that streams along
// the current diagonal) so the CPU test-suite can compare it with the oracle without a GPU: the 32 lanes of every phase run one
// after the other.  The product (libagcgpu.so) only contains the device build.
#define LZD_COUNTERS
#include "../../agc_b200/csrc/lz_diag_core.cuh"
#include <cstring>
#include <vector>
extern "C" {
// text / ref: 1 byte per symbol (0..3).  ht: the reference's hash table widened to u32 (0xFFFFFFFF = empty).  When is_rc the packed
// store holds the REVERSE COMPLEMENT of text and the view reads it back in place, as the device does for rc segments.
// lead = bases of other data placed before the segment in the store (unaligned starts).  Returns the delta length, -2 = overflow.
__attribute__((visibility("default"))) long lzd_host_encode(const unsigned char* text, unsigned n, const unsigned char* ref, unsigned m,
                                                            const unsigned* ht, unsigned ht_size, int is_short, unsigned mml, int is_rc, unsigned lead,
                                                            unsigned char* out, unsigned cap)
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
    static LzdScratch S;
    memset(&S, 0xA5, sizeof S);
    return (long)lzd_encode_segment(a, S, mml, out, cap);
}
// rounds, windows, path_tokens, stops, defers, opens, multi, probes since the last call
__attribute__((visibility("default"))) void lzd_host_counters(unsigned long long* out8) { memcpy(out8, &g_lzd_cnt, sizeof g_lzd_cnt); memset(&g_lzd_cnt, 0, sizeof g_lzd_cnt); }
__attribute__((visibility("default"))) unsigned lzd_host_scratch_bytes() { return (unsigned)sizeof(LzdScratch); }
}
