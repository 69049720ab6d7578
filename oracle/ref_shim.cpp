// TEST INFRASTRUCTURE ONLY. C shim over the *unmodified* reference CLZDiff_V2
// (/root/reference/src/common/lz_diff.cpp, compiled where it lies by oracle/Makefile.ref).
// Exposes Encode / Estimate / GetCodingCostVector so tests can pin oracle/agc_oracle.c
// and the CUDA kernels against the reference itself. Never linked into the product.
#include "lz_diff.h"
#include <cstring>
extern "C" {
// returns encoded size; out may be NULL to query size only
long ref_lz_encode(const uint8_t* ref, long m, const uint8_t* text, long n, int min_match_len, uint8_t* out, long cap)
{
    CLZDiff_V2 lz(min_match_len);
    lz.SetMinMatchLen(min_match_len);
    contig_t r(ref, ref + m), t(text, text + n), e;
    lz.Prepare(r);
    lz.Encode(t, e);
    if (out && (long)e.size() <= cap) memcpy(out, e.data(), e.size());
    return (long)e.size();
}
long ref_lz_estimate(const uint8_t* ref, long m, const uint8_t* text, long n, int min_match_len, unsigned bound)
{
    CLZDiff_V2 lz(min_match_len);
    lz.SetMinMatchLen(min_match_len);
    contig_t r(ref, ref + m), t(text, text + n);
    lz.Prepare(r);
    return (long)lz.Estimate(t, bound);
}
long ref_lz_cost_vector(const uint8_t* ref, long m, const uint8_t* text, long n, int min_match_len, int prefix_costs, uint32_t* out)
{
    CLZDiff_V2 lz(min_match_len);
    lz.SetMinMatchLen(min_match_len);
    contig_t r(ref, ref + m), t(text, text + n);
    lz.Prepare(r);
    lz.AssureIndex();
    std::vector<uint32_t> v;
    lz.GetCodingCostVector(t, v, prefix_costs != 0);
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
    return (long)v.size();
}
long ref_lz_decode(const uint8_t* ref, long m, const uint8_t* enc, long e, int min_match_len, uint8_t* out, long cap)
{
    CLZDiff_V2 lz(min_match_len);
    contig_t r(ref, ref + m), en(enc, enc + e), d;
    lz.SetMinMatchLen(min_match_len);
    lz.Decode(r, en, d);
    if ((long)d.size() <= cap) memcpy(out, d.data(), d.size());
    return (long)d.size();
}
}
