// TEST INFRASTRUCTURE ONLY. C shim over the *unmodified* reference CLZDiff_V2
// (/root/reference/src/common/lz_diff.cpp, compiled where it lies by oracle/Makefile.ref).
// Exposes Encode / Estimate / GetCodingCostVector so tests can pin oracle/agc_oracle.c
// and the CUDA kernels against the reference itself. Never linked into the product.
#include "lz_diff.h"
#include <cstring>
#include <ctime>
#include <vector>
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
// per-stage CPU microbenchmark: CLZDiff_V2::Encode alone -- the index is prepared once, then every text
// (text i = texts[offs[i] .. offs[i+1])) is encoded `reps` times; returns the nanoseconds spent in Encode, *out_bytes the delta bytes
long ref_lz_encode_many_ns(const uint8_t* ref, long m, const uint8_t* texts, const long* offs, long n_texts, int min_match_len, int reps,
                           long* out_bytes)
{
    CLZDiff_V2 lz(min_match_len);
    lz.SetMinMatchLen(min_match_len);
    contig_t r(ref, ref + m);
    lz.Prepare(r);
    lz.AssureIndex();
    std::vector<contig_t> t;
    for (long i = 0; i < n_texts; ++i) t.emplace_back(texts + offs[i], texts + offs[i + 1]);
    contig_t e; long bytes = 0;
    struct timespec a, b;
    clock_gettime(CLOCK_MONOTONIC, &a);
    for (int k = 0; k < reps; ++k) for (auto& x : t) { lz.Encode(x, e); bytes += (long)e.size(); }
    clock_gettime(CLOCK_MONOTONIC, &b);
    if (out_bytes) *out_bytes = bytes;
    return (b.tv_sec - a.tv_sec) * 1000000000L + (b.tv_nsec - a.tv_nsec);
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
