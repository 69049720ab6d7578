// lz_chunk_core.cuh -- the chunk parser and the stitcher of the chunk-parallel LZ-diff encoder as host/device source: the
// kernels of kernels_lz_chunk.cu instantiate it on the device; tests/lzc_host builds the same source for the CPU so the test
// suite can compare chunk parse + stitch with the oracle without a GPU (test infrastructure; the product has no host path).
#pragma once
#include <stdint.h>
#include "lz_chunk.cuh"
#ifdef __CUDACC__
#define LZC_HD __host__ __device__ __forceinline__
#else
#define LZC_HD inline
#endif
#ifndef AGC_EMPTY32
#define AGC_EMPTY32 0xffffffffu
#endif
// warp votes of the parser's scheduler: the whole warp calls lzc_parse_chunk together on the device; one lane on the host
#ifdef __CUDA_ARCH__
#define LZC_BALLOT(p) __ballot_sync(0xffffffffu, (p))
#define LZC_POPC(x) __popc(x)
#else
#define LZC_BALLOT(p) ((p) ? 1u : 0u)
#define LZC_POPC(x) ((x) ? 1 : 0)
#endif

// ---- portable forms of the intrinsics the device code uses
LZC_HD uint32_t lzc_min(uint32_t a, uint32_t b) { return a < b ? a : b; }
LZC_HD uint32_t lzc_max(uint32_t a, uint32_t b) { return a > b ? a : b; }
LZC_HD uint32_t lzc_clz64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__clzll((long long)x);
#else
    return x ? (uint32_t)__builtin_clzll(x) : 64u;
#endif
}
LZC_HD uint32_t lzc_ctz64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)(__ffsll((long long)x) - 1);
#else
    return (uint32_t)__builtin_ctzll(x);
#endif
}
LZC_HD uint64_t lzc_be64(uint64_t x)        // first base (first byte, top two bits) to bits 63:62
{
#ifdef __CUDA_ARCH__
    uint32_t lo = __byte_perm((uint32_t)x, 0, 0x0123), hi = __byte_perm((uint32_t)(x >> 32), 0, 0x0123);
    return ((uint64_t)lo << 32) | hi;
#else
    return __builtin_bswap64(x);
#endif
}
LZC_HD uint64_t lzc_rev2(uint64_t x)        // reverse the order of the 32 2-bit groups
{
#ifdef __CUDA_ARCH__
    uint64_t r = __brevll(x);
#else
    uint64_t r = x;
    r = ((r >> 1) & 0x5555555555555555ULL) | ((r & 0x5555555555555555ULL) << 1);
    r = ((r >> 2) & 0x3333333333333333ULL) | ((r & 0x3333333333333333ULL) << 2);
    r = ((r >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((r & 0x0F0F0F0F0F0F0F0FULL) << 4);
    r = __builtin_bswap64(r);
#endif
    return ((r & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((r & 0x5555555555555555ULL) << 1);
}
LZC_HD uint64_t lzc_murmur64(uint64_t h)    // MurMur64Hash (src/common/utils.h:164-176)
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}
// 32 bases starting at base index g (may be negative, > -32: the slots before base 0 are garbage the callers never use)
LZC_HD uint64_t lzc_win_s(const uint64_t* P, int64_t g)
{
    if (g < 0) { if (g <= -32) return 0; return lzc_be64(P[0]) >> (uint32_t)(2 * (-g)); }
    const uint64_t i = (uint64_t)g >> 5; const uint32_t sh = (uint32_t)(g & 31) * 2u;
    const uint64_t a = lzc_be64(P[i]), b = lzc_be64(P[i + 1]);
    return (a << sh) | ((b >> 1) >> (63u - sh));
}
#ifdef __CUDA_ARCH__
__device__ __forceinline__ uint64_t lzc_lds64(uint32_t a) { uint64_t v; asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lzc_lds16(uint32_t a) { uint16_t v; asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a)); return v; }
__device__ __forceinline__ uint32_t lzc_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
#else
inline uint64_t lzc_lds64(uint32_t) { return 0; }
inline uint32_t lzc_lds16(uint32_t) { return 0; }
inline uint32_t lzc_lds32(uint32_t) { return 0; }
#endif

LZC_HD uint32_t lzc_put_int(uint8_t* dst, int64_t xs)            // append_int (lz_diff.h:229-262)
{
    uint32_t n = 0;
    if (xs == 0) { dst[0] = '0'; return 1; }
    if (xs < 0) { dst[n++] = '-'; xs = -xs; }
    uint32_t x = (uint32_t)xs;
    const uint32_t nd = x < 10 ? 1 : x < 100 ? 2 : x < 1000 ? 3 : x < 10000 ? 4 : x < 100000 ? 5 : x < 1000000 ? 6 : x < 10000000 ? 7
                      : x < 100000000 ? 8 : x < 1000000000 ? 9 : 10;
    for (uint32_t k = nd; k-- > 0;) { const uint32_t q = x / 10u; dst[n + k] = (uint8_t)('0' + (x - q * 10u)); x = q; }
    return n + nd;
}
// encode_match (lz_diff.cpp:631-643)
LZC_HD uint32_t lzc_put_match(uint8_t* dst, int64_t dif, bool with_len, uint32_t lenv)
{
    uint32_t L = lzc_put_int(dst, dif);
    if (with_len) { dst[L++] = ','; L += lzc_put_int(dst + L, (int64_t)lenv); }
    dst[L++] = '.';
    return L;
}

// coding_cost_match (lz_diff.h:159-172): always counts the length field
LZC_HD uint32_t lzc_int_len(uint32_t x)
{
    return x < 10 ? 1 : x < 100 ? 2 : x < 1000 ? 3 : x < 10000 ? 4 : x < 100000 ? 5 : x < 1000000 ? 6 : x < 10000000 ? 7
         : x < 100000000 ? 8 : x < 1000000000 ? 9 : 10;
}
LZC_HD uint32_t lzc_match_cost(uint32_t mp, uint32_t len, uint32_t pred, uint32_t mml)
{
    const int dif = (int)mp - (int)pred;
    return (dif >= 0 ? lzc_int_len((uint32_t)dif) : lzc_int_len((uint32_t)-dif) + 1u) + lzc_int_len(len - mml) + 2u;
}

// view of one (text segment, reference) pair for a single thread
template <bool STAGED, bool HT_STAGED = STAGED>
struct LzcView {
    const uint64_t* T; int64_t gs; uint32_t n, rc;
    const uint64_t* R; uint32_t r_s;            // reference words: generic pointer / shared-memory address
    const void* ht; uint32_t ht_s;              // hash table
    uint32_t mask, is_short, m;

    LZC_HD uint64_t twin(int64_t pos) const
    {
        if (!rc) return lzc_win_s(T, gs + pos);
        return ~lzc_rev2(lzc_win_s(T, gs + (int64_t)n - pos - 32));
    }
    LZC_HD uint64_t rword(uint32_t i) const { return STAGED ? lzc_lds64(r_s + 8u * i) : R[i]; }
    LZC_HD uint64_t rwin(int64_t pos) const
    {
        if (pos < 0) { if (pos <= -32) return 0; return rwin0() >> (uint32_t)(2 * (-pos)); }
        const uint32_t i = (uint32_t)pos >> 5, sh = ((uint32_t)pos & 31u) * 2u;
        const uint64_t a = lzc_be64(rword(i)), b = lzc_be64(rword(i + 1));
        return (a << sh) | ((b >> 1) >> (63u - sh));
    }
    LZC_HD uint64_t rwin0() const { return lzc_be64(rword(0)); }
    // the 16 reference bases that start at the indexed position 4 * v (a whole byte of the packed reference), as 32 big-endian bits:
    // the cheap first test of a slot (find_best_match's key compare)
    LZC_HD uint32_t rquick(uint32_t v) const
    {
#ifdef __CUDA_ARCH__
        const uint32_t o = v & 3u;
        uint32_t w0, w1;
        if (STAGED) { w0 = lzc_lds32(r_s + (v & ~3u)); w1 = lzc_lds32(r_s + (v & ~3u) + 4u); }
        else { const uint32_t* w = (const uint32_t*)R + (v >> 2); w0 = w[0]; w1 = w[1]; }
        return __byte_perm(w0, w1, 0x0123u + o * 0x1111u);
#else
        const uint8_t* b = (const uint8_t*)R + v;
        return ((uint32_t)b[0] << 24) | ((uint32_t)b[1] << 16) | ((uint32_t)b[2] << 8) | (uint32_t)b[3];
#endif
    }
    // two consecutive 32-base windows of the reference from three words
    LZC_HD void rwin2(int64_t pos, uint64_t& r0, uint64_t& r1) const
    {
        if (pos < 0) { r0 = rwin(pos); r1 = rwin(pos + 32); return; }
        const uint32_t i = (uint32_t)pos >> 5, sh = ((uint32_t)pos & 31u) * 2u;
        const uint64_t a = lzc_be64(rword(i)), b = lzc_be64(rword(i + 1)), c = lzc_be64(rword(i + 2));
        r0 = (a << sh) | ((b >> 1) >> (63u - sh));
        r1 = (b << sh) | ((c >> 1) >> (63u - sh));
    }
    // The 64 text bases of the 16-byte aligned block of the packed store that holds text position p (< n), in text order (w0 then
    // w1), one 16-byte load; rel = text position of the block's first base minus p (-63 .. 0).
    LZC_HD void tblock(uint32_t p, uint64_t& w0, uint64_t& w1, int32_t& rel) const
    {
        const uint64_t s = rc ? (uint64_t)(gs + (int64_t)n - 1 - (int64_t)p) : (uint64_t)(gs + (int64_t)p);
        const uint64_t blk = s >> 6;
#ifdef __CUDA_ARCH__
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(T) + blk);
        const uint64_t lo = ((uint64_t)__byte_perm(q.x, 0, 0x0123) << 32) | __byte_perm(q.y, 0, 0x0123);
        const uint64_t hi = ((uint64_t)__byte_perm(q.z, 0, 0x0123) << 32) | __byte_perm(q.w, 0, 0x0123);
#else
        const uint64_t lo = lzc_be64(T[2 * blk]), hi = lzc_be64(T[2 * blk + 1]);
#endif
        if (!rc) { w0 = lo; w1 = hi; rel = -(int32_t)(s & 63u); }
        else { w0 = ~lzc_rev2(hi); w1 = ~lzc_rev2(lo); rel = -(int32_t)(63u - (uint32_t)(s & 63u)); }
    }
    LZC_HD uint32_t slot(uint32_t s) const
    {
        if (is_short) { const uint32_t v = HT_STAGED ? lzc_lds16(ht_s + 2u * s) : (uint32_t)((const uint16_t*)ht)[s]; return v == 0xffffu ? AGC_EMPTY32 : v; }
        return HT_STAGED ? lzc_lds32(ht_s + 4u * s) : ((const uint32_t*)ht)[s];
    }
    LZC_HD uint32_t tsym(uint32_t q) const { return (uint32_t)(twin(q) >> 62); }
    LZC_HD uint32_t rsym(uint32_t q) const { return (uint32_t)(rwin(q) >> 62); }

    // matching_length(text + tp, ref + rp, maxlen)
    LZC_HD uint32_t lcp_fwd(uint32_t tp, uint32_t rp, uint32_t maxlen) const
    {
        for (uint32_t off = 0; off < maxlen; off += 32) {
            const uint64_t x = twin((int64_t)tp + off) ^ rwin((int64_t)rp + off);
            if (x) { const uint32_t l = off + (lzc_clz64(x) >> 1); return l < maxlen ? l : maxlen; }
        }
        return maxlen;
    }
    // backward: text[tp-1-j] == ref[rp-1-j], j < lim
    LZC_HD uint32_t lcp_bwd(uint32_t tp, uint32_t rp, uint32_t lim) const
    {
        for (uint32_t off = 0; off < lim; off += 32) {
            const uint64_t x = twin((int64_t)tp - off - 32) ^ rwin((int64_t)rp - off - 32);
            if (x) { const uint32_t l = off + (lzc_ctz64(x) >> 1); return l < lim ? l : lim; }
        }
        return lim;
    }
};

// find_best_match16/32 (lz_diff.cpp:287-372) exactly as the sequential code evaluates it (all candidates, full extensions)
template <class View>
LZC_HD bool lzc_best_match_full(const View& a, uint32_t h, uint64_t x, uint32_t p, uint32_t np, uint32_t kl, uint32_t mml,
                                    uint32_t& o_hp, uint32_t& o_b, uint32_t& o_f, bool& np_limited)
{
    uint32_t best_b = 0, best_f = 0, best_hp = 0, mtu = mml;
    np_limited = false;
    for (uint32_t t = 0; t < 64; ++t) {
        const uint32_t v = a.slot((h + t) & a.mask);
        if (v == AGC_EMPTY32) break;
        const uint32_t hp = v * 4u;
        if ((a.rwin(hp) >> (64 - 2 * kl)) != x) continue;
        const uint32_t f = a.lcp_fwd(p, hp, lzc_min(a.n - p, a.m - hp));
        const uint32_t lim = lzc_min(np, hp);
        const uint32_t b = lim ? a.lcp_bwd(p, hp, lim) : 0u;
        if (b == np && np < hp) np_limited = true;              // more previous literals could have extended this candidate
        if (b + f > mtu) { best_b = b; best_f = f; best_hp = hp; mtu = b + f; }
    }
    o_hp = best_hp; o_b = best_b; o_f = best_f;
    return best_b + best_f >= mml;
}

// ------------------------------------------------------------------------------------------------ phase 1: one lane, a stream of chunks
// COSTS: instead of token bytes a chunk writes GetCodingCostVector entries (lz_diff.cpp:159-284: the same parse) for what is final
// in its own view -- 1 for every literal, the token cost of its later matches at their first (prefix) or last position; the
// vector is zero-filled before, the stitcher adds the first / merged matches and clears what the true parse covers differently.
struct LzcItem {                   // one chunk handed to a lane
    uint32_t c0, c1;               // text range
    uint8_t* out;                  // its token slab
    uint32_t* v; uint32_t prefix;  // COSTS: the segment's cost vector, prefix_costs
    LzcRec* rec;                   // where its record goes
};

// Warp-cooperative state machine over a stream of chunks: fetch(a, item) sets the text fields of the view (gs, n, rc) and the item,
// false = no chunk left.  A lane is EXTENDING (stepping a forward extension, 64 bases per step -- almost all of the work), in NEED of
// the rare path (resolve a finished extension: backward part, decision, token output; probe positions until the next extension
// starts; finish a chunk and fetch the next one) or DONE (no chunk left).  The lanes of a warp are all at different places of their
// chunks, so left alone every iteration would pay for every path with a handful of lanes active in each.  Instead every lane does
// ONE unit of work per iteration -- an extension step, the probe of one position, the resolve of one candidate, or the end of a chunk
// plus the fetch of the next -- and the warp votes which kind of unit runs: the one most lanes wait for (the others park for that
// iteration).  Each path then runs with many lanes active, and a lane that finishes a chunk takes the next one at once, so no lane
// idles while others still work.  (Host build: one lane, the votes are trivial.)
template <bool STAGED, bool COSTS, class Fetch>
LZC_HD void lzc_parse_stream(LzcView<STAGED> a, uint32_t mml, Fetch& fetch)
{
    const uint32_t kl = mml - 3u, m = a.m;
    LzcItem it; it.c0 = it.c1 = 0; it.out = nullptr; it.v = nullptr; it.prefix = 0; it.rec = nullptr;
    LzcRec R;
    uint32_t n = 0, c0 = 0, c1 = 0, i = 0, np = 0, pred = 0, olen = 0, flags = 0, prefix = 0;
    bool have_first = false, neq_checked = false, ended_open = false, have_item = false;
    int32_t end_diag = 0;
    uint8_t* out = nullptr; uint32_t* v = nullptr;
    // lane states; every iteration the warp serves the state most lanes are in (see above)
    enum { ST_EXT = 0, ST_PROBE = 1, ST_RESOLVE = 2, ST_FIN = 3, ST_DONE = 4 };
    int st = ST_FIN;                                      // "finish the (non-existent) chunk and fetch one"
    uint32_t e_hp = 0, e_off = 0, e_lim = 0, e_maxlen = 0, e_f = 0, e_b = 0;
    bool e_multi = false, e_ok = false, e_npl = false;    // RESOLVE of a multi-candidate probe: everything is known already
    for (;;) {
        const uint32_t n_ext = LZC_POPC(LZC_BALLOT(st == ST_EXT)), n_probe = LZC_POPC(LZC_BALLOT(st == ST_PROBE));
        const uint32_t n_res = LZC_POPC(LZC_BALLOT(st == ST_RESOLVE)), n_fin = LZC_POPC(LZC_BALLOT(st == ST_FIN));
        if (!(n_ext | n_probe | n_res | n_fin)) break;
        int pick = ST_EXT; uint32_t best = n_ext;
        if (n_probe > best) { pick = ST_PROBE; best = n_probe; }
        if (n_res > best) { pick = ST_RESOLVE; best = n_res; }
        if (n_fin > best) { pick = ST_FIN; best = n_fin; }
        if (st != pick) continue;                         // this lane parks

        if (pick == ST_EXT) {
            // matching_length(text + i, ref + e_hp, e_lim): aligned 64-base blocks of the text (a single 16-byte load each, no
            // shifting on the text side) against two windows of the reference; two blocks per unit
            bool fin = false;
#pragma unroll
            for (int rep = 0; rep < 2 && !fin; ++rep) {
                uint64_t w0, w1, r0, r1; int32_t rel;
                a.tblock(i + e_off, w0, w1, rel);
                const int32_t o0 = (int32_t)e_off + rel;          // block start relative to i; negative only on the first step
                a.rwin2((int64_t)e_hp + o0, r0, r1);
                uint64_t x0 = w0 ^ r0, x1 = w1 ^ r1;
                if (o0 < 0) {                                     // bases before the first compared position do not count
                    const uint32_t skip = (uint32_t)(-o0);
                    if (skip >= 32u) { x0 = 0; if (skip > 32u) x1 &= ~0ull >> (2u * (skip - 32u)); }
                    else x0 &= ~0ull >> (2u * skip);
                }
                fin = true;
                if (x0) e_f = (uint32_t)(o0 + (int32_t)(lzc_clz64(x0) >> 1));
                else if (x1) e_f = (uint32_t)(o0 + 32 + (int32_t)(lzc_clz64(x1) >> 1));
                else { e_off = (uint32_t)(o0 + 64); if (e_off < e_lim) fin = false; else e_f = e_lim; }
            }
            if (fin) { if (e_f > e_lim) e_f = e_lim; e_multi = false; st = ST_RESOLVE; }
            continue;
        }

        if (pick == ST_PROBE) {
            // one text position: candidates = slots until the first empty one whose key equals the text's (find_best_match's
            // "f_len >= key_len"); the first 16 bases of a slot's key are tested with one unaligned 32-bit read
            const uint64_t x = a.twin(i) >> (64 - 2 * kl);
            const uint32_t h = (uint32_t)lzc_murmur64(x) & a.mask;
            const uint32_t xq = kl >= 16u ? (uint32_t)(x >> (2u * kl - 32u)) : (uint32_t)(x << (32u - 2u * kl));
            uint32_t ncand = 0, hp0 = 0;
            for (uint32_t t = 0; t < 64; ++t) {
                const uint32_t sv = a.slot((h + t) & a.mask);
                if (sv == AGC_EMPTY32) break;
                const uint32_t dq = a.rquick(sv) ^ xq;
                if (kl >= 16u ? dq != 0u : (dq >> (32u - 2u * kl)) != 0u) continue;
                if (kl > 16u && (a.rwin(sv * 4u) >> (64 - 2 * kl)) != x) continue;
                if (!ncand) hp0 = sv * 4u;
                if (++ncand > 1) break;
            }
            if (ncand == 1) {
                e_hp = hp0; e_off = 0;
                e_maxlen = lzc_min(n - i, m - hp0);
                e_lim = lzc_min(e_maxlen, lzc_max(c1 - i, mml + 1u));       // cap: enough to decide "b + f > min_match_len" whatever b is
                st = ST_EXT;
                continue;
            }
            if (ncand > 1) {                                                // rare: the sequential evaluation of all candidates
                e_ok = lzc_best_match_full(a, h, x, i, np, kl, mml, e_hp, e_b, e_f, e_npl);
                e_multi = true; st = ST_RESOLVE;
                continue;
            }
            if (!neq_checked) { neq_checked = true; if (a.tsym(i) != a.rsym(i)) flags |= LZC_NEQ; }
            ++i; ++np;                                                      // literal
            if (!(i < c1 && i + kl < n)) st = ST_FIN;
            continue;
        }

        if (pick == ST_RESOLVE) {
            // the candidate(s) of position i are evaluated: backward part, decision, token output
            bool ok, open = false, np_limited;
            uint32_t hp = e_hp, b, f = e_f;
            const bool multi = e_multi;
            if (multi) { ok = e_ok; b = e_b; np_limited = e_npl; }
            else {
                const uint32_t lim = lzc_min(np, hp);
                b = lim ? a.lcp_bwd(i, hp, lim) : 0u;
                np_limited = (b == np && np < hp);
                ok = b + f > mml;
                open = ok && f == e_lim && e_lim < e_maxlen;
            }
            if (!ok) {
                // a candidate that failed although its backward extension was cut short by no_prev_literals: with the true (possibly
                // larger) count it might have succeeded -- only matters before the chunk's first match
                if (!have_first && np_limited) flags |= LZC_SENS;
                if (!neq_checked) { neq_checked = true; if (a.tsym(i) != a.rsym(i)) flags |= LZC_NEQ; }
                ++i; ++np;
                st = (i < c1 && i + kl < n) ? ST_PROBE : ST_FIN;
                continue;
            }
            const uint32_t ts = i - b, mp = hp - b, len = b + f;
            np -= b;                                                         // literals [ts - np, ts) stay pending
            if (!have_first) {
                have_first = true;
                flags |= LZC_HAS_FIRST | (open ? LZC_FIRST_OPEN : 0u) | (multi ? LZC_FIRST_MULTI : 0u);
                if (np_limited && ts == c0) flags |= LZC_FIRST_BLIM;        // back extension reached the chunk start: may go on before it
                if (multi && np_limited) flags |= LZC_SENS;                 // candidate choice depends on the true no_prev_literals
                R.lit0 = np; R.first_p = i; R.first_ts = ts; R.first_mp = mp; R.first_len = len;
                if (COSTS) for (uint32_t j = 0; j < np; ++j) v[ts - np + j] = 1u;
                else for (uint32_t j = 0; j < np; ++j) out[olen + j] = (uint8_t)('A' + a.tsym(ts - np + j));
                olen += np;
                if (!(n == m && ts == c0 && mp == c0)) { if (!neq_checked && np) { neq_checked = true; if (a.tsym(ts - np) != a.rsym(ts - np)) flags |= LZC_NEQ; } }
            } else {
                const uint32_t pred_now = pred + np;
                if (COSTS) {
                    for (uint32_t j = 0; j < np; ++j) v[ts - np + j] = 1u;
                    olen += np;
                    if (open) { flags |= LZC_END_OPEN; R.open_ts = ts; R.open_mp = mp; R.open_predb = pred_now; }
                    else {
                        // a chunk only writes inside its own range (its neighbour writes its own view there at the same time):
                        // the cost of a match that ends beyond the chunk is left to the stitcher
                        const uint32_t cpos = prefix ? ts : ts + len - 1u, tc = lzc_match_cost(mp, len, pred_now, mml);
                        if (cpos < c1) v[cpos] = tc; else { R.pad[0] = tc; R.pad[1] = cpos; }
                        olen += 1u;
                    }
                } else {
                    const bool bang = (mp == pred_now);
                    for (uint32_t j = 0; j < np; ++j) {
                        const uint32_t q = ts - np + j, sy = a.tsym(q);
                        uint8_t ch = (uint8_t)('A' + sy);
                        const uint32_t d = np - j;                           // distance back from the match (lz_diff.cpp:772)
                        if (bang && d < mp && sy == a.rsym(mp - d)) ch = '!';
                        out[olen + j] = ch;
                    }
                    olen += np;
                    if (open) { flags |= LZC_END_OPEN; R.open_ts = ts; R.open_mp = mp; R.open_predb = pred_now; }
                    else {
                        const bool to_end = (ts + len == n) && (mp + len == m);
                        olen += lzc_put_match(out + olen, (int64_t)(int)mp - (int64_t)(int)pred_now, !to_end, len - mml);
                    }
                }
            }
            pred = mp + len; i = ts + len; np = 0; end_diag = (int32_t)mp - (int32_t)ts;
            if (open) { ended_open = true; st = ST_FIN; }
            else st = (i < c1 && i + kl < n) ? ST_PROBE : ST_FIN;
            continue;
        }

        // ST_FIN: close the chunk (if there is one) and take the next
        if (have_item) {
            // tail of the text (positions the sequential loop never probes) and literals pending at the chunk's end
            if (!ended_open && i < c1 && i + kl >= n) { const uint32_t e = lzc_min(c1, n); np += e - i; i = e; }
            if (np) {
                if (!neq_checked) { neq_checked = true; if (a.tsym(i - np) != a.rsym(i - np)) flags |= LZC_NEQ; }
                if (COSTS) for (uint32_t j = 0; j < np; ++j) v[i - np + j] = 1u;
                else for (uint32_t j = 0; j < np; ++j) out[olen + j] = (uint8_t)('A' + a.tsym(i - np + j));
                olen += np;
                if (!have_first) R.lit0 = np;
            }
            // "this chunk equals the reference at the same positions": one diagonal-0 match from the chunk's first position to its end
            if (n == m && have_first && R.lit0 == 0 && R.first_ts == c0 && R.first_mp == c0 && i >= lzc_min(c1, n) && olen == 0 && !(flags & LZC_END_OPEN))
                flags |= LZC_EQ;
            R.flags = flags; R.bytes = olen; R.end_i = i; R.end_np = np; R.end_pred = pred + np; R.end_diag = end_diag;
            *it.rec = R;
            have_item = false;
        }
        if (!fetch(a, it)) { st = ST_DONE; continue; }
        n = a.n; c0 = it.c0; c1 = it.c1; out = it.out; v = it.v; prefix = it.prefix;
        i = c0; np = 0; pred = 0; olen = 0; flags = 0;
        have_first = false; neq_checked = (n != m); ended_open = false; end_diag = 0; have_item = true;
        R.lit0 = 0; R.first_p = R.first_ts = R.first_mp = R.first_len = 0; R.open_ts = R.open_mp = R.open_predb = 0;
        R.pad[0] = 0; R.pad[1] = 0xffffffffu;             // cost mode: (cost, position) of a token cost the stitcher has to write
        st = (i < c1 && i + kl < n) ? ST_PROBE : ST_FIN;
    }
}

// one chunk through the stream parser (host build and tests)
template <bool STAGED, bool COSTS = false>
LZC_HD void lzc_parse_chunk(const LzcView<STAGED>& a, uint32_t c0, uint32_t c1, uint32_t mml, uint8_t* out, LzcRec& R,
                            uint32_t* v = nullptr, uint32_t prefix = 0)
{
    struct One {
        uint32_t c0, c1; uint8_t* out; uint32_t* v; uint32_t prefix; LzcRec* rec; bool given;
        LZC_HD bool operator()(LzcView<STAGED>&, LzcItem& it)
        {
            if (given) return false;
            given = true; it.c0 = c0; it.c1 = c1; it.out = out; it.v = v; it.prefix = prefix; it.rec = rec;
            return true;
        }
    } one = { c0, c1, out, v, prefix, &R, false };
    lzc_parse_stream<STAGED, COSTS>(a, mml, one);
}


// ------------------------------------------------------------------------------------------------ phase 2: one thread, one segment
// rec / cslab: records and slabs of the segment's chunks (chunk k at rec[k], cslab + k * LZC_CSLAB).  Returns the number of bytes
// written to out, -1 when the segment has to go to the sequential kernel, -2 when out_cap was too small.
// COSTS: no bytes; the stitcher completes the cost vector v the chunks started (see lzc_parse_chunk) and returns 0.
template <class View, bool COSTS = false>
LZC_HD int64_t lzc_stitch_segment(const View& a, const LzcReq& q, uint32_t mml, const LzcRec* rec, const uint8_t* cslab, uint8_t* out, uint32_t cap,
                                  uint32_t* v = nullptr, uint32_t prefix = 0)
{
    const uint32_t n = q.n, m = a.m;
    const uint32_t nch = q.nch;
    int fail = 0; bool ovf = false;          // fail: why the segment goes to the sequential kernel (diagnostic codes)
    uint32_t keep_pos = 0xffffffffu;         // COSTS: position of the last token cost written here (clear() must not erase it)
    auto clear = [&](uint32_t from, uint32_t to) {      // COSTS: positions [from, to) are covered by a match of the true parse
        if (COSTS) for (uint32_t p = from; p < to; ++p) if (p != keep_pos) v[p] = 0u;
    };

    // equal sequences (lz_diff.cpp:678-680): every chunk is one diagonal-0 match over its whole range
    if (!COSTS && n == m) {
        bool all_eq = true, any_neq = false;
        for (uint32_t k = 0; k < nch; ++k) { const uint32_t f = rec[k].flags; all_eq &= (f & LZC_EQ) != 0; any_neq |= (f & LZC_NEQ) != 0; }
        if (nch && all_eq) return 0;
        if (!any_neq) fail = 1;                     // neither proven equal nor proven different: the sequential kernel decides
    }

    uint32_t o = 0;                                    // bytes written
    // TRUE state of the sequential parse: `pos` = first text position not covered by an emitted token, `np` pending literals
    // (already in the output as raw letters), `pred` = pred_pos at pos.  pos may lie beyond the start of the next chunk when a
    // match crossed the boundary (cov_diag = its diagonal).  `open`: a match verified up to the current chunk, end unknown.
    uint32_t pos = 0, np = 0, pred = 0;
    uint32_t cov_start = 0;                            // text start of the match that ends at pos
    int64_t cov_diag = 0;
    bool open = false;
    uint32_t o_ts = 0, o_mp = 0, o_predb = 0;

    auto put_lit_text = [&](uint32_t from, uint32_t to) {          // raw literals of text positions [from, to)
        if (COSTS) { for (uint32_t p = from; p < to; ++p) v[p] = 1u; return; }
        for (uint32_t p = from; p < to; ++p) { if (o < cap) out[o] = (uint8_t)('A' + a.tsym(p)); else ovf = true; ++o; }
    };
    auto copy_bytes = [&](const uint8_t* src, uint32_t cnt) {
        if (COSTS) return;                                          // the chunk wrote its entries itself
        if ((uint64_t)o + cnt > cap) { ovf = true; o += cnt; return; }
        for (uint32_t j = 0; j < cnt; ++j) out[o + j] = src[j];
        o += cnt;
    };
    auto put_match = [&](uint32_t ts, uint32_t mp, uint32_t len, uint32_t predb) {
        if (COSTS) { keep_pos = prefix ? ts : ts + len - 1u; v[keep_pos] = lzc_match_cost(mp, len, predb, mml); return; }
        uint8_t buf[24];
        const bool to_end = (ts + len == n) && (mp + len == m);
        const uint32_t L = lzc_put_match(buf, (int64_t)(int)mp - (int64_t)(int)predb, !to_end, len - mml);
        copy_bytes(buf, L);
    };
    // the '!' rewrite (lz_diff.cpp:769-779) over the literal run that ends the output
    auto bang = [&](uint32_t mp) {
        if (COSTS || ovf) return;
        for (uint32_t d = 1; d < o && d < mp; ++d) {
            const uint8_t ch = out[o - d];
            if (ch < 'A' || ch > 'Z') break;
            if ((uint32_t)(ch - 'A') == a.rsym(mp - d)) out[o - d] = '!';
        }
    };

    for (uint32_t k = 0; k < nch && !fail; ++k) {
        const LzcRec R = rec[k];
        const uint32_t c0 = k * q.chunk, c1 = lzc_min(n, c0 + q.chunk);
        const uint8_t* cb = cslab + (uint64_t)k * LZC_CSLAB;
        const bool has_first = (R.flags & LZC_HAS_FIRST) != 0;
        // state after the chunk's own last token
        auto take_end = [&]() {
            if (R.flags & LZC_END_OPEN) { open = true; o_ts = R.open_ts; o_mp = R.open_mp; o_predb = R.open_predb; }
            else {
                pos = R.end_i; np = R.end_np; pred = R.end_pred; cov_diag = R.end_diag; cov_start = 0;
                if (COSTS && R.pad[1] != 0xffffffffu) { v[R.pad[1]] = R.pad[0]; keep_pos = R.pad[1]; }    // the chunk's last match ends beyond it
                else if (COSTS && !prefix && R.end_np == 0 && has_first) keep_pos = R.end_i - 1u;
            }
        };

        if (open) {
            const int64_t diag = (int64_t)o_mp - (int64_t)o_ts;
            if (has_first && R.first_ts == c0 && (int64_t)R.first_mp - (int64_t)R.first_ts == diag) {
                // the chunk's first match is the continuation of the open one
                if (R.flags & LZC_FIRST_OPEN) continue;                  // still open: the whole chunk lies inside the match (it wrote nothing)
                const uint32_t e = R.first_ts + R.first_len;
                put_match(o_ts, o_mp, e - o_ts, o_predb);
                open = false;
                copy_bytes(cb + R.lit0, R.bytes - R.lit0);
                take_end();
                continue;
            }
            // otherwise the open match has to end inside this chunk, before anything the chunk found
            const uint32_t e_max = o_ts + lzc_min(n - o_ts, m - o_mp);
            const uint32_t stop = lzc_min(has_first ? R.first_p : c1, e_max);
            uint32_t e = c0;
            if (stop > c0) e = c0 + a.lcp_fwd(c0, (uint32_t)((int64_t)c0 + diag), stop - c0);
            if (e == stop && e != e_max) {
                if (!has_first && stop == c1) { clear(c0, c1); continue; }   // the chunk saw nothing (index entries dropped): still open
                fail = 2; break;
            }
            put_match(o_ts, o_mp, e - o_ts, o_predb);
            open = false; pos = e; np = 0; pred = o_mp + (e - o_ts); cov_diag = diag; cov_start = o_ts;
        }
        if (pos >= c1 && c1 > c0) { clear(lzc_max(c0, cov_start), c1); continue; }   // the whole chunk lies inside an emitted match
        if (pos == c0 && np > 0 && (R.flags & LZC_SENS)) {
            // A probe before the chunk's first match failed only because the chunk did not know the literals pending before it:
            // redo those probes sequentially with the true count.  Either none succeeds (the chunk's parse stands) or the first
            // success is emitted here and the chunk continues behind it like behind any match that crossed its start.
            const uint32_t kl = mml - 3u;
            const uint32_t lim = has_first ? R.first_p : R.end_i;
            for (uint32_t p = c0; p < lim && p + kl < n; ++p) {
                const uint64_t x = a.twin(p) >> (64 - 2 * kl);
                const uint32_t h = (uint32_t)lzc_murmur64(x) & a.mask;
                uint32_t hp, b, f; bool npl;
                if (!lzc_best_match_full(a, h, x, p, np + (p - c0), kl, mml, hp, b, f, npl)) continue;
                put_lit_text(c0, p);
                np += p - c0; pred += p - c0;
                const uint32_t ts = p - b, mp = hp - b, len = b + f;
                o -= b; np -= b; pred -= b;
                if (mp == pred) bang(mp);
                put_match(ts, mp, len, pred);
                clear(ts, p);                                            // literals the match took back
                pos = ts + len; np = 0; pred = mp + len; cov_diag = (int64_t)mp - (int64_t)ts; cov_start = ts;
                break;
            }
            if (pos >= c1 && c1 > c0) { clear(lzc_max(c0, cov_start), c1); continue; }
        }
        const uint32_t from = pos;                                       // >= c0
        if (from > c0) {
            clear(lzc_max(c0, cov_start), from);                         // whatever the chunk saw there lies inside a match
            // np == 0 here.  The chunk's first match may be the covered piece itself (same diagonal, same end)
            if (has_first && R.first_ts == c0 && (int64_t)R.first_mp - (int64_t)R.first_ts == cov_diag && !(R.flags & LZC_FIRST_OPEN)
                && R.first_ts + R.first_len == from) {
                copy_bytes(cb + R.lit0, R.bytes - R.lit0);
                take_end();
                continue;
            }
            if (has_first && (R.first_p < from || (R.flags & LZC_FIRST_MULTI))) { fail = 3; break; }
        } else if (np > 0 && (R.flags & LZC_FIRST_MULTI) && ((R.flags & LZC_FIRST_BLIM) || (R.flags & LZC_SENS))) { fail = 4; break; }

        if (!has_first) {                                                // literals only
            if (from == c0) copy_bytes(cb, R.bytes); else put_lit_text(from, R.end_i);
            const uint32_t cnt = R.end_i - from;
            np += cnt; pred += cnt; pos = R.end_i;
            continue;
        }
        uint32_t ts = R.first_ts, mp = R.first_mp, len = R.first_len;
        if (from > ts) {                                                 // the backward extension stops at the covered range
            const uint32_t adj = from - ts; ts += adj; mp += adj; len -= adj;
            if (!(len > mml)) { fail = 5; break; }                   // ... and the match may not survive that
        }
        if (from == c0) copy_bytes(cb, R.lit0); else put_lit_text(from, ts);
        { const uint32_t cnt = ts - from; np += cnt; pred += cnt; }
        if (from == c0 && (R.flags & LZC_FIRST_BLIM) && np > 0) {
            // the backward extension continues into the literals before the chunk (lz_diff.cpp:306-309)
            uint32_t j = 0;
            while (j < np && mp > 0 && ts > 0 && a.tsym(ts - 1) == a.rsym(mp - 1)) { --ts; --mp; ++len; ++j; }
            o -= j; np -= j; pred -= j;
            clear(ts, ts + j);
        }
        if (mp == pred) bang(mp);
        if (R.flags & LZC_FIRST_OPEN) { open = true; o_ts = ts; o_mp = mp; o_predb = pred; continue; }
        put_match(ts, mp, len, pred);
        copy_bytes(cb + R.lit0, R.bytes - R.lit0);
        take_end();
    }
    if (open) fail = 6;                                 // cannot happen (the last chunk never leaves a match open)
    if (fail) return -10 - fail;
    if (ovf) return -2;
    return COSTS ? 0 : (int64_t)o;
}
