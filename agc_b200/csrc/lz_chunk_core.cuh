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

// view of one (text segment, reference) pair for a single thread
template <bool STAGED>
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
    LZC_HD uint32_t slot(uint32_t s) const
    {
        if (is_short) { const uint32_t v = STAGED ? lzc_lds16(ht_s + 2u * s) : (uint32_t)((const uint16_t*)ht)[s]; return v == 0xffffu ? AGC_EMPTY32 : v; }
        return STAGED ? lzc_lds32(ht_s + 4u * s) : ((const uint32_t*)ht)[s];
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
template <bool STAGED>
LZC_HD bool lzc_best_match_full(const LzcView<STAGED>& a, uint32_t h, uint64_t x, uint32_t p, uint32_t np, uint32_t kl, uint32_t mml,
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

// ------------------------------------------------------------------------------------------------ phase 1: one thread, one chunk
template <bool STAGED>
LZC_HD void lzc_parse_chunk(const LzcView<STAGED>& a, uint32_t c0, uint32_t c1, uint32_t mml, uint8_t* __restrict__ out, LzcRec& R, bool active = true)
{
    const uint32_t kl = mml - 3u, n = a.n, m = a.m;
    uint32_t i = c0, np = 0, pred = 0, olen = 0, flags = 0;
    bool have_first = false, neq_checked = (n != m), ended_open = false;
    int32_t end_diag = 0;
    R.lit0 = 0; R.first_p = R.first_ts = R.first_mp = R.first_len = 0; R.open_ts = R.open_mp = R.open_predb = 0;
    // Warp-cooperative state machine.  A lane is EXTENDING (stepping a forward extension, 64 bases per step -- almost all of the
    // work), in NEED of the rare path (resolve a finished extension: backward part, decision, token output; then probe positions
    // until the next extension starts) or DONE.  The lanes of a warp are all at different places of their chunks, so left alone
    // every iteration would pay for both paths with a handful of lanes active in each.  Instead the warp votes: extension steps run
    // while at least half of the lanes extend; lanes that need the rare path park until 16 of them wait (or nobody extends) and are
    // then served together.  Both paths run with most lanes active.  (Host build: one lane, the votes are trivial.)
    enum { ST_NEED = 0, ST_EXT = 1, ST_DONE = 2 };
    int st = active ? ST_NEED : ST_DONE;
    bool pending = false;                                 // a finished extension waits for its resolve
    uint32_t e_hp = 0, e_off = 0, e_lim = 0, e_maxlen = 0, e_f = 0;
    for (;;) {
        const uint32_t m_ext = LZC_BALLOT(st == ST_EXT), m_need = LZC_BALLOT(st == ST_NEED);
        if (!(m_ext | m_need)) break;
        if (m_ext != 0 && LZC_POPC(m_need) < 16) {
            if (st == ST_EXT) {
                // matching_length(text + i, ref + e_hp, e_lim), two 32-base windows per step
                const uint64_t x0 = a.twin((int64_t)i + e_off) ^ a.rwin((int64_t)e_hp + e_off);
                const uint64_t x1 = a.twin((int64_t)i + e_off + 32) ^ a.rwin((int64_t)e_hp + e_off + 32);
                bool fin = true;
                if (x0) e_f = e_off + (lzc_clz64(x0) >> 1);
                else if (x1) e_f = e_off + 32u + (lzc_clz64(x1) >> 1);
                else { e_off += 64; if (e_off < e_lim) fin = false; else e_f = e_lim; }
                if (fin) { if (e_f > e_lim) e_f = e_lim; pending = true; st = ST_NEED; }
            }
            continue;
        }
        while (st == ST_NEED) {
            bool ok = false, open = false;
            uint32_t hp = 0, b = 0, f = 0;
            bool np_limited = false, multi = false;
            if (pending) {
                pending = false;
                hp = e_hp; f = e_f;
                const uint32_t lim = lzc_min(np, hp);
                b = lim ? a.lcp_bwd(i, hp, lim) : 0u;
                np_limited = (b == np && np < hp);
                ok = b + f > mml;
                open = ok && f == e_lim && e_lim < e_maxlen;
            } else {
                if (!(i < c1 && i + kl < n)) { st = ST_DONE; break; }
                const uint64_t x = a.twin(i) >> (64 - 2 * kl);
                const uint32_t h = (uint32_t)lzc_murmur64(x) & a.mask;
                // candidates: slots until the first empty one whose key equals the text's (find_best_match's "f_len >= key_len")
                uint32_t ncand = 0, hp0 = 0;
                for (uint32_t t = 0; t < 64; ++t) {
                    const uint32_t v = a.slot((h + t) & a.mask);
                    if (v == AGC_EMPTY32) break;
                    if ((a.rwin(v * 4u) >> (64 - 2 * kl)) == x) { if (!ncand) hp0 = v * 4u; ++ncand; if (ncand > 1) break; }
                }
                if (ncand == 1) {
                    e_hp = hp0; e_off = 0;
                    e_maxlen = lzc_min(n - i, m - hp0);
                    e_lim = lzc_min(e_maxlen, lzc_max(c1 - i, mml + 1u));   // cap: enough to decide "b + f > min_match_len" whatever b is
                    st = ST_EXT;
                    break;
                }
                if (ncand > 1) {
                    multi = true;
                    ok = lzc_best_match_full<STAGED>(a, h, x, i, np, kl, mml, hp, b, f, np_limited);
                }
            }
            if (!ok) {
                // a candidate that failed although its backward extension was cut short by no_prev_literals: with the true (possibly
                // larger) count it might have succeeded -- only matters before the chunk's first match
                if (!have_first && np_limited) flags |= LZC_SENS;
                if (!neq_checked) { neq_checked = true; if (a.tsym(i) != a.rsym(i)) flags |= LZC_NEQ; }
                ++i; ++np;
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
                for (uint32_t j = 0; j < np; ++j) out[olen + j] = (uint8_t)('A' + a.tsym(ts - np + j));
                olen += np;
                if (!(n == m && ts == c0 && mp == c0)) { if (!neq_checked && np) { neq_checked = true; if (a.tsym(ts - np) != a.rsym(ts - np)) flags |= LZC_NEQ; } }
            } else {
                const uint32_t pred_now = pred + np;
                const bool bang = (mp == pred_now);
                for (uint32_t j = 0; j < np; ++j) {
                    const uint32_t q = ts - np + j, sy = a.tsym(q);
                    uint8_t ch = (uint8_t)('A' + sy);
                    const uint32_t d = np - j;                               // distance back from the match (lz_diff.cpp:772)
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
            pred = mp + len; i = ts + len; np = 0; end_diag = (int32_t)mp - (int32_t)ts;
            if (open) { ended_open = true; st = ST_DONE; }
        }
    }
    if (!active) return;
    // tail of the text (positions the sequential loop never probes) and literals pending at the chunk's end: raw letters
    if (!ended_open && i < c1 && i + kl >= n) { const uint32_t e = lzc_min(c1, n); np += e - i; i = e; }
    if (np) {
        if (!neq_checked) { neq_checked = true; if (a.tsym(i - np) != a.rsym(i - np)) flags |= LZC_NEQ; }
        for (uint32_t j = 0; j < np; ++j) out[olen + j] = (uint8_t)('A' + a.tsym(i - np + j));
        olen += np;
        if (!have_first) R.lit0 = np;
    }
    // "this chunk equals the reference at the same positions": one diagonal-0 match from the chunk's first position to its end
    if (n == m && have_first && R.lit0 == 0 && R.first_ts == c0 && R.first_mp == c0 && i >= lzc_min(c1, n) && olen == 0 && !(flags & LZC_END_OPEN))
        flags |= LZC_EQ;
    R.flags = flags; R.bytes = olen; R.end_i = i; R.end_np = np; R.end_pred = pred + np; R.end_diag = end_diag; R.pad[0] = R.pad[1] = 0;
}


// ------------------------------------------------------------------------------------------------ phase 2: one thread, one segment
// rec / cslab: records and slabs of the segment's chunks (chunk k at rec[k], cslab + k * LZC_CSLAB).  Returns the number of bytes
// written to out, -1 when the segment has to go to the sequential kernel, -2 when out_cap was too small.
template <class View>
LZC_HD int64_t lzc_stitch_segment(const View& a, const LzcReq& q, uint32_t mml, const LzcRec* rec, const uint8_t* cslab, uint8_t* out, uint32_t cap)
{
    const uint32_t n = q.n, m = a.m;
    const uint32_t nch = q.nch;
    int fail = 0; bool ovf = false;          // fail: why the segment goes to the sequential kernel (diagnostic codes)

    // equal sequences (lz_diff.cpp:678-680): every chunk is one diagonal-0 match over its whole range
    if (n == m) {
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
    int64_t cov_diag = 0;
    bool open = false;
    uint32_t o_ts = 0, o_mp = 0, o_predb = 0;

    auto put_lit_text = [&](uint32_t from, uint32_t to) {          // raw literals of text positions [from, to)
        for (uint32_t p = from; p < to; ++p) { if (o < cap) out[o] = (uint8_t)('A' + a.tsym(p)); else ovf = true; ++o; }
    };
    auto copy_bytes = [&](const uint8_t* src, uint32_t cnt) {
        if ((uint64_t)o + cnt > cap) { ovf = true; o += cnt; return; }
        for (uint32_t j = 0; j < cnt; ++j) out[o + j] = src[j];
        o += cnt;
    };
    auto put_match = [&](uint32_t ts, uint32_t mp, uint32_t len, uint32_t predb) {
        uint8_t buf[24];
        const bool to_end = (ts + len == n) && (mp + len == m);
        const uint32_t L = lzc_put_match(buf, (int64_t)(int)mp - (int64_t)(int)predb, !to_end, len - mml);
        copy_bytes(buf, L);
    };
    // the '!' rewrite (lz_diff.cpp:769-779) over the literal run that ends the output
    auto bang = [&](uint32_t mp) {
        if (ovf) return;
        for (uint32_t d = 1; d < o && d < mp; ++d) {
            const uint8_t ch = out[o - d];
            if (ch < 'A' || ch > 'Z') break;
            if ((uint32_t)(ch - 'A') == a.rsym(mp - d)) out[o - d] = '!';
        }
    };

    for (uint32_t k = 0; k < nch && !fail; ++k) {
        const LzcRec R = rec[k];
        const uint32_t c0 = k * LZC_CHUNK, c1 = lzc_min(n, c0 + LZC_CHUNK);
        const uint8_t* cb = cslab + (uint64_t)k * LZC_CSLAB;
        const bool has_first = (R.flags & LZC_HAS_FIRST) != 0;
        // state after the chunk's own last token
        auto take_end = [&]() {
            if (R.flags & LZC_END_OPEN) { open = true; o_ts = R.open_ts; o_mp = R.open_mp; o_predb = R.open_predb; }
            else { pos = R.end_i; np = R.end_np; pred = R.end_pred; cov_diag = R.end_diag; }
        };

        if (open) {
            const int64_t diag = (int64_t)o_mp - (int64_t)o_ts;
            if (has_first && R.first_ts == c0 && (int64_t)R.first_mp - (int64_t)R.first_ts == diag) {
                // the chunk's first match is the continuation of the open one
                if (R.flags & LZC_FIRST_OPEN) continue;                  // still open: the whole chunk lies inside the match
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
                if (!has_first && stop == c1) continue;                  // the chunk saw nothing (index entries dropped): still open
                fail = 2; break;
            }
            put_match(o_ts, o_mp, e - o_ts, o_predb);
            open = false; pos = e; np = 0; pred = o_mp + (e - o_ts); cov_diag = diag;
        }
        if (pos >= c1 && c1 > c0) continue;                              // the whole chunk lies inside an emitted match
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
                pos = ts + len; np = 0; pred = mp + len; cov_diag = (int64_t)mp - (int64_t)ts;
                break;
            }
            if (pos >= c1 && c1 > c0) continue;
        }
        const uint32_t from = pos;                                       // >= c0
        if (from > c0) {
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
    return (int64_t)o;
}
