// kernels_lz.cu -- reference-segment store, LZ index build and the LZ-diff kernels (sm_100a).
//
// Restates, warp-cooperatively, CLZDiff_V2 of the reference (src/common/lz_diff.{h,cpp}):
//   k_extract_ref / k_index_*  : CLZDiffBase::Prepare + prepare_index + make_index16/32 (lz_diff.cpp:48-149,375-428).
//        Slot layout parity (sequential first-come linear probing with the 64-try drop rule) is kept by inserting in
//        reference order, one warp per reference; the 32 probe slots of one insertion are read in parallel.
//   k_lz<MODE>                 : MODE 0 = Encode (669-798), 1 = Estimate (839-946), 2 = GetCodingCostVector (159-284).
//        One warp parses one segment.  Each round the 32 lanes probe 32 consecutive text positions (hash + chain walk,
//        find_best_match's "f_len >= key_len" filter is an exact code compare); the first position that yields an
//        accepted match ends the round, everything before it is a literal.  Match extension (compare_fwd /
//        matching_length and the backward loop, lz_diff.cpp:287-372) is a warp-wide XOR + clz/ffs over 1024 bases
//        per step on 2-bit packed data.  A CTA handles requests of ONE group: the group's packed reference and its
//        hash table are staged into shared memory once per CTA with TMA bulk copies (cp.async.bulk + mbarrier).
//   Segments holding non-ACGT symbols (N runs, IUPAC) take the same parse over 1-byte symbols (ByteAcc).
//
// Integer/byte work only; the roofline is HBM: B_lz = ceil(n/4) + ceil(m/4) + e bytes per segment.
#include "internal.cuh"
#include "lz_chunk.cuh"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define FULL 0xffffffffu
#define LZ_THREADS 512
#define LZ_STAGE_LIMIT (110u * 1024u)

// ------------------------------------------------------------------------------------------------ hash table view
struct HT {
    const void* tab;
    uint32_t mask;
    uint32_t is_short;
    __device__ __forceinline__ uint32_t get(uint32_t slot) const
    {
        if (is_short) { uint32_t v = ((const uint16_t*)tab)[slot]; return v == 0xffffu ? AGC_EMPTY32 : v; }
        return ((const uint32_t*)tab)[slot];
    }
};

// ------------------------------------------------------------------------------------------------ accessors
struct PackedAcc {
    static constexpr bool BYTES = false;
    const uint64_t* T; int64_t gs; uint32_t n; uint32_t rc;
    const uint64_t* R; uint32_t m;

    __device__ __forceinline__ uint64_t twin(int64_t pos) const
    {
        if (!rc) return agc_win_s(T, gs + pos);
        return ~agc_rev2(agc_win_s(T, gs + (int64_t)n - pos - 32));
    }
    __device__ __forceinline__ uint64_t rwin(int64_t pos) const { return agc_win_s(R, pos); }
    __device__ __forceinline__ bool tcode(uint32_t p, uint32_t kl, uint64_t& x) const { x = twin(p) >> (64 - 2 * kl); return true; }
    __device__ __forceinline__ bool rcode_eq(uint32_t rp, uint64_t x, uint32_t kl) const { return (rwin(rp) >> (64 - 2 * kl)) == x; }
    // cheap reject for the probe loops: indexed reference positions are multiples of 4 bases = whole bytes, so the first
    // min(kl, 16) bases of the key at slot value v are the (unaligned, big-endian) 32 bits at byte v: two word loads + one
    // PRMT.  xq = the same bits of the text key (key_quick).  true = "may be equal" (rcode_eq decides).
    __device__ __forceinline__ uint32_t key_quick(uint64_t x, uint32_t kl) const { return kl >= 16 ? (uint32_t)(x >> (2 * kl - 32)) : (uint32_t)(x << (32 - 2 * kl)); }
    __device__ __forceinline__ bool rcode_quick(uint32_t v, uint32_t xq, uint32_t kl) const
    {
        const uint32_t* w = (const uint32_t*)R + (v >> 2);
        const uint32_t o = v & 3u;
        const uint32_t r = __byte_perm(w[0], w[1], 0x0123u + o * 0x1111u);
        const uint32_t d = r ^ xq;
        return kl >= 16 ? d == 0 : (d >> (32 - 2 * kl)) == 0;
    }
    __device__ __forceinline__ uint32_t tsym(uint32_t q) const { return (uint32_t)(twin(q) >> 62); }
    __device__ __forceinline__ uint32_t rsym(uint32_t q) const { return (uint32_t)(rwin(q) >> 62); }
    __device__ __forceinline__ uint32_t nrun(uint32_t, uint32_t, uint32_t) const { return 0; }

    // matching_length(text+tp, ref+rp, maxlen): 32 lanes x 32 bases per step
    __device__ uint32_t lcp_fwd(uint32_t tp, uint32_t rp, uint32_t maxlen, uint32_t lane) const
    {
        for (uint32_t base = 0; base < maxlen; base += 1024) {
            uint32_t off = base + 32 * lane, idx = 0;
            bool stop = true;
            if (off < maxlen) {
                uint64_t xr = twin((int64_t)tp + off) ^ rwin((int64_t)rp + off);
                idx = xr ? ((uint32_t)__clzll((long long)xr) >> 1) : 32u;
                uint32_t rem = maxlen - off;
                if (idx > rem) idx = rem;
                stop = idx < 32;
            }
            uint32_t mk = __ballot_sync(FULL, stop);
            if (mk) { uint32_t L = __ffs(mk) - 1; return base + 32 * L + __shfl_sync(FULL, idx, L); }
        }
        return maxlen;
    }
    // backward extension: text[tp-1-j] == ref[rp-1-j], j < lim
    __device__ uint32_t lcp_bwd(uint32_t tp, uint32_t rp, uint32_t lim, uint32_t lane) const
    {
        for (uint32_t base = 0; base < lim; base += 1024) {
            uint32_t off = base + 32 * lane, idx = 0;
            bool stop = true;
            if (off < lim) {
                uint64_t xr = twin((int64_t)tp - off - 32) ^ rwin((int64_t)rp - off - 32);
                idx = xr ? ((uint32_t)(__ffsll((long long)xr) - 1) >> 1) : 32u;
                uint32_t rem = lim - off;
                if (idx > rem) idx = rem;
                stop = idx < 32;
            }
            uint32_t mk = __ballot_sync(FULL, stop);
            if (mk) { uint32_t L = __ffs(mk) - 1; return base + 32 * L + __shfl_sync(FULL, idx, L); }
        }
        return lim;
    }
};

struct ByteAcc {
    static constexpr bool BYTES = true;
    const uint8_t* T; uint32_t n;
    const uint8_t* R; uint32_t m;        // R has key_len bytes of 31 after m
    __device__ __forceinline__ bool tcode(uint32_t p, uint32_t kl, uint64_t& x) const
    {
        x = 0;
        for (uint32_t i = 0; i < kl; ++i) { uint32_t s = T[p + i]; if (s > 3) return false; x = (x << 2) + s; }
        return true;
    }
    __device__ __forceinline__ bool rcode_eq(uint32_t rp, uint64_t x, uint32_t kl) const
    {
        uint64_t y = 0;
        for (uint32_t i = 0; i < kl; ++i) { uint32_t s = R[rp + i]; if (s > 3) return false; y = (y << 2) + s; }
        return y == x;
    }
    __device__ __forceinline__ uint32_t key_quick(uint64_t, uint32_t) const { return 0; }
    __device__ __forceinline__ bool rcode_quick(uint32_t, uint32_t, uint32_t) const { return true; }
    __device__ __forceinline__ uint32_t tsym(uint32_t q) const { return T[q]; }
    __device__ __forceinline__ uint32_t rsym(uint32_t q) const { return R[q]; }
    // get_Nrun_len (lz_diff.h:122-132)
    __device__ uint32_t nrun(uint32_t p, uint32_t maxlen, uint32_t lane) const
    {
        if (T[p] != 4 || T[p + 1] != 4 || T[p + 2] != 4) return 0;
        for (uint32_t base = 3; ; base += 32) {
            uint32_t q = base + lane;
            bool stop = !(q < maxlen && T[p + q] == 4);
            uint32_t mk = __ballot_sync(FULL, stop);
            if (mk) return base + __ffs(mk) - 1;
        }
    }
    __device__ uint32_t lcp_fwd(uint32_t tp, uint32_t rp, uint32_t maxlen, uint32_t lane) const
    {
        for (uint32_t base = 0; base < maxlen; base += 128) {
            uint32_t off = base + 4 * lane, idx = 0;
            bool stop = true;
            if (off < maxlen) {
                uint32_t lim = maxlen - off; if (lim > 4) lim = 4;
                while (idx < lim && T[tp + off + idx] == R[rp + off + idx]) ++idx;
                stop = idx < 4;
            }
            uint32_t mk = __ballot_sync(FULL, stop);
            if (mk) { uint32_t L = __ffs(mk) - 1; return base + 4 * L + __shfl_sync(FULL, idx, L); }
        }
        return maxlen;
    }
    __device__ uint32_t lcp_bwd(uint32_t tp, uint32_t rp, uint32_t lim, uint32_t lane) const
    {
        for (uint32_t base = 0; base < lim; base += 128) {
            uint32_t off = base + 4 * lane, idx = 0;
            bool stop = true;
            if (off < lim) {
                uint32_t l4 = lim - off; if (l4 > 4) l4 = 4;
                while (idx < l4 && T[tp - 1 - off - idx] == R[rp - 1 - off - idx]) ++idx;
                stop = idx < 4;
            }
            uint32_t mk = __ballot_sync(FULL, stop);
            if (mk) { uint32_t L = __ffs(mk) - 1; return base + 4 * L + __shfl_sync(FULL, idx, L); }
        }
        return lim;
    }
};

// ------------------------------------------------------------------------------------------------ small helpers
__device__ __forceinline__ uint32_t uint_len_v2(uint32_t x)     // lz_diff.h:385-403 (caps at 8)
{
    return x < 10 ? 1 : x < 100 ? 2 : x < 1000 ? 3 : x < 10000 ? 4 : x < 100000 ? 5 : x < 1000000 ? 6 : x < 10000000 ? 7 : 8;
}
__device__ __forceinline__ uint32_t int_len_base(uint32_t x)    // lz_diff.h:179-191
{
    return x < 10 ? 1 : x < 100 ? 2 : x < 1000 ? 3 : x < 10000 ? 4 : x < 100000 ? 5 : x < 1000000 ? 6 : x < 10000000 ? 7
         : x < 100000000 ? 8 : x < 1000000000 ? 9 : 10;
}
// append_int (lz_diff.h:229-262)
// (32-bit arithmetic: every value AGC prints here is a position difference or a length below 2^32; a 64-bit divide costs
// ~10x more instructions on the device)
__device__ __forceinline__ uint32_t put_int(uint8_t* dst, int64_t xs)
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

struct Sink {
    // MODE 0
    uint8_t* out; uint32_t olen; uint32_t cap; uint32_t ovf;
    // MODE 1
    uint32_t est; uint32_t bound;
    // MODE 2
    uint32_t* v; int prefix;
};

template <class A>
__device__ void emit_literals(const A& a, Sink& s, uint32_t start, uint32_t count, bool bang, uint32_t mp, uint32_t lane)
{
    if (count == 0) return;
    if ((uint64_t)s.olen + count > s.cap) { s.ovf = 1; s.olen += count; return; }
    uint32_t d0 = 0xffffffffu;
    if (A::BYTES && bang) {             // the '!' scan stops at the first byte outside 'A'..'Z' (lz_diff.cpp:774-775)
        for (uint32_t j = lane; j < count; j += 32) if (a.tsym(start + j) > 25u) d0 = min(d0, count - j);
#pragma unroll
        for (int o = 16; o; o >>= 1) d0 = min(d0, __shfl_xor_sync(FULL, d0, o));
    }
    for (uint32_t j = lane; j < count; j += 32) {
        uint32_t sy = a.tsym(start + j);
        uint8_t ch = (uint8_t)('A' + sy);
        if (bang) {
            uint32_t d = count - j;      // distance back from the match (i of the reference loop, lz_diff.cpp:772)
            if (d < s.olen + count && d < mp && d < d0 && sy == a.rsym(mp - d)) ch = '!';
        }
        s.out[s.olen + j] = ch;
    }
    s.olen += count;
}

__device__ void emit_match(Sink& s, int64_t dif, bool with_len, uint32_t lenv, uint32_t lane)
{
    uint32_t L = 0;
    if (lane == 0) {
        uint8_t buf[24];
        L = put_int(buf, dif);
        if (with_len) { buf[L++] = ','; L += put_int(buf + L, (int64_t)lenv); }
        buf[L++] = '.';
        if ((uint64_t)s.olen + L <= s.cap) for (uint32_t i = 0; i < L; ++i) s.out[s.olen + i] = buf[i];
    }
    L = __shfl_sync(FULL, L, 0);
    if ((uint64_t)s.olen + L > s.cap) s.ovf = 1;
    s.olen += L;
}

__device__ void emit_nrun(Sink& s, uint32_t nr, uint32_t lane)     // encode_Nrun (lz_diff.h:152-157)
{
    uint32_t L = 0;
    if (lane == 0) {
        uint8_t buf[16];
        buf[L++] = 30; L += put_int(buf + L, (int64_t)nr - 4); buf[L++] = 4;
        if ((uint64_t)s.olen + L <= s.cap) for (uint32_t i = 0; i < L; ++i) s.out[s.olen + i] = buf[i];
    }
    L = __shfl_sync(FULL, L, 0);
    if ((uint64_t)s.olen + L > s.cap) s.ovf = 1;
    s.olen += L;
}

__device__ __forceinline__ void fill_u32(uint32_t* v, uint32_t start, uint32_t count, uint32_t val, uint32_t lane)
{
    for (uint32_t j = lane; j < count; j += 32) v[start + j] = val;
}
__device__ __forceinline__ void span_cost(uint32_t* v, uint32_t start, uint32_t count, uint32_t tc, int prefix, uint32_t lane)
{
    for (uint32_t j = lane; j < count; j += 32) v[start + j] = (prefix ? j == 0 : j == count - 1) ? tc : 0u;
}

// find_best_match16/32 (lz_diff.cpp:287-372) for text position p whose code is x; the probe chain is read 32 slots at a time
template <class A>
__device__ bool eval_chain(const A& a, const HT& ht, uint32_t hpos, uint64_t x, uint32_t p, uint32_t npl, uint32_t kl,
                           uint32_t mml, uint32_t lane, uint32_t& o_rp, uint32_t& o_b, uint32_t& o_f)
{
    uint32_t best_b = 0, best_f = 0, best_rp = 0, mtu = mml;
    const uint32_t xq = a.key_quick(x, kl);
    for (uint32_t t0 = 0; t0 < 64; t0 += 32) {
        uint32_t v = ht.get((hpos + t0 + lane) & ht.mask);
        uint32_t em = __ballot_sync(FULL, v == AGC_EMPTY32);
        uint32_t cnt = em ? (uint32_t)__ffs(em) - 1 : 32u;
        bool cand = lane < cnt && a.rcode_quick(v, xq, kl) && a.rcode_eq(v * 4u, x, kl);
        uint32_t cm = __ballot_sync(FULL, cand);
        while (cm) {
            uint32_t l = __ffs(cm) - 1; cm &= cm - 1;
            uint32_t hp = __shfl_sync(FULL, v, l) * 4u;
            uint32_t maxlen = min(a.n - p, a.m - hp);
            uint32_t f = a.lcp_fwd(p, hp, maxlen, lane);
            uint32_t lim = min(npl, hp);
            uint32_t b = lim ? a.lcp_bwd(p, hp, lim, lane) : 0u;
            if (b + f > mtu) { best_b = b; best_f = f; best_rp = hp; mtu = b + f; }
        }
        if (em) break;
    }
    o_rp = best_rp; o_b = best_b; o_f = best_f;
    return best_b + best_f >= mml;
}

// The greedy parse shared by Encode / Estimate / GetCodingCostVector.
template <class A, int MODE>
__device__ void lz_parse(const A& a, const HT& ht, uint32_t mml, uint32_t lane, Sink& s)
{
    const uint32_t kl = mml - 3u, n = a.n, m = a.m;
    if (MODE != 2 && n == m) {                       // equal sequences (lz_diff.cpp:678-680, 848-850)
        if (a.lcp_fwd(0, 0, n, lane) == n) { s.olen = 0; s.est = 0; return; }
    }
    uint32_t i = 0, pred = 0, np = 0;                // np = no_prev_literals = length of the pending literal run ending at i
    bool lit_streak = false;                         // the previous round found no match
    while (i + kl < n) {
        uint32_t p = i + lane;
        bool active = p + kl < n;
        uint32_t ev = 0; uint64_t x = 0;
        if (active && !a.tcode(p, kl, x)) ev = 1;
        uint32_t nact = __popc(__ballot_sync(FULL, active));
        uint32_t c = 0;                              // lanes [0,c) of this round are already committed as literals
        bool restart = false;
        // The probe walk of a round costs as many steps as the longest chain among the probing lanes, and after a mismatch the
        // next match starts within a few positions (the first one that is a multiple of 4 on the reference): the first 8
        // positions are probed on their own, the other 24 only if none of them ends the round.
        // (inside a stretch without matches -- a text compared with the wrong candidate, a novel insertion -- the 8-lane phase only adds
        // a round trip: after a round that found nothing all 32 lanes probe at once)
        const uint32_t ngrp = lit_streak ? 1u : 2u;
        for (uint32_t grp = 0; grp < ngrp && !restart; ++grp) {
        const bool mine = lit_streak ? true : (grp == 0 ? lane < 8 : lane >= 8);
        if (grp == 1 && nact <= 8) break;
        if (active && mine && ev == 0) {
            uint32_t hp = (uint32_t)agc_murmur64(x) & ht.mask;
            const uint32_t xq = a.key_quick(x, kl);
            // four slots per step: their loads are independent (one round trip to the table -- L2 for groups too large for shared
            // memory -- instead of four), the tests keep the slot order
            bool done = false;
            for (uint32_t t = 0; t < 64 && !done; t += 4) {
                const uint32_t v0 = ht.get((hp + t) & ht.mask), v1 = ht.get((hp + t + 1) & ht.mask);
                const uint32_t v2 = ht.get((hp + t + 2) & ht.mask), v3 = ht.get((hp + t + 3) & ht.mask);
                const uint32_t vv[4] = { v0, v1, v2, v3 };
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (vv[k] == AGC_EMPTY32) { done = true; break; }
                    if (a.rcode_quick(vv[k], xq, kl) && a.rcode_eq(vv[k] * 4u, x, kl)) { ev = 2; done = true; break; }
                }
            }
        }
        uint32_t evmask = __ballot_sync(FULL, mine && ev != 0);
        while (evmask) {
            uint32_t L = __ffs(evmask) - 1; evmask &= evmask - 1;
            uint32_t evL = __shfl_sync(FULL, ev, L);
            uint64_t xL = __shfl_sync(FULL, x, L);
            uint32_t pL = i + L, lits = L - c;
            if (MODE == 1) {                         // bound test at the top of every reference iteration (lz_diff.cpp:868)
                if (s.est > s.bound) return;
                if ((uint64_t)s.est + lits > s.bound) { s.est = s.bound + 1; return; }
                s.est += lits;
            }
            np += lits; pred += lits; c = L;
            if (evL == 1) {
                uint32_t nr = a.nrun(pL, n - pL, lane);
                if (nr >= 4) {
                    if (MODE == 0) { emit_literals(a, s, pL - np, np, false, 0, lane); emit_nrun(s, nr, lane); }
                    if (MODE == 1) s.est += 2 + uint_len_v2(nr);
                    if (MODE == 2) { fill_u32(s.v, pL - np, np, 1, lane); span_cost(s.v, pL, nr, 2 + int_len_base(nr - 4), s.prefix, lane); }
                    i = pL + nr; np = 0; restart = true;
                    break;
                }
                if (MODE == 1) ++s.est;
                ++np; ++pred; c = L + 1;
                continue;
            }
            uint32_t rp, b, f;
            if (!eval_chain(a, ht, (uint32_t)agc_murmur64(xL) & ht.mask, xL, pL, np, kl, mml, lane, rp, b, f)) {
                if (MODE == 1) ++s.est;
                ++np; ++pred; c = L + 1;
                continue;
            }
            if (MODE == 1) {                         // Estimate does not rewind len_bck (reference quirk, lz_diff.cpp:926-935)
                bool to_end = (pL + b + f == n) && (rp + b + f == m);
                int dif = (int)rp - (int)pred;
                uint32_t r = dif >= 0 ? uint_len_v2((uint32_t)dif) : 1 + uint_len_v2((uint32_t)-dif);
                if (!to_end) r += 1 + uint_len_v2(b + f - mml);
                s.est += r + 1;
                pred = rp + b + f; i = pL + b + f; np = 0; restart = true;
                break;
            }
            uint32_t mp = rp - b, ts = pL - b, len = b + f;      // rewind (lz_diff.cpp:757-767)
            pred -= b; np -= b;
            if (MODE == 0) {
                emit_literals(a, s, ts - np, np, mp == pred, mp, lane);
                bool to_end = (ts + len == n) && (mp + len == m);
                emit_match(s, (int64_t)(int)mp - (int64_t)(int)pred, !to_end, len - mml, lane);
            }
            if (MODE == 2) {
                fill_u32(s.v, ts - np, np, 1, lane);
                int dif = (int)mp - (int)pred;
                uint32_t tc = (dif >= 0 ? int_len_base((uint32_t)dif) : int_len_base((uint32_t)-dif) + 1) + int_len_base(len - mml) + 2;
                span_cost(s.v, ts, len, tc, s.prefix, lane);
            }
            pred = mp + len; i = ts + len; np = 0; restart = true;
            break;
        }
        }
        lit_streak = !restart;
        if (!restart) {
            uint32_t r = nact - c;
            if (MODE == 1 && r) {
                if (s.est > s.bound) return;
                if ((uint64_t)s.est + (r - 1) > s.bound) { s.est = s.bound + 1; return; }
                s.est += r;
            }
            np += r; pred += r; i += nact;
        }
    }
    if (MODE == 0) { np += n - i; emit_literals(a, s, n - np, np, false, 0, lane); }
    if (MODE == 1) s.est += n - i;                   // u32 wrap when the un-rewound i overshoots, as in the reference
    if (MODE == 2) { np += n - i; fill_u32(s.v, n - np, np, 1, lane); }
}

// ------------------------------------------------------------------------------------------------ TMA bulk staging
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
:: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(smem_u32(bar)), "r"(phase) : "memory");
}

// ------------------------------------------------------------------------------------------------ LZ kernels
template <int MODE>
__global__ void __launch_bounds__(LZ_THREADS, 2) k_lz_packed(
    const uint64_t* __restrict__ P, const GroupRefDev* __restrict__ groups, const LzReqDev* __restrict__ reqs,
    const LzUnit* __restrict__ units, uint32_t mml, uint32_t stage_limit, uint8_t* __restrict__ slab,
    uint32_t* __restrict__ res, uint32_t* __restrict__ costv, int prefix, uint32_t* __restrict__ err)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t next_req;
    const LzUnit u = units[blockIdx.x];
    const GroupRefDev g = groups[u.group];
    const uint32_t ht_bytes = g.ht_size * ((g.flags & GRF_SHORT) ? 2u : 4u);
    const bool stage = (g.packed_bytes + ht_bytes <= stage_limit);
    const uint8_t* refp = g.packed;
    const void* htp = g.ht;
    if (threadIdx.x == 0) { next_req = 0; if (stage) mbar_init(&bar, 1); }
    __syncthreads();
    if (stage) {
        if (threadIdx.x == 0) {
            mbar_expect_tx(&bar, g.packed_bytes + ht_bytes);
            bulk_g2s(smem, g.packed, g.packed_bytes, &bar);
            bulk_g2s(smem + g.packed_bytes, g.ht, ht_bytes, &bar);
        }
        mbar_wait(&bar, 0);
        refp = smem; htp = smem + g.packed_bytes;
    }
    const uint32_t lane = threadIdx.x & 31;
    HT ht; ht.tab = htp; ht.mask = g.ht_size - 1; ht.is_short = g.flags & GRF_SHORT;
    while (true) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(&next_req, 1u);
        r = __shfl_sync(FULL, r, 0);
        if (r >= u.count) break;
        const LzReqDev q = reqs[u.first + r];
        if (lane == 0) {                                 // pull the whole packed segment into L2 while the parse works on its head
            const uint64_t b0 = (q.gstart >> 2) & ~(uint64_t)15;
            const uint32_t nb = (uint32_t)(((q.gstart + q.n + 3) >> 2) - b0 + 15) & ~15u;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"((const uint8_t*)P + b0), "r"(nb) : "memory");
        }
        PackedAcc a; a.T = P; a.gs = (int64_t)q.gstart; a.n = q.n; a.rc = q.is_rc; a.R = (const uint64_t*)refp; a.m = g.m;
        Sink s; s.out = slab + q.out_off; s.olen = 0; s.cap = q.out_cap; s.ovf = 0; s.est = 0; s.bound = q.bound;
        s.v = costv + q.out_off; s.prefix = MODE == 2 ? (int)q.bound : prefix;       // cost vectors: prefix_costs travels per request
        lz_parse<PackedAcc, MODE>(a, ht, mml, lane, s);
        if (lane == 0) {
            res[q.orig] = MODE == 1 ? s.est : s.olen;
            if (s.ovf) atomicOr(err, 1u);
        }
    }
}

// byte-symbol variant: one warp per request; text/reference symbols and the index live in global memory
struct ByteReq { const uint8_t* text; const uint8_t* ref; const void* ht; uint32_t n, m, ht_size, is_short, bound, orig; uint64_t out_off; uint32_t out_cap, pad; };

template <int MODE>
__global__ void __launch_bounds__(128) k_lz_bytes(const ByteReq* __restrict__ reqs, uint32_t n_req, uint32_t mml,
                                                  uint8_t* __restrict__ slab, uint32_t* __restrict__ res,
                                                  uint32_t* __restrict__ costv, int prefix, uint32_t* __restrict__ err)
{
    uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= n_req) return;
    const ByteReq q = reqs[w];
    ByteAcc a; a.T = q.text; a.n = q.n; a.R = q.ref; a.m = q.m;
    HT ht; ht.tab = q.ht; ht.mask = q.ht_size - 1; ht.is_short = q.is_short;
    Sink s; s.out = slab + q.out_off; s.olen = 0; s.cap = q.out_cap; s.ovf = 0; s.est = 0; s.bound = q.bound;
    s.v = costv + q.out_off; s.prefix = MODE == 2 ? (int)q.bound : prefix;
    lz_parse<ByteAcc, MODE>(a, ht, mml, lane, s);
    if (lane == 0) { res[q.orig] = MODE == 1 ? s.est : s.olen; if (s.ovf) atomicOr(err, 1u); }
}

// ------------------------------------------------------------------------------------------------ reference store
struct RefJob { uint64_t gstart; uint32_t n, is_rc; uint8_t* packed; uint8_t* codes; void* ht; uint32_t group, ht_cap; };

// copy a (possibly reverse-complemented) segment into a byte-aligned packed reference; one thread per 32-base word
__global__ void k_extract_ref(const uint64_t* __restrict__ P, const RefJob* __restrict__ jobs)
{
    const RefJob j = jobs[blockIdx.y];
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nwords = (j.n + 31) / 32;
    if (w >= nwords) return;
    PackedAcc a; a.T = P; a.gs = (int64_t)j.gstart; a.n = j.n; a.rc = j.is_rc; a.R = nullptr; a.m = 0;
    uint64_t v = a.twin((int64_t)w * 32);
    uint32_t rem = j.n - w * 32;
    if (rem < 32) v &= ~0ULL << (64 - 2 * rem);
    ((uint64_t*)j.packed)[w] = agc_be64(v);
}

__device__ __forceinline__ uint64_t ht_size_for(uint64_t cnt)     // lz_diff.cpp:117-127
{
    uint64_t hs = (uint64_t)((double)cnt / 0.7);
    while (hs & (hs - 1)) hs &= hs - 1;
    hs <<= 1;
    if (hs < 8) hs = 8;
    return hs;
}

// make_index16/32 (lz_diff.cpp:375-428): one warp per reference, insertions in reference order.
template <bool BYTES>
__global__ void __launch_bounds__(128) k_index(const RefJob* __restrict__ jobs, uint32_t n_jobs, GroupRefDev* __restrict__ groups, uint32_t mml)
{
    uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= n_jobs) return;
    const RefJob j = jobs[w];
    if (BYTES != (j.codes != nullptr)) return;
    const uint32_t kl = mml - 3u, m = j.n;
    const bool is_short = (m / 4u) < 65535u;
    // number of indexable positions (lz_diff.cpp:87-104)
    uint32_t cnt = 0;
    if (!BYTES) cnt = m >= kl ? (m - kl) / 4u + 1u : 0u;
    else {
        for (uint32_t i = 4 * lane; i + kl <= m; i += 128) {
            bool ok = true;
            for (uint32_t t = 0; t < kl; ++t) if (j.codes[i + t] > 3) { ok = false; break; }
            cnt += ok;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
    }
    const uint32_t hs = (uint32_t)ht_size_for(cnt), mask = hs - 1;
    volatile uint16_t* t16 = (volatile uint16_t*)j.ht;
    volatile uint32_t* t32 = (volatile uint32_t*)j.ht;
    const uint64_t* R = (const uint64_t*)j.packed;
    for (uint32_t i0 = 0; i0 < m; i0 += 128) {
        uint32_t i = i0 + 4 * lane;
        bool valid = i + kl <= m;
        uint64_t x = 0;
        if (valid) {
            if (!BYTES) x = agc_win(R, i) >> (64 - 2 * kl);
            else for (uint32_t t = 0; t < kl; ++t) { uint32_t sy = j.codes[i + t]; if (sy > 3) { valid = false; break; } x = (x << 2) + sy; }
        }
        uint32_t hp = (uint32_t)agc_murmur64(x) & mask;
        uint32_t vm = __ballot_sync(FULL, valid);
        while (vm) {
            uint32_t l = __ffs(vm) - 1; vm &= vm - 1;
            uint32_t pos = __shfl_sync(FULL, hp, l);
            uint32_t val = (i0 >> 2) + l;
            for (uint32_t t0 = 0; t0 < 64; t0 += 32) {
                uint32_t slot = (pos + t0 + lane) & mask;
                uint32_t cur = is_short ? (uint32_t)t16[slot] : t32[slot];
                bool empty = is_short ? cur == 0xffffu : cur == AGC_EMPTY32;
                uint32_t em = __ballot_sync(FULL, empty);
                if (em) {
                    if (lane == (uint32_t)__ffs(em) - 1) { if (is_short) t16[slot] = (uint16_t)val; else t32[slot] = val; }
                    break;
                }
            }
            __syncwarp();
        }
    }
    if (lane == 0) {
        GroupRefDev g;
        g.packed = j.packed; g.ht = j.ht; g.codes = j.codes; g.m = m; g.ht_size = hs;
        g.flags = GRF_PRESENT | (is_short ? GRF_SHORT : 0u) | (BYTES ? GRF_DIRTY : 0u);
        g.packed_bytes = ((m + 3) / 4 + 15) / 16 * 16 + 32;
        groups[j.group] = g;
    }
}

// make_index16/32 for references without non-ACGT symbols, all insertions of one reference in parallel (one CTA per reference).
// The stored value of a key IS its insertion rank (ref_pos / 4), and sequential first-come linear probing has one defining
// property: key k sits in the first slot from its home that no key j < k occupies (64 tries, else it is dropped).  So the table
// is built with atomicMin on the slot values: a key that meets a larger value takes the slot and carries the displaced key on
// (its try count = its displacement from its own home), a key that meets a smaller value moves on.  The fixed point satisfies
// the defining property for every key, hence equals the sequential layout (tests: test_index_layout_parity).
// Tables of up to LZ_IDX_SMEM_SLOTS slots are built in shared memory and written out once (u16 or u32 slots).
#define LZ_IDX_THREADS 512
#define LZ_IDX_SMEM_SLOTS 32768u
__global__ void __launch_bounds__(LZ_IDX_THREADS) k_index_par(const RefJob* __restrict__ jobs, GroupRefDev* __restrict__ groups, uint32_t mml,
                                                              uint32_t* __restrict__ gtab_all)
{
    extern __shared__ uint32_t s_tab[];
    const RefJob j = jobs[blockIdx.x];
    if (j.codes != nullptr) return;                               // references with non-ACGT symbols: k_index<true>
    const uint32_t kl = mml - 3u, m = j.n;
    const bool is_short = (m / 4u) < 65535u;
    const uint32_t cnt = m >= kl ? (m - kl) / 4u + 1u : 0u;
    const uint32_t hs = (uint32_t)ht_size_for(cnt), mask = hs - 1;
    // where the u32 working table lives: shared memory (small tables), the final table itself (u32 slots), or a scratch slice
    // (u16 tables too large for shared memory; ht_cap = this job's offset into the scratch)
    uint32_t* tab = hs <= LZ_IDX_SMEM_SLOTS ? s_tab : (is_short ? gtab_all + j.ht_cap : (uint32_t*)j.ht);
    for (uint32_t i = threadIdx.x; i < hs; i += LZ_IDX_THREADS) tab[i] = AGC_EMPTY32;
    __syncthreads();
    const uint64_t* R = (const uint64_t*)j.packed;
    auto home = [&](uint32_t v) { return (uint32_t)agc_murmur64(agc_win(R, (uint64_t)v * 4u) >> (64 - 2 * kl)) & mask; };
    for (uint32_t v0 = threadIdx.x; v0 < cnt; v0 += LZ_IDX_THREADS) {
        uint32_t cur = v0, s = home(cur), tries = 0;
        while (tries < 64) {
            const uint32_t old = atomicMin(&tab[s], cur);
            if (old == AGC_EMPTY32) break;                        // placed in an empty slot
            if (old > cur) { cur = old; tries = ((s - home(cur)) & mask) + 1u; }   // took the slot: the displaced key moves on
            else ++tries;                                         // an earlier key owns it
            s = (s + 1u) & mask;
        }
    }
    __syncthreads();
    if (is_short) { uint16_t* o = (uint16_t*)j.ht; for (uint32_t i = threadIdx.x; i < hs; i += LZ_IDX_THREADS) { const uint32_t v = tab[i]; o[i] = v == AGC_EMPTY32 ? (uint16_t)0xffffu : (uint16_t)v; } }
    else if (tab != (uint32_t*)j.ht) { uint32_t* o = (uint32_t*)j.ht; for (uint32_t i = threadIdx.x; i < hs; i += LZ_IDX_THREADS) o[i] = tab[i]; }
    if (threadIdx.x == 0) {
        GroupRefDev g;
        g.packed = j.packed; g.ht = j.ht; g.codes = nullptr; g.m = m; g.ht_size = hs;
        g.flags = GRF_PRESENT | (is_short ? GRF_SHORT : 0u);
        g.packed_bytes = ((m + 3) / 4 + 15) / 16 * 16 + 32;
        groups[j.group] = g;
    }
}

// ------------------------------------------------------------------------------------------------ output gather
__global__ void __launch_bounds__(1024) k_excl_scan_u32(const uint32_t* __restrict__ in, uint32_t n, uint64_t* __restrict__ out)
{
    __shared__ uint64_t s_w[32];
    __shared__ uint64_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (uint32_t start = 0; start < n; start += 1024) {
        uint32_t i = start + threadIdx.x;
        uint64_t v = i < n ? in[i] : 0, inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(FULL, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
        if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            uint64_t wv = s_w[threadIdx.x], x = wv;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint64_t t = __shfl_up_sync(FULL, x, o); if (threadIdx.x >= o) x += t; }
            s_w[threadIdx.x] = x - wv;
        }
        __syncthreads();
        uint64_t ex = carry + s_w[threadIdx.x >> 5] + inc - v;
        if (i < n) out[i] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) carry = ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = carry;
}

__global__ void k_gather(const uint8_t* __restrict__ slab, const uint64_t* __restrict__ src_off, const uint64_t* __restrict__ dst_off,
                         uint32_t n, uint8_t* __restrict__ dense)
{
    uint32_t w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= n) return;
    uint64_t d0 = dst_off[w], len = dst_off[w + 1] - d0;
    const uint8_t* s = slab + src_off[w];
    for (uint64_t i = lane; i < len; i += 32) dense[d0 + i] = s[i];
}

// ------------------------------------------------------------------------------------------------ reference packing
// periodicity probe of store_in_archive(ref) (src/common/segment.h:226-249) on a clean packed reference:
// use_tuples = no lag in 4..31 has 2*cnt >= cur (cnt = equal symbol pairs at that lag, cur = n - lag).
__global__ void __launch_bounds__(32) k_ref_probe_packed(const GroupRefDev* __restrict__ groups, const uint32_t* __restrict__ ids,
                                                        uint32_t n, uint8_t* __restrict__ use_tuples)
{
    if (blockIdx.x >= n) return;
    const GroupRefDev g = groups[ids[blockIdx.x]];
    if (g.flags & GRF_DIRTY) return;
    const uint64_t* R = (const uint64_t*)g.packed;
    const uint32_t lane = threadIdx.x, m = g.m;
    bool periodic = false;
    for (uint32_t lag = 4; lag < 32 && !periodic; ++lag) {
        if (m <= lag) break;
        uint32_t pairs = m - lag, cnt = 0;
        for (uint32_t j0 = 32 * lane; j0 < pairs; j0 += 1024) {
            uint64_t x = agc_win(R, j0) ^ agc_win(R, j0 + lag);
            uint64_t eq = ~(x | (x >> 1)) & 0x5555555555555555ULL;
            uint32_t rem = pairs - j0;
            if (rem < 32) eq &= ~0ULL << (64 - 2 * rem);
            cnt += __popcll(eq);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(FULL, cnt, o);
        if (2ull * cnt >= pairs) periodic = true;
    }
    if (lane == 0) use_tuples[blockIdx.x] = periodic ? 0 : 1;
}

// bytes2tuples_impl<4,4> (segment.h:108-138) from the packed form: floor(m/4) whole bytes, one right-aligned
// remainder byte (always present), marker (4<<4) + m%4.
__global__ void k_ref_tuples_packed(const GroupRefDev* __restrict__ groups, const uint32_t* __restrict__ ids, uint32_t n,
                                    const uint64_t* __restrict__ out_off, uint8_t* __restrict__ out)
{
    if (blockIdx.y >= n) return;
    const GroupRefDev g = groups[ids[blockIdx.y]];
    if (g.flags & GRF_DIRTY) return;
    uint32_t full = g.m / 4, total = full + 2;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint8_t v;
    if (i < full) v = g.packed[i];
    else if (i == full) { uint32_t r = g.m & 3u; v = r ? (uint8_t)(g.packed[full] >> (8 - 2 * r)) : 0; }
    else v = (uint8_t)((4u << 4) + (g.m & 3u));
    out[out_off[blockIdx.y] + i] = v;
}

// same probe on 1-byte symbols (references with non-ACGT symbols): cnt counts every equal pair, cur only ACGT positions
// (segment.h:231-241); also returns the largest symbol (selects the tuple width, segment.h:73-91)
__global__ void __launch_bounds__(32) k_ref_probe_bytes(const uint8_t* __restrict__ d, uint32_t m, uint8_t* __restrict__ use_tuples,
                                                       uint8_t* __restrict__ max_sym)
{
    const uint32_t lane = threadIdx.x;
    uint32_t me = 0;
    for (uint32_t j = lane; j < m; j += 32) me = max(me, (uint32_t)d[j]);
#pragma unroll
    for (int o = 16; o; o >>= 1) me = max(me, __shfl_xor_sync(FULL, me, o));
    bool periodic = false;
    for (uint32_t lag = 4; lag < 32 && !periodic; ++lag) {
        if (m <= lag) break;
        uint32_t cnt = 0, cur = 0;
        for (uint32_t j = lane; j + lag < m; j += 32) { cnt += d[j] == d[j + lag]; cur += d[j] < 4; }
#pragma unroll
        for (int o = 16; o; o >>= 1) { cnt += __shfl_xor_sync(FULL, cnt, o); cur += __shfl_xor_sync(FULL, cur, o); }
        if (cur && 2ull * cnt >= cur) periodic = true;
    }
    if (lane == 0) { *use_tuples = periodic ? 0 : 1; *max_sym = (uint8_t)me; }
}
// bytes2tuples_impl<NO_BYTES, MULT> (segment.h:108-138) / raw + 0x10 marker (86-90)
__global__ void k_ref_tuples_bytes(const uint8_t* __restrict__ d, uint32_t m, uint32_t nb, uint32_t mult, uint8_t* __restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (nb == 1) { if (i < m) out[i] = d[i]; else if (i == m) out[i] = 0x10; return; }
    uint32_t full = m / nb;
    if (i > full + 1) return;
    uint32_t c = 0;
    if (i < full) for (uint32_t j = 0; j < nb; ++j) c = c * mult + d[i * nb + j];
    else if (i == full) for (uint32_t q = full * nb; q < m; ++q) c = c * mult + d[q];
    else c = (nb << 4) + (m % nb);
    out[i] = (uint8_t)c;
}

// ================================================================================================ host side
static uint64_t clean_ht_size(uint32_t m, uint32_t mml)
{
    uint32_t kl = mml - 3;
    uint64_t cnt = m >= kl ? (m - kl) / 4 + 1 : 0;
    uint64_t hs = (uint64_t)((double)cnt / 0.7);
    while (hs & (hs - 1)) hs &= hs - 1;
    hs <<= 1; if (hs < 8) hs = 8;
    return hs;
}

static int ensure_groups(agcgpu_ctx* ctx, uint32_t max_group)
{
    if (ctx->h_groups.size() <= max_group) {
        GroupRefDev z; memset(&z, 0, sizeof(z));
        ctx->h_groups.resize((size_t)max_group + 1, z);
    }
    size_t need = ctx->h_groups.size() * sizeof(GroupRefDev);
    if (ctx->d_groups.cap < need) {
        size_t newcap = std::max(need * 2, (size_t)1 << 16);
        if (int r = agc_reserve(ctx, ctx->d_groups, newcap, true)) return r;
    }
    return 0;
}

static int build_refs(agcgpu_ctx* ctx, std::vector<RefJob>& jobs, bool from_segments)
{
    if (jobs.empty()) return 0;
    const uint32_t mml = ctx->prm.min_match_len;
    uint32_t max_group = 0, max_n = 0;
    for (auto& j : jobs) { max_group = std::max(max_group, j.group); max_n = std::max(max_n, j.n); }
    if (int r = ensure_groups(ctx, max_group)) return r;
    if (int r = agc_reserve(ctx, ctx->scr_misc, jobs.size() * sizeof(RefJob))) return r;
    CK(cudaMemcpyAsync(ctx->scr_misc.p, jobs.data(), jobs.size() * sizeof(RefJob), cudaMemcpyHostToDevice, ctx->st));
    const RefJob* d_jobs = (const RefJob*)ctx->scr_misc.p;
    if (from_segments && max_n) {
        for (size_t y0 = 0; y0 < jobs.size(); y0 += 32768) {
            uint32_t ny = (uint32_t)std::min<size_t>(32768, jobs.size() - y0);
            dim3 grid(((max_n + 31) / 32 + 127) / 128, ny);
            k_extract_ref<<<grid, 128, 0, ctx->st>>>((const uint64_t*)ctx->packed.p, d_jobs + y0);
            CKL();
        }
        for (auto& j : jobs) if (j.codes) if (int r = agc_expand_segment(ctx, j.gstart, j.n, j.is_rc, j.codes, mml - 3)) return r;
    }
    uint32_t nb = (uint32_t)((jobs.size() + 3) / 4);
    {
        // scratch slices for u16 tables that do not fit in shared memory (references of 115 k .. 262 k symbols)
        uint64_t scratch = 0;
        for (auto& j : jobs) {
            const uint64_t hs = clean_ht_size(j.n, mml);
            if (!j.codes && (j.n / 4) < 65535 && hs > LZ_IDX_SMEM_SLOTS) { j.ht_cap = (uint32_t)scratch; scratch += hs; }
        }
        if (scratch) {
            if (scratch >= 0xffffffffull) return agc_fail(ctx, AGCGPU_ENOMEM, "index build: too many large references in one batch");
            if (int r = agc_reserve(ctx, ctx->scr_dense, scratch * 4 + 64)) return r;
            CK(cudaMemcpyAsync(ctx->scr_misc.p, jobs.data(), jobs.size() * sizeof(RefJob), cudaMemcpyHostToDevice, ctx->st));
        }
        static bool attr_set = false;
        if (!attr_set) { cudaFuncSetAttribute(k_index_par, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(LZ_IDX_SMEM_SLOTS * 4)); attr_set = true; }
        k_index_par<<<(uint32_t)jobs.size(), LZ_IDX_THREADS, LZ_IDX_SMEM_SLOTS * 4, ctx->st>>>(d_jobs, (GroupRefDev*)ctx->d_groups.p, mml, (uint32_t*)ctx->scr_dense.p);
        CKL();
    }
    bool any_dirty = false;
    for (auto& j : jobs) any_dirty |= j.codes != nullptr;
    if (any_dirty) { k_index<true><<<nb, 128, 0, ctx->st>>>(d_jobs, (uint32_t)jobs.size(), (GroupRefDev*)ctx->d_groups.p, mml); CKL(); }
    CK(cudaStreamSynchronize(ctx->st));
    // mirror descriptors on the host (ht_size of dirty references is computed on the device)
    for (auto& j : jobs) {
        CK(cudaMemcpy(&ctx->h_groups[j.group], (GroupRefDev*)ctx->d_groups.p + j.group, sizeof(GroupRefDev), cudaMemcpyDeviceToHost));
        ctx->stats.d2h_bytes += sizeof(GroupRefDev);
    }
    return 0;
}

static int alloc_ref_job(agcgpu_ctx* ctx, RefJob& j, bool dirty)
{
    const uint32_t mml = ctx->prm.min_match_len;
    size_t pbytes = ((size_t)(j.n + 3) / 4 + 15) / 16 * 16 + 32;
    uint64_t hs = clean_ht_size(j.n, mml);
    bool is_short = (j.n / 4) < 65535;
    size_t hbytes = hs * (is_short ? 2 : 4);
    j.packed = (uint8_t*)agc_arena_alloc(ctx, pbytes);
    j.ht = agc_arena_alloc(ctx, hbytes);
    j.codes = dirty ? (uint8_t*)agc_arena_alloc(ctx, (size_t)j.n + mml) : nullptr;
    if (!j.packed || !j.ht || (dirty && !j.codes)) return agc_fail(ctx, AGCGPU_ENOMEM, "reference arena allocation failed");
    j.ht_cap = (uint32_t)hs;
    CK(cudaMemsetAsync(j.packed, 0, pbytes, ctx->st));
    CK(cudaMemsetAsync(j.ht, 0xff, hbytes, ctx->st));
    return 0;
}

int agc_refs_from_segments(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n)
{
    std::vector<RefJob> jobs(n);
    for (uint32_t i = 0; i < n; ++i) {
        const agcgpu_seg_req& q = reqs[i];
        if (q.contig >= ctx->n_contigs) return agc_fail(ctx, AGCGPU_EINVAL, "put_reference: contig %u not resident", q.contig);
        uint64_t clen = ctx->h_cstart[q.contig + 1] - ctx->h_cstart[q.contig];
        if (q.start + q.len > clen) return agc_fail(ctx, AGCGPU_EINVAL, "put_reference: segment outside contig");
        RefJob& j = jobs[i];
        j.gstart = ctx->h_cstart[q.contig] + q.start; j.n = q.len; j.is_rc = q.is_rc; j.group = q.group_id;
        if (int r = alloc_ref_job(ctx, j, agc_segment_dirty(ctx, j.gstart, j.n))) return r;
    }
    return build_refs(ctx, jobs, true);
}

// pack host symbols: one thread per output byte
__global__ void k_pack_codes(const uint8_t* __restrict__ codes, uint32_t n, uint8_t* __restrict__ packed)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i * 4 >= n) return;
    uint32_t v = 0;
    for (uint32_t b = 0; b < 4; ++b) { uint32_t q = i * 4 + b; uint32_t s = q < n ? codes[q] : 0; if (s > 3) s = 0; v |= s << (6 - 2 * b); }
    packed[i] = (uint8_t)v;
}

int agc_ref_from_host(agcgpu_ctx* ctx, uint32_t group, const uint8_t* symbols, uint32_t len)
{
    const uint32_t mml = ctx->prm.min_match_len;
    bool dirty = false;
    for (uint32_t i = 0; i < len; ++i) if (symbols[i] > 3) { dirty = true; break; }
    std::vector<RefJob> jobs(1);
    RefJob& j = jobs[0];
    j.gstart = 0; j.n = len; j.is_rc = 0; j.group = group;
    if (int r = alloc_ref_job(ctx, j, true)) return r;       // codes buffer doubles as the upload staging
    std::vector<uint8_t> tmp((size_t)len + mml, 31);
    if (len) memcpy(tmp.data(), symbols, len);
    CK(cudaMemcpyAsync(j.codes, tmp.data(), tmp.size(), cudaMemcpyHostToDevice, ctx->st));
    ctx->stats.h2d_bytes += tmp.size();
    if (len) { k_pack_codes<<<((len + 3) / 4 + 255) / 256, 256, 0, ctx->st>>>(j.codes, len, j.packed); CKL(); }
    CK(cudaStreamSynchronize(ctx->st));
    if (!dirty) j.codes = nullptr;
    return build_refs(ctx, jobs, false);
}

// 1 byte / symbol copy of a clean packed reference + pad bytes of 31
__global__ void k_expand_ref(const uint8_t* __restrict__ packed, uint32_t m, uint8_t* __restrict__ dst, uint32_t pad)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) dst[i] = (packed[i >> 2] >> (6 - 2 * (i & 3))) & 3u;
    else if (i < m + pad) dst[i] = 31;
}

// make sure group g has a 1-byte/symbol copy of its reference (needed when a dirty text meets a clean reference)
static int ensure_ref_codes(agcgpu_ctx* ctx, uint32_t gid)
{
    GroupRefDev& g = ctx->h_groups[gid];
    if (g.codes) return 0;
    const uint32_t pad = ctx->prm.min_match_len;
    uint8_t* codes = (uint8_t*)agc_arena_alloc(ctx, (size_t)g.m + pad);
    if (!codes) return agc_fail(ctx, AGCGPU_ENOMEM, "reference arena allocation failed");
    k_expand_ref<<<(g.m + pad + 255) / 256, 256, 0, ctx->st>>>(g.packed, g.m, codes, pad);
    CKL();
    g.codes = codes;     // host mirror only: the device descriptor keeps codes == nullptr for clean references
    return 0;
}

template <int MODE>
static void launch_packed(agcgpu_ctx* ctx, uint32_t n_units, size_t smem, const LzReqDev* d_req, const LzUnit* d_units,
                          uint8_t* slab, uint32_t* res, uint32_t* costv, int prefix, uint32_t* err)
{
    cudaFuncSetAttribute(k_lz_packed<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k_lz_packed<MODE><<<n_units, LZ_THREADS, smem, ctx->st>>>((const uint64_t*)ctx->packed.p, (const GroupRefDev*)ctx->d_groups.p,
        d_req, d_units, ctx->prm.min_match_len, (uint32_t)smem, slab, res, costv, prefix, err);
}
template <int MODE>
static void launch_bytes(agcgpu_ctx* ctx, const ByteReq* d_req, uint32_t n, uint8_t* slab, uint32_t* res, uint32_t* costv,
                         int prefix, uint32_t* err)
{
    k_lz_bytes<MODE><<<(n + 3) / 4, 128, 0, ctx->st>>>(d_req, n, ctx->prm.min_match_len, slab, res, costv, prefix, err);
}

// mode 0 encode (out_bytes/out_offsets), 1 estimate (out_u32[n]), 2 cost vector (n == 1, out_u32[len])
int agc_lz_run(agcgpu_ctx* ctx, int mode, const agcgpu_seg_req* reqs, uint32_t n, int prefix_costs,
               uint8_t* out_bytes, uint64_t out_cap, uint64_t* out_offsets, uint32_t* out_u32)
{
    if (n == 0) { if (mode == 0 && out_offsets) out_offsets[0] = 0; return 0; }
    const uint32_t mml = ctx->prm.min_match_len;
    // ---- classify + order by group
    std::vector<LzReqDev> packed_reqs; packed_reqs.reserve(n);
    std::vector<uint32_t> dirty_idx;
    std::vector<uint64_t> slab_off(n);
    uint64_t slab_total = 0, alg_bytes = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const agcgpu_seg_req& q = reqs[i];
        if (q.contig >= ctx->n_contigs) return agc_fail(ctx, AGCGPU_EINVAL, "lz: contig %u not resident", q.contig);
        uint64_t clen = ctx->h_cstart[q.contig + 1] - ctx->h_cstart[q.contig];
        if (q.start + q.len > clen) return agc_fail(ctx, AGCGPU_EINVAL, "lz: segment %u outside contig", i);
        if (q.group_id >= ctx->h_groups.size() || !(ctx->h_groups[q.group_id].flags & GRF_PRESENT))
            return agc_fail(ctx, AGCGPU_EINVAL, "lz: group %u has no reference", q.group_id);
        const GroupRefDev& g = ctx->h_groups[q.group_id];
        uint64_t gstart = ctx->h_cstart[q.contig] + q.start;
        uint64_t cap = mode == 0 ? ((uint64_t)q.len * 3) / 2 + 32 : (mode == 2 ? q.len : 0);
        slab_off[i] = slab_total;
        slab_total += mode == 0 ? (cap + 15) / 16 * 16 : cap;
        alg_bytes += (q.len + 3) / 4 + (g.m + 3) / 4;
        bool dirty = (g.flags & GRF_DIRTY) || agc_segment_dirty(ctx, gstart, q.len);
        if (dirty) { dirty_idx.push_back(i); continue; }
        LzReqDev d; d.gstart = gstart; d.n = q.len; d.is_rc = q.is_rc; d.group = q.group_id; d.bound = q.bound;
        d.out_off = slab_off[i]; d.out_cap = (uint32_t)std::min<uint64_t>(cap, 0xffffffffu); d.orig = i;
        packed_reqs.push_back(d);
    }
    std::stable_sort(packed_reqs.begin(), packed_reqs.end(), [](const LzReqDev& a, const LzReqDev& b) { return a.group < b.group; });
    const size_t resident = 2 * (size_t)ctx->n_sm;            // CTAs of the sequential kernel that fit on the device
    bool chunk_timed = false;
    // ---- device buffers
    if (int r = agc_reserve(ctx, ctx->scr_out, slab_total * (mode == 2 ? 4 : 1) + 64)) return r;
    if (int r = agc_reserve(ctx, ctx->scr_sizes, (size_t)n * 4 + 64)) return r;
    if (int r = agc_reserve(ctx, ctx->counters, 64)) return r;
    CK(cudaMemsetAsync(ctx->counters.p, 0, 64, ctx->st));
    CK(cudaMemsetAsync(ctx->scr_sizes.p, 0, (size_t)n * 4, ctx->st));
    uint8_t* slab = (uint8_t*)ctx->scr_out.p;
    uint32_t* costv = (uint32_t*)ctx->scr_out.p;
    uint32_t* res = (uint32_t*)ctx->scr_sizes.p;
    uint32_t* err = (uint32_t*)ctx->counters.p + 8;
    CK(cudaEventRecord(ctx->ev0, ctx->st));
    // the sequential (warp per segment) kernel over packed_reqs[sel]; sel == nullptr: all of them
    auto run_sequential = [&](const std::vector<LzReqDev>& rq) -> int {
        if (rq.empty()) return 0;
        std::vector<LzUnit> un;
        size_t smem_seq = 0;
        const uint32_t umax = (uint32_t)std::max<size_t>(16, ((rq.size() + resident - 1) / resident + 15) / 16 * 16);
        for (size_t a = 0; a < rq.size();) {
            size_t b = a;
            while (b < rq.size() && rq[b].group == rq[a].group) ++b;
            const GroupRefDev& g = ctx->h_groups[rq[a].group];
            size_t need = (size_t)g.packed_bytes + (size_t)g.ht_size * ((g.flags & GRF_SHORT) ? 2 : 4);
            if (need <= LZ_STAGE_LIMIT) smem_seq = std::max(smem_seq, need);
            size_t cnt = b - a, nun = (cnt + umax - 1) / umax, per = (cnt + nun - 1) / nun;
            for (size_t s0 = a; s0 < b; s0 += per) {
                LzUnit u; u.group = rq[a].group; u.first = (uint32_t)s0; u.count = (uint32_t)std::min(per, b - s0); u.pad = 0;
                un.push_back(u);
            }
            a = b;
        }
        if (int r = agc_reserve(ctx, ctx->scr_req, rq.size() * sizeof(LzReqDev))) return r;
        if (int r = agc_reserve(ctx, ctx->scr_units, un.size() * sizeof(LzUnit))) return r;
        CK(cudaMemcpyAsync(ctx->scr_req.p, rq.data(), rq.size() * sizeof(LzReqDev), cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->scr_units.p, un.data(), un.size() * sizeof(LzUnit), cudaMemcpyHostToDevice, ctx->st));
        ctx->stats.h2d_bytes += rq.size() * sizeof(LzReqDev) + un.size() * sizeof(LzUnit);
        const LzReqDev* d_req = (const LzReqDev*)ctx->scr_req.p;
        const LzUnit* d_units = (const LzUnit*)ctx->scr_units.p;
        size_t smem = std::max<size_t>(smem_seq, 1024);
        if (mode == 0) launch_packed<0>(ctx, (uint32_t)un.size(), smem, d_req, d_units, slab, res, costv, prefix_costs, err);
        else if (mode == 1) launch_packed<1>(ctx, (uint32_t)un.size(), smem, d_req, d_units, slab, res, costv, prefix_costs, err);
        else launch_packed<2>(ctx, (uint32_t)un.size(), smem, d_req, d_units, slab, res, costv, prefix_costs, err);
        CKL();
        return 0;
    };
    static const bool no_chunks = getenv("AGCGPU_LZ_SEQUENTIAL") != nullptr;          // diagnostics: force the sequential kernel
    // ---- encode: warp per segment streaming along the diagonal (kernels_lz_diag.cu) for every segment of ordinary length; what is
    // left (segments above LZD_MAX_N bases: contigs without splitters) takes the chunk-parallel kernels below
    static const bool no_diag = getenv("AGCGPU_LZ_NO_DIAG") != nullptr;               // diagnostics: chunk-parallel kernels for everything
    bool diag_timed = false;
    if (mode == 0 && !no_chunks && !no_diag && !packed_reqs.empty()) {
        // One warp walks a segment alone.  Launch (1): groups whose reference AND hash table fit in shared memory (32 warps, one CTA per
        // SM) -- every ordinary 60 kb group.  Launch (2), optional: groups of which only the reference fits (references of 92 kb and more
        // have a 128+ KB table): 16 warps per CTA, the index probes go to L2.  Segments above LZD_MAX_N bases, segments much
        // longer than their reference and references beyond 96 KB packed take the chunk-parallel kernels.  (Variant (2) for everything: no faster at 0.1 % SNPs, 0.415 vs 0.423 ms,
        // and 23 % slower at 1 %.)
        static const int ht_mode = getenv("AGCGPU_LZ_DIAG_HT") ? atoi(getenv("AGCGPU_LZ_DIAG_HT")) : 1;      // 0: variant (2) for every group
        const size_t stage_max_full = ht_mode ? (size_t)227 * 1024 - agc_lzd_scratch_bytes(1) - 1024 : 0;
        const size_t stage_max_ref = (size_t)96 * 1024;
        // launch (2) is OFF by default: on C3's 120-260 kb merged groups at 1 % SNPs the probes through L2 make it 3-4x slower than the
        // chunk-parallel kernels (5-12 ms vs 2-3 ms per one-sample launch); AGCGPU_LZ_DIAG_REFONLY=1 switches it on
        static const bool ref_only = !ht_mode || (getenv("AGCGPU_LZ_DIAG_REFONLY") && atoi(getenv("AGCGPU_LZ_DIAG_REFONLY")));
        std::vector<LzReqDev> dq[2], rest;
        for (const LzReqDev& q : packed_reqs) {
            const GroupRefDev& g = ctx->h_groups[q.group];
            const size_t full = (size_t)g.packed_bytes + (size_t)g.ht_size * ((g.flags & GRF_SHORT) ? 2 : 4);
            // (a text much longer than its reference -- a merged segment against an ordinary group -- is mostly literals: the one warp
            // would probe them 32 at a time, ~0.1 us per position; the chunk kernels spread them over a lane per 512 positions)
            if (q.n > LZD_MAX_N || q.n > g.m + 8192u) rest.push_back(q);
            else if (full <= stage_max_full) dq[1].push_back(q);
            else if (ref_only && g.packed_bytes <= stage_max_ref) dq[0].push_back(q);
            else rest.push_back(q);
        }
        size_t d_off = 0;
        const size_t n_diag = dq[0].size() + dq[1].size();
        if (n_diag) if (int r = agc_reserve(ctx, ctx->scr_rec, n_diag * (sizeof(LzReqDev) + sizeof(LzUnit)) + 512)) return r;
        std::vector<LzUnit> un[2];
        LzReqDev* d_req[2] = { nullptr, nullptr }; LzUnit* d_un[2] = { nullptr, nullptr };
        size_t stage[2] = { 0, 0 };
        for (int hs = 1; hs >= 0; --hs) {                          // requests and units of both launches go up first ...
            const std::vector<LzReqDev>& v = dq[hs];
            if (v.empty()) continue;
            const uint32_t unit_max = hs ? 32u : 16u;
            for (size_t a2 = 0; a2 < v.size();) {
                size_t b2 = a2;
                while (b2 < v.size() && v[b2].group == v[a2].group) ++b2;
                const GroupRefDev& g = ctx->h_groups[v[a2].group];
                stage[hs] = std::max(stage[hs], (size_t)g.packed_bytes + (hs ? (size_t)g.ht_size * ((g.flags & GRF_SHORT) ? 2 : 4) : 0));
                const size_t cnt = b2 - a2, nun = (cnt + unit_max - 1) / unit_max, per = (cnt + nun - 1) / nun;      // one request per warp and round
                for (size_t s0 = a2; s0 < b2; s0 += per) {
                    LzUnit u; u.group = v[a2].group; u.first = (uint32_t)s0; u.count = (uint32_t)std::min(per, b2 - s0); u.pad = 0;
                    un[hs].push_back(u);
                }
                a2 = b2;
            }
            d_req[hs] = (LzReqDev*)((uint8_t*)ctx->scr_rec.p + d_off);
            d_un[hs] = (LzUnit*)(d_req[hs] + v.size());
            d_off += (v.size() * sizeof(LzReqDev) + un[hs].size() * sizeof(LzUnit) + 255) / 256 * 256;
            CK(cudaMemcpyAsync(d_req[hs], v.data(), v.size() * sizeof(LzReqDev), cudaMemcpyHostToDevice, ctx->st));
            CK(cudaMemcpyAsync(d_un[hs], un[hs].data(), un[hs].size() * sizeof(LzUnit), cudaMemcpyHostToDevice, ctx->st));
            ctx->stats.h2d_bytes += v.size() * sizeof(LzReqDev) + un[hs].size() * sizeof(LzUnit);
        }
        if (n_diag) CK(cudaEventRecord(ctx->ev0, ctx->st));        // ... so that the timed interval holds the kernels only
        for (int hs = 1; hs >= 0; --hs) {
            if (dq[hs].empty()) continue;
            if (int r = agc_lzd_launch(ctx, d_req[hs], d_un[hs], (uint32_t)un[hs].size(), stage[hs], hs, slab, res, err)) return r;
            ctx->stats.lz_diag_segments += (uint32_t)dq[hs].size();
            diag_timed = true;
        }
        if (diag_timed && rest.empty()) { CK(cudaEventRecord(ctx->ev1, ctx->st)); chunk_timed = true; }
        packed_reqs.swap(rest);
    }
    if (!packed_reqs.empty() && (mode == 0 || mode == 2) && !no_chunks) {
        // ---- chunk-parallel encode / cost vectors (kernels_lz_chunk.cu): thread per chunk, then thread per segment; segments the
        // stitcher could not prove identical to the sequential parse are redone by the sequential kernel
        const size_t nr = packed_reqs.size();
        std::vector<LzcReq> cr(nr);
        std::vector<LzcUnit> cu;
        uint64_t n_chunks = 0; size_t smem_c = 0;
        const uint32_t UNIT_ITEMS = 4 * LZC_THREADS;
        // chunk size: large batches take LZC_CHUNK; a batch that would not even give every resident lane a chunk (one sample of a
        // collection) is latency bound -- smaller chunks mean more lanes and fewer tokens per lane
        uint64_t big_chunks = 0;
        for (size_t i = 0; i < nr; ++i) big_chunks += (packed_reqs[i].n + LZC_CHUNK - 1) / LZC_CHUNK;
        uint32_t chunk = LZC_CHUNK;
        while (chunk > LZC_CHUNK_MIN && big_chunks * (LZC_CHUNK / chunk) < 2ull * 1024 * (uint64_t)ctx->n_sm) chunk >>= 1;
        for (size_t a = 0; a < nr;) {
            size_t b = a;
            while (b < nr && packed_reqs[b].group == packed_reqs[a].group) ++b;
            const GroupRefDev& g = ctx->h_groups[packed_reqs[a].group];
            size_t need = (size_t)g.packed_bytes + (size_t)g.ht_size * ((g.flags & GRF_SHORT) ? 2 : 4);
            if (need <= LZC_STAGE_LIMIT) smem_c = std::max(smem_c, need);
            uint32_t base = 0;
            for (size_t i = a; i < b; ++i) {
                const LzReqDev& q = packed_reqs[i];
                LzcReq& c = cr[i];
                c.gstart = q.gstart; c.n = q.n; c.is_rc = q.is_rc; c.group = q.group; c.chunk_first = (uint32_t)n_chunks;
                c.chunk = chunk; c.pad = 0; c.nch = std::max<uint32_t>(1u, (q.n + chunk - 1) / chunk); c.unit_base = base; c.out_off = q.out_off; c.out_cap = mode == 2 ? q.bound : q.out_cap; c.orig = q.orig;
                base += c.nch; n_chunks += c.nch;
            }
            // all requests of the group share one running chunk count; a unit is a slice of UNIT_ITEMS chunks of it
            for (uint32_t i0 = 0; i0 < base; i0 += UNIT_ITEMS) {
                LzcUnit u; u.group = packed_reqs[a].group; u.first = (uint32_t)a; u.count = (uint32_t)(b - a); u.item0 = i0;
                u.n_items = std::min(UNIT_ITEMS, base - i0); u.pad = 0;
                cu.push_back(u);
            }
            a = b;
        }
        if (n_chunks >= 0xffffffffull) return agc_fail(ctx, AGCGPU_EINVAL, "lz: batch too large (%llu chunks)", (unsigned long long)n_chunks);
        if (int r = agc_reserve(ctx, ctx->scr_req, nr * sizeof(LzcReq))) return r;
        if (int r = agc_reserve(ctx, ctx->scr_units, cu.size() * sizeof(LzcUnit))) return r;
        if (int r = agc_reserve(ctx, ctx->scr_chunk, n_chunks * (uint64_t)LZC_CSLAB + 256)) return r;
        if (int r = agc_reserve(ctx, ctx->scr_rec, n_chunks * sizeof(LzcRec) + nr * 4 + 256)) return r;
        CK(cudaMemcpyAsync(ctx->scr_req.p, cr.data(), nr * sizeof(LzcReq), cudaMemcpyHostToDevice, ctx->st));
        CK(cudaMemcpyAsync(ctx->scr_units.p, cu.data(), cu.size() * sizeof(LzcUnit), cudaMemcpyHostToDevice, ctx->st));
        ctx->stats.h2d_bytes += nr * sizeof(LzcReq) + cu.size() * sizeof(LzcUnit);
        if (mode == 2) CK(cudaMemsetAsync(costv, 0, slab_total * 4, ctx->st));      // the chunks only write the non-zero entries
        if (!diag_timed) CK(cudaEventRecord(ctx->ev0, ctx->st));
        LzcRec* d_rec = (LzcRec*)ctx->scr_rec.p;
        uint32_t* d_fb = (uint32_t*)(d_rec + n_chunks);
        uint32_t* cnt2 = (uint32_t*)ctx->counters.p;                       // [0] segments for the sequential kernel, [1] overflow
        if (int r = agc_lzc_launch(ctx, (const LzcReq*)ctx->scr_req.p, (uint32_t)nr, (const LzcUnit*)ctx->scr_units.p, (uint32_t)cu.size(),
                                   std::max<size_t>(smem_c, 1024), (uint8_t*)ctx->scr_chunk.p, d_rec, slab, res, d_fb, cnt2, mode == 2 ? costv : nullptr)) return r;
        CK(cudaEventRecord(ctx->ev1, ctx->st));
        uint32_t h_cnt[2] = { 0, 0 };
        CK(cudaMemcpyAsync(h_cnt, cnt2, 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        if (h_cnt[1]) return agc_fail(ctx, AGCGPU_EOVERFLOW, "lz encode: per-segment output bound exceeded");
        ctx->stats.lz_chunk_segments += nr; ctx->stats.lz_sequential_segments += h_cnt[0];
        if (h_cnt[0]) {
            std::vector<uint32_t> h_fb(nr);
            CK(cudaMemcpy(h_fb.data(), d_fb, nr * 4, cudaMemcpyDeviceToHost));
            std::vector<LzReqDev> redo;
            for (size_t i = 0; i < nr; ++i) if (h_fb[i]) redo.push_back(packed_reqs[i]);
            if (int r = run_sequential(redo)) return r;
        }
        chunk_timed = true; ctx->last_lzc_chunks = n_chunks;
    } else if (!packed_reqs.empty()) {
        if (!diag_timed) CK(cudaEventRecord(ctx->ev0, ctx->st));
        if (int r = run_sequential(packed_reqs)) return r;
    }
    if (!chunk_timed) CK(cudaEventRecord(ctx->ev1, ctx->st));
    if (!dirty_idx.empty()) {
        // byte path: expand text (and, for clean references, the reference) to 1 byte / symbol
        size_t tbytes = 0;
        for (uint32_t i : dirty_idx) tbytes += ((size_t)reqs[i].len + 63) / 64 * 64;
        if (int r = agc_reserve(ctx, ctx->scr_bytes, tbytes + 64)) return r;
        std::vector<ByteReq> br(dirty_idx.size());
        size_t off = 0;
        for (size_t k = 0; k < dirty_idx.size(); ++k) {
            uint32_t i = dirty_idx[k];
            const agcgpu_seg_req& q = reqs[i];
            if (int r = ensure_ref_codes(ctx, q.group_id)) return r;
            const GroupRefDev& g = ctx->h_groups[q.group_id];
            uint8_t* t = (uint8_t*)ctx->scr_bytes.p + off;
            off += ((size_t)q.len + 63) / 64 * 64;
            if (int r = agc_expand_segment(ctx, ctx->h_cstart[q.contig] + q.start, q.len, q.is_rc, t, 0)) return r;
            ByteReq& b = br[k];
            b.text = t; b.ref = g.codes; b.ht = g.ht; b.n = q.len; b.m = g.m; b.ht_size = g.ht_size; b.is_short = g.flags & GRF_SHORT;
            b.bound = q.bound; b.orig = i; b.out_off = slab_off[i];
            uint64_t cap = mode == 0 ? ((uint64_t)q.len * 3) / 2 + 32 : q.len;
            b.out_cap = (uint32_t)std::min<uint64_t>(cap, 0xffffffffu); b.pad = 0;
        }
        if (int r = agc_reserve(ctx, ctx->scr_misc, br.size() * sizeof(ByteReq))) return r;
        CK(cudaMemcpyAsync(ctx->scr_misc.p, br.data(), br.size() * sizeof(ByteReq), cudaMemcpyHostToDevice, ctx->st));
        const ByteReq* d_br = (const ByteReq*)ctx->scr_misc.p;
        if (mode == 0) launch_bytes<0>(ctx, d_br, (uint32_t)br.size(), slab, res, costv, prefix_costs, err);
        else if (mode == 1) launch_bytes<1>(ctx, d_br, (uint32_t)br.size(), slab, res, costv, prefix_costs, err);
        else launch_bytes<2>(ctx, d_br, (uint32_t)br.size(), slab, res, costv, prefix_costs, err);
        CKL();
    }
    // ---- results
    uint32_t h_err = 0;
    if (mode == 1) {
        CK(cudaMemcpyAsync(out_u32, res, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        ctx->stats.d2h_bytes += (size_t)n * 4;
    } else if (mode == 2) {
        if (out_u32) {                                   // single vector to the host (agcgpu_lz_cost_vector)
            CK(cudaMemcpyAsync(out_u32, costv, (size_t)reqs[0].len * 4, cudaMemcpyDeviceToHost, ctx->st));
            CK(cudaStreamSynchronize(ctx->st));
            ctx->stats.d2h_bytes += (size_t)reqs[0].len * 4;
        }                                                // else: vector i stays at scr_out (u32 index = sum of the lengths before it)
    } else if (!out_offsets) {
        // device-only encode (cost-split path): delta i stays in the slab at last_slab_off[i], its size in scr_sizes[i]
        CK(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        if (h_err) return agc_fail(ctx, AGCGPU_EOVERFLOW, "lz encode: per-segment output bound exceeded");
        ctx->last_slab_off = slab_off;
        cudaEventElapsedTime(&ctx->stats.last_lz_kernel_ms, ctx->ev0, ctx->ev1);
        return 0;
    } else {
        if (int r = agc_reserve(ctx, ctx->scr_offs, ((size_t)n + 1) * 16)) return r;
        uint64_t* d_dst = (uint64_t*)ctx->scr_offs.p;
        uint64_t* d_src = d_dst + (n + 1);
        CK(cudaMemcpyAsync(d_src, slab_off.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->st));
        k_excl_scan_u32<<<1, 1024, 0, ctx->st>>>(res, n, d_dst);
        CKL();
        CK(cudaMemcpyAsync(out_offsets, d_dst, ((size_t)n + 1) * 8, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(&h_err, err, 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        if (h_err) return agc_fail(ctx, AGCGPU_EOVERFLOW, "lz encode: per-segment output bound exceeded");
        uint64_t total = out_offsets[n];
        if (out_bytes && total > out_cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "lz encode: need %llu bytes, caller gave %llu",
                                                          (unsigned long long)total, (unsigned long long)out_cap);
        // out_bytes == nullptr (sharded encode): the dense deltas stay in scr_dense behind `out_cap` bytes the caller fills in
        const uint64_t lead = out_bytes ? 0 : out_cap;
        if (total || lead) {
            if (int r = agc_reserve(ctx, ctx->scr_dense, lead + total + 64)) return r;
            if (total) { k_gather<<<(n + 3) / 4, 128, 0, ctx->st>>>(slab, d_src, d_dst, n, (uint8_t*)ctx->scr_dense.p + lead); CKL(); }
            if (out_bytes && total) {
                CK(cudaMemcpyAsync(out_bytes, ctx->scr_dense.p, total, cudaMemcpyDeviceToHost, ctx->st));
                CK(cudaStreamSynchronize(ctx->st));
            }
        }
        ctx->stats.d2h_bytes += (out_bytes ? total : 0) + ((size_t)n + 1) * 8;
        alg_bytes += total;
    }
    if (mode == 2 && !out_u32) return 0;                 // nothing was synchronised: the caller queues its reduction behind the launch
    cudaEventElapsedTime(&ctx->stats.last_lz_kernel_ms, ctx->ev0, ctx->ev1);
    ctx->stats.lz_alg_bytes = alg_bytes;
    if (mode == 1 && getenv("AGCGPU_TRACE_LZ") && *getenv("AGCGPU_TRACE_LZ")) {
        fprintf(stderr, "[agcgpu] lz estimate call: %u requests, kernels %.3f ms:", n, ctx->stats.last_lz_kernel_ms);
        for (uint32_t i = 0; i < n && i < 12; ++i) fprintf(stderr, " (n %u m %u bound %u)", reqs[i].len, ctx->h_groups[reqs[i].group_id].m, reqs[i].bound);
        fprintf(stderr, "\n");
    }
    if (mode == 0 && getenv("AGCGPU_TRACE_LZ") && *getenv("AGCGPU_TRACE_LZ")) {
        uint32_t mx = 0, mn = ~0u, nrc = 0; uint64_t sum = 0;
        for (uint32_t i = 0; i < n; ++i) { mx = std::max(mx, reqs[i].len); mn = std::min(mn, reqs[i].len); sum += reqs[i].len; nrc += reqs[i].is_rc != 0; }
        fprintf(stderr, "[agcgpu] lz encode call: %u requests (%u rc), %llu bases, len %u..%u, kernels %.3f ms\n", n, nrc, (unsigned long long)sum, mn, mx, ctx->stats.last_lz_kernel_ms);
    }
    if (mode == 0) { ctx->stats.lz_alg_bytes_total += alg_bytes; ctx->stats.lz_kernel_ms_total += ctx->stats.last_lz_kernel_ms; ctx->stats.lz_encode_launches++; }
    return 0;
}

// ------------------------------------------------------------------------------------------------ missing-middle split
// find_cand_segment_with_missing_middle_splitter (agc_compressor.cpp:1540-1610) after the two get_coding_cost calls: v1 (reversed
// when the segment was coded in the other orientation) is cumulated from the left, v2 from the right, and the first position
// with the smallest sum wins.  One CTA per decision; thread t owns the strip [t*S, (t+1)*S).
struct SplitJob { uint64_t off1, off2; uint32_t len, rev1, rev2, pad; };

__global__ void __launch_bounds__(1024) k_split_reduce(const uint32_t* __restrict__ costv, const SplitJob* __restrict__ jobs,
                                                       uint32_t* __restrict__ out_pos, uint32_t* __restrict__ out_sum)
{
    __shared__ uint32_t s_a[1024], s_b[1024];
    __shared__ unsigned long long s_best[32];
    const SplitJob j = jobs[blockIdx.x];
    const uint32_t len = j.len, t = threadIdx.x;
    const uint32_t* v1 = costv + j.off1; const uint32_t* v2 = costv + j.off2;
    const uint32_t S = (len + 1023u) / 1024u;
    const uint32_t lo = min(len, t * S), hi = min(len, lo + S);
    auto u1 = [&](uint32_t i) { return j.rev1 ? v1[len - 1 - i] : v1[i]; };
    auto u2 = [&](uint32_t i) { return j.rev2 ? v2[len - 1 - i] : v2[i]; };
    uint32_t a = 0, b = 0;
    for (uint32_t i = lo; i < hi; ++i) { a += u1(i); b += u2(i); }
    s_a[t] = a; s_b[t] = b;
    __syncthreads();
    // exclusive prefix of the strip sums of v1, exclusive suffix of those of v2 (Hillis-Steele over 1024 entries, u32 wrap-around
    // like the reference's partial_sum on vector<uint32_t>)
    for (uint32_t o = 1; o < 1024; o <<= 1) {
        uint32_t xa = t >= o ? s_a[t - o] : 0u, xb = t + o < 1024 ? s_b[t + o] : 0u;
        __syncthreads();
        s_a[t] += xa; s_b[t] += xb;
        __syncthreads();
    }
    uint32_t run1 = s_a[t] - a, rest2 = s_b[t];          // sum of u1 before the strip; sum of u2 from the strip's first entry on
    uint32_t best = 0xffffffffu, bpos = 0;
    for (uint32_t i = lo; i < hi; ++i) {
        run1 += u1(i);
        const uint32_t cs = run1 + rest2;
        if (cs < best) { best = cs; bpos = i; }
        rest2 -= u2(i);
    }
    // first minimum: smallest (sum, position) pair; a strip that saw nothing keeps (~0, ~0)
    unsigned long long key = lo < hi ? ((unsigned long long)best << 32) | bpos : ~0ull;
    if (lo < hi && best == 0xffffffffu) key = ~0ull;     // "cs < ~0u" never fired: not a candidate (the reference starts from ~0u too)
#pragma unroll
    for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, key, o); key = x < key ? x : key; }
    if ((t & 31) == 0) s_best[t >> 5] = key;
    __syncthreads();
    if (t < 32) {
        key = s_best[t];
#pragma unroll
        for (int o = 16; o; o >>= 1) { unsigned long long x = __shfl_xor_sync(FULL, key, o); key = x < key ? x : key; }
        if (t == 0) {
            if (key == ~0ull) { out_pos[blockIdx.x] = 0; out_sum[blockIdx.x] = 0xffffffffu; }
            else { out_pos[blockIdx.x] = (uint32_t)key; out_sum[blockIdx.x] = (uint32_t)(key >> 32); }
        }
    }
}

int agc_lz_cost_split(agcgpu_ctx* ctx, const agcgpu_split_req* reqs, uint32_t n, uint32_t* out_pos, uint32_t* out_sum)
{
    // sub-batches bounded by the size of the cost vectors (2 x len x 4 bytes per decision)
    const uint64_t budget = 1ull << 30;
    for (uint32_t a = 0; a < n;) {
        uint32_t b = a; uint64_t bytes = 0;
        while (b < n && (b == a || bytes + (uint64_t)reqs[b].len * 8 <= budget)) { bytes += (uint64_t)reqs[b].len * 8; ++b; }
        const uint32_t cnt = b - a;
        std::vector<agcgpu_seg_req> sr(2 * (size_t)cnt);
        std::vector<SplitJob> jobs(cnt);
        uint64_t off = 0;
        for (uint32_t i = 0; i < cnt; ++i) {
            const agcgpu_split_req& q = reqs[a + i];
            for (int h = 0; h < 2; ++h) {
                agcgpu_seg_req& s = sr[2 * (size_t)i + h];
                const uint32_t f = h ? q.flags >> 3 : q.flags;
                s.contig = q.contig; s.start = q.start; s.len = q.len; s.is_rc = f & 1u; s.group_id = h ? q.group2 : q.group1;
                s.bound = (f >> 1) & 1u; s.reserved = 0;
            }
            jobs[i].off1 = off; jobs[i].off2 = off + q.len; jobs[i].len = q.len;
            jobs[i].rev1 = (q.flags >> 2) & 1u; jobs[i].rev2 = (q.flags >> 5) & 1u; jobs[i].pad = 0;
            off += 2ull * q.len;
        }
        // the cost vectors come from the chunk-parallel parse in cost mode (agc_lz_run mode 2) and stay on the device
        if (int r = agc_lz_run(ctx, 2, sr.data(), 2 * cnt, 0, nullptr, 0, nullptr, nullptr)) return r;
        const uint32_t* d_costs = (const uint32_t*)ctx->scr_out.p;
        if (int r = agc_reserve(ctx, ctx->scr_offs, cnt * (sizeof(SplitJob) + 8) + 64)) return r;
        SplitJob* d_jobs = (SplitJob*)ctx->scr_offs.p;
        uint32_t* d_pos = (uint32_t*)(d_jobs + cnt); uint32_t* d_sum = d_pos + cnt;
        CK(cudaMemcpyAsync(d_jobs, jobs.data(), cnt * sizeof(SplitJob), cudaMemcpyHostToDevice, ctx->st));
        k_split_reduce<<<cnt, 1024, 0, ctx->st>>>(d_costs, d_jobs, d_pos, d_sum);
        CKL();
        CK(cudaMemcpyAsync(out_pos + a, d_pos, cnt * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaMemcpyAsync(out_sum + a, d_sum, cnt * 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        if (getenv("AGCGPU_TRACE_SPLIT")) fprintf(stderr, "[agcgpu] cost split: %u decisions, %llu cost entries, lz kernels %.2f ms, sequential segments so far %llu of %llu\n", cnt,
                                                  (unsigned long long)off, ctx->stats.last_lz_kernel_ms, (unsigned long long)ctx->stats.lz_sequential_segments, (unsigned long long)ctx->stats.lz_chunk_segments);
        ctx->stats.d2h_bytes += cnt * 8ull; ctx->stats.h2d_bytes += cnt * sizeof(SplitJob);
        a = b;
    }
    return 0;
}

// LZ-diff encoding split across the ranks of the NCCL communicator: every rank holds the same request list, encodes
// a contiguous share balanced by bases, and the shares are all-gathered device-to-device: block of rank r =
// [u64 offsets of its deltas (count+1)][dense deltas].  One D2H of the gathered blocks on every rank.
int agc_lz_encode_sharded(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_offsets)
{
    const uint32_t W = agc_comm_world(), me = agc_comm_rank();
    std::vector<uint64_t> cum((size_t)n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) cum[i + 1] = cum[i] + reqs[i].len + 64;
    std::vector<uint32_t> cutp(W + 1, n);
    cutp[0] = 0;
    for (uint32_t r = 1; r < W; ++r) cutp[r] = (uint32_t)(std::lower_bound(cum.begin(), cum.end(), cum.back() / W * r) - cum.begin());
    for (uint32_t r = 1; r <= W; ++r) if (cutp[r] < cutp[r - 1]) cutp[r] = cutp[r - 1];
    cutp[W] = n;
    const uint32_t lo = cutp[me], cnt = cutp[me + 1] - lo;
    const uint64_t hdr = ((uint64_t)(cnt + 1) * 8 + 15) / 16 * 16;
    std::vector<uint64_t> my_offs((size_t)cnt + 1, 0);
    uint64_t my_bytes = hdr;
    int local = 0;                                       // a local failure still enters the collective (all ranks fail together)
    if (cnt) {
        local = agc_lz_run(ctx, 0, reqs + lo, cnt, 0, nullptr, hdr, my_offs.data(), nullptr);
        if (!local) my_bytes = hdr + my_offs[cnt];
    } else local = agc_reserve(ctx, ctx->scr_dense, hdr + 64);
    if (!local && cudaMemcpyAsync(ctx->scr_dense.p, my_offs.data(), (size_t)(cnt + 1) * 8, cudaMemcpyHostToDevice, ctx->st) != cudaSuccess) local = AGCGPU_ECUDA;
    std::vector<uint64_t> sizes; uint64_t stride = 0;
    if (int r = agc_comm_allgatherv(ctx, ctx->scr_dense.p, my_bytes, local, sizes, &stride)) return r;
    std::vector<uint8_t> host((size_t)stride * W);
    if (stride) CK(cudaMemcpyAsync(host.data(), ctx->scr_gather.p, (size_t)stride * W, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.d2h_bytes += (size_t)stride * W;
    out_offsets[0] = 0;
    uint64_t o = 0;
    for (uint32_t r = 0; r < W; ++r) {
        const uint32_t c = cutp[r + 1] - cutp[r];
        const uint64_t h = ((uint64_t)(c + 1) * 8 + 15) / 16 * 16;
        if (sizes[r] < h) return agc_fail(ctx, AGCGPU_ECUDA, "sharded encode: truncated block from rank %u", r);
        const uint64_t* offs = (const uint64_t*)(host.data() + (size_t)stride * r);
        if (sizes[r] != h + offs[c]) return agc_fail(ctx, AGCGPU_ECUDA, "sharded encode: block of rank %u has the wrong size", r);
        if (o + offs[c] > out_cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "lz encode: output buffer too small");
        for (uint32_t i = 0; i < c; ++i) out_offsets[cutp[r] + i + 1] = o + offs[i + 1];
        if (offs[c]) memcpy(out + o, host.data() + (size_t)stride * r + h, offs[c]);
        o += offs[c];
    }
    return 0;
}

// store_in_archive(ref) without the zstd call (segment.h:218-255)
int agc_pack_refs(agcgpu_ctx* ctx, const uint32_t* group_ids, uint32_t n, uint8_t* out, uint64_t out_cap,
                  uint64_t* out_offsets, uint8_t* out_use_tuples)
{
    if (n == 0) { out_offsets[0] = 0; return 0; }
    std::vector<uint64_t> off(n + 1, 0);
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t gid = group_ids[i];
        if (gid >= ctx->h_groups.size() || !(ctx->h_groups[gid].flags & GRF_PRESENT))
            return agc_fail(ctx, AGCGPU_EINVAL, "pack_ref: group %u has no reference", gid);
    }
    if (int r = agc_reserve(ctx, ctx->scr_misc, (size_t)n * 32 + 256)) return r;
    uint32_t* d_ids = (uint32_t*)ctx->scr_misc.p;
    uint64_t* d_off = (uint64_t*)((uint8_t*)ctx->scr_misc.p + (((size_t)n * 4 + 15) / 16) * 16);
    uint8_t* d_use = (uint8_t*)(d_off + n + 1);
    uint8_t* d_max = d_use + n;
    CK(cudaMemcpyAsync(d_ids, group_ids, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemsetAsync(d_use, 0, (size_t)2 * n, ctx->st));
    k_ref_probe_packed<<<n, 32, 0, ctx->st>>>((const GroupRefDev*)ctx->d_groups.p, d_ids, n, d_use);
    CKL();
    for (uint32_t i = 0; i < n; ++i) {
        const GroupRefDev& g = ctx->h_groups[group_ids[i]];
        if (g.flags & GRF_DIRTY) { k_ref_probe_bytes<<<1, 32, 0, ctx->st>>>(g.codes, g.m, d_use + i, d_max + i); CKL(); }
    }
    std::vector<uint8_t> h_max(n);
    CK(cudaMemcpyAsync(out_use_tuples, d_use, n, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(h_max.data(), d_max, n, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    uint32_t max_m = 0;
    auto tuple_mode = [&](uint32_t i, uint32_t& nb, uint32_t& mult) {       // bytes2tuples (segment.h:73-91)
        uint8_t me = (ctx->h_groups[group_ids[i]].flags & GRF_DIRTY) ? h_max[i] : 0;
        if (me < 4) { nb = 4; mult = 4; } else if (me < 6) { nb = 3; mult = 6; } else if (me < 16) { nb = 2; mult = 16; } else { nb = 1; mult = 0; }
    };
    for (uint32_t i = 0; i < n; ++i) {
        const GroupRefDev& g = ctx->h_groups[group_ids[i]];
        uint32_t nb, mult; tuple_mode(i, nb, mult);
        uint64_t sz = !out_use_tuples[i] ? (uint64_t)g.m : (nb == 1 ? (uint64_t)g.m + 1 : (uint64_t)g.m / nb + 2);
        off[i + 1] = off[i] + sz;
        max_m = std::max(max_m, g.m);
    }
    if (off[n] > out_cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "pack_ref: need %llu bytes", (unsigned long long)off[n]);
    if (int r = agc_reserve(ctx, ctx->scr_dense, off[n] + 64)) return r;
    uint8_t* dense = (uint8_t*)ctx->scr_dense.p;
    CK(cudaMemcpyAsync(d_off, off.data(), ((size_t)n + 1) * 8, cudaMemcpyHostToDevice, ctx->st));
    for (uint32_t y0 = 0; y0 < n; y0 += 32768) {
        uint32_t ny = std::min<uint32_t>(32768, n - y0);
        dim3 grid((max_m / 4 + 2 + 255) / 256, ny);
        k_ref_tuples_packed<<<grid, 256, 0, ctx->st>>>((const GroupRefDev*)ctx->d_groups.p, d_ids + y0, ny, d_off + y0, dense);
        CKL();
    }
    for (uint32_t i = 0; i < n; ++i) {
        const GroupRefDev& g = ctx->h_groups[group_ids[i]];
        const bool dirty = g.flags & GRF_DIRTY;
        if (!out_use_tuples[i]) {          // failed the probe: stored as raw symbols (level 19)
            if (!g.m) continue;
            if (dirty) CK(cudaMemcpyAsync(dense + off[i], g.codes, g.m, cudaMemcpyDeviceToDevice, ctx->st));
            else { k_expand_ref<<<(g.m + 255) / 256, 256, 0, ctx->st>>>(g.packed, g.m, dense + off[i], 0); CKL(); }
        } else if (dirty) {
            uint32_t nb, mult; tuple_mode(i, nb, mult);
            uint32_t total = nb == 1 ? g.m + 1 : g.m / nb + 2;
            k_ref_tuples_bytes<<<(total + 255) / 256, 256, 0, ctx->st>>>(g.codes, g.m, nb, mult, dense + off[i]);
            CKL();
        }
    }
    CK(cudaMemcpyAsync(out, dense, off[n], cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.d2h_bytes += off[n] + n;
    memcpy(out_offsets, off.data(), ((size_t)n + 1) * 8);
    return 0;
}
