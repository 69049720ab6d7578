// lz_diag_core.cuh -- CLZDiff_V2::Encode (src/common/lz_diff.cpp:669-798) as a warp-per-segment parse that STREAMS along the
// current diagonal.  Host/device source: kernels_lz_diag.cu instantiates it on the device (32 lanes); tests/lzd_host builds the
// same source for the CPU (the 32 lanes of every phase run one after the other) so the test suite can compare it with the oracle
// without a GPU (test infrastructure; the product has no host path).
//
// The sequential parse alternates two kinds of work.  (1) Between two matches it probes text positions one by one until the hash
// index yields a candidate that extends to more than min_match_len symbols.  (2) It extends that match until the first mismatch.
// For a segment that resembles its reference (a sample of the same species: SNPs every 100..1000 bases, an indel every few kb)
// nearly all tokens lie on ONE diagonal (match_pos - text_pos) for thousands of bases, and what happens after a mismatch at q is
// fully determined by (a) the positions of the following mismatches on that diagonal and (b) what the hash index returns for the
// few positions q, q+1, ... up to the next indexed position (a multiple of 4 on the reference) whose k-mer is clean:
//   literal(q) [literals up to the candidate] -> candidate on the same diagonal, backward extension takes the literals back up
//   to the last mismatch -> match until the next mismatch.
// So a warp takes a WINDOW of LZD_WIN text positions on the diagonal of the match it just found and
//   1. compares window and reference (XOR of 2-bit packed words, 32 bases per lane and step) into a mismatch bitmap,
//   2. lists the mismatch positions (each is a possible state "match ended here, no pending literals"),
//   3. resolves every state in parallel: REAL probes of the hash index for q, q+1, ... (a group of lanes per state) until a
//      candidate on the diagonal is accepted (b + f > min_match_len decided from the mismatch list) -- anything else the index
//      returns (a candidate on another diagonal, several candidates, nothing within LZD_MAX_T positions) makes the state a STOP,
//   4. follows the chain of states from the window's first mismatch and writes their tokens (literals with the '!' rewrite of
//      lz_diff.cpp:769-779, "0,len." matches) side by side.
// A STOP state, the end of text or reference, or a region too dense to list hands the exact state (i, pred_pos, 0 literals) back
// to the general round: 32 lanes probe 32 consecutive positions, one lane walks them in order with the sequential code
// (lzc_best_match_full for several candidates), the first accepted candidate opens the next diagonal.  Every decision is taken
// from the same comparisons and index reads the sequential code makes, so the delta is the reference's byte string.
#pragma once
#include "lz_chunk_core.cuh"

#ifdef __CUDA_ARCH__
#define LZD_LANE (threadIdx.x & 31u)
#define LZD_STEP 32u
#define LZD_SYNC() __syncwarp()
#define LZD_IS0 ((threadIdx.x & 31u) == 0u)
#else
#define LZD_LANE 0u
#define LZD_STEP 1u
#define LZD_SYNC() ((void)0)
#define LZD_IS0 true
#endif
// every lane of the warp runs the body once with L = its lane id (host: L = 0..31 in turn)
#define LZD_LANES(L) for (uint32_t L = LZD_LANE; L < 32u; L += LZD_STEP)
// items k = 0..n-1 dealt to the lanes
#define LZD_ITEMS(k, n) for (uint32_t k = LZD_LANE; k < (n); k += LZD_STEP)

#ifndef LZD_ITERS
#define LZD_ITERS 32u              // steps of 2048 bases (32 lanes x one 64-base block) per window at most
#endif
#define LZD_MAXM 128u              // mismatches listed per window (a denser window ends at the LZD_MAXM-th)
#define LZD_FULL 96u               // a window stops taking steps once it lists this many
#define LZD_MAX_T 48u              // positions probed behind a mismatch before the state is handed to the general round
#define LZD_MAX_ROUNDS 40u

enum { LZD_PENDING = 0, LZD_ACCEPT = 1, LZD_STOP = 2, LZD_DEFER = 3 };
enum { LZD_END_OPEN = 0xffffu, LZD_END_LIMIT = 0xfffeu };

struct LzdScratch {                // per warp (shared memory on the device)
    uint32_t mis[LZD_MAXM];        // mismatch positions of the window, ascending
    uint32_t ts[LZD_MAXM];         // state j accepted: start of its match
    uint16_t endi[LZD_MAXM];       // ... index of the mismatch that ends it (LZD_END_OPEN / LZD_END_LIMIT)
    uint16_t tb[LZD_MAXM];         // ... bytes of its token
    uint16_t off[LZD_MAXM];        // ... on the chain: byte offset of the token in the output of this window
    uint8_t stat[LZD_MAXM];
    uint8_t cur[LZD_MAXM];         // next offset to probe behind the mismatch
    uint8_t plist[LZD_MAXM];       // states still pending (two lists, used in turn)
    uint8_t plist2[LZD_MAXM];
    uint8_t path[LZD_MAXM];
    uint32_t cnt[32];              // probe results of a round
    uint32_t hp[32];
    uint32_t flag, n_pend, n_path, scan_acc;
    // result of a general round / of a window's chain walk (written by lane 0, read by all)
    uint32_t r_kind, r_i, r_pred, r_np, r_ts, r_mp, r_len, r_scan, r_lit0, r_litn, r_bang;
    int32_t r_dif;
};

LZC_HD uint32_t lzd_popc(uint32_t x)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__popc(x);
#else
    return (uint32_t)__builtin_popcount(x);
#endif
}
LZC_HD uint32_t lzd_ffs0(uint32_t x)             // index of the lowest set bit (x != 0)
{
#ifdef __CUDA_ARCH__
    return (uint32_t)__ffs((int)x) - 1u;
#else
    return (uint32_t)__builtin_ctz(x);
#endif
}
// x: XOR of two 32-base windows (base 0 in bits 63:62) -> bit b set iff base b differs
LZC_HD uint32_t lzd_mismatch_bits(uint64_t x)
{
    if (!x) return 0u;
    uint64_t y = (x | (x >> 1)) & 0x5555555555555555ULL;
    y = (y | (y >> 1)) & 0x3333333333333333ULL;
    y = (y | (y >> 2)) & 0x0F0F0F0F0F0F0F0FULL;
    y = (y | (y >> 4)) & 0x00FF00FF00FF00FFULL;
    y = (y | (y >> 8)) & 0x0000FFFF0000FFFFULL;
    y = (y | (y >> 16)) & 0x00000000FFFFFFFFULL;
    uint32_t v = (uint32_t)y;                    // bit k = group k from the least significant end = base 31 - k
#ifdef __CUDA_ARCH__
    return __brev(v);
#else
    v = ((v >> 1) & 0x55555555u) | ((v & 0x55555555u) << 1);
    v = ((v >> 2) & 0x33333333u) | ((v & 0x33333333u) << 2);
    v = ((v >> 4) & 0x0F0F0F0Fu) | ((v & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(v);
#endif
}

// one text position against the index: number of slots (0, 1, 2 = "two or more") whose key equals the text's, hp0 = the first one
// (the candidate loop of find_best_match16/32, lz_diff.cpp:297-326: slots until the first empty one, at most 64)
template <class View>
LZC_HD uint32_t lzd_probe(const View& a, uint32_t p, uint32_t kl, uint32_t& hp0)
{
    const uint64_t x = a.twin(p) >> (64 - 2 * kl);
    const uint32_t h = (uint32_t)lzc_murmur64(x) & a.mask;
    const uint32_t xq = kl >= 16u ? (uint32_t)(x >> (2u * kl - 32u)) : (uint32_t)(x << (32u - 2u * kl));
    uint32_t ncand = 0;
    for (uint32_t t = 0; t < 64; ++t) {
        const uint32_t sv = a.slot((h + t) & a.mask);
        if (sv == AGC_EMPTY32) break;
        const uint32_t dq = a.rquick(sv) ^ xq;
        if (kl >= 16u ? dq != 0u : (dq >> (32u - 2u * kl)) != 0u) continue;
        if (kl > 16u && (a.rwin(sv * 4u) >> (64 - 2 * kl)) != x) continue;
        if (!ncand) hp0 = sv * 4u;
        if (++ncand > 1) break;
    }
    return ncand;
}

struct LzdOut { uint8_t* out; uint32_t olen, cap; bool ovf; };
#if defined(LZD_COUNTERS) && !defined(__CUDA_ARCH__)
struct LzdCounters { unsigned long long rounds, windows, path_tokens, stops, defers, opens, multi, probes; };
static LzdCounters g_lzd_cnt;
#define LZD_COUNT(f, v) (g_lzd_cnt.f += (v))
#else
#define LZD_COUNT(f, v) ((void)0)
#endif

// literals of text positions [start, start + count), with the '!' rewrite when the match that follows continues the predicted
// position (lz_diff.cpp:769-779): d = distance back from the match
template <class View>
LZC_HD void lzd_emit_literals(const View& a, LzdOut& o, uint32_t start, uint32_t count, bool bang, uint32_t mp)
{
    if (!count) return;
    if ((uint64_t)o.olen + count > o.cap) { o.ovf = true; o.olen += count; return; }
    LZD_ITEMS(j, count) {
        const uint32_t sy = a.tsym(start + j);
        uint8_t ch = (uint8_t)('A' + sy);
        if (bang) {
            const uint32_t d = count - j;
            if (d < o.olen + count && d < mp && sy == a.rsym(mp - d)) ch = '!';
        }
        o.out[o.olen + j] = ch;
    }
    o.olen += count;
    LZD_SYNC();
}
// encode_match (lz_diff.cpp:631-643); every lane gets the new length, lane 0 writes
LZC_HD void lzd_emit_match(LzdOut& o, int64_t dif, bool with_len, uint32_t lenv)
{
    uint8_t buf[24];
    const uint32_t L = lzc_put_match(buf, dif, with_len, lenv);
    if ((uint64_t)o.olen + L > o.cap) o.ovf = true;
    else if (LZD_IS0) for (uint32_t k = 0; k < L; ++k) o.out[o.olen + k] = buf[k];
    o.olen += L;
}

// "equal sequences" (lz_diff.cpp:678-680): n == m and every symbol equal
template <class View>
LZC_HD bool lzd_equal(const View& a, LzdScratch& S)
{
    const uint32_t n = a.n;
    if (LZD_IS0) S.flag = 0;
    LZD_SYNC();
    for (uint32_t w0 = 0; w0 < n; w0 += 1024u) {                 // (unequal segments leave after a step or two)
        LZD_LANES(L) {
            const uint32_t p = w0 + 32u * L;
            if (p < n) {
                uint64_t x = a.twin(p) ^ a.rwin(p);
                const uint32_t valid = n - p;
                if (valid < 32u) x &= ~0ull << (64u - 2u * valid);
                if (x) S.flag = 1;
            }
        }
        LZD_SYNC();
        if (S.flag) return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------ the general round
// 32 lanes probe positions i .. i+31, lane 0 walks them with the sequential semantics.  Results in S.r_*:
//   r_kind 0: all probed positions are literals (state advanced)
//          1: a single candidate was accepted; it is verified up to r_scan and left OPEN (its end is found by the windows)
//          2: a match chosen among several candidates, complete (r_len)
// r_i / r_pred / r_np: the state before the match's literals are written (np pending literals ending at r_ts);
// literals to write: [r_lit0, r_lit0 + r_litn), r_bang; the match: r_ts, r_mp, r_dif.
template <class View>
LZC_HD void lzd_general_round(const View& a, LzdScratch& S, uint32_t mml, uint32_t i, uint32_t pred, uint32_t np)
{
    const uint32_t kl = mml - 3u, n = a.n, m = a.m;
    const uint32_t cnt = lzc_min(32u, n - kl - i);                  // positions p with p + kl < n
    LZD_COUNT(rounds, 1);
    LZD_LANES(L) {
        uint32_t hp0 = 0, nc = 0;
        if (L < cnt) nc = lzd_probe(a, i + L, kl, hp0);
        S.cnt[L] = nc; S.hp[L] = hp0;
    }
    LZD_SYNC();
    if (LZD_IS0) {
        uint32_t kind = 0, L = 0;
        for (; L < cnt; ++L) {
            const uint32_t nc = S.cnt[L];
            if (nc == 0) { ++np; ++pred; continue; }
            const uint32_t p = i + L;
            if (nc == 1) {
                const uint32_t hp = S.hp[L];
                const uint32_t maxlen = lzc_min(n - p, m - hp);
                const uint32_t fcap = a.lcp_fwd(p, hp, lzc_min(maxlen, mml + 1u));
                const uint32_t lim = lzc_min(np, hp);
                const uint32_t b = lim ? a.lcp_bwd(p, hp, lim) : 0u;
                if (b + fcap > mml) {
                    np -= b; pred -= b;
                    S.r_ts = p - b; S.r_mp = hp - b; S.r_scan = p + fcap; S.r_len = 0;
                    kind = 1;
                    break;
                }
                ++np; ++pred;
                continue;
            }
            const uint64_t x = a.twin(p) >> (64 - 2 * kl);
            const uint32_t h = (uint32_t)lzc_murmur64(x) & a.mask;
            uint32_t hp, b, f; bool npl;
            if (!lzc_best_match_full(a, h, x, p, np, kl, mml, hp, b, f, npl)) { ++np; ++pred; continue; }
            np -= b; pred -= b;
            S.r_ts = p - b; S.r_mp = hp - b; S.r_len = b + f; S.r_scan = 0;
            kind = 2;
            LZD_COUNT(multi, 1);
            break;
        }
        S.r_kind = kind;
        if (kind == 0) { S.r_i = i + cnt; S.r_pred = pred; S.r_np = np; }
        else {
            S.r_i = S.r_ts; S.r_pred = pred; S.r_np = np;
            S.r_lit0 = S.r_ts - np; S.r_litn = np; S.r_bang = (S.r_mp == pred) ? 1u : 0u;
            S.r_dif = (int32_t)S.r_mp - (int32_t)pred;
        }
    }
    LZD_SYNC();
}

// ------------------------------------------------------------------------------------------------ windows along a diagonal
// exclusive prefix sum of v over the lanes, inside an LZD_LANES body (every lane calls it); `total` is valid after the body
LZC_HD uint32_t lzd_scan(LzdScratch& S, uint32_t v, uint32_t L, uint32_t& total)
{
#ifdef __CUDA_ARCH__
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if (L >= (uint32_t)o) inc += t; }
    total = __shfl_sync(0xffffffffu, inc, 31);
    return inc - v;
#else
    if (L == 0) S.scan_acc = 0;
    const uint32_t r = S.scan_acc;
    S.scan_acc += v; total = S.scan_acc;
    return r;
#endif
}
LZC_HD bool lzd_any(uint32_t pred)               // device: warp vote; host: "maybe" (the callers only use it to skip work)
{
#ifdef __CUDA_ARCH__
    return __any_sync(0xffffffffu, pred) != 0;
#else
    (void)pred; return true;
#endif
}

// Entry: an open match (o_ts, o_mp; its literals are written, its token "dif[,len]." is not) verified up to text position scan.
// Exit: the state (i, pred, 0 pending literals) where the general round has to go on; every token before it is written.
template <class View>
LZC_HD void lzd_follow_diagonal(const View& a, LzdScratch& S, uint32_t mml, LzdOut& o, uint32_t o_ts, uint32_t o_mp, int32_t o_dif,
                                uint32_t scan, uint32_t& i_out, uint32_t& pred_out)
{
    const uint32_t kl = mml - 3u, n = a.n, m = a.m;
    const int64_t d = (int64_t)o_mp - (int64_t)o_ts;
    const uint32_t lim_pos = (uint32_t)((int64_t)n < (int64_t)m - d ? (int64_t)n : (int64_t)m - d);    // the diagonal ends here
    bool open = true;                            // a match whose token is still to be written (o_ts, o_dif)
    uint32_t w0 = scan;
    for (;;) {
        LZD_COUNT(windows, 1);
        // ---- 1. compare text and reference from w0 on, 64-base blocks of the packed store (one 16-byte load per lane and step),
        // and list the mismatch positions in order; the window ends at the end of the diagonal, after LZD_ITERS steps or when the
        // list is (nearly) full
        uint32_t nm = 0, wend = w0;
        bool at_limit = (w0 >= lim_pos);
        if (!at_limit) {
            uint64_t tw0, tw1; int32_t rel0;
            a.tblock(w0, tw0, tw1, rel0);
            const int64_t b0 = (int64_t)w0 + rel0;               // text position of the first base of the block that holds w0
            for (uint32_t it = 0; it < LZD_ITERS; ++it) {
                uint32_t total = 0;
                LZD_LANES(L) {
                    const int64_t bs = b0 + 64 * (int64_t)(it * 32u + L);
                    // y0 / y1: one bit (the lower of its two) per differing base of the block's first / second 32 bases
                    uint64_t y0 = 0, y1 = 0;
                    if (bs < (int64_t)lim_pos) {
                        uint64_t x0, x1, r0, r1; int32_t rel;
                        const uint32_t pp = bs < (int64_t)w0 ? w0 : (uint32_t)bs;
                        a.tblock(pp, x0, x1, rel);
                        a.rwin2(bs + d, r0, r1);
                        x0 ^= r0; x1 ^= r1;
                        if (x0 | x1) {
                            y0 = (x0 | (x0 >> 1)) & 0x5555555555555555ULL; y1 = (x1 | (x1 >> 1)) & 0x5555555555555555ULL;
                            if (bs < (int64_t)w0 || bs + 64 > (int64_t)lim_pos) {          // first / last block: bases outside [w0, lim_pos)
                                const uint32_t lo = bs < (int64_t)w0 ? (uint32_t)((int64_t)w0 - bs) : 0u;
                                const int64_t hi64 = (int64_t)lim_pos - bs;
                                const uint32_t hi = hi64 < 64 ? (uint32_t)hi64 : 64u;
                                // base b of a word sits at bit 62 - 2b
                                if (lo >= 32u) { y0 = 0; if (lo > 32u) y1 &= ~0ull >> (2u * (lo - 32u)); } else if (lo) y0 &= ~0ull >> (2u * lo);
                                if (hi <= 32u) { y1 = 0; if (hi < 32u) y0 &= ~(~0ull >> (2u * hi)); } else if (hi < 64u) y1 &= ~(~0ull >> (2u * (hi - 32u)));
                            }
                        }
                    }
                    const uint32_t c = lzd_popc((uint32_t)y0) + lzd_popc((uint32_t)(y0 >> 32)) + lzd_popc((uint32_t)y1) + lzd_popc((uint32_t)(y1 >> 32));
                    if (lzd_any(c != 0u)) {
                        uint32_t at = nm + lzd_scan(S, c, L, total);
                        while (y0 && at < LZD_MAXM) { const uint32_t z = lzc_clz64(y0); y0 &= ~(0x8000000000000000ULL >> z); S.mis[at++] = (uint32_t)(bs + (z >> 1)); }
                        while (y1 && at < LZD_MAXM) { const uint32_t z = lzc_clz64(y1); y1 &= ~(0x8000000000000000ULL >> z); S.mis[at++] = (uint32_t)(bs + 32 + (z >> 1)); }
                    }
                }
                nm += total;
                const int64_t covered = b0 + 64 * (int64_t)((it + 1u) * 32u);
                wend = covered >= (int64_t)lim_pos ? lim_pos : (uint32_t)covered;
                if (wend == lim_pos) { at_limit = true; break; }
                if (nm >= LZD_FULL) break;
            }
            LZD_SYNC();
            if (nm >= LZD_MAXM) { nm = LZD_MAXM - 1u; wend = S.mis[LZD_MAXM - 1u]; at_limit = false; }   // dense: the window ends at a known mismatch
        }
        if (nm == 0) {
            if (!at_limit) { w0 = wend; continue; }                 // nothing happens in this window: the open match runs on
            if (open) {                                             // the open match reaches the end of the diagonal
                const uint32_t len = lim_pos - o_ts;
                const bool to_end = (o_ts + len == n) && (o_mp + len == m);
                lzd_emit_match(o, (int64_t)o_dif, !to_end, len - mml);
            }
            i_out = lim_pos; pred_out = (uint32_t)((int64_t)lim_pos + d);
            return;
        }
        // ---- 2. resolve the states: state j = "a match ended at mis[j], no pending literals".  One offset per state and round
        // while many states are pending; then the few left (mismatches close to each other) get a group of lanes each.
        LZD_ITEMS(j, nm) { S.stat[j] = LZD_PENDING; S.cur[j] = 0; }
        LZD_SYNC();
        // decision for state j at offset t given the probe result r (0 none, 1 one candidate on the diagonal, 2 anything else,
        // 3 not probed: beyond what this window knows); returns the new status
        auto decide = [&](uint32_t j, uint32_t t, uint32_t r) -> uint32_t {
            const uint32_t p = S.mis[j] + t;
            if (r == 3u) return (p + kl >= lim_pos) ? LZD_STOP : LZD_DEFER;     // end of the diagonal: the general round finishes
            if (r == 2u) return LZD_STOP;
            if (r == 0u) return LZD_PENDING;                                    // literal
            // a candidate on the diagonal: its k-mer is clean, so the last mismatch lies before p
            uint32_t jj = j;
            while (jj + 1u < nm && S.mis[jj + 1u] < p) ++jj;
            const uint32_t b = p - S.mis[jj] - 1u;
            if (jj + 1u < nm) {
                if (b + (S.mis[jj + 1u] - p) > mml) { S.ts[j] = S.mis[jj] + 1u; S.endi[j] = (uint16_t)(jj + 1u); return LZD_ACCEPT; }
                return LZD_PENDING;
            }
            if (at_limit) {
                if (b + (lim_pos - p) > mml) { S.ts[j] = S.mis[jj] + 1u; S.endi[j] = (uint16_t)LZD_END_LIMIT; return LZD_ACCEPT; }
                return LZD_PENDING;
            }
            if (b + (wend - p) > mml) { S.ts[j] = S.mis[jj] + 1u; S.endi[j] = (uint16_t)LZD_END_OPEN; return LZD_ACCEPT; }
            return LZD_DEFER;
        };
        auto probe_at = [&](uint32_t p) -> uint32_t {
            if (!(p + kl < lim_pos && p < wend)) return 3u;
            uint32_t hp0 = 0;
            const uint32_t nc = lzd_probe(a, p, kl, hp0);
            LZD_COUNT(probes, 1);
            return nc == 0 ? 0u : (nc == 1u && (int64_t)hp0 == (int64_t)p + d) ? 1u : 2u;
        };
        // pending states live in a list that is compacted after every pass (plist / plist2 in turn), so the lanes stay dense
        uint8_t* cur_list = S.plist; uint8_t* nxt_list = S.plist2;
        LZD_ITEMS(j, nm) cur_list[j] = (uint8_t)j;
        LZD_SYNC();
        uint32_t npend = nm;
        for (uint32_t round = 0; round < LZD_MAX_ROUNDS && npend; ++round) {
            const uint32_t G = npend > 16u ? 1u : npend > 8u ? 2u : npend > 4u ? 4u : npend > 2u ? 8u : npend > 1u ? 16u : 32u;   // lanes (offsets) per state
            const uint32_t per = 32u / G;
            uint32_t kept = 0;
            for (uint32_t g0 = 0; g0 < npend; g0 += per) {
                if (G > 1u) {
                    LZD_LANES(L) {
                        uint32_t r = 3u;
                        const uint32_t s = g0 + L / G;
                        if (s < npend) { const uint32_t j = cur_list[s]; r = probe_at(S.mis[j] + S.cur[j] + (L % G)); }
                        S.cnt[L] = r;
                    }
                    LZD_SYNC();
                }
                uint32_t total = 0;
                LZD_LANES(L) {
                    const uint32_t s = g0 + L / G;
                    uint32_t keep = 0, j = 0;
                    if ((L % G) == 0u && s < npend) {
                        j = cur_list[s];
                        uint32_t t = S.cur[j], st = LZD_PENDING;
                        if (G == 1u) { st = decide(j, t, probe_at(S.mis[j] + t)); ++t; }
                        else for (uint32_t g = 0; g < G && st == LZD_PENDING; ++g, ++t) st = decide(j, t, S.cnt[L + g]);
                        if (st == LZD_PENDING && t >= LZD_MAX_T) st = LZD_STOP;
                        S.cur[j] = (uint8_t)t; S.stat[j] = (uint8_t)st;
                        keep = st == LZD_PENDING ? 1u : 0u;
                    }
                    const uint32_t at = kept + lzd_scan(S, keep, L, total);
                    if (keep) nxt_list[at] = (uint8_t)j;
                }
                kept += total;
                LZD_SYNC();
            }
            npend = kept;
            uint8_t* tmp = cur_list; cur_list = nxt_list; nxt_list = tmp;
        }
        // ---- 3. bytes of every accepted state's token: literals [mis[j], ts[j]) + "0" [",len"] "."
        LZD_ITEMS(j, nm) {
            uint32_t tb = 0;
            if (S.stat[j] == LZD_ACCEPT) {
                tb = S.ts[j] - S.mis[j];
                const uint32_t e = S.endi[j];
                if (e != LZD_END_OPEN) {
                    const uint32_t endp = e == LZD_END_LIMIT ? lim_pos : S.mis[e];
                    const bool to_end = (endp == n) && ((int64_t)endp + d == (int64_t)m);
                    tb += 1u + (to_end ? 0u : 1u + lzc_int_len(endp - S.ts[j] - mml)) + 1u;
                }
            }
            S.tb[j] = (uint16_t)tb;
        }
        LZD_SYNC();
        // ---- 4. the chain of states from the window's first mismatch: lane 0 follows the pointers, the offsets are a prefix sum
        if (LZD_IS0) {
            uint32_t k = 0, j = 0, kind = 0;     // kind: 1 stop at state j (STOP / DEFER / still pending), 2 open at the window's end, 3 diagonal ended
            for (;;) {
                if (S.stat[j] != LZD_ACCEPT) { kind = 1; break; }
                S.path[k++] = (uint8_t)j;
                const uint32_t e = S.endi[j];
                if (e >= LZD_END_LIMIT) { kind = e == LZD_END_OPEN ? 2u : 3u; break; }
                j = e;
            }
            S.n_path = k; S.r_kind = kind; S.r_i = j;
            LZD_COUNT(path_tokens, k);
            if (kind == 1) { if (S.stat[j] == LZD_DEFER) LZD_COUNT(defers, 1); else LZD_COUNT(stops, 1); }
            if (kind == 2) LZD_COUNT(opens, 1);
        }
        LZD_SYNC();
        uint32_t bytes = 0;
        {
            const uint32_t npath0 = S.n_path;
            for (uint32_t c0 = 0; c0 < npath0; c0 += 32u) {
                uint32_t total = 0;
                LZD_LANES(L) {
                    const uint32_t c = c0 + L;
                    const uint32_t jj = c < npath0 ? S.path[c] : 0u;
                    const uint32_t v = c < npath0 ? (uint32_t)S.tb[jj] : 0u;
                    const uint32_t at = bytes + lzd_scan(S, v, L, total);
                    if (c < npath0) S.off[jj] = (uint16_t)at;
                }
                bytes += total;
            }
        }
        LZD_SYNC();
        // the open match ends at the first mismatch of the window
        if (open) {
            const uint32_t len = S.mis[0] - o_ts;
            lzd_emit_match(o, (int64_t)o_dif, true, len - mml);       // (not "to the end": a mismatch follows)
            open = false;
        } else if (S.mis[0] != w0) {
            // cannot happen: a window without an open match starts at a state
            i_out = w0; pred_out = (uint32_t)((int64_t)w0 + d); return;
        }
        const uint32_t npath = S.n_path, kind = S.r_kind, jstop = S.r_i;
        if ((uint64_t)o.olen + bytes > o.cap) o.ovf = true;
        else {
            LZD_ITEMS(c, npath) {
                const uint32_t jj = S.path[c];
                uint8_t* dst = o.out + o.olen + S.off[jj];
                const uint32_t tsj = S.ts[jj];
                uint32_t nx = jj;
                for (uint32_t p = S.mis[jj]; p < tsj; ++p) {
                    if (nx < nm && S.mis[nx] == p) { *dst++ = (uint8_t)('A' + a.tsym(p)); ++nx; }
                    else *dst++ = (uint8_t)'!';
                }
                const uint32_t e = S.endi[jj];
                if (e != LZD_END_OPEN) {
                    const uint32_t endp = e == LZD_END_LIMIT ? lim_pos : S.mis[e];
                    const bool to_end = (endp == n) && ((int64_t)endp + d == (int64_t)m);
                    lzc_put_match(dst, 0, !to_end, endp - tsj - mml);
                }
            }
        }
        o.olen += bytes;
        LZD_SYNC();
        if (kind == 2) {                         // the last state's match runs beyond the window
            const uint32_t jj = S.path[npath - 1u];
            open = true; o_ts = S.ts[jj]; o_mp = (uint32_t)((int64_t)o_ts + d); o_dif = 0;
            w0 = wend;
            continue;
        }
        if (kind == 3) { i_out = lim_pos; pred_out = (uint32_t)((int64_t)lim_pos + d); return; }
        // stop at state jstop: DEFER = the window could not decide it -> a new window from there; otherwise back to the general round
        const uint32_t qs = S.mis[jstop];
        if (S.stat[jstop] == LZD_DEFER && qs > w0) { w0 = qs; continue; }
        i_out = qs; pred_out = (uint32_t)((int64_t)qs + d);
        return;
    }
}

// ------------------------------------------------------------------------------------------------ one segment
// returns the delta length, -2 = out_cap too small
template <class View>
LZC_HD int64_t lzd_encode_segment(const View& a, LzdScratch& S, uint32_t mml, uint8_t* out, uint32_t cap)
{
    const uint32_t kl = mml - 3u, n = a.n, m = a.m;
    if (n == m && lzd_equal(a, S)) return 0;
    LzdOut o; o.out = out; o.olen = 0; o.cap = cap; o.ovf = false;
    uint32_t i = 0, pred = 0, np = 0;
    while (i + kl < n) {
        lzd_general_round(a, S, mml, i, pred, np);
        const uint32_t kind = S.r_kind;
        i = S.r_i; pred = S.r_pred; np = S.r_np;
        if (kind == 0) continue;
        const uint32_t ts = S.r_ts, mp = S.r_mp, len = S.r_len, scan = S.r_scan;
        const int32_t dif = S.r_dif;
        lzd_emit_literals(a, o, S.r_lit0, S.r_litn, S.r_bang != 0, mp);
        np = 0;
        if (kind == 2) {
            const bool to_end = (ts + len == n) && (mp + len == m);
            lzd_emit_match(o, (int64_t)dif, !to_end, len - mml);
            pred = mp + len; i = ts + len;
            continue;
        }
        lzd_follow_diagonal(a, S, mml, o, ts, mp, dif, scan, i, pred);
    }
    np += n - i;
    lzd_emit_literals(a, o, n - np, np, false, 0);
    return o.ovf ? -2 : (int64_t)o.olen;
}
