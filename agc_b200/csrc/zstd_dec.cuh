// zstd_dec.cuh -- decoder for zstd frames (RFC 8878), the inverse of the residual coder: what ZSTD_decompressDCtx of the
// vendored zstd 1.5.5 does for CSegment::unpack / get (src/common/segment.cpp:500-577, 220-399) and
// CCollection_V3::zstd_decompress.  Written once for host and device (ZD_FN): kernels_zstd.cu builds the device side
// (one thread decodes one frame; frames of a batch run side by side), tests/zstd_host builds the same source for the
// CPU suite, which checks it against frames written by the reference's libzstd at every level AGC uses.
//
// Covered: single-segment and windowed frames without a dictionary, raw / RLE / compressed blocks, raw / RLE / Huffman
// (1 or 4 streams, FSE-compressed or direct weights) / treeless literals, predefined / RLE / FSE / repeat sequence tables,
// repeat offsets, optional content checksum (skipped, not verified).  Returns the decoded size or a negative ZD_E* code.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZD_FN __host__ __device__ inline
#else
#define ZD_FN inline
#endif

namespace zd {

enum { ZD_ESRC = -1 /* truncated or malformed frame */, ZD_EDST = -2 /* output buffer too small */, ZD_EUNSUP = -3 /* dictionary / skippable frame */ };

struct FseEntry { uint16_t base; uint8_t sym; uint8_t nb; };        // next state = base + read(nb)
struct Work {                                                       // ~150 KB per frame in flight
    FseEntry ll[512], of[256], ml[512], wt[64];
    uint16_t huf[2048];                                             // symbol | nbBits << 8, indexed by the next huf_log bits
    uint32_t huf_log;
    int have_huf, have_ll, have_of, have_ml;
    uint32_t ll_log, of_log, ml_log;
    uint32_t rep[3];
    uint16_t next[64];                                              // FSE table construction scratch (symbolNext)
    int16_t norm[64];
    uint8_t lit[(128 << 10) + 32];
};

ZD_FN uint32_t hb(uint32_t v) { uint32_t r = 0; while (v >>= 1) ++r; return r; }

// ---- forward bit reader (FSE table descriptions) ------------------------------------------------------------------
struct FwdBits { const uint8_t* p; const uint8_t* end; uint64_t acc; uint32_t n; };
ZD_FN void fb_init(FwdBits& b, const uint8_t* p, const uint8_t* end) { b.p = p; b.end = end; b.acc = 0; b.n = 0; }
ZD_FN uint32_t fb_peek(FwdBits& b, uint32_t bits) { while (b.n < bits && b.p < b.end) { b.acc |= (uint64_t)*b.p++ << b.n; b.n += 8; } return (uint32_t)(b.acc & ((1ull << bits) - 1)); }
ZD_FN void fb_skip(FwdBits& b, uint32_t bits) { b.acc >>= bits; b.n = b.n >= bits ? b.n - bits : 0; }

// ---- backward bit reader (Huffman and sequence streams) ------------------------------------------------------------
struct BackBits { const uint8_t* start; int64_t pos; };             // pos = index of the next bit to read, counted from the stream start; < 0 = exhausted
ZD_FN int bb_init(BackBits& b, const uint8_t* p, uint32_t n)
{
    if (n == 0 || p[n - 1] == 0) return -1;
    b.start = p; b.pos = (int64_t)(n - 1) * 8 + hb(p[n - 1]);       // position of the end mark; the payload lies below it
    return 0;
}
ZD_FN uint32_t bb_read(BackBits& b, uint32_t bits)                  // bits <= 32; bits below the stream start read as 0
{
    if (bits == 0) return 0;
    int64_t lo = b.pos - bits;                                      // value = bits [lo, pos)
    b.pos = lo;
    uint64_t v = 0;
    int64_t first = lo < 0 ? 0 : lo;
    int64_t byte0 = first >> 3, byte1 = (lo + bits - 1) >> 3;
    if (lo + (int64_t)bits <= 0) return 0;
    for (int64_t i = byte1; i >= byte0; --i) v = (v << 8) | b.start[i];
    int64_t sh = lo - byte0 * 8;                                    // may be negative when lo < 0
    if (sh >= 0) v >>= sh; else v <<= -sh;
    return (uint32_t)(v & ((bits == 32) ? 0xffffffffull : ((1ull << bits) - 1)));
}

// ---- FSE ----------------------------------------------------------------------------------------------------------
// FSE_readNCount: normalized counts from a forward bit stream; returns bytes consumed or < 0
ZD_FN int read_ncount(int16_t* norm, uint32_t max_sym, uint32_t max_log, uint32_t* n_sym, uint32_t* table_log, const uint8_t* src, uint32_t n)
{
    FwdBits b; fb_init(b, src, src + n);
    if (n < 1) return ZD_ESRC;
    uint32_t log = fb_peek(b, 4) + 5; fb_skip(b, 4);
    if (log > max_log) return ZD_ESRC;
    *table_log = log;
    int32_t remaining = (1 << log) + 1, threshold = 1 << log;
    uint32_t nb = log + 1, sym = 0, bits_used = 4;
    while (remaining > 1 && sym <= max_sym) {
        int32_t max = (2 * threshold - 1) - remaining, count;
        uint32_t v = fb_peek(b, nb);
        if ((int32_t)(v & (threshold - 1)) < max) { count = v & (threshold - 1); fb_skip(b, nb - 1); bits_used += nb - 1; }
        else { count = v & (2 * threshold - 1); if (count >= threshold) count -= max; fb_skip(b, nb); bits_used += nb; }
        --count;                                                    // the stored value is count + 1; -1 = "less than 1"
        remaining -= count < 0 ? -count : count;
        norm[sym++] = (int16_t)count;
        if (count == 0) {                                           // runs of zero-probability symbols: 2-bit repeat flags
            while (true) {
                uint32_t r = fb_peek(b, 2); fb_skip(b, 2); bits_used += 2;
                for (uint32_t i = 0; i < r && sym <= max_sym; ++i) norm[sym++] = 0;
                if (r != 3) break;
            }
        }
        while (remaining < threshold && threshold > 1) { --nb; threshold >>= 1; }
    }
    if (remaining != 1 || sym > max_sym + 1) return ZD_ESRC;
    *n_sym = sym;
    uint32_t bytes = (bits_used + 7) >> 3;
    return bytes > n ? ZD_ESRC : (int)bytes;
}

// FSE_buildDTable
ZD_FN void build_fse(FseEntry* t, const int16_t* norm, uint32_t n_sym, uint32_t log, uint16_t* next)
{
    const uint32_t size = 1u << log;
    uint32_t high = size - 1;
    for (uint32_t s = 0; s < n_sym; ++s) {
        if (norm[s] == -1) { t[high--].sym = (uint8_t)s; next[s] = 1; }
        else next[s] = (uint16_t)norm[s];
    }
    const uint32_t step = (size >> 1) + (size >> 3) + 3, mask = size - 1;
    uint32_t pos = 0;
    for (uint32_t s = 0; s < n_sym; ++s)
        for (int i = 0; i < norm[s]; ++i) { t[pos].sym = (uint8_t)s; do pos = (pos + step) & mask; while (pos > high); }
    for (uint32_t u = 0; u < size; ++u) {
        uint32_t s = t[u].sym, ns = next[s]++;
        t[u].nb = (uint8_t)(log - hb(ns));
        t[u].base = (uint16_t)((ns << t[u].nb) - size);
    }
}

ZD_FN void build_rle(FseEntry* t, uint8_t sym) { t[0].sym = sym; t[0].nb = 0; t[0].base = 0; }

// predefined distributions (RFC 8878 3.1.1.3.2.2)
ZD_FN void default_norm(int which, int16_t* norm, uint32_t* n_sym, uint32_t* log)
{
    const int16_t LL[36] = { 4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1 };
    const int16_t ML[53] = { 1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                             1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1 };
    const int16_t OF[29] = { 1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1 };
    if (which == 0) { for (int i = 0; i < 36; ++i) norm[i] = LL[i]; *n_sym = 36; *log = 6; }
    else if (which == 1) { for (int i = 0; i < 29; ++i) norm[i] = OF[i]; *n_sym = 29; *log = 5; }
    else { for (int i = 0; i < 53; ++i) norm[i] = ML[i]; *n_sym = 53; *log = 6; }
}

// ---- Huffman --------------------------------------------------------------------------------------------------------
// tree description -> decoding table; returns bytes consumed or < 0
ZD_FN int read_huf(Work& w, const uint8_t* src, uint32_t n)
{
    if (n < 1) return ZD_ESRC;
    uint8_t wts[256]; uint32_t nw = 0, used;
    const uint32_t h = src[0];
    if (h >= 128) {                                                 // direct: 4 bits per weight
        nw = h - 127; used = 1 + (nw + 1) / 2;
        if (used > n) return ZD_ESRC;
        for (uint32_t i = 0; i < nw; ++i) wts[i] = (i & 1) ? (src[1 + i / 2] & 15) : (src[1 + i / 2] >> 4);
    } else {                                                        // FSE-compressed weights, two interleaved states
        used = 1 + h;
        if (used > n || h < 1) return ZD_ESRC;
        uint32_t n_sym, log;
        int hdr = read_ncount(w.norm, 12, 6, &n_sym, &log, src + 1, h);
        if (hdr < 0) return hdr;
        build_fse(w.wt, w.norm, n_sym, log, w.next);
        BackBits b; if (bb_init(b, src + 1 + hdr, h - hdr)) return ZD_ESRC;
        uint32_t s1 = bb_read(b, log), s2 = bb_read(b, log);
        while (true) {                                              // HUF_readStats / FSE_decompress tail rules
            if (nw >= 255) return ZD_ESRC;
            wts[nw++] = w.wt[s1].sym;
            if (b.pos - (int64_t)w.wt[s1].nb < 0) { if (nw >= 255) return ZD_ESRC; wts[nw++] = w.wt[s2].sym; break; }
            s1 = w.wt[s1].base + bb_read(b, w.wt[s1].nb);
            if (nw >= 255) return ZD_ESRC;
            wts[nw++] = w.wt[s2].sym;
            if (b.pos - (int64_t)w.wt[s2].nb < 0) { if (nw >= 255) return ZD_ESRC; wts[nw++] = w.wt[s1].sym; break; }
            s2 = w.wt[s2].base + bb_read(b, w.wt[s2].nb);
        }
    }
    uint32_t total = 0;
    for (uint32_t i = 0; i < nw; ++i) { if (wts[i] > 11) return ZD_ESRC; if (wts[i]) total += 1u << (wts[i] - 1); }
    if (total == 0) return ZD_ESRC;
    const uint32_t log = hb(total) + 1;
    if (log > 11) return ZD_ESRC;
    const uint32_t rest = (1u << log) - total;                      // the last weight is implied: must be a power of two
    if (rest == 0 || (rest & (rest - 1))) return ZD_ESRC;
    wts[nw++] = (uint8_t)(hb(rest) + 1);
    // canonical order: by weight ascending (longest codes first in the table), symbols ascending within a weight
    uint32_t start[13] = { 0 }, cnt[13] = { 0 };
    for (uint32_t i = 0; i < nw; ++i) cnt[wts[i]]++;
    uint32_t at = 0;
    for (uint32_t wv = 1; wv <= log; ++wv) { start[wv] = at; at += cnt[wv] << (wv - 1); }
    for (uint32_t i = 0; i < nw; ++i) {
        uint32_t wv = wts[i];
        if (!wv) continue;
        uint32_t span = 1u << (wv - 1), nb = log + 1 - wv;
        for (uint32_t j = 0; j < span; ++j) w.huf[start[wv] + j] = (uint16_t)(i | (nb << 8));
        start[wv] += span;
    }
    w.huf_log = log; w.have_huf = 1;
    return (int)used;
}

ZD_FN int huf_stream(const Work& w, const uint8_t* src, uint32_t n, uint8_t* dst, uint32_t dn)
{
    BackBits b; if (bb_init(b, src, n)) return ZD_ESRC;
    const uint32_t log = w.huf_log;
    for (uint32_t i = 0; i < dn; ++i) {
        int64_t save = b.pos;
        uint32_t v = bb_read(b, log);                               // peek: the code is the top nb bits of the next log bits
        uint16_t e = w.huf[v];
        b.pos = save - (e >> 8);
        dst[i] = (uint8_t)e;
    }
    return b.pos == 0 ? 0 : ZD_ESRC;                                // the stream must be consumed exactly
}

// ---- blocks ---------------------------------------------------------------------------------------------------------
ZD_FN int decode_literals(Work& w, const uint8_t* src, uint32_t n, const uint8_t** lit, uint32_t* lit_n)
{
    if (n < 1) return ZD_ESRC;
    const uint32_t type = src[0] & 3, sf = (src[0] >> 2) & 3;
    if (type < 2) {                                                 // raw / RLE
        uint32_t hs, rn;
        if ((sf & 1) == 0) { hs = 1; rn = src[0] >> 3; }
        else if (sf == 1) { if (n < 2) return ZD_ESRC; hs = 2; rn = (src[0] >> 4) | ((uint32_t)src[1] << 4); }
        else { if (n < 3) return ZD_ESRC; hs = 3; rn = (src[0] >> 4) | ((uint32_t)src[1] << 4) | ((uint32_t)src[2] << 12); }
        if (rn > (128u << 10)) return ZD_ESRC;
        if (type == 0) { if (hs + rn > n) return ZD_ESRC; *lit = src + hs; *lit_n = rn; return (int)(hs + rn); }
        if (hs + 1 > n) return ZD_ESRC;
        for (uint32_t i = 0; i < rn; ++i) w.lit[i] = src[hs];
        *lit = w.lit; *lit_n = rn;
        return (int)(hs + 1);
    }
    uint32_t hs, rn, cn, streams = 4;
    if (n < 3) return ZD_ESRC;
    if (sf < 2) { hs = 3; uint32_t v = (src[0] >> 4) | ((uint32_t)src[1] << 4) | ((uint32_t)src[2] << 12); rn = v & 1023; cn = v >> 10; if (sf == 0) streams = 1; }
    else if (sf == 2) { if (n < 4) return ZD_ESRC; hs = 4; uint32_t v = (src[0] >> 4) | ((uint32_t)src[1] << 4) | ((uint32_t)src[2] << 12) | ((uint32_t)src[3] << 20); rn = v & 16383; cn = v >> 14; }
    else { if (n < 5) return ZD_ESRC; hs = 5; uint64_t v = (src[0] >> 4) | ((uint64_t)src[1] << 4) | ((uint64_t)src[2] << 12) | ((uint64_t)src[3] << 20) | ((uint64_t)src[4] << 28); rn = (uint32_t)(v & 262143); cn = (uint32_t)(v >> 18); }
    if (rn > (128u << 10) || hs + cn > n) return ZD_ESRC;
    const uint8_t* p = src + hs; uint32_t left = cn;
    if (type == 2) { int t = read_huf(w, p, left); if (t < 0) return t; p += t; left -= t; }
    else if (!w.have_huf) return ZD_ESRC;
    if (streams == 1) { int r = huf_stream(w, p, left, w.lit, rn); if (r < 0) return r; }
    else {
        if (left < 6) return ZD_ESRC;
        uint32_t s1 = p[0] | (p[1] << 8), s2 = p[2] | (p[3] << 8), s3 = p[4] | (p[5] << 8);
        if (6 + s1 + s2 + s3 > left) return ZD_ESRC;
        uint32_t s4 = left - 6 - s1 - s2 - s3, q = (rn + 3) / 4;
        if (3 * q > rn) return ZD_ESRC;
        const uint8_t* d = p + 6;
        int r;
        if ((r = huf_stream(w, d, s1, w.lit, q)) < 0) return r;
        if ((r = huf_stream(w, d + s1, s2, w.lit + q, q)) < 0) return r;
        if ((r = huf_stream(w, d + s1 + s2, s3, w.lit + 2 * q, q)) < 0) return r;
        if ((r = huf_stream(w, d + s1 + s2 + s3, s4, w.lit + 3 * q, rn - 3 * q)) < 0) return r;
    }
    *lit = w.lit; *lit_n = rn;
    return (int)(hs + cn);
}

ZD_FN int seq_table(Work& w, int which, uint32_t mode, FseEntry* t, uint32_t* log, int* have, const uint8_t* src, uint32_t n)
{
    const uint32_t max_sym[3] = { 35, 31, 52 }, max_log[3] = { 9, 8, 9 };
    if (mode == 0) { uint32_t ns; default_norm(which, w.norm, &ns, log); build_fse(t, w.norm, ns, *log, w.next); *have = 1; return 0; }
    if (mode == 1) { if (n < 1 || src[0] > max_sym[which]) return ZD_ESRC; build_rle(t, src[0]); *log = 0; *have = 1; return 1; }
    if (mode == 2) {
        uint32_t ns;
        int used = read_ncount(w.norm, max_sym[which], max_log[which], &ns, log, src, n);
        if (used < 0) return used;
        build_fse(t, w.norm, ns, *log, w.next); *have = 1;
        return used;
    }
    return *have ? 0 : ZD_ESRC;                                     // repeat mode
}

ZD_FN int decode_block(Work& w, const uint8_t* src, uint32_t n, uint8_t* dst_base, uint64_t dst_pos, uint64_t dst_cap)
{
    const uint8_t* lit; uint32_t lit_n;
    int used = decode_literals(w, src, n, &lit, &lit_n);
    if (used < 0) return used;
    const uint8_t* p = src + used; uint32_t left = n - used;
    if (left < 1) return ZD_ESRC;
    uint32_t nseq = p[0];
    if (nseq == 0) { p += 1; left -= 1; }
    else if (nseq < 128) { p += 1; left -= 1; }
    else if (nseq < 255) { if (left < 2) return ZD_ESRC; nseq = ((nseq - 128) << 8) + p[1]; p += 2; left -= 2; }
    else { if (left < 3) return ZD_ESRC; nseq = p[1] + ((uint32_t)p[2] << 8) + 0x7F00; p += 3; left -= 3; }
    uint64_t out = dst_pos;
    uint32_t lit_pos = 0;
    if (nseq) {
        if (left < 1) return ZD_ESRC;
        const uint32_t modes = p[0]; p += 1; left -= 1;
        if (modes & 3) return ZD_ESRC;
        int t;
        if ((t = seq_table(w, 0, modes >> 6, w.ll, &w.ll_log, &w.have_ll, p, left)) < 0) return t;
        p += t; left -= t;
        if ((t = seq_table(w, 1, (modes >> 4) & 3, w.of, &w.of_log, &w.have_of, p, left)) < 0) return t;
        p += t; left -= t;
        if ((t = seq_table(w, 2, (modes >> 2) & 3, w.ml, &w.ml_log, &w.have_ml, p, left)) < 0) return t;
        p += t; left -= t;
        const uint32_t LL_base[36] = { 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 0x80, 0x100, 0x200, 0x400,
                                       0x800, 0x1000, 0x2000, 0x4000, 0x8000, 0x10000 };
        const uint8_t LL_bits[36] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16 };
        const uint32_t ML_base[53] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32, 33, 34,
                                       35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 0x83, 0x103, 0x203, 0x403, 0x803, 0x1003, 0x2003, 0x4003, 0x8003, 0x10003 };
        const uint8_t ML_bits[53] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3,
                                      4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16 };
        BackBits b; if (bb_init(b, p, left)) return ZD_ESRC;
        uint32_t sl = bb_read(b, w.ll_log), so = bb_read(b, w.of_log), sm = bb_read(b, w.ml_log);
        for (uint32_t i = 0; i < nseq; ++i) {
            const uint32_t llc = w.ll[sl].sym, ofc = w.of[so].sym, mlc = w.ml[sm].sym;
            if (llc > 35 || mlc > 52 || ofc > 31) return ZD_ESRC;
            uint32_t ofv = (ofc ? (1u << ofc) : 1u) + bb_read(b, ofc);
            const uint32_t mlen = ML_base[mlc] + bb_read(b, ML_bits[mlc]);
            const uint32_t llen = LL_base[llc] + bb_read(b, LL_bits[llc]);
            uint32_t offset;
            if (ofv > 3) { offset = ofv - 3; w.rep[2] = w.rep[1]; w.rep[1] = w.rep[0]; w.rep[0] = offset; }
            else {
                uint32_t idx = ofv - 1 + (llen == 0);               // 0..3
                if (idx == 0) offset = w.rep[0];
                else {
                    offset = idx == 3 ? w.rep[0] - 1 : w.rep[idx];
                    if (offset == 0) return ZD_ESRC;
                    if (idx != 1) w.rep[2] = w.rep[1];
                    w.rep[1] = w.rep[0]; w.rep[0] = offset;
                }
            }
            if (i + 1 < nseq) {                                     // state updates: LL, ML, OF
                sl = w.ll[sl].base + bb_read(b, w.ll[sl].nb);
                sm = w.ml[sm].base + bb_read(b, w.ml[sm].nb);
                so = w.of[so].base + bb_read(b, w.of[so].nb);
            }
            if (b.pos < 0) return ZD_ESRC;
            if (lit_pos + llen > lit_n) return ZD_ESRC;
            if (out + llen + mlen > dst_cap) return ZD_EDST;
            for (uint32_t j = 0; j < llen; ++j) dst_base[out + j] = lit[lit_pos + j];
            out += llen; lit_pos += llen;
            if (offset > out) return ZD_ESRC;
            for (uint32_t j = 0; j < mlen; ++j) dst_base[out + j] = dst_base[out + j - offset];     // byte by byte: overlaps repeat
            out += mlen;
        }
        if (b.pos != 0) return ZD_ESRC;
    }
    const uint32_t tail = lit_n - lit_pos;
    if (out + tail > dst_cap) return ZD_EDST;
    for (uint32_t j = 0; j < tail; ++j) dst_base[out + j] = lit[lit_pos + j];
    out += tail;
    if (out - dst_pos > (128u << 10)) return ZD_ESRC;
    return (int)(out - dst_pos);
}

// Frame_Content_Size of a frame header, -1 when the header does not carry it (frames written by ZSTD_compressCCtx always do)
ZD_FN int64_t frame_content_size(const uint8_t* src, uint64_t n)
{
    if (n < 6 || (src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24)) != 0xFD2FB528u) return -1;
    const uint32_t fhd = src[4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, did = fhd & 3;
    const uint32_t did_bytes = did == 3 ? 4 : did;
    uint64_t p = 5 + (single ? 0 : 1) + did_bytes;
    const uint32_t fcs_bytes = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (fcs_bytes == 0 || p + fcs_bytes > n) return -1;
    uint64_t fcs = 0;
    for (uint32_t i = 0; i < fcs_bytes; ++i) fcs |= (uint64_t)src[p + i] << (8 * i);
    if (fcs_bytes == 2) fcs += 256;
    return (int64_t)fcs;
}

// one frame; `w` needs no initialisation
ZD_FN int64_t decompress_frame(const uint8_t* src, uint64_t n, uint8_t* dst, uint64_t cap, Work& w)
{
    if (n < 6) return ZD_ESRC;
    if ((src[0] | (src[1] << 8) | (src[2] << 16) | ((uint32_t)src[3] << 24)) != 0xFD2FB528u) return ZD_EUNSUP;
    const uint32_t fhd = src[4], fcs_flag = fhd >> 6, single = (fhd >> 5) & 1, checksum = (fhd >> 2) & 1, did = fhd & 3;
    if (fhd & 8) return ZD_ESRC;
    if (did) return ZD_EUNSUP;
    uint64_t p = 5;
    if (!single) p += 1;                                            // window descriptor: the whole output is addressable here
    const uint32_t fcs_bytes = fcs_flag == 0 ? single : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
    if (p + fcs_bytes > n) return ZD_ESRC;
    uint64_t fcs = 0; bool have_fcs = fcs_bytes != 0;
    for (uint32_t i = 0; i < fcs_bytes; ++i) fcs |= (uint64_t)src[p + i] << (8 * i);
    if (fcs_bytes == 2) fcs += 256;
    p += fcs_bytes;
    w.have_huf = w.have_ll = w.have_of = w.have_ml = 0;
    w.rep[0] = 1; w.rep[1] = 4; w.rep[2] = 8;
    uint64_t out = 0;
    while (true) {
        if (p + 3 > n) return ZD_ESRC;
        const uint32_t bh = src[p] | (src[p + 1] << 8) | ((uint32_t)src[p + 2] << 16);
        p += 3;
        const uint32_t last = bh & 1, type = (bh >> 1) & 3, bs = bh >> 3;
        if (type == 0) {
            if (p + bs > n) return ZD_ESRC;
            if (out + bs > cap) return ZD_EDST;
            for (uint32_t i = 0; i < bs; ++i) dst[out + i] = src[p + i];
            out += bs; p += bs;
        } else if (type == 1) {
            if (p + 1 > n) return ZD_ESRC;
            if (out + bs > cap) return ZD_EDST;
            for (uint32_t i = 0; i < bs; ++i) dst[out + i] = src[p];
            out += bs; p += 1;
        } else if (type == 2) {
            if (p + bs > n || bs > (128u << 10)) return ZD_ESRC;
            int r = decode_block(w, src + p, bs, dst, out, cap);
            if (r < 0) return r;
            out += (uint32_t)r; p += bs;
        } else return ZD_ESRC;
        if (last) break;
    }
    if (checksum) { if (p + 4 > n) return ZD_ESRC; p += 4; }
    if (have_fcs && fcs != out) return ZD_ESRC;
    return (int64_t)out;
}

}  // namespace zd
