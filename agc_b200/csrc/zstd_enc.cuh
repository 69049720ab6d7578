// zstd_enc.cuh -- residual coder: a from-scratch producer of zstd frames that are byte-identical to what the
// reference's vendored libzstd (3rd_party/zstd/lib, "1.5.5" @3d1332f) emits for ZSTD_compressCCtx(level) on the input
// classes AGC feeds it (SURVEY a24 + Appendix A): strategies btopt / btultra / btultra2, single-shot, no dictionary,
// content size in the header, no checksum.  The vendored source is the specification; every routine cites the
// function it restates (paths relative to 3rd_party/zstd/lib/compress unless noted).
//
// The code is written once for host and device (ZE_FN): the device build is what libagcgpu ships (kernels_zstd.cu);
// the host build exists only so tests/ can diff it against the reference's libzstd on a CPU box.  One *warp* parses one input:
// scalar control flow is warp-uniform (every lane computes the same value), array-wide steps are lane-strided; the match
// finder's tree walks run one per thread on all threads of the CTA ("window engine" below).
//
// The header can be included more than once with different ZE_NS (namespace) / ZE_WN_W (slots of the match-finder window):
// kernels_zstd.cu instantiates a wide coder (512-slot window, 16 warps per frame) for large inputs and a narrow one (32-slot
// window, one warp and ~20 KB of shared memory per frame) for the small frames that dominate at scale.
#ifndef ZE_NS
#define ZE_NS ze
#endif
#ifndef ZE_WN_W
#define ZE_WN_W 512
#endif
#ifndef ZE_COMMON_DEFS
#define ZE_COMMON_DEFS
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define ZE_FN __host__ __device__ __forceinline__
#define ZE_FN_NOINLINE __host__ __device__ __noinline__
#else
#define ZE_FN inline
#define ZE_FN_NOINLINE inline
#endif

// warp-cooperative helpers: on the device all 32 lanes of a warp run the coder with identical (uniform) scalar state and
// split array-wide steps between them; on the host the same code runs with one "lane".
#if defined(__CUDA_ARCH__)
#define ZE_LANE (threadIdx.x & 31u)
#define ZE_LANES 32u
#define ze_ballot(p) __ballot_sync(0xffffffffu, (p))
#define ze_shfl(v, l) __shfl_sync(0xffffffffu, (v), (l))
#define ze_sync() __syncwarp()
#define ze_ffs(m) ((unsigned)__ffs((int)(m)))
#define ze_match_any(v) __match_any_sync(0xffffffffu, (v))
#define ze_reduce_or(v) __reduce_or_sync(0xffffffffu, (v))
#define ze_ctz64(x) ((unsigned)__ffsll((long long)(x)) - 1u)
#define ze_shfl_up(v, d) __shfl_up_sync(0xffffffffu, (v), (d))
#define ze_reduce_max(v) __reduce_max_sync(0xffffffffu, (unsigned)(v))
#define ze_reduce_add(v) __reduce_add_sync(0xffffffffu, (unsigned)(v))
#define ze_popc(v) ((unsigned)__popc(v))
#define ze_hibit(v) (31u - (unsigned)__clz((int)(v)))
#else
#define ZE_LANE 0u
#define ZE_LANES 1u
#define ze_ballot(p) ((p) ? 1u : 0u)
#define ze_shfl(v, l) (v)
#define ze_sync() ((void)0)
#define ze_ffs(m) ((unsigned)__builtin_ffs((int)(m)))
#define ze_match_any(v) 1u
#define ze_reduce_or(v) (v)
#define ze_ctz64(x) ((unsigned)__builtin_ctzll(x))
#define ze_shfl_up(v, d) (v)
#define ze_reduce_max(v) ((unsigned)(v))
#define ze_reduce_add(v) ((unsigned)(v))
#define ze_popc(v) ((unsigned)__builtin_popcount(v))
#define ze_hibit(v) (31u - (unsigned)__builtin_clz(v))
#endif

typedef uint8_t u8; typedef uint16_t u16; typedef uint32_t u32; typedef uint64_t u64; typedef int32_t i32; typedef int16_t i16;
#endif  // ZE_COMMON_DEFS

namespace ZE_NS {

// optional phase profile of the device build (-DZE_PROF): clock64 ticks and event counts per frame, see kernels_zstd.cu
#if defined(ZE_PROF) && defined(__CUDA_ARCH__)
#define ZE_PROF_N 32
#define ZE_T(var) long long var = clock64()
#define ZE_ACC(w, slot, var) ((w).prof[slot] += (u64)(clock64() - (var)))
#define ZE_CNT(w, slot, v) ((w).prof[slot] += (u64)(v))
#else
#define ZE_T(var) ((void)0)
#define ZE_ACC(w, slot, var) ((void)0)
#define ZE_CNT(w, slot, v) ((void)0)
#endif

// ---------------------------------------------------------------------------------------------------- constants
enum { ST_BTLAZY2 = 6, ST_BTOPT = 7, ST_BTULTRA = 8, ST_BTULTRA2 = 9 };
enum { SET_BASIC = 0, SET_RLE = 1, SET_COMPRESSED = 2, SET_REPEAT = 3 };
enum { REP_NONE = 0, REP_CHECK = 1, REP_VALID = 2 };
static const u32 OPT_NUM = 1u << 12, OPT_SIZE = OPT_NUM + 3;
static const u32 BLOCK_MAX = 128u << 10;
static const i32 MAX_PRICE = 1 << 30;
static const u32 MaxLL = 35, MaxML = 52, MaxOff = 31, DefaultMaxOff = 28;
static const u32 BITCOST_MULT = 256;

#if defined(__CUDA_ARCH__)
#define ZE_TABLE static __device__ const
#else
#define ZE_TABLE static const
#endif
ZE_TABLE u8 kLLbits[36] = { 0,0,0,0,0,0,0,0, 0,0,0,0,0,0,0,0, 1,1,1,1,2,2,3,3, 4,6,7,8,9,10,11,12, 13,14,15,16 };
ZE_TABLE u8 kMLbits[53] = { 0,0,0,0,0,0,0,0, 0,0,0,0,0,0,0,0, 0,0,0,0,0,0,0,0, 0,0,0,0,0,0,0,0, 1,1,1,1,2,2,3,3, 4,4,5,7,8,9,10,11, 12,13,14,15,16 };
ZE_TABLE i16 kLLnorm[36] = { 4,3,2,2,2,2,2,2, 2,2,2,2,2,1,1,1, 2,2,2,2,2,2,2,2, 2,3,2,1,1,1,1,1, -1,-1,-1,-1 };
ZE_TABLE i16 kMLnorm[53] = { 1,4,3,2,2,2,2,2, 2,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,-1,-1, -1,-1,-1,-1,-1 };
ZE_TABLE i16 kOFnorm[29] = { 1,1,1,1,1,1,2,2, 2,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1, -1,-1,-1,-1,-1 };
ZE_TABLE u8 kLLcode[64] = { 0,1,2,3,4,5,6,7, 8,9,10,11,12,13,14,15, 16,16,17,17,18,18,19,19, 20,20,20,20,21,21,21,21,
                            22,22,22,22,22,22,22,22, 23,23,23,23,23,23,23,23, 24,24,24,24,24,24,24,24, 24,24,24,24,24,24,24,24 };
ZE_TABLE u8 kMLcode[128] = { 0,1,2,3,4,5,6,7,8,9,10,11,12,13,14,15, 16,17,18,19,20,21,22,23,24,25,26,27,28,29,30,31,
                             32,32,33,33,34,34,35,35,36,36,36,36,37,37,37,37, 38,38,38,38,38,38,38,38,39,39,39,39,39,39,39,39,
                             40,40,40,40,40,40,40,40,40,40,40,40,40,40,40,40, 41,41,41,41,41,41,41,41,41,41,41,41,41,41,41,41,
                             42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42, 42,42,42,42,42,42,42,42,42,42,42,42,42,42,42,42 };
// zstd_compress_sequences.c:17-50 kInverseProbabilityLog256 : -log2(x/256) in 8.8 fixed point
ZE_TABLE u16 kInvProbLog256[256] = {
    0,2048,1792,1642,1536,1453,1386,1329,1280,1236,1197,1162,1130,1100,1073,1047,1024,1001,980,960,941,923,906,889,
    874,859,844,830,817,804,791,779,768,756,745,734,724,714,704,694,685,676,667,658,650,642,633,626,
    618,610,603,595,588,581,574,567,561,554,548,542,535,529,523,517,512,506,500,495,489,484,478,473,
    468,463,458,453,448,443,438,434,429,424,420,415,411,407,402,398,394,390,386,382,377,373,370,366,
    362,358,354,350,347,343,339,336,332,329,325,322,318,315,311,308,305,302,298,295,292,289,286,282,
    279,276,273,270,267,264,261,258,256,253,250,247,244,241,239,236,233,230,228,225,222,220,217,215,
    212,209,207,204,202,199,197,194,192,190,187,185,182,180,178,175,173,171,168,166,164,162,159,157,
    155,153,151,149,146,144,142,140,138,136,134,132,130,128,126,123,121,119,117,115,114,112,110,108,
    106,104,102,100,98,96,94,93,91,89,87,85,83,82,80,78,76,74,73,71,69,67,66,64,
    62,61,59,57,55,54,52,50,49,47,46,44,42,41,39,37,36,34,33,31,30,28,26,25,
    23,22,20,19,17,16,14,13,11,10,8,7,5,4,2,1 };
// FSE_normalizeCount rtbTable (fse_compress.c:463)
ZE_TABLE u32 kRtb[8] = { 0, 473195, 504333, 520860, 550000, 700000, 750000, 830000 };

// clevels.h rows actually reachable from AGC (levels 13,17,18,19) : {W,C,H,S,L,T,strategy} for >256K, <=256K, <=128K, <=16K
ZE_TABLE u8 kLevels[4][4][7] = {
    { {22,22,22,4,5,32,ST_BTLAZY2}, {18,18,19,4,4,16,ST_BTOPT},    {17,18,17,3,4,12,ST_BTOPT},    {14,15,14,5,3,32,ST_BTULTRA} },     // 13
    { {23,23,22,5,4,64,ST_BTOPT},   {18,19,19,8,3,0,ST_BTULTRA},   {17,18,17,8,3,0,ST_BTULTRA},   {14,15,15,6,3,128,ST_BTULTRA2} },   // 17 (T=256 patched below)
    { {23,23,22,6,3,64,ST_BTULTRA}, {18,19,19,6,3,128,ST_BTULTRA2},{17,18,17,10,3,0,ST_BTULTRA},  {14,15,15,7,3,0,ST_BTULTRA2} },     // 18
    { {23,24,22,7,3,0,ST_BTULTRA2}, {18,19,19,8,3,0,ST_BTULTRA2},  {17,18,17,5,3,0,ST_BTULTRA2},  {14,15,15,8,3,0,ST_BTULTRA2} } };   // 19
ZE_TABLE u16 kLevelsT[4][4] = { {32,16,12,32}, {64,256,256,128}, {64,128,512,256}, {256,256,256,256} };

struct Params { u32 windowLog, chainLog, hashLog, searchLog, minMatch, targetLength, strategy; u32 blockSize; int splitter; int supported; };

ZE_FN u32 highbit(u32 v) {
#if defined(__CUDA_ARCH__)
    return 31u - (u32)__clz((int)v);
#else
    return 31u - (u32)__builtin_clz(v);
#endif
}

// ZSTD_getCParams_internal (zstd_compress.c:7027) + ZSTD_adjustCParams_internal (:1462) + derived switches (:255, resetCCtx blockSize)
ZE_FN Params get_params(int level, u64 n)
{
    Params p;
    int li = level == 13 ? 0 : level == 17 ? 1 : level == 18 ? 2 : level == 19 ? 3 : -1;
    p.supported = li >= 0;
    if (li < 0) li = 3;
    u32 tid = (n <= (256u << 10)) + (n <= (128u << 10)) + (n <= (16u << 10));
    p.windowLog = kLevels[li][tid][0]; p.chainLog = kLevels[li][tid][1]; p.hashLog = kLevels[li][tid][2];
    p.searchLog = kLevels[li][tid][3]; p.minMatch = kLevels[li][tid][4]; p.targetLength = kLevelsT[li][tid]; p.strategy = kLevels[li][tid][6];
    if (n >= (1ull << 30)) p.supported = 0;
    {   u32 t = (u32)n;
        u32 srcLog = t < 64 ? 6 : highbit(t - 1) + 1;
        if (p.windowLog > srcLog) p.windowLog = srcLog;
        u32 dw = p.windowLog;
        u32 cycleLog = p.chainLog - 1;                            // bt strategies
        if (p.hashLog > dw + 1) p.hashLog = dw + 1;
        if (cycleLog > dw) p.chainLog -= (cycleLog - dw);
        if (p.windowLog < 10) p.windowLog = 10;
    }
    p.splitter = p.strategy >= ST_BTOPT && p.windowLog >= 17;         // ZSTD_resolveBlockSplitterMode (:255)
    u64 ws = 1ull << p.windowLog; if (ws > n) ws = n; if (ws < 1) ws = 1;
    p.blockSize = (u32)(ws < BLOCK_MAX ? ws : BLOCK_MAX);
    return p;
}

// ---------------------------------------------------------------------------------------------------- state
struct Opt { i32 price; u32 off, mlen, litlen; u32 rep[3]; };
struct Match { u32 off, len; };
struct Seq { u32 offBase; u16 litLength, mlBase; };
struct FseCT { u16 tableLog, maxSym; u16 state[512]; i32 dFind[53]; u32 dBits[53]; };
struct HufCT { u8 nbBits[256]; u16 val[256]; u32 tableLog, maxSym; };
struct Entropy { HufCT huf; i32 hufRepeat; FseCT ll, of, ml; i32 llRep, ofRep, mlRep; };
struct BlockState { Entropy e; u32 rep[3]; };
struct HufNode { u32 count; u16 parent; u8 byte, nbBits; };
struct SeqStore { Seq* seqStart; Seq* seq; u8* litStart; u8* lit; u8* llCode; u8* mlCode; u8* ofCode; u32 longType; u32 longPos; };   // longType 0 none 1 LL 2 ML
struct FseMeta { int llType, ofType, mlType; u32 tablesSize, lastCountSize; u8 buf[133]; };
struct HufMeta { int hType; u32 desSize; u8 des[128]; };

struct Win;
struct Work {
    // inputs
    const u8* src; u32 srcSize; Params cp;
    // window / match state (zstd_compress_internal.h ZSTD_window_t, ZSTD_matchState_t)
    u32 baseOff;             // index of src[0]  (2 initially, += blockSize after initStats_ultra)
    u32 lowLimit, dictLimit, nextToUpdate, hashLog3;
    u32* hashTable; u32* chainTable; u32* hashTable3;
    // optimal parser state (optState_t)
    Opt* opt; Match* matches;
    u32* litFreq; u32* llFreq; u32* mlFreq; u32* ofFreq;
    u32* histTab;            // 256 counters in the low-latency scratch (device: lane-parallel histograms), may be null
    u32* priceTab;           // prices of the current statistics: lit[256], ll code[36], ml code[53], off code[32] (refresh_prices)
    u32 litSum, llSum, mlSum, ofSum, litSumBP, llSumBP, mlSumBP, ofSumBP; int pricePredef;
    // sequences
    SeqStore ss; u32 maxNbSeq;
    BlockState* bs[2]; int prevIdx;
    int isFirstBlock;
    // scratch
    u32* count;              // 256
    HufNode* huffNode;       // 512 + 2
    u32* rankPos;            // 192*2
    u8* scratch;             // >= 1024 bytes
    FseCT* tmpCT;            // weights / cost probes
    HufCT* tmpHuf;
    FseMeta* fseMeta; HufMeta* hufMeta;
    u32* partitions;         // 196+1
    u32* dummySlot;          // sink for the binary-tree walks (must live in the same address space as the tables)
    Win* win;                // window engine of the match finder (below)
    int error;
#if defined(ZE_PROF) && defined(__CUDA_ARCH__)
    u64 prof[ZE_PROF_N];
#endif
};

ZE_FN void wr16(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); }
ZE_FN void wr24(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); }
ZE_FN void wr32(u8* p, u32 v) { p[0] = (u8)v; p[1] = (u8)(v >> 8); p[2] = (u8)(v >> 16); p[3] = (u8)(v >> 24); }

// ZSTD_count (zstd_compress_internal.h:752-772): common prefix length of a and b, a bounded by lim.
// Lane l compares bytes [4l, 4l+4) of each 4*LANES-byte step; the first lane that sees a mismatch (or the limit) decides.
ZE_FN u32 ld32u(const u8* p)                                      // unaligned 4-byte little-endian load (may touch 3 bytes past p+3)
{
#if defined(__CUDA_ARCH__)
    const u32* q = (const u32*)((uintptr_t)p & ~(uintptr_t)3);
    return __funnelshift_r(q[0], q[1], 8u * (u32)((uintptr_t)p & 3u));
#else
    return (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
#endif
}
ZE_FN u64 ld64u(const u8* p)                                      // unaligned 8-byte little-endian load (may touch 15 bytes past p)
{
#if defined(__CUDA_ARCH__)
    const u64* q = (const u64*)((uintptr_t)p & ~(uintptr_t)7);
    u32 sh = 8u * (u32)((uintptr_t)p & 7u);
    u64 lo = q[0], hi = q[1];
    return sh ? (lo >> sh) | (hi << (64 - sh)) : lo;
#else
    u64 v = 0; for (int i = 7; i >= 0; --i) v = (v << 8) | p[i]; return v;
#endif
}
ZE_FN void ld64w(const u8* p, u32& lo, u32& hi)                   // unaligned 8 bytes as two little-endian words (may touch 11 bytes past p)
{
#if defined(__CUDA_ARCH__)
    const u32* q = (const u32*)((uintptr_t)p & ~(uintptr_t)3);
    u32 sh = 8u * (u32)((uintptr_t)p & 3u);
    u32 w0 = q[0], w1 = q[1], w2 = q[2];
    lo = __funnelshift_r(w0, w1, sh); hi = __funnelshift_r(w1, w2, sh);
#else
    lo = (u32)p[0] | ((u32)p[1] << 8) | ((u32)p[2] << 16) | ((u32)p[3] << 24);
    hi = (u32)p[4] | ((u32)p[5] << 8) | ((u32)p[6] << 16) | ((u32)p[7] << 24);
#endif
}
ZE_FN void ld64w2(const u8* p, u32& lo, u32& hi)                  // same bytes through two aligned 8-byte loads (may touch 15 bytes past p)
{
#if defined(__CUDA_ARCH__)
    const uint2* q = (const uint2*)((uintptr_t)p & ~(uintptr_t)7);
    const u32 o = (u32)((uintptr_t)p & 7u), sh = 8u * (o & 3u);
    const uint2 v0 = q[0], v1 = q[1];
    const u32 w0 = o < 4 ? v0.x : v0.y, w1 = o < 4 ? v0.y : v1.x, w2 = o < 4 ? v1.x : v1.y;
    lo = __funnelshift_r(w0, w1, sh); hi = __funnelshift_r(w1, w2, sh);
#else
    ld64w(p, lo, hi);
#endif
}
struct U2 { u32 x, y; };
ZE_FN U2 ld_pair(const u32* p)                                    // 8-byte aligned pair
{
#if defined(__CUDA_ARCH__)
    uint2 v = *reinterpret_cast<const uint2*>(p); U2 r; r.x = v.x; r.y = v.y; return r;
#else
    U2 r; r.x = p[0]; r.y = p[1]; return r;
#endif
}
ZE_FN void st_pair(u32* p, u32 x, u32 y)
{
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint2*>(p) = make_uint2(x, y);
#else
    p[0] = x; p[1] = y;
#endif
}
ZE_FN u32 rd32(const u8* p) { return ld32u(p); }                // MEM_readLE32 / 64 on input text (reads may run a few bytes past p + 3 / p + 7)
ZE_FN u64 rd64(const u8* p) { return ld64u(p); }
ZE_FN u32 count_eq(const u8* a, const u8* b, const u8* lim)
{
    if (a >= lim) return 0;
    const u32 nmax = (u32)(lim - a);
    for (u32 base = 0; base < nmax; base += 4 * ZE_LANES) {
        u32 off = base + 4 * ZE_LANE, idx = 0;
        bool stop = true;
        if (off < nmax) {
            u32 rem = nmax - off;
            if (rem >= 4) {
                u32 x = ld32u(a + off) ^ ld32u(b + off);
                idx = x ? ((ze_ffs(x) - 1) >> 3) : 4;
            } else { while (idx < rem && a[off + idx] == b[off + idx]) ++idx; }
            stop = idx < 4;
        }
        u32 mk = ze_ballot(stop);
        if (mk) { u32 L = ze_ffs(mk) - 1; return base + 4 * L + ze_shfl(idx, L); }
    }
    return nmax;
}

// ZSTD_hashPtr (zstd_compress_internal.h:825): mls 5 / 6 use the 64-bit multiplicative hashes, everything else hash4
ZE_FN u32 hash_ptr(const u8* p, u32 hBits, u32 mls)
{
    if (mls == 5) return (u32)(((rd64(p) << 24) * 889523592379ULL) >> (64 - hBits));
    if (mls == 6) return (u32)(((rd64(p) << 16) * 227718039650203ULL) >> (64 - hBits));
    return (rd32(p) * 2654435761U) >> (32 - hBits);
}
ZE_FN u32 hash3_ptr(const u8* p, u32 h) { return ((rd32(p) << 8) * 506832829U) >> (32 - h); }
ZE_FN u32 LLcode(u32 ll) { return ll > 63 ? highbit(ll) + 19 : kLLcode[ll]; }
ZE_FN u32 MLcode(u32 ml) { return ml > 127 ? highbit(ml) + 36 : kMLcode[ml]; }

// ZSTD_updateRep (zstd_compress_internal.h:713)
ZE_FN void update_rep(u32* rep, u32 offBase, u32 ll0)
{
    if (offBase > 3) { rep[2] = rep[1]; rep[1] = rep[0]; rep[0] = offBase - 3; }
    else {
        u32 rc = offBase - 1 + ll0;
        if (rc > 0) {
            u32 cur = (rc == 3) ? rep[0] - 1 : rep[rc];
            rep[2] = (rc >= 2) ? rep[1] : rep[2];
            rep[1] = rep[0];
            rep[0] = cur;
        }
    }
}

// ---------------------------------------------------------------------------------------------------- binary tree match finder (zstd_opt.c)
ZE_FN u32 lowest_match_index(const Work& w, u32 curr)             // ZSTD_getLowestMatchIndex, no dictionary
{
    u32 maxDist = 1u << w.cp.windowLog;
    return (curr - w.lowLimit > maxDist) ? curr - maxDist : w.lowLimit;
}

// ZSTD_insertBt1 (zstd_opt.c:441-560), noDict
ZE_FN_NOINLINE u32 insert_bt1(Work& w, u32 curr, const u8* iend, u32 target, u32 mls)
{
    const u8* base = w.src - w.baseOff;
    const u8* ip = base + curr;
    u32 h = hash_ptr(ip, w.cp.hashLog, mls);
    u32* bt = w.chainTable;
    u32 btLog = w.cp.chainLog - 1, btMask = (1u << btLog) - 1;
    u32 matchIndex = w.hashTable[h];
    u32 clSmaller = 0, clLarger = 0;
    u32 btLow = btMask >= curr ? 0 : curr - btMask;
    u32* smallerPtr = bt + 2 * (curr & btMask);
    u32* largerPtr = smallerPtr + 1;
    u32* const dummyPtr = w.dummySlot;
    u32 windowLow = lowest_match_index(w, target);
    u32 matchEndIdx = curr + 8 + 1;
    u32 bestLength = 8;
    u32 nbCompares = 1u << w.cp.searchLog;
    w.hashTable[h] = curr;
    for (; nbCompares && matchIndex >= windowLow; --nbCompares) {
        u32* nextPtr = bt + 2 * (matchIndex & btMask);
        u32 ml = clSmaller < clLarger ? clSmaller : clLarger;
        const u8* match = base + matchIndex;
        ml += count_eq(ip + ml, match + ml, iend);
        if (ml > bestLength) { bestLength = ml; if (ml > matchEndIdx - matchIndex) matchEndIdx = matchIndex + ml; }
        if (ip + ml == iend) break;
        if (match[ml] < ip[ml]) {
            *smallerPtr = matchIndex; clSmaller = ml;
            if (matchIndex <= btLow) { smallerPtr = dummyPtr; break; }
            smallerPtr = nextPtr + 1; matchIndex = nextPtr[1];
        } else {
            *largerPtr = matchIndex; clLarger = ml;
            if (matchIndex <= btLow) { largerPtr = dummyPtr + 1; break; }
            largerPtr = nextPtr; matchIndex = nextPtr[0];
        }
    }
    *largerPtr = 0;
    *smallerPtr = 0;
    u32 positions = 0;
    if (bestLength > 384) { positions = bestLength - 384; if (positions > 192) positions = 192; }
    u32 adv = matchEndIdx - (curr + 8);
    return positions > adv ? positions : adv;
}

// ---- window engine ----------------------------------------------------------------------------------------------
// The tree walk of ZSTD_insertBt1 / ZSTD_insertBtAndGetAllMatches is a chain of dependent loads, and one walk per position
// is the whole cost of the match finder.  The engine takes the next WN_W positions at once:
//   build  (all threads of the CTA, one position = "slot" per thread, read-only on the tree as it is in memory):
//     walk    the slot's search path on the snapshot tree, recording (node, match length, side) per step;
//     pairs   for every earlier slot of the same hash bucket: common prefix length and order of the two positions;
//     resolve what the walk WOULD be once all earlier slots are inserted.  The binary tree of a bucket is the Cartesian
//             tree of its nodes (key = suffix order, newest node on top), so the search path of q consists of the nodes
//             that have no newer node between them and q.  Inserting the earlier slots p (all newer than the snapshot)
//             therefore (1) puts in front the p's that are not shadowed by a newer p on their side, and (2) removes the
//             snapshot nodes x that have some p between x and q -- exactly the nodes on the common prefix of the two
//             recorded paths that lie on p's side.  No tree access is needed for this.
//   consume (the parser's warp, in position order): inserts only advance a cursor (their tree links are written later),
//             queries read the slot's resolved path and apply ZSTD_insertBtAndGetAllMatches' rules to it.
//   commit (all threads): write the links of the consumed slots, bucket by bucket in position order.
// Whatever breaks the model -- a walk that ran out of compares or reached the end of the block, more than WN_D earlier
// slots in one bucket, positions the reference skips -- ends the window there (commit, then a fresh build), and a slot that
// is unusable even as the first of a fresh window takes the sequential walk, so the tree and the matches are always the
// ones the reference order of operations produces.
static const u32 WN_W = ZE_WN_W, WN_CAP = 32, WN_D = 12, WN_Q = 6, WN_KEYS = 2 * ZE_WN_W, WN_HOPS = 64;
enum { WF_OVF = 1, WF_IEND = 2, WF_BUDGET = 4, WF_BAD = 8, WF_TRUNC = 16 };
enum { WJ_EXIT = 0, WJ_BUILD = 1, WJ_COMMIT = 2, WJ_RESOLVE = 3 };

struct Win {
    u32 job;
    u32 base, count, next, endIdx, baseOff;        // slots = positions [base, base + count); slots < next are consumed
    u32 upto;                                       // WJ_COMMIT: slots [0, upto)
    u32 jfrom, jto;                                 // WJ_BUILD: slots [jfrom, jto)
    u32 built, pending;                             // slots < built are resolved; pending: a posted job has not been waited for yet
    u32 maxChain;                                   // deepest bucket chain of the window
    const u8* text;                                 // text[position]
    u32* hashTable; u32* bt;
    u32 btMask, hashLog, mls, lowLimit, budget;
    u32 qbase;                                      // a query's initial best length (minimum match - 1)
    u64 phase[8];                                   // -DZE_PROF: ticks per build phase as thread 0 sees them
    u32 hash[WN_W + 4];
    u16 ktab[WN_KEYS], prevKey[WN_W];               // bucket links: latest slot per key (hash & (WN_KEYS-1)) / the previous slot of the same key, +1, 0 = none
    u32 keep[WN_W];                                 // recorded steps that stay on the path
    u32 adv[WN_W];                                  // ZSTD_insertBt1's return value
    u8 wflags[WN_W], sflags[WN_W];                  // flags after the walk / after walk + pairs (resolve starts from the latter)
    u8 nq[WN_W];                                    // matches a query at this slot reports when no repcode beats them; 0xff = take the general path
    u32 qm[WN_W][WN_Q + 1][2];                      // their (offBase, length), lengths increasing
    u8 n[WN_W], flags[WN_W], np[WN_W], tm[WN_W], tn[WN_W], chain[WN_W], dead[WN_W];   // dead: a position the reference skips (never inserted)   // recorded steps, WF_*, visited earlier slots, path length, linked steps, earlier slots in the bucket
    // rows are padded to an odd number of words / 8-byte pairs: a warp reads one column (thread = slot) without bank conflicts
    u16 cs[WN_W][WN_D + 1];                            // earlier slots of the same bucket, newest first
    u8 pp[WN_W][WN_D];                              // the visited ones (indices into cs), newest first
    u32 pl[WN_W][WN_D + 1];                           // common prefix with cs[k] | (cs[k] is the smaller one) << 31
    u32 rec[WN_W][WN_CAP + 1][2];                      // node, match length | (node is smaller) << 31
};

// Who runs the window jobs.  Device, wide CTA: the helper warps (threads 32..) only -- the parser's warp posts a job with
// bar.arrive on barrier 1 and, when it needs the result, waits on barrier 2, so a job can run while the parser goes on; the
// helpers separate the phases of a job among themselves on barrier 3.  Device, one-warp CTA and host: the caller runs the job.
#if defined(__CUDA_ARCH__)
#define ZE_JOB_TID (blockDim.x > 32 ? threadIdx.x - 32u : threadIdx.x)
#define ZE_JOB_NT (blockDim.x > 32 ? blockDim.x - 32u : 32u)
#else
#define ZE_JOB_TID 0u
#define ZE_JOB_NT 1u
#endif
ZE_FN void ze_bar_sync(u32 id, u32 nt)
{
#if defined(__CUDA_ARCH__)
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nt) : "memory");
#else
    (void)id; (void)nt;
#endif
}
ZE_FN void ze_bar_arrive(u32 id, u32 nt)
{
#if defined(__CUDA_ARCH__)
    __threadfence_block();
    asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(nt) : "memory");
#else
    (void)id; (void)nt;
#endif
}
ZE_FN void ze_job_sync()                                          // between the phases of a job
{
#if defined(__CUDA_ARCH__)
    if (blockDim.x > 32) ze_bar_sync(3, blockDim.x - 32u); else __syncwarp();
#endif
}

ZE_FN u32 nth_set_bit(u32 mask, u32 n)                            // position of the n-th (0-based) set bit
{
#if defined(__CUDA_ARCH__)
    return __fns(mask, 0, (int)n + 1);
#else
    for (u32 i = 0; i < 32; ++i) if ((mask >> i) & 1u) { if (n == 0) return i; --n; }
    return 32;
#endif
}
// common prefix of a and b, at most maxlen bytes; ca / cb = the bytes after it (lane-private).  The first 8 bytes decide most
// comparisons; longer ones continue 32 bytes per round trip (all loads of a step are issued before the first compare).
ZE_FN u32 lcp_private(const u8* a, const u8* b, u32 maxlen, u32& ca, u32& cb)
{
    u32 ml = 0;
    if (maxlen >= 8) {
        u32 alo, ahi, blo, bhi;
        ld64w(a, alo, ahi); ld64w2(b, blo, bhi);
        u32 xlo = alo ^ blo, xhi = ahi ^ bhi;
        if (xlo | xhi) {
            bool inLo = xlo != 0;
            u32 x = inLo ? xlo : xhi, t = (ze_ffs(x) - 1) >> 3;
            ca = ((inLo ? alo : ahi) >> (8 * t)) & 255u; cb = ((inLo ? blo : bhi) >> (8 * t)) & 255u;
            return (inLo ? 0 : 4) + t;
        }
        ml = 8;
    }
#if defined(__CUDA_ARCH__)
    while (maxlen - ml >= 32) {
        const u32* qa = (const u32*)((uintptr_t)(a + ml) & ~(uintptr_t)3); const u32 sa = 8u * (u32)((uintptr_t)(a + ml) & 3u);
        const u32* qb = (const u32*)((uintptr_t)(b + ml) & ~(uintptr_t)3); const u32 sb = 8u * (u32)((uintptr_t)(b + ml) & 3u);
        u32 wa[9], wb[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) { wa[i] = qa[i]; wb[i] = qb[i]; }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const u32 va = __funnelshift_r(wa[i], wa[i + 1], sa), vb = __funnelshift_r(wb[i], wb[i + 1], sb), x = va ^ vb;
            if (x) { const u32 t = (ze_ffs(x) - 1) >> 3; ca = (va >> (8 * t)) & 255u; cb = (vb >> (8 * t)) & 255u; return ml + 4 * i + t; }
        }
        ml += 32;
    }
#endif
    while (maxlen - ml >= 8) {
        u32 alo, ahi, blo, bhi;
        ld64w(a + ml, alo, ahi); ld64w2(b + ml, blo, bhi);
        u32 xlo = alo ^ blo, xhi = ahi ^ bhi;
        if (xlo | xhi) {
            bool inLo = xlo != 0;
            u32 x = inLo ? xlo : xhi, t = (ze_ffs(x) - 1) >> 3;
            ca = ((inLo ? alo : ahi) >> (8 * t)) & 255u; cb = ((inLo ? blo : bhi) >> (8 * t)) & 255u;
            return ml + (inLo ? 0 : 4) + t;
        }
        ml += 8;
    }
    while (ml < maxlen && a[ml] == b[ml]) ++ml;
    ca = cb = 0;
    if (ml < maxlen) { ca = a[ml]; cb = b[ml]; }
    return ml;
}

ZE_FN void win_walk(Win& W, u32 l)
{
    const u8* text = W.text; const u32 q = W.base + l; const u8* ip = text + q;
    const u32 h = hash_ptr(ip, W.hashLog, W.mls);
    W.hash[l] = h;
    u32 mi = W.hashTable[h];
    u32 n = 0, fl = 0, nb = W.budget, clS = 0, clL = 0;
    const u32 remTot = W.endIdx - q, low = W.lowLimit, btMask = W.btMask;
    const u32* bt = W.bt;
    while (nb && mi >= low) {
        if (n == WN_CAP) { fl |= WF_OVF; break; }
        U2 nx = ld_pair(bt + 2 * (mi & btMask));
        u32 ml = clS < clL ? clS : clL, ca, cm;
        ml += lcp_private(ip + ml, text + mi + ml, remTot - ml, ca, cm);
        const bool atEnd = ml == remTot, smaller = cm < ca;
        W.rec[l][n][0] = mi; W.rec[l][n][1] = atEnd ? ml : (ml | (smaller ? 0x80000000u : 0u)); ++n;
        if (atEnd) { fl |= WF_IEND; break; }
        if (smaller) { clS = ml; mi = nx.y; } else { clL = ml; mi = nx.x; }
        --nb;
    }
    if (nb == 0 && mi >= low) fl |= WF_BUDGET;
    W.n[l] = (u8)n; W.wflags[l] = (u8)fl; W.dead[l] = 0;
}

// bucket links of a window: prevKey[l] = the nearest earlier slot with the same key.  Slots are taken 32 at a time in position
// order (one warp per group, __match_any inside the group, a table of the latest slot per key across groups), so the links
// cost one pass instead of a scan over all earlier slots per slot.
ZE_FN void win_links(Win& W)
{
    const u32 tid = ZE_JOB_TID, nt = ZE_JOB_NT;
    for (u32 i = tid; i < WN_KEYS; i += nt) W.ktab[i] = 0;
    ze_job_sync();
#if defined(__CUDA_ARCH__)
    const u32 nwarps = nt >> 5, wj = tid >> 5, lane = tid & 31u;
    for (u32 g = 0; 32u * g < W.count; ++g) {
        if (wj == g % nwarps) {
            const u32 l = 32u * g + lane; const bool act = l < W.count;
            const u32 key = act ? (W.hash[l] & (WN_KEYS - 1u)) : (0x80000000u | lane);
            u32 pk = act ? W.ktab[key] : 0u;
            const u32 mk = __match_any_sync(0xffffffffu, key), lower = mk & ((1u << lane) - 1u);
            if (lower) pk = 32u * g + (31u - (u32)__clz((int)lower)) + 1u;
            if (act) { W.prevKey[l] = (u16)pk; if ((mk >> lane) == 1u) W.ktab[key] = (u16)(l + 1u); }
        }
        ze_job_sync();
    }
#else
    for (u32 l = 0; l < W.count; ++l) { const u32 key = W.hash[l] & (WN_KEYS - 1u); W.prevKey[l] = W.ktab[key]; W.ktab[key] = (u16)(l + 1u); }
#endif
}

ZE_FN void win_pairs(Win& W, u32 l)
{
    const u32 h = W.hash[l];
    u32 c = 0, fl = W.wflags[l];
    // earlier slots of the bucket, newest first: follow the links of the key, keep the live slots with the same hash
    {   u32 j = W.prevKey[l], hops = 0;
        while (j) {
            const u32 sidx = j - 1;
            if (W.hash[sidx] == h && !W.dead[sidx]) { if (c == WN_D) { fl |= WF_BAD; break; } W.cs[l][c++] = (u16)sidx; }
            j = W.prevKey[sidx];
            if (++hops == WN_HOPS && j) { fl |= WF_BAD; break; }      // too many dead / foreign slots on the way: give up on this slot
        }
    }
    W.chain[l] = (u8)c;
#if defined(__CUDA_ARCH__)
    if (c > W.maxChain) atomicMax(&W.maxChain, c);
#else
    if (c > W.maxChain) W.maxChain = c;
#endif
    const u8* text = W.text; const u32 q = W.base + l, remTot = W.endIdx - q;
    for (u32 k = 0; k < c; ++k) {
        u32 ca, cb;
        u32 ml = lcp_private(text + q, text + W.base + W.cs[l][k], remTot, ca, cb);
        if (ml == remTot) fl |= WF_BAD;
        W.pl[l][k] = ml | (cb < ca ? 0x80000000u : 0u);
    }
    W.sflags[l] = (u8)fl;
}

ZE_FN void win_resolve(Win& W, u32 l)
{
    u32 fl = W.sflags[l];
    const u32 nst = W.n[l], c = W.chain[l], q = W.base + l, B = W.budget;
    u32 live = 0;
    u32 np = 0, bSl = 0, bLl = 0, dropS = 0, dropL = 0;
    i32 bS = -1, bL = -1;
    for (u32 k = 0; k < c; ++k) {
        if (W.dead[W.cs[l][k]]) continue;
        ++live;
        const u32 v = W.pl[l][k], lcp = v & 0x7fffffffu; const bool sm = (v >> 31) != 0;
        const i32 b = sm ? bS : bL; const u32 bl = sm ? bSl : bLl;
        bool vis;
        if (b < 0) vis = true;
        else if (lcp != bl) vis = lcp > bl;
        else { const bool below = (W.pl[W.cs[l][b]][k - (u32)b - 1] >> 31) != 0; vis = sm ? !below : below; }   // order of the two earlier slots
        if (vis) { W.pp[l][np++] = (u8)k; if (sm) { bS = (i32)k; bSl = lcp; } else { bL = (i32)k; bLl = lcp; } }
        const u32 p = W.cs[l][k], npz = W.n[p], lim = nst < npz ? nst : npz;
        u32 cp = 0;
        while (cp < lim && W.rec[l][cp][0] == W.rec[p][cp][0] && ((W.rec[l][cp][1] ^ W.rec[p][cp][1]) >> 31) == 0) ++cp;
        if (sm) { if (cp > dropS) dropS = cp; } else if (cp > dropL) dropL = cp;
    }
    u32 keep = 0;
    for (u32 j = 0; j < nst; ++j) { const bool sj = (W.rec[l][j][1] >> 31) != 0; if (j >= (sj ? dropS : dropL)) keep |= 1u << j; }
    if ((fl & WF_IEND) && live > 0) fl |= WF_BAD;
    u32 m = np + ze_popc(keep);
    bool trunc = false;
    const bool incomplete = (fl & (WF_OVF | WF_BUDGET)) != 0;
    if (m > B) { m = B; trunc = true; }
    else if (incomplete) { if (m == B) trunc = true; else fl |= WF_BAD; }
    u32 tn = m;
    if ((fl & WF_IEND) && nst <= m) { trunc = true; tn = m - 1; }
    // ZSTD_insertBt1's bookkeeping over the resolved path
    u32 best = 8, endI = q + 9, cnt = 0;
    u32 qb = W.qbase, nq = 0; const u32 remTot = W.endIdx - q;       // ... and ZSTD_insertBtAndGetAllMatches' list for a query without repcode matches
    for (u32 e = 0; e < np && cnt < m; ++e, ++cnt) {
        const u32 k = W.pp[l][e], mi = W.base + W.cs[l][k], ml = W.pl[l][k] & 0x7fffffffu;
        if (ml > best) { best = ml; if (ml > endI - mi) endI = mi + ml; }
        if (ml > qb) { qb = ml; if (nq < WN_Q && ml <= OPT_NUM && ml != remTot) { W.qm[l][nq][0] = q - mi + 3; W.qm[l][nq][1] = ml; ++nq; } else nq = 0xff; }
    }
    for (u32 j = 0; j < nst && cnt < m; ++j) if ((keep >> j) & 1u) {
        const u32 mi = W.rec[l][j][0], ml = W.rec[l][j][1] & 0x7fffffffu;
        if (ml > best) { best = ml; if (ml > endI - mi) endI = mi + ml; }
        if (ml > qb) { qb = ml; if (nq < WN_Q && ml <= OPT_NUM && ml != remTot) { W.qm[l][nq][0] = q - mi + 3; W.qm[l][nq][1] = ml; ++nq; } else nq = 0xff; }
        ++cnt;
    }
    if ((fl & WF_IEND) && nst <= m) nq = 0xff;
    u32 positions = 0;
    if (best > 384) { positions = best - 384; if (positions > 192) positions = 192; }
    const u32 a2 = endI - (q + 8);
    if (trunc) fl |= WF_TRUNC;
    W.nq[l] = (u8)(nq > WN_Q ? 0xff : nq); W.keep[l] = keep; W.np[l] = (u8)np; W.tm[l] = (u8)m; W.tn[l] = (u8)tn; W.adv[l] = positions > a2 ? positions : a2; W.flags[l] = (u8)fl;
}

ZE_FN void win_commit_slot(Win& W, u32 l)
{
    const u32 q = W.base + l, btMask = W.btMask; u32* bt = W.bt;
    u32* sp = bt + 2 * (q & btMask); u32* lp = sp + 1;
    W.hashTable[W.hash[l]] = q;
    const u32 np = W.np[l], nst = W.n[l], tn = W.tn[l], keep = W.keep[l];
    u32 cnt = 0;
    for (u32 e = 0; e < np && cnt < tn; ++e, ++cnt) {
        const u32 k = W.pp[l][e], mi = W.base + W.cs[l][k]; u32* nextPtr = bt + 2 * (mi & btMask);
        if (W.pl[l][k] >> 31) { *sp = mi; sp = nextPtr + 1; } else { *lp = mi; lp = nextPtr; }
    }
    for (u32 j = 0; j < nst && cnt < tn; ++j) if ((keep >> j) & 1u) {
        const u32 mi = W.rec[l][j][0]; u32* nextPtr = bt + 2 * (mi & btMask);
        if (W.rec[l][j][1] >> 31) { *sp = mi; sp = nextPtr + 1; } else { *lp = mi; lp = nextPtr; }
        ++cnt;
    }
    *lp = 0; *sp = 0;
}

// one job, executed by the job threads (see above)
ZE_FN_NOINLINE void win_run(Win& W, u32 job)
{
    const u32 tid = ZE_JOB_TID, nt = ZE_JOB_NT;
    if (job == WJ_BUILD || job == WJ_RESOLVE) {
#if defined(ZE_PROF) && defined(__CUDA_ARCH__)
#define ZE_PH(i) do { if (tid == 0) { long long t_ = clock64(); W.phase[i] += (u64)(t_ - t_ph); t_ph = t_; } } while (0)
        long long t_ph = clock64();
#else
#define ZE_PH(i) ((void)0)
#endif
        // WJ_RESOLVE: some slots died, the later ones get their bucket chains (which may now reach further back) and paths again
        const u32 from = job == WJ_BUILD ? W.jfrom : W.next, to = job == WJ_BUILD ? W.jto : W.count;
        if (job == WJ_BUILD) {
            for (u32 l = from + tid; l < to; l += nt) win_walk(W, l);
            ZE_PH(0);
            ze_job_sync();
            ZE_PH(1);
            win_links(W);
        }
        for (u32 l = from + tid; l < to; l += nt) win_pairs(W, l);
        ZE_PH(2);
        ze_job_sync();
        ZE_PH(3);
        for (u32 l = from + tid; l < to; l += nt) win_resolve(W, l);
        ZE_PH(4);
        ze_job_sync();
        ZE_PH(5);
        for (u32 l = from + tid; l < to; l += nt) {                  // a truncated walk detaches nodes: later slots of the bucket cannot be resolved
            u32 fl = W.flags[l];
            for (u32 k = 0; k < W.chain[l]; ++k) if (!W.dead[W.cs[l][k]] && (W.flags[W.cs[l][k]] & WF_TRUNC)) fl |= WF_BAD;
            W.flags[l] = (u8)fl;
        }
    } else if (job == WJ_COMMIT) {
        const u32 rounds = W.maxChain;
        for (u32 r = 0; r <= rounds; ++r) {                            // slots of one bucket in position order, buckets side by side
            for (u32 l = tid; l < W.upto; l += nt) if (W.chain[l] == r && !W.dead[l]) win_commit_slot(W, l);
            ze_job_sync();
        }
    }
}
// wait for the posted job
ZE_FN void win_wait(Win& W)
{
    if (!W.pending) return;
#if defined(__CUDA_ARCH__)
    if (blockDim.x > 32) ze_bar_sync(2, blockDim.x);
#endif
    ze_sync();
    if (ZE_LANE == 0) { if (W.job == WJ_BUILD) W.built = W.jto; W.pending = 0; }
    ze_sync();
}
// hand a job to the job threads; the caller's parameters in W must be written (by lane 0) before
ZE_FN void win_post(Win& W, u32 job)
{
    win_wait(W);
    ze_sync();
    if (ZE_LANE == 0) { W.job = job; W.pending = 1; }
    ze_sync();
#if defined(__CUDA_ARCH__)
    if (blockDim.x > 32) ze_bar_arrive(1, blockDim.x);
    else { win_run(W, job); __syncwarp(); }
#else
    win_run(W, job);
#endif
}
ZE_FN void win_dispatch(Win& W, u32 job) { win_post(W, job); win_wait(W); }

// write the links of the consumed slots and close the window
ZE_FN_NOINLINE void win_flush(Work& w)
{
    Win& W = *w.win;
    win_wait(W);
    if (W.count && W.next > 0 && W.baseOff == w.baseOff) {
        ZE_T(t_cm); ZE_CNT(w, 14, 1);
        if (ZE_LANE == 0) W.upto = W.next;
        ze_sync();
        win_dispatch(W, WJ_COMMIT);
        ZE_ACC(w, 5, t_cm);
    }
    ze_sync();
    if (ZE_LANE == 0) { W.count = 0; W.next = 0; }
    ze_sync();
}

// the reference skips the positions [first unconsumed slot, newpos): they never enter the tree
ZE_FN_NOINLINE void win_skip_to(Work& w, u32 newpos)
{
    Win& W = *w.win;
    const u32 t = newpos - W.base;
    if (!W.count || newpos < W.base + W.next || t + 8 >= W.count) { win_flush(w); return; }
    ZE_T(t_rs); ZE_CNT(w, 17, 1);
    for (u32 j = W.next + ZE_LANE; j < t; j += ZE_LANES) W.dead[j] = 1;
    ze_sync();
    if (ZE_LANE == 0) W.next = t;
    ze_sync();
    win_dispatch(W, WJ_RESOLVE);
    ZE_ACC(w, 18, t_rs);
}

// is `pos` the next unconsumed slot of the current window?  Otherwise close it and build a new one starting at pos
// (false when the walk limits would depend on the position: near the window / tree-buffer edge and at the block's tail)
ZE_FN_NOINLINE bool win_ready(Work& w, u32 pos, const u8* iend, u32 mls)
{
    Win& W = *w.win;
    const u8* base = w.src - w.baseOff;
    const u32 endIdx = (u32)(iend - base);
    if (W.count && W.endIdx == endIdx && W.baseOff == w.baseOff && W.base + W.next == pos && W.next < W.count) return true;
    win_flush(w);
    const u32 btMask = (1u << (w.cp.chainLog - 1)) - 1, maxDist = 1u << w.cp.windowLog;
    if (endIdx > btMask || endIdx - w.lowLimit > maxDist || w.lowLimit < 1) return false;
    if (pos + 8 > endIdx) return false;
    ZE_T(t_bd); ZE_CNT(w, 8, 1);
    // insert-heavy parses (btopt: long skips) fill a large window; where nearly every position is a query and long repeats cut
    // windows short (btultra*), a smaller one costs less to build
    u32 cap = (WN_W >= 512 && w.cp.strategy != ST_BTOPT) ? WN_W / 2 : WN_W;
    if (cap > ZE_JOB_NT && ZE_JOB_NT >= 32) cap = ZE_JOB_NT;          // one slot per job thread
    u32 cnt = endIdx - 8 - pos + 1; if (cnt > cap) cnt = cap;
    // (Building a window in stages while the parser consumes the earlier ones was measured and lost: the stages are latency
    // bound, so k stages cost k times the walk latency, and the job warps slow the parser's warp on the same SM.)
    win_wait(W);
    if (ZE_LANE == 0) {
        W.base = pos; W.count = cnt; W.next = 0; W.endIdx = endIdx; W.baseOff = w.baseOff; W.maxChain = 0; W.jfrom = 0; W.jto = cnt;
        W.text = base; W.hashTable = w.hashTable; W.bt = w.chainTable; W.btMask = btMask; W.hashLog = w.cp.hashLog; W.mls = mls;
        W.lowLimit = w.lowLimit; W.budget = 1u << w.cp.searchLog; W.qbase = (mls == 3 ? 3u : 4u) - 1u;
    }
    ze_sync();
    win_dispatch(W, WJ_BUILD);
    ze_sync();
    ZE_ACC(w, 4, t_bd);
    return true;
}

// ZSTD_updateTree_internal (zstd_opt.c:562-582)
ZE_FN_NOINLINE void update_tree(Work& w, u32 target, const u8* iend, u32 mls)
{
    u32 idx = w.nextToUpdate;
    ZE_T(t_ut);
    Win& W = *w.win;
    while (idx < target) {
        if (!(W.count && idx - W.base == W.next && W.next < W.count) && !win_ready(w, idx, iend, mls)) { ZE_CNT(w, 9, 1); idx += insert_bt1(w, idx, iend, target, mls); continue; }
        const u32 l = idx - W.base;
        u32 hi = target - W.base; if (hi > W.count) hi = W.count;
        u32 s = l;                                            // first slot of [l, hi) that is not a plain insert
        while (s < hi) {
            const u32 j = s + ZE_LANE;
            const bool stop = j < hi && ((W.flags[j] & WF_BAD) || W.adv[j] != 1);
            const u32 mk = ze_ballot(stop);
            if (mk) { s += ze_ffs(mk) - 1; break; }
            s += ZE_LANES;
        }
        if (s > hi) s = hi;
        ZE_CNT(w, 13, s - l);
        ze_sync();
        if (ZE_LANE == 0) W.next = s;
        ze_sync();
        idx = W.base + s;
        if (s == hi) continue;
        if (W.flags[s] & WF_BAD) {
            ZE_CNT(w, 15, 1);
            win_flush(w);
            if (s == 0) { ZE_CNT(w, 9, 1); idx += insert_bt1(w, idx, iend, target, mls); }
        } else {                                              // the reference skips positions after this insert
            ZE_CNT(w, 16, 1);
            idx += W.adv[s];
            ze_sync();
            if (ZE_LANE == 0) W.next = s + 1;
            ze_sync();
            win_skip_to(w, idx);
        }
    }
    w.nextToUpdate = target;
    ZE_ACC(w, 3, t_ut);
}

// ZSTD_btGetAllMatches_internal + ZSTD_insertBtAndGetAllMatches (zstd_opt.c:590-820), noDict
ZE_FN u32 get_all_matches_impl(Work& w, Match* matches, u32* nextToUpdate3, const u8* ip, const u8* iLimit, const u32* rep, u32 ll0, u32 lengthToBeat)
{
    const u8* base = w.src - w.baseOff;
    u32 curr = (u32)(ip - base);
    u32 mls = w.cp.minMatch < 3 ? 3 : (w.cp.minMatch > 6 ? 6 : w.cp.minMatch);
    if (curr < w.nextToUpdate) return 0;                 // skipped area
    if (curr != w.nextToUpdate) update_tree(w, curr, iLimit, mls);

    u32 sufficient_len = w.cp.targetLength < OPT_NUM - 1 ? w.cp.targetLength : OPT_NUM - 1;
    u32 minMatch = (mls == 3) ? 3 : 4;
    u32 h = 0, matchIndex = 0;
    u32* bt = w.chainTable;
    u32 btLog = w.cp.chainLog - 1, btMask = (1u << btLog) - 1;
    u32 clSmaller = 0, clLarger = 0;
    u32 dictLimit = w.dictLimit;
    u32 btLow = btMask >= curr ? 0 : curr - btMask;
    u32 windowLow = lowest_match_index(w, curr);
    u32 matchLow = windowLow ? windowLow : 1;
    u32* smallerPtr = bt + 2 * (curr & btMask);
    u32* largerPtr = smallerPtr + 1;
    u32 matchEndIdx = curr + 8 + 1;
    u32* const dummyPtr = w.dummySlot;
    u32 mnum = 0;
    u32 nbCompares = 1u << w.cp.searchLog;
    u32 bestLength = lengthToBeat - 1;

    // repcodes.  Device: the first 40 bytes of all three candidates are compared in one round trip (lanes 0-9 / 10-19 / 20-29 take
    // 4 bytes each of candidate 0 / 1 / 2), which settles most of them; a longer one continues with count_eq.  (The narrow coder
    // only fetches the first word of each candidate that way: its inputs rarely have repcode matches, and the check is cheaper.)
    // Then they are examined in the reference's order.
    {   u32 repLenAll[3] = { 0, 0, 0 };
#if defined(__CUDA_ARCH__)
        if (WN_W < 512) {
            const u32 t0 = rd32(ip);
            u32 off = 0; bool eqL = false;
            if (ZE_LANE < 3) {
                const u32 rc = ll0 + ZE_LANE;
                off = (rc == 3) ? rep[0] - 1 : rep[rc];
                if (off - 1 < curr - dictLimit && curr - off >= windowLow) {
                    const u32 tr = rd32(ip - off);
                    eqL = minMatch == 3 ? ((t0 << 8) == (tr << 8)) : (t0 == tr);
                }
            }
            const u32 eqMask = ze_ballot(eqL) & 7u;
            for (u32 g = 0; g < 3; ++g) if ((eqMask >> g) & 1u) {
                const u32 o = ze_shfl(off, g);
                repLenAll[g] = count_eq(ip + minMatch, ip + minMatch - o, iLimit) + minMatch;
            }
        } else
        {   const u32 grp = ZE_LANE / 10u, k = ZE_LANE - grp * 10u;
            const u32 rem = (u32)(iLimit - ip);
            u32 off = 0; bool ok = false;
            if (grp < 3) {
                const u32 rc = ll0 + grp;
                off = (rc == 3) ? rep[0] - 1 : rep[rc];
                ok = off - 1 < curr - dictLimit && curr - off >= windowLow;
            }
            u32 nb = 0;                                        // equal bytes among this lane's four
            if (ok && 4 * k < rem) {
                const u32 x = rd32(ip + 4 * k) ^ rd32(ip + 4 * k - off);
                nb = x ? ((ze_ffs(x) - 1) >> 3) : 4;
                if (4 * k + nb > rem) nb = rem - 4 * k;
            }
            const u32 stopMask = ze_ballot(grp < 3 && nb < 4);
            const u32 okMask = ze_ballot(ok);
            for (u32 g = 0; g < 3; ++g) {
                if (!((okMask >> (10 * g)) & 1u)) continue;
                const u32 sm = (stopMask >> (10 * g)) & 0x3ffu;
                u32 len;
                if (sm) { const u32 f = ze_ffs(sm) - 1; len = 4 * f + ze_shfl(nb, 10 * g + f); }
                else {
                    const u32 o = ze_shfl(off, 10 * g);
                    len = 40 + count_eq(ip + 40, ip + 40 - o, iLimit);
                }
                repLenAll[g] = len >= minMatch ? len : 0;
            }
        }
#endif
        u32 lastR = 3 + ll0;
        for (u32 rc = ll0; rc < lastR; ++rc) {
            u32 repLen = 0;
#if defined(__CUDA_ARCH__)
            repLen = repLenAll[rc - ll0];
#else
            u32 repOffset = (rc == 3) ? rep[0] - 1 : rep[rc];
            u32 repIndex = curr - repOffset;
            if (repOffset - 1 < curr - dictLimit) {
                bool eq = minMatch == 3 ? ((rd32(ip) << 8) == (rd32(ip - repOffset) << 8)) : (rd32(ip) == rd32(ip - repOffset));
                if ((repIndex >= windowLow) & eq) repLen = count_eq(ip + minMatch, ip + minMatch - repOffset, iLimit) + minMatch;
            }
            (void)repLenAll;
#endif
            if (repLen > bestLength) {
                bestLength = repLen;
                matches[mnum].off = rc - ll0 + 1; matches[mnum].len = repLen; ++mnum;
                if ((repLen > sufficient_len) | (ip + repLen == iLimit)) return mnum;
            }
        }
    }
    // hash3
    if (mls == 3 && bestLength < mls) {
        // ZSTD_insertAndFindFirstIndexHash3 (zstd_opt.c:408-430)
        u32 idx = *nextToUpdate3;
        u32 hash3 = hash3_ptr(ip, w.hashLog3);
        while (idx < curr) { w.hashTable3[hash3_ptr(base + idx, w.hashLog3)] = idx; ++idx; }
        *nextToUpdate3 = curr;
        u32 matchIndex3 = w.hashTable3[hash3];
        if ((matchIndex3 >= matchLow) & (curr - matchIndex3 < (1u << 18))) {
            u32 mlen = count_eq(ip, base + matchIndex3, iLimit);
            if (mlen >= mls) {
                bestLength = mlen;
                matches[0].off = (curr - matchIndex3) + 3; matches[0].len = mlen; mnum = 1;
                if ((mlen > sufficient_len) | (ip + mlen == iLimit)) { w.nextToUpdate = curr + 1; return 1; }
            }
        }
    }
    {   // common case: the position is the next slot of the current window and its match list was resolved with the window
        Win& W = *w.win;
        const u32 l = curr - W.base;
        if (W.count && l == W.next && l < W.count && !(W.flags[l] & WF_BAD) && W.nq[l] != 0xff && lengthToBeat - 1 == W.qbase) {
            ZE_T(t_rp);
            const u32 nq = W.nq[l];
            u32 endMax = 0;
            for (u32 c = 0; c < nq; c += ZE_LANES) {
                const u32 i = c + ZE_LANE;
                u32 off = 0, len = 0;
                if (i < nq) { off = W.qm[l][i][0]; len = W.qm[l][i][1]; }
                const bool take = i < nq && len > bestLength;
                const u32 tk = ze_ballot(take);
                if (take) { const u32 k = mnum + ze_popc(tk & ((1u << ZE_LANE) - 1u)); matches[k].off = off; matches[k].len = len; }
                const u32 e = ze_reduce_max(take ? curr + 3 - off + len : 0u);
                if (e > endMax) endMax = e;
                mnum += ze_popc(tk);
            }
            if (endMax > matchEndIdx) matchEndIdx = endMax;
            const u32 ntu = matchEndIdx - 8;
            ze_sync();
            if (ZE_LANE == 0) W.next = l + 1;
            ze_sync();
            if (ntu != curr + 1) { ZE_CNT(w, 16, 1); win_skip_to(w, ntu); }
            w.nextToUpdate = ntu;
            ZE_ACC(w, 7, t_rp); ZE_CNT(w, 11, 1);
            return mnum;
        }
    }
    if (win_ready(w, curr, iLimit, mls) && (w.win->flags[curr - w.win->base] & WF_BAD) && curr != w.win->base) {
        win_flush(w);                                         // the window ends here: try this position as the first slot of a fresh one
        win_ready(w, curr, iLimit, mls);
    }
    if (win_ready(w, curr, iLimit, mls)) {                     // apply the query's rules to the slot's resolved path
        ZE_T(t_rp);
        Win& W = *w.win;
        const u32 l = curr - W.base;
        const u32 fl = W.flags[l], nst = W.n[l], np = W.np[l], m = W.tm[l], keep = W.keep[l];
        bool usable = !(fl & WF_BAD);
        if (usable && (fl & WF_IEND) && nst <= m && (W.rec[l][nst - 1][1] & 0x7fffffffu) <= bestLength) usable = false;   // the reference compares past iLimit here
        if (usable) {
            // ZE_LANES steps of the resolved path at a time; which steps set a new best length is a prefix maximum over the lanes
            u32 linked = m; bool broke = false;
            for (u32 c = 0; c < m && !broke; c += ZE_LANES) {
                const u32 e = c + ZE_LANE;
                const bool valid = e < m;
                u32 mi = 0, ml = 0;
                if (valid) {
                    if (e < np) { const u32 k = W.pp[l][e]; mi = W.base + W.cs[l][k]; ml = W.pl[l][k] & 0x7fffffffu; }
                    else { const u32 j = nth_set_bit(keep, e - np); mi = W.rec[l][j][0]; ml = W.rec[l][j][1] & 0x7fffffffu; }
                }
                u32 incl = valid ? ml : 0;
                for (u32 d = 1; d < ZE_LANES; d <<= 1) { u32 t = ze_shfl_up(incl, d); if (ZE_LANE >= d && t > incl) incl = t; }
                u32 excl = ze_shfl_up(incl, 1); if (ZE_LANE == 0) excl = 0;
                if (excl < bestLength) excl = bestLength;
                const bool rec = valid && ml > excl;
                const u32 brk = ze_ballot(rec && ((ml > OPT_NUM) | (ip + ml == iLimit)));
                const u32 lim = brk ? ze_ffs(brk) - 1 : 32;                          // step that ends the walk (reported, not linked)
                const u32 uptoM = lim >= 31 ? 0xffffffffu : ((2u << lim) - 1u);
                const u32 below = (1u << ZE_LANE) - 1u;
                const u32 recmask = ze_ballot(rec) & uptoM;
                const bool myrec = (recmask >> ZE_LANE) & 1u;
                if (myrec) { const u32 k = mnum + ze_popc(recmask & below); matches[k].off = (curr - mi) + 3; matches[k].len = ml; }
                const u32 mxEnd = ze_reduce_max(myrec ? mi + ml : 0u), mxLen = ze_reduce_max(myrec ? ml : 0u);
                if (mxEnd > matchEndIdx) matchEndIdx = mxEnd;
                if (mxLen > bestLength) bestLength = mxLen;
                mnum += ze_popc(recmask);
                if (brk) { broke = true; linked = c + lim; }
            }
            ze_sync();
            const u32 ntu = matchEndIdx - 8;
            if (ZE_LANE == 0) { W.next = l + 1; if (broke && linked < W.tn[l]) W.tn[l] = (u8)linked; }
            ze_sync();
            if (broke) win_flush(w); else if (ntu != curr + 1) { ZE_CNT(w, 16, 1); win_skip_to(w, ntu); }
            w.nextToUpdate = ntu;
            ZE_ACC(w, 7, t_rp); ZE_CNT(w, 11, 1);
            return mnum;
        }
    }
    win_flush(w);                                             // sequential walk on the tree in memory
    h = hash_ptr(ip, w.cp.hashLog, mls);
    matchIndex = w.hashTable[h];
    w.hashTable[h] = curr;
    ZE_T(t_sq); ZE_CNT(w, 10, 1);
    for (; nbCompares && matchIndex >= matchLow; --nbCompares) {
        u32* nextPtr = bt + 2 * (matchIndex & btMask);
        u32 ml = clSmaller < clLarger ? clSmaller : clLarger;
        const u8* match = base + matchIndex;
        ml += count_eq(ip + ml, match + ml, iLimit);
        if (ml > bestLength) {
            if (ml > matchEndIdx - matchIndex) matchEndIdx = matchIndex + ml;
            bestLength = ml;
            matches[mnum].off = (curr - matchIndex) + 3; matches[mnum].len = ml; ++mnum;
            if ((ml > OPT_NUM) | (ip + ml == iLimit)) break;
        }
        if (match[ml] < ip[ml]) {
            *smallerPtr = matchIndex; clSmaller = ml;
            if (matchIndex <= btLow) { smallerPtr = dummyPtr; break; }
            smallerPtr = nextPtr + 1; matchIndex = nextPtr[1];
        } else {
            *largerPtr = matchIndex; clLarger = ml;
            if (matchIndex <= btLow) { largerPtr = dummyPtr + 1; break; }
            largerPtr = nextPtr; matchIndex = nextPtr[0];
        }
    }
    *largerPtr = 0;
    *smallerPtr = 0;
    w.nextToUpdate = matchEndIdx - 8;
    ZE_ACC(w, 6, t_sq);
    return mnum;
}
ZE_FN u32 get_all_matches(Work& w, Match* matches, u32* nextToUpdate3, const u8* ip, const u8* iLimit, const u32* rep, u32 ll0, u32 lengthToBeat)
{
    ZE_T(t_gm); ZE_CNT(w, 12, 1);
    u32 r = get_all_matches_impl(w, matches, nextToUpdate3, ip, iLimit, rep, ll0, lengthToBeat);
    ZE_ACC(w, 2, t_gm);
    return r;
}

// ---------------------------------------------------------------------------------------------------- price model (zstd_opt.c:30-380)
ZE_FN u32 bit_weight(u32 stat) { return highbit(stat + 1) * BITCOST_MULT; }
ZE_FN u32 frac_weight(u32 raw) { u32 stat = raw + 1, hb = highbit(stat); return hb * BITCOST_MULT + ((stat << 8) >> hb); }
ZE_FN u32 weight(u32 stat, int optLevel) { return optLevel ? frac_weight(stat) : bit_weight(stat); }

ZE_FN void set_base_prices(Work& w, int optLevel)
{
    w.litSumBP = weight(w.litSum, optLevel); w.llSumBP = weight(w.llSum, optLevel);
    w.mlSumBP = weight(w.mlSum, optLevel); w.ofSumBP = weight(w.ofSum, optLevel);
}
ZE_FN_NOINLINE u32 downscale(u32* t, u32 last, u32 shift, int base1)
{
    u32 sum = 0;
    for (u32 s = 0; s <= last; ++s) { u32 b = base1 ? 1 : (t[s] > 0); u32 n = b + (t[s] >> shift); sum += n; t[s] = n; }
    return sum;
}
ZE_FN_NOINLINE u32 scale_stats(u32* t, u32 last, u32 logTarget)
{
    u32 prev = 0; for (u32 s = 0; s <= last; ++s) prev += t[s];
    u32 factor = prev >> logTarget;
    if (factor <= 1) return prev;
    return downscale(t, last, highbit(factor), 1);
}
// ZSTD_rescaleFreqs (zstd_opt.c:140-258), no dictionary
ZE_FN_NOINLINE void rescale_freqs(Work& w, const u8* src, u32 srcSize, int optLevel)
{
    w.pricePredef = 0;
    if (w.llSum == 0) {
        if (srcSize <= 8) w.pricePredef = 1;
        for (u32 i = 0; i < 256; ++i) w.litFreq[i] = 0;
        for (u32 i = 0; i < srcSize; ++i) w.litFreq[src[i]]++;
        w.litSum = downscale(w.litFreq, 255, 8, 0);
        for (u32 i = 0; i <= MaxLL; ++i) w.llFreq[i] = 1;
        w.llFreq[0] = 4; w.llFreq[1] = 2; w.llSum = MaxLL + 1 + 3 + 1;
        for (u32 i = 0; i <= MaxML; ++i) w.mlFreq[i] = 1;
        w.mlSum = MaxML + 1;
        {   const u8 b[32] = { 6,2,1,1,2,3,4,4, 4,3,2,1,1,1,1,1, 1,1,1,1,1,1,1,1, 1,1,1,1,1,1,1,1 };
            u32 s = 0; for (u32 i = 0; i <= MaxOff; ++i) { w.ofFreq[i] = b[i]; s += b[i]; } w.ofSum = s; }
    } else {
        w.litSum = scale_stats(w.litFreq, 255, 12);
        w.llSum = scale_stats(w.llFreq, MaxLL, 11);
        w.mlSum = scale_stats(w.mlFreq, MaxML, 11);
        w.ofSum = scale_stats(w.ofFreq, MaxOff, 11);
    }
    set_base_prices(w, optLevel);
}
ZE_FN u32 lit_cost1(const Work& w, u8 lit, int optLevel)        // ZSTD_rawLiteralsCost(p, 1)
{
    if (w.pricePredef) return 6 * BITCOST_MULT;
    u32 price = w.litSumBP, maxp = w.litSumBP - BITCOST_MULT;
    u32 lp = weight(w.litFreq[lit], optLevel);
    if (lp > maxp) lp = maxp;
    return price - lp;
}
ZE_FN u32 ll_price(const Work& w, u32 ll, int optLevel)         // ZSTD_litLengthPrice
{
    if (w.pricePredef) return weight(ll, optLevel);
    u32 extra = 0;
    if (ll == BLOCK_MAX) { extra = BITCOST_MULT; ll = BLOCK_MAX - 1; }
    u32 c = LLcode(ll);
    return extra + kLLbits[c] * BITCOST_MULT + w.llSumBP - weight(w.llFreq[c], optLevel);
}
ZE_FN u32 match_price(const Work& w, u32 offBase, u32 ml, int optLevel)   // ZSTD_getMatchPrice
{
    u32 offCode = highbit(offBase), mlBase = ml - 3;
    if (w.pricePredef) return weight(mlBase, optLevel) + (16 + offCode) * BITCOST_MULT;
    u32 price = offCode * BITCOST_MULT + (w.ofSumBP - weight(w.ofFreq[offCode], optLevel));
    if (optLevel < 2 && offCode >= 20) price += (offCode - 19) * 2 * BITCOST_MULT;
    u32 mc = MLcode(mlBase);
    price += kMLbits[mc] * BITCOST_MULT + (w.mlSumBP - weight(w.mlFreq[mc], optLevel));
    price += BITCOST_MULT / 5;
    return price;
}
// Price tables.  The statistics only change when sequences are stored, so a price is a table lookup plus the base price of
// its category: priceTab = lit[256] weight(litFreq), then per code extraBits * 256 - weight(freq) for ll[36], ml[53], off[32]
// (offset codes >= 20 carry btopt's handicap).  refresh_prices rebuilds everything (block start); update_stats_t patches the
// entries one stored sequence touches.  Same unsigned arithmetic as ZSTD_litLengthPrice / ZSTD_getMatchPrice, regrouped.
struct BP { u32 lit, ll, ml, of; };
ZE_FN BP base_prices(const Work& w) { BP b; b.lit = w.litSumBP; b.ll = w.llSumBP; b.ml = w.mlSumBP; b.of = w.ofSumBP; return b; }
ZE_FN u32 of_entry(const Work& w, u32 c, int optLevel) { u32 v = c * BITCOST_MULT - weight(w.ofFreq[c], optLevel); if (optLevel < 2 && c >= 20) v += (c - 19) * 2 * BITCOST_MULT; return v; }
ZE_FN_NOINLINE void refresh_prices(Work& w, int optLevel)
{
    if (w.pricePredef) return;
    u32* T = w.priceTab;
    for (u32 i = ZE_LANE; i < 256 + 36 + 53 + 32; i += ZE_LANES) {
        u32 v;
        if (i < 256) v = weight(w.litFreq[i], optLevel);
        else if (i < 256 + 36) { const u32 c = i - 256; v = kLLbits[c] * BITCOST_MULT - weight(w.llFreq[c], optLevel); }
        else if (i < 256 + 36 + 53) { const u32 c = i - 292; v = kMLbits[c] * BITCOST_MULT - weight(w.mlFreq[c], optLevel); }
        else v = of_entry(w, i - 345, optLevel);
        T[i] = v;
    }
    ze_sync();
}
ZE_FN u32 lit_cost1_t(const Work& w, const BP& bp, u8 lit)
{
    if (w.pricePredef) return 6 * BITCOST_MULT;
    u32 lp = w.priceTab[lit]; const u32 maxp = bp.lit - BITCOST_MULT;
    if (lp > maxp) lp = maxp;
    return bp.lit - lp;
}
ZE_FN u32 ll_price_t(const Work& w, const BP& bp, u32 ll, int optLevel)
{
    if (w.pricePredef) return weight(ll, optLevel);
    u32 extra = 0;
    if (ll == BLOCK_MAX) { extra = BITCOST_MULT; ll = BLOCK_MAX - 1; }
    return extra + w.priceTab[256 + LLcode(ll)] + bp.ll;
}
ZE_FN u32 match_price_t(const Work& w, const BP& bp, u32 offBase, u32 ml, int optLevel)
{
    if (w.pricePredef) return match_price(w, offBase, ml, optLevel);
    return w.priceTab[345 + highbit(offBase)] + bp.of + w.priceTab[292 + MLcode(ml - 3)] + bp.ml + BITCOST_MULT / 5;
}
// ZSTD_updateStats (zstd_opt.c:356) + the table entries it invalidates
ZE_FN_NOINLINE void update_stats_t(Work& w, u32 ll, const u8* lits, u32 offBase, u32 ml, int optLevel)
{
    for (u32 u = 0; u < ll; ++u) w.litFreq[lits[u]] += 2;
    w.litSum += ll * 2;
    const u32 lc = LLcode(ll), oc = highbit(offBase), mc = MLcode(ml - 3);
    w.llFreq[lc]++; w.llSum++;
    w.ofFreq[oc]++; w.ofSum++;
    w.mlFreq[mc]++; w.mlSum++;
    if (w.pricePredef) return;
    u32* T = w.priceTab;
    ze_sync();
    for (u32 u = ZE_LANE; u < ll; u += ZE_LANES) T[lits[u]] = weight(w.litFreq[lits[u]], optLevel);
    T[256 + lc] = kLLbits[lc] * BITCOST_MULT - weight(w.llFreq[lc], optLevel);
    T[292 + mc] = kMLbits[mc] * BITCOST_MULT - weight(w.mlFreq[mc], optLevel);
    T[345 + oc] = of_entry(w, oc, optLevel);
    ze_sync();
}
ZE_FN_NOINLINE void update_stats(Work& w, u32 ll, const u8* lits, u32 offBase, u32 ml)
{
    for (u32 u = 0; u < ll; ++u) w.litFreq[lits[u]] += 2;
    w.litSum += ll * 2;
    w.llFreq[LLcode(ll)]++; w.llSum++;
    w.ofFreq[highbit(offBase)]++; w.ofSum++;
    w.mlFreq[MLcode(ml - 3)]++; w.mlSum++;
}
// ZSTD_storeSeq (zstd_compress_internal.h:649-706)
ZE_FN_NOINLINE void store_seq(SeqStore& ss, u32 ll, const u8* lits, u32 offBase, u32 ml)
{
    for (u32 i = ZE_LANE; i < ll; i += ZE_LANES) ss.lit[i] = lits[i];
    ze_sync();
    ss.lit += ll;
    u32 pos = (u32)(ss.seq - ss.seqStart);
    if (ll > 0xFFFF) { ss.longType = 1; ss.longPos = pos; }
    ss.seq->litLength = (u16)ll;
    ss.seq->offBase = offBase;
    u32 mlBase = ml - 3;
    if (mlBase > 0xFFFF) { ss.longType = 2; ss.longPos = pos; }
    ss.seq->mlBase = (u16)mlBase;
    ss.seq++;
}

// ---------------------------------------------------------------------------------------------------- optimal parser
// ZSTD_compressBlock_opt_generic (zstd_opt.c:1075-1437), noDict, no LDM.  Returns the size of the last literals run.
ZE_FN_NOINLINE u32 compress_block_opt_impl(Work& w, u32* rep, const u8* src, u32 srcSize, int optLevel)
{
    const u8* istart = src; const u8* ip = istart; const u8* anchor = istart;
    const u8* iend = istart + srcSize; const u8* ilimit = iend - 8;
    const u8* prefixStart = w.src - w.baseOff + w.dictLimit;
    u32 sufficient_len = w.cp.targetLength < OPT_NUM - 1 ? w.cp.targetLength : OPT_NUM - 1;
    u32 minMatch = (w.cp.minMatch == 3) ? 3 : 4;
    u32 nextToUpdate3 = w.nextToUpdate;
    Opt* opt = w.opt; Match* matches = w.matches;
    Opt lastStretch; lastStretch.price = 0; lastStretch.off = lastStretch.mlen = lastStretch.litlen = 0; lastStretch.rep[0] = lastStretch.rep[1] = lastStretch.rep[2] = 0;

    rescale_freqs(w, src, srcSize, optLevel);
    refresh_prices(w, optLevel);
    BP bp = base_prices(w);
    ip += (ip == prefixStart);

    while (ip < ilimit) {
        u32 cur, last_pos = 0;
        bool shortest = false;
        {   u32 litlen = (u32)(ip - anchor), ll0 = !litlen;
            u32 nbMatches = get_all_matches(w, matches, &nextToUpdate3, ip, iend, rep, ll0, minMatch);
            if (!nbMatches) { ip++; continue; }
            opt[0].mlen = 0; opt[0].litlen = litlen; opt[0].price = (i32)ll_price_t(w, bp, litlen, optLevel);
            opt[0].rep[0] = rep[0]; opt[0].rep[1] = rep[1]; opt[0].rep[2] = rep[2];
            {   u32 maxML = matches[nbMatches - 1].len, maxOff = matches[nbMatches - 1].off;
                if (maxML > sufficient_len) {
                    lastStretch.litlen = 0; lastStretch.mlen = maxML; lastStretch.off = maxOff;
                    cur = 0; last_pos = maxML; shortest = true;
                }
            }
            if (!shortest) {
                u32 pos;
                for (pos = 1; pos < minMatch; pos++) { opt[pos].price = MAX_PRICE; opt[pos].mlen = 0; opt[pos].litlen = litlen + pos; }
                {   const i32 ll0p = (i32)ll_price_t(w, bp, 0, optLevel), p0 = opt[0].price;
                    for (u32 m = 0; m < nbMatches; m++) {
                        u32 offBase = matches[m].off, end = matches[m].len;
                        for (u32 q = pos + ZE_LANE; q <= end; q += ZE_LANES) {          // lanes take different lengths
                            i32 mp = (i32)match_price_t(w, bp, offBase, q, optLevel);
                            opt[q].mlen = q; opt[q].off = offBase; opt[q].litlen = 0;
                            opt[q].price = p0 + mp + ll0p;
                        }
                        if (end >= pos) pos = end + 1;
                    }
                    ze_sync();
                }
                last_pos = pos - 1;
                opt[pos].price = MAX_PRICE;
            }
        }
        if (!shortest) {
            for (cur = 1; cur <= last_pos; cur++) {
                const u8* inr = ip + cur;
                {   u32 litlen = opt[cur - 1].litlen + 1;
                    i32 price = opt[cur - 1].price + (i32)lit_cost1_t(w, bp, ip[cur - 1])
                              + ((i32)ll_price_t(w, bp, litlen, optLevel) - (i32)ll_price_t(w, bp, litlen - 1, optLevel));
                    if (price <= opt[cur].price) {
                        Opt prevMatch = opt[cur];
                        opt[cur] = opt[cur - 1];
                        opt[cur].litlen = litlen; opt[cur].price = price;
                        if (optLevel >= 1 && prevMatch.litlen == 0
                            && ((i32)ll_price_t(w, bp, 1, optLevel) - (i32)ll_price_t(w, bp, 0, optLevel)) < 0
                            && ip + cur < iend) {
                            i32 with1 = prevMatch.price + (i32)lit_cost1_t(w, bp, ip[cur])
                                      + ((i32)ll_price_t(w, bp, 1, optLevel) - (i32)ll_price_t(w, bp, 0, optLevel));
                            i32 withMore = price + (i32)lit_cost1_t(w, bp, ip[cur])
                                         + ((i32)ll_price_t(w, bp, litlen + 1, optLevel) - (i32)ll_price_t(w, bp, litlen, optLevel));
                            if (with1 < withMore && with1 < opt[cur + 1].price) {
                                u32 prev = cur - prevMatch.mlen;
                                u32 nr[3] = { opt[prev].rep[0], opt[prev].rep[1], opt[prev].rep[2] };
                                update_rep(nr, prevMatch.off, opt[prev].litlen == 0);
                                opt[cur + 1] = prevMatch;
                                opt[cur + 1].rep[0] = nr[0]; opt[cur + 1].rep[1] = nr[1]; opt[cur + 1].rep[2] = nr[2];
                                opt[cur + 1].litlen = 1; opt[cur + 1].price = with1;
                                if (last_pos < cur + 1) last_pos = cur + 1;
                            }
                        }
                    }
                }
                if (opt[cur].litlen == 0) {
                    u32 prev = cur - opt[cur].mlen;
                    u32 nr[3] = { opt[prev].rep[0], opt[prev].rep[1], opt[prev].rep[2] };
                    update_rep(nr, opt[cur].off, opt[prev].litlen == 0);
                    opt[cur].rep[0] = nr[0]; opt[cur].rep[1] = nr[1]; opt[cur].rep[2] = nr[2];
                }
                if (inr > ilimit) continue;
                if (cur == last_pos) break;
                if (optLevel == 0 && opt[cur + 1].price <= opt[cur].price + (i32)(BITCOST_MULT / 2)) continue;
                {   u32 ll0 = (opt[cur].litlen == 0);
                    i32 basePrice = opt[cur].price + (i32)ll_price_t(w, bp, 0, optLevel);
                    u32 nbMatches = get_all_matches(w, matches, &nextToUpdate3, inr, iend, opt[cur].rep, ll0, minMatch);
                    if (!nbMatches) continue;
                    {   u32 longestML = matches[nbMatches - 1].len;
                        if (longestML > sufficient_len || cur + longestML >= OPT_NUM || ip + cur + longestML >= iend) {
                            lastStretch.mlen = longestML; lastStretch.off = matches[nbMatches - 1].off; lastStretch.litlen = 0;
                            last_pos = cur + longestML; shortest = true;
                            break;
                        }
                    }
                    for (u32 m = 0; m < nbMatches; m++) {
                        u32 offset = matches[m].off, lastML = matches[m].len;
                        u32 startML = m > 0 ? matches[m - 1].len + 1 : minMatch;
                        if (optLevel == 0) {
                            // btopt scans the lengths downwards and stops at the first one that does not improve its position.
                            // Every length has its own position, so the tests are independent: lanes take ZE_LANES lengths at a
                            // time and the first failing one (ballot) ends the scan exactly where the sequential loop would.
                            if (lastML >= startML) {
                                const u32 top = cur + lastML;
                                if (top > last_pos) {
                                    for (u32 q = last_pos + 1 + ZE_LANE; q <= top; q += ZE_LANES) { opt[q].price = MAX_PRICE; opt[q].litlen = 1; }
                                    ze_sync();
                                    last_pos = top;
                                }
                                for (u32 c = 0; c <= lastML - startML; c += ZE_LANES) {
                                    const u32 d = c + ZE_LANE; const bool in = d <= lastML - startML;
                                    const u32 mlen = lastML - d, pos = cur + mlen;
                                    i32 price = 0; bool ok = false;
                                    if (in) { price = basePrice + (i32)match_price_t(w, bp, offset, mlen, optLevel); ok = price < opt[pos].price; }
                                    const u32 fail = ze_ballot(in && !ok);
                                    const u32 upto = fail ? ze_ffs(fail) - 1 : 32;
                                    if (in && ZE_LANE < upto) { opt[pos].mlen = mlen; opt[pos].off = offset; opt[pos].litlen = 0; opt[pos].price = price; }
                                    if (fail) break;
                                }
                                ze_sync();
                            }
                        } else if (lastML >= startML) {            // btultra(2): every length is tried -> lanes take different lengths
                            u32 top = cur + lastML;
                            if (top > last_pos) {
                                for (u32 q = last_pos + 1 + ZE_LANE; q <= top; q += ZE_LANES) { opt[q].price = MAX_PRICE; opt[q].litlen = 1; }
                                ze_sync();
                                last_pos = top;
                            }
                            for (u32 mlen = startML + ZE_LANE; mlen <= lastML; mlen += ZE_LANES) {
                                u32 pos = cur + mlen;
                                i32 price = basePrice + (i32)match_price_t(w, bp, offset, mlen, optLevel);
                                if (price < opt[pos].price) { opt[pos].mlen = mlen; opt[pos].off = offset; opt[pos].litlen = 0; opt[pos].price = price; }
                            }
                            ze_sync();
                        }
                    }
                }
                opt[last_pos + 1].price = MAX_PRICE;
            }
            if (!shortest) { lastStretch = opt[last_pos]; cur = last_pos - lastStretch.mlen; }
        }
        // _shortestPath
        if (lastStretch.mlen == 0) { ip += last_pos; continue; }
        if (lastStretch.litlen == 0) {
            u32 nr[3] = { opt[cur].rep[0], opt[cur].rep[1], opt[cur].rep[2] };
            update_rep(nr, lastStretch.off, opt[cur].litlen == 0);
            rep[0] = nr[0]; rep[1] = nr[1]; rep[2] = nr[2];
        } else {
            rep[0] = lastStretch.rep[0]; rep[1] = lastStretch.rep[1]; rep[2] = lastStretch.rep[2];
            cur -= lastStretch.litlen;
        }
        {   u32 storeEnd = cur + 2, storeStart, stretchPos = cur;
            if (lastStretch.litlen > 0) {           // (sic) the reference falls through into the next block: kept
                opt[storeEnd].litlen = lastStretch.litlen; opt[storeEnd].mlen = 0;
                storeStart = storeEnd - 1; opt[storeStart] = lastStretch;
            }
            opt[storeEnd] = lastStretch; storeStart = storeEnd;
            while (1) {
                Opt nextStretch = opt[stretchPos];
                opt[storeStart].litlen = nextStretch.litlen;
                if (nextStretch.mlen == 0) break;
                storeStart--;
                opt[storeStart] = nextStretch;
                stretchPos -= nextStretch.litlen + nextStretch.mlen;
            }
            for (u32 sp = storeStart; sp <= storeEnd; sp++) {
                u32 llen = opt[sp].litlen, mlen = opt[sp].mlen, offBase = opt[sp].off, advance = llen + mlen;
                if (mlen == 0) { ip = anchor + llen; continue; }
                update_stats_t(w, llen, anchor, offBase, mlen, optLevel);
                store_seq(w.ss, llen, anchor, offBase, mlen);
                anchor += advance; ip = anchor;
            }
            set_base_prices(w, optLevel);
            bp = base_prices(w);
        }
    }
    win_flush(w);                                             // the next block (other limits, maybe another index base) starts a fresh window
    return (u32)(iend - anchor);
}
ZE_FN u32 compress_block_opt(Work& w, u32* rep, const u8* src, u32 srcSize, int optLevel)
{
    ZE_T(t_bo);
    u32 r = compress_block_opt_impl(w, rep, src, srcSize, optLevel);
    ZE_ACC(w, 1, t_bo);
    return r;
}


// ---------------------------------------------------------------------------------------------------- btlazy2 (zstd_lazy.c)
// Level 13 on inputs above 256 KB (tuple-packed references of very long segments) selects ZSTD_btlazy2: the lazy parser at
// depth 2 over the "delayed update" binary tree.  Rare in AGC's call population, so it is restated as it is -- sequential,
// every lane of the parser's warp executes the same scalar steps -- without a window engine.
// ZSTD_updateDUBT (zstd_lazy.c:29-65): positions enter the tree as an unsorted chain through the hash bucket
ZE_FN_NOINLINE void update_dubt(Work& w, u32 target, u32 mls)
{
    const u8* base = w.src - w.baseOff;
    u32* bt = w.chainTable; const u32 btMask = (1u << (w.cp.chainLog - 1)) - 1;
    for (u32 idx = w.nextToUpdate; idx < target; ++idx) {
        const u32 h = hash_ptr(base + idx, w.cp.hashLog, mls);
        const u32 matchIndex = w.hashTable[h];
        ze_sync();
        w.hashTable[h] = idx;
        bt[2 * (idx & btMask)] = matchIndex;                 // next candidate
        bt[2 * (idx & btMask) + 1] = 1;                       // ZSTD_DUBT_UNSORTED_MARK
        ze_sync();
    }
    w.nextToUpdate = target;
}
// ZSTD_insertDUBT1 (zstd_lazy.c:68-164), noDict: sort one already chained position into the tree
ZE_FN_NOINLINE void insert_dubt1(Work& w, u32 curr, const u8* iend, u32 nbCompares, u32 btLow)
{
    const u8* base = w.src - w.baseOff;
    u32* bt = w.chainTable; const u32 btMask = (1u << (w.cp.chainLog - 1)) - 1;
    const u8* ip = base + curr;
    u32 clSmaller = 0, clLarger = 0;
    u32* smallerPtr = bt + 2 * (curr & btMask);
    u32* largerPtr = smallerPtr + 1;
    u32 matchIndex = *smallerPtr;
    u32* const dummyPtr = w.dummySlot;
    const u32 maxDist = 1u << w.cp.windowLog;
    const u32 windowLow = (curr - w.lowLimit > maxDist) ? curr - maxDist : w.lowLimit;
    ze_sync();
    for (; nbCompares && matchIndex > windowLow; --nbCompares) {
        u32* nextPtr = bt + 2 * (matchIndex & btMask);
        u32 ml = clSmaller < clLarger ? clSmaller : clLarger;
        const u8* match = base + matchIndex;
        ml += count_eq(ip + ml, match + ml, iend);
        if (ip + ml == iend) break;
        const u32 n0 = nextPtr[0], n1 = nextPtr[1];
        ze_sync();
        if (match[ml] < ip[ml]) {
            *smallerPtr = matchIndex; clSmaller = ml;
            if (matchIndex <= btLow) { smallerPtr = dummyPtr; break; }
            smallerPtr = nextPtr + 1; matchIndex = n1;
        } else {
            *largerPtr = matchIndex; clLarger = ml;
            if (matchIndex <= btLow) { largerPtr = dummyPtr + 1; break; }
            largerPtr = nextPtr; matchIndex = n0;
        }
        ze_sync();
    }
    ze_sync();
    *smallerPtr = 0; *largerPtr = 0;
    ze_sync();
}
// ZSTD_BtFindBestMatch + ZSTD_DUBT_findBestMatch (zstd_lazy.c:243-404), noDict.  *offBase must hold the caller's current value
ZE_FN_NOINLINE u32 dubt_find_best(Work& w, const u8* ip, const u8* iend, u32* offBasePtr, u32 mls)
{
    const u8* base = w.src - w.baseOff;
    const u32 curr = (u32)(ip - base);
    if (curr < w.nextToUpdate) return 0;                     // skipped area
    update_dubt(w, curr, mls);
    u32* bt = w.chainTable; const u32 btMask = (1u << (w.cp.chainLog - 1)) - 1;
    const u32 h = hash_ptr(ip, w.cp.hashLog, mls);
    u32 matchIndex = w.hashTable[h];
    const u32 windowLow = lowest_match_index(w, curr);
    const u32 btLow = btMask >= curr ? 0 : curr - btMask;
    const u32 unsortLimit = btLow > windowLow ? btLow : windowLow;
    u32* nextCandidate = bt + 2 * (matchIndex & btMask);
    u32* unsortedMark = nextCandidate + 1;
    u32 nbCompares = 1u << w.cp.searchLog, nbCandidates = nbCompares, previousCandidate = 0;
    ze_sync();
    // reach the end of the unsorted candidates, turning their marks into a reversed chain
    while (matchIndex > unsortLimit && *unsortedMark == 1 && nbCandidates > 1) {
        const u32 nxt = *nextCandidate;
        ze_sync();
        *unsortedMark = previousCandidate;
        previousCandidate = matchIndex;
        matchIndex = nxt;
        nextCandidate = bt + 2 * (matchIndex & btMask);
        unsortedMark = nextCandidate + 1;
        nbCandidates--;
        ze_sync();
    }
    if (matchIndex > unsortLimit && *unsortedMark == 1) { ze_sync(); *nextCandidate = 0; *unsortedMark = 0; }
    ze_sync();
    // batch sort the stacked candidates
    matchIndex = previousCandidate;
    while (matchIndex) {
        const u32 nextIdx = bt[2 * (matchIndex & btMask) + 1];
        ze_sync();
        insert_dubt1(w, matchIndex, iend, nbCandidates, unsortLimit);
        matchIndex = nextIdx;
        nbCandidates++;
    }
    // find the longest match (and insert curr)
    u32 clSmaller = 0, clLarger = 0;
    u32* smallerPtr = bt + 2 * (curr & btMask);
    u32* largerPtr = smallerPtr + 1;
    u32* const dummyPtr = w.dummySlot;
    u32 matchEndIdx = curr + 8 + 1, bestLength = 0;
    matchIndex = w.hashTable[h];
    ze_sync();
    w.hashTable[h] = curr;
    for (; nbCompares && matchIndex > windowLow; --nbCompares) {
        u32* nextPtr = bt + 2 * (matchIndex & btMask);
        u32 ml = clSmaller < clLarger ? clSmaller : clLarger;
        const u8* match = base + matchIndex;
        ml += count_eq(ip + ml, match + ml, iend);
        if (ml > bestLength) {
            if (ml > matchEndIdx - matchIndex) matchEndIdx = matchIndex + ml;
            if ((i32)(4 * (ml - bestLength)) > (i32)(highbit(curr - matchIndex + 1) - highbit(*offBasePtr))) { bestLength = ml; *offBasePtr = (curr - matchIndex) + 3; }
            if (ip + ml == iend) break;
        }
        const u32 n0 = nextPtr[0], n1 = nextPtr[1];
        ze_sync();
        if (match[ml] < ip[ml]) {
            *smallerPtr = matchIndex; clSmaller = ml;
            if (matchIndex <= btLow) { smallerPtr = dummyPtr; break; }
            smallerPtr = nextPtr + 1; matchIndex = n1;
        } else {
            *largerPtr = matchIndex; clLarger = ml;
            if (matchIndex <= btLow) { largerPtr = dummyPtr + 1; break; }
            largerPtr = nextPtr; matchIndex = n0;
        }
        ze_sync();
    }
    ze_sync();
    *smallerPtr = 0; *largerPtr = 0;
    ze_sync();
    w.nextToUpdate = matchEndIdx - 8;
    return bestLength;
}
// ZSTD_compressBlock_btlazy2 = ZSTD_compressBlock_lazy_generic(search_binaryTree, depth 2, noDict) (zstd_lazy.c:1516-1777)
ZE_FN_NOINLINE u32 compress_block_btlazy2(Work& w, u32* rep, const u8* src, u32 srcSize)
{
    const u8* istart = src; const u8* ip = istart; const u8* anchor = istart;
    const u8* iend = istart + srcSize; const u8* ilimit = iend - 8;
    const u8* base = w.src - w.baseOff;
    const u8* prefixLowest = base + w.dictLimit;
    const u32 mls = w.cp.minMatch < 4 ? 4 : (w.cp.minMatch > 6 ? 6 : w.cp.minMatch);
    u32 offset_1 = rep[0], offset_2 = rep[1], offsetSaved1 = 0, offsetSaved2 = 0;
    ip += (ip == prefixLowest);
    {   const u32 curr = (u32)(ip - base), maxDist = 1u << w.cp.windowLog;
        const u32 windowLow = (curr - w.dictLimit > maxDist) ? curr - maxDist : w.dictLimit;     // ZSTD_getLowestPrefixIndex
        const u32 maxRep = curr - windowLow;
        if (offset_2 > maxRep) { offsetSaved2 = offset_2; offset_2 = 0; }
        if (offset_1 > maxRep) { offsetSaved1 = offset_1; offset_1 = 0; }
    }
    while (ip < ilimit) {
        u32 matchLength = 0, offBase = 1;                     // REPCODE1_TO_OFFBASE
        const u8* start = ip + 1;
        if ((offset_1 > 0) & (rd32(ip + 1 - offset_1) == rd32(ip + 1))) matchLength = count_eq(ip + 1 + 4, ip + 1 + 4 - offset_1, iend) + 4;
        {   u32 offFound = 999999999u;
            const u32 ml2 = dubt_find_best(w, ip, iend, &offFound, mls);
            if (ml2 > matchLength) { matchLength = ml2; start = ip; offBase = offFound; }
        }
        if (matchLength < 4) { ip += ((u32)(ip - anchor) >> 8) + 1; continue; }       // kSearchStrength
        while (ip < ilimit) {                                 // depth 1 and 2
            ip++;
            if (offBase && ((offset_1 > 0) & (rd32(ip) == rd32(ip - offset_1)))) {
                const u32 mlRep = count_eq(ip + 4, ip + 4 - offset_1, iend) + 4;
                const i32 gain2 = (i32)(mlRep * 3), gain1 = (i32)(matchLength * 3 - highbit(offBase) + 1);
                if (mlRep >= 4 && gain2 > gain1) { matchLength = mlRep; offBase = 1; start = ip; }
            }
            {   u32 ofb = 999999999u;
                const u32 ml2 = dubt_find_best(w, ip, iend, &ofb, mls);
                const i32 gain2 = (i32)(ml2 * 4 - highbit(ofb)), gain1 = (i32)(matchLength * 4 - highbit(offBase) + 4);
                if (ml2 >= 4 && gain2 > gain1) { matchLength = ml2; offBase = ofb; start = ip; continue; }
            }
            if (ip < ilimit) {
                ip++;
                if (offBase && ((offset_1 > 0) & (rd32(ip) == rd32(ip - offset_1)))) {
                    const u32 mlRep = count_eq(ip + 4, ip + 4 - offset_1, iend) + 4;
                    const i32 gain2 = (i32)(mlRep * 4), gain1 = (i32)(matchLength * 4 - highbit(offBase) + 1);
                    if (mlRep >= 4 && gain2 > gain1) { matchLength = mlRep; offBase = 1; start = ip; }
                }
                {   u32 ofb = 999999999u;
                    const u32 ml2 = dubt_find_best(w, ip, iend, &ofb, mls);
                    const i32 gain2 = (i32)(ml2 * 4 - highbit(ofb)), gain1 = (i32)(matchLength * 4 - highbit(offBase) + 7);
                    if (ml2 >= 4 && gain2 > gain1) { matchLength = ml2; offBase = ofb; start = ip; continue; }
                }
            }
            break;
        }
        if (offBase > 3) {                                    // catch up
            const u32 off = offBase - 3;
            while (((start > anchor) & (start - off > prefixLowest)) && start[-1] == (start - off)[-1]) { start--; matchLength++; }
            offset_2 = offset_1; offset_1 = off;
        }
        {   const u32 litLength = (u32)(start - anchor);
            store_seq(w.ss, litLength, anchor, offBase, matchLength);
            anchor = ip = start + matchLength;
        }
        while (((ip <= ilimit) & (offset_2 > 0)) && rd32(ip) == rd32(ip - offset_2)) {      // immediate repcode
            matchLength = count_eq(ip + 4, ip + 4 - offset_2, iend) + 4;
            { const u32 t = offset_2; offset_2 = offset_1; offset_1 = t; }
            store_seq(w.ss, 0, anchor, 1, matchLength);
            ip += matchLength; anchor = ip;
        }
    }
    offsetSaved2 = (offsetSaved1 != 0 && offset_1 != 0) ? offsetSaved1 : offsetSaved2;
    rep[0] = offset_1 ? offset_1 : offsetSaved1;
    rep[1] = offset_2 ? offset_2 : offsetSaved2;
    return (u32)(iend - anchor);
}

// ---------------------------------------------------------------------------------------------------- bit stream (common/bitstream.h)
struct BitW { u64 cont; u32 pos; u8* start; u8* ptr; u8* end; };
ZE_FN bool bit_init(BitW& b, u8* dst, u64 cap) { b.cont = 0; b.pos = 0; b.start = b.ptr = dst; b.end = dst + cap - 8; return cap > 8; }
ZE_FN void bit_add(BitW& b, u64 v, u32 n) { if (n) b.cont |= (v & ((1ull << n) - 1)) << b.pos; b.pos += n; }
ZE_FN void bit_flush(BitW& b)
{
    u32 nb = b.pos >> 3;
    for (u32 i = 0; i < 8; ++i) b.ptr[i] = (u8)(b.cont >> (8 * i));
    b.ptr += nb; if (b.ptr > b.end) b.ptr = b.end;
    b.pos &= 7; b.cont = nb >= 8 ? 0 : b.cont >> (nb * 8);
}
ZE_FN u32 bit_close(BitW& b) { bit_add(b, 1, 1); bit_flush(b); if (b.ptr >= b.end) return 0; return (u32)(b.ptr - b.start) + (b.pos > 0); }

// ---------------------------------------------------------------------------------------------------- FSE (fse_compress.c)
ZE_FN u32 fse_min_table_log(u32 srcSize, u32 maxSym) { u32 a = highbit(srcSize) + 1, b = highbit(maxSym) + 2; return a < b ? a : b; }
ZE_FN u32 fse_optimal_table_log(u32 maxLog, u32 srcSize, u32 maxSym, u32 minus)      // FSE_optimalTableLog_internal (:343)
{
    u32 maxBitsSrc = highbit(srcSize - 1) - minus, tl = maxLog, minBits = fse_min_table_log(srcSize, maxSym);
    if (tl == 0) tl = 11;
    if (maxBitsSrc < tl) tl = maxBitsSrc;
    if (minBits > tl) tl = minBits;
    if (tl < 5) tl = 5;
    if (tl > 12) tl = 12;
    return tl;
}
// FSE_normalizeM2 (:366) ; returns false on failure
ZE_FN_NOINLINE bool fse_normalize_m2(i16* norm, u32 tableLog, const u32* count, u64 total, u32 maxSym, i16 lowProb)
{
    const i16 NYA = -2;
    u32 distributed = 0, ToDistribute;
    u32 lowThreshold = (u32)(total >> tableLog);
    u32 lowOne = (u32)((total * 3) >> (tableLog + 1));
    for (u32 s = 0; s <= maxSym; s++) {
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = lowProb; distributed++; total -= count[s]; continue; }
        if (count[s] <= lowOne) { norm[s] = 1; distributed++; total -= count[s]; continue; }
        norm[s] = NYA;
    }
    ToDistribute = (1u << tableLog) - distributed;
    if (ToDistribute == 0) return true;
    if ((total / ToDistribute) > lowOne) {
        lowOne = (u32)((total * 3) / (ToDistribute * 2));
        for (u32 s = 0; s <= maxSym; s++)
            if (norm[s] == NYA && count[s] <= lowOne) { norm[s] = 1; distributed++; total -= count[s]; }
        ToDistribute = (1u << tableLog) - distributed;
    }
    if (distributed == maxSym + 1) {
        u32 maxV = 0, maxC = 0;
        for (u32 s = 0; s <= maxSym; s++) if (count[s] > maxC) { maxV = s; maxC = count[s]; }
        norm[maxV] += (i16)ToDistribute;
        return true;
    }
    if (total == 0) {
        for (u32 s = 0; ToDistribute > 0; s = (s + 1) % (maxSym + 1)) if (norm[s] > 0) { ToDistribute--; norm[s]++; }
        return true;
    }
    {   u64 vStepLog = 62 - tableLog, mid = (1ull << (vStepLog - 1)) - 1;
        u64 rStep = ((((u64)1 << vStepLog) * ToDistribute) + mid) / (u32)total;
        u64 tmpTotal = mid;
        for (u32 s = 0; s <= maxSym; s++) if (norm[s] == NYA) {
            u64 end = tmpTotal + (count[s] * rStep);
            u32 sStart = (u32)(tmpTotal >> vStepLog), sEnd = (u32)(end >> vStepLog), wgt = sEnd - sStart;
            if (wgt < 1) return false;
            norm[s] = (i16)wgt; tmpTotal = end;
        }
    }
    return true;
}
// FSE_normalizeCount (:450).  returns tableLog, 0 for the rle special case, ~0u on error
ZE_FN_NOINLINE u32 fse_normalize(i16* norm, u32 tableLog, const u32* count, u32 total, u32 maxSym, u32 useLowProb)
{
    if (tableLog == 0) tableLog = 11;
    if (tableLog < 5 || tableLog > 12) return ~0u;
    if (tableLog < fse_min_table_log(total, maxSym)) return ~0u;
    i16 lowProb = useLowProb ? -1 : 1;
    u64 scale = 62 - tableLog, step = ((u64)1 << 62) / total, vStep = 1ull << (scale - 20);
    i32 still = 1 << tableLog;
    u32 largest = 0; i16 largestP = 0;
    u32 lowThreshold = total >> tableLog;
    for (u32 s = 0; s <= maxSym; s++) {
        if (count[s] == total) return 0;
        if (count[s] == 0) { norm[s] = 0; continue; }
        if (count[s] <= lowThreshold) { norm[s] = lowProb; still--; }
        else {
            i16 proba = (i16)((count[s] * step) >> scale);
            if (proba < 8) { u64 restToBeat = vStep * kRtb[proba]; proba += (count[s] * step) - ((u64)proba << scale) > restToBeat; }
            if (proba > largestP) { largestP = proba; largest = s; }
            norm[s] = proba; still -= proba;
        }
    }
    if (-still >= (norm[largest] >> 1)) { if (!fse_normalize_m2(norm, tableLog, count, total, maxSym, lowProb)) return ~0u; }
    else norm[largest] += (i16)still;
    return tableLog;
}
// FSE_writeNCount_generic (:233), output buffer always large enough here.  returns size, 0 on error
ZE_FN_NOINLINE u32 fse_write_ncount(u8* out0, const i16* norm, u32 maxSym, u32 tableLog)
{
    u8* out = out0;
    i32 nbBits, tableSize = 1 << tableLog, remaining, threshold;
    u32 bitStream = 0; i32 bitCount = 0; u32 symbol = 0, alphabetSize = maxSym + 1; int previousIs0 = 0;
    bitStream += (tableLog - 5) << bitCount; bitCount += 4;
    remaining = tableSize + 1; threshold = tableSize; nbBits = (i32)tableLog + 1;
    while (symbol < alphabetSize && remaining > 1) {
        if (previousIs0) {
            u32 start = symbol;
            while (symbol < alphabetSize && !norm[symbol]) symbol++;
            if (symbol == alphabetSize) break;
            while (symbol >= start + 24) { start += 24; bitStream += 0xFFFFu << bitCount; out[0] = (u8)bitStream; out[1] = (u8)(bitStream >> 8); out += 2; bitStream >>= 16; }
            while (symbol >= start + 3) { start += 3; bitStream += 3u << bitCount; bitCount += 2; }
            bitStream += (symbol - start) << bitCount; bitCount += 2;
            if (bitCount > 16) { out[0] = (u8)bitStream; out[1] = (u8)(bitStream >> 8); out += 2; bitStream >>= 16; bitCount -= 16; }
        }
        {   i32 cnt = norm[symbol++];
            i32 mx = (2 * threshold - 1) - remaining;
            remaining -= cnt < 0 ? -cnt : cnt;
            cnt++;
            if (cnt >= threshold) cnt += mx;
            bitStream += (u32)cnt << bitCount;
            bitCount += nbBits; bitCount -= (cnt < mx);
            previousIs0 = (cnt == 1);
            if (remaining < 1) return 0;
            while (remaining < threshold) { nbBits--; threshold >>= 1; }
        }
        if (bitCount > 16) { out[0] = (u8)bitStream; out[1] = (u8)(bitStream >> 8); out += 2; bitStream >>= 16; bitCount -= 16; }
    }
    if (remaining != 1) return 0;
    out[0] = (u8)bitStream; out[1] = (u8)(bitStream >> 8); out += (bitCount + 7) / 8;
    return (u32)(out - out0);
}
// FSE_buildCTable_wksp (:56).  tableSymbol scratch needs tableSize bytes, cumul maxSym+2 u16 (taken from `scratch`)
ZE_FN_NOINLINE void fse_build_ctable(FseCT& ct, const i16* norm, u32 maxSym, u32 tableLog, u8* scratch)
{
    u32 tableSize = 1u << tableLog, tableMask = tableSize - 1, step = (tableSize >> 1) + (tableSize >> 3) + 3, maxSV1 = maxSym + 1;
    u16* cumul = (u16*)scratch;                       // maxSV1 + 1 entries (<= 54)
    u8* tableSymbol = scratch + 128;                  // tableSize entries
    u32 high = tableSize - 1;
    ct.tableLog = (u16)tableLog; ct.maxSym = (u16)maxSym;
    cumul[0] = 0;
    for (u32 u = 1; u <= maxSV1; u++) {
        if (norm[u - 1] == -1) { cumul[u] = cumul[u - 1] + 1; tableSymbol[high--] = (u8)(u - 1); }
        else cumul[u] = cumul[u - 1] + (u16)norm[u - 1];
    }
    cumul[maxSV1] = (u16)(tableSize + 1);
    {   // symbol spreading: both reference code paths visit positions 0, step, 2*step ... skipping the low-prob area
        u32 position = 0;
        for (u32 sy = 0; sy < maxSV1; sy++) {
            i32 freq = norm[sy];
            for (i32 i = 0; i < freq; i++) {
                tableSymbol[position] = (u8)sy;
                position = (position + step) & tableMask;
                while (position > high) position = (position + step) & tableMask;
            }
        }
    }
    for (u32 u = 0; u < tableSize; u++) { u8 sy = tableSymbol[u]; ct.state[cumul[sy]++] = (u16)(tableSize + u); }
    {   u32 total = 0;
        for (u32 sy = 0; sy <= maxSym; sy++) {
            i32 n = norm[sy];
            if (n == 0) { ct.dBits[sy] = ((tableLog + 1) << 16) - (1u << tableLog); ct.dFind[sy] = 0; }
            else if (n == -1 || n == 1) { ct.dBits[sy] = (tableLog << 16) - (1u << tableLog); ct.dFind[sy] = (i32)(total - 1); total++; }
            else {
                u32 maxBitsOut = tableLog - highbit((u32)n - 1), minStatePlus = (u32)n << maxBitsOut;
                ct.dBits[sy] = (maxBitsOut << 16) - minStatePlus; ct.dFind[sy] = (i32)(total - (u32)n); total += (u32)n;
            }
        }
    }
}
ZE_FN_NOINLINE void fse_build_rle(FseCT& ct, u8 sym)              // FSE_buildCTable_rle (:531)
{
    ct.tableLog = 0; ct.maxSym = sym; ct.state[0] = 0; ct.state[1] = 0; ct.dBits[sym] = 0; ct.dFind[sym] = 0;
}
struct FseState { u32 value; const FseCT* ct; };
ZE_FN void fse_init2(FseState& st, const FseCT& ct, u32 sym)       // FSE_initCState2 (common/fse.h:452)
{
    st.ct = &ct;
    u32 nbBitsOut = (ct.dBits[sym] + (1u << 15)) >> 16;
    u32 v = (nbBitsOut << 16) - ct.dBits[sym];
    st.value = ct.state[(i32)(v >> nbBitsOut) + ct.dFind[sym]];
}
ZE_FN void fse_encode(BitW& b, FseState& st, u32 sym)              // FSE_encodeSymbol (:463)
{
    u32 nbBitsOut = (st.value + st.ct->dBits[sym]) >> 16;
    bit_add(b, st.value, nbBitsOut);
    st.value = st.ct->state[(i32)(st.value >> nbBitsOut) + st.ct->dFind[sym]];
}
ZE_FN void fse_flush_state(BitW& b, const FseState& st) { bit_add(b, st.value, st.ct->tableLog); bit_flush(b); }
ZE_FN u32 fse_bit_cost(const FseCT& ct, u32 sym, u32 accuracyLog)   // FSE_bitCost (:494)
{
    u32 tableLog = ct.tableLog, minNbBits = ct.dBits[sym] >> 16, threshold = (minNbBits + 1) << 16, tableSize = 1u << tableLog;
    u32 deltaFromThreshold = threshold - (ct.dBits[sym] + tableSize);
    u32 normalizedDelta = (deltaFromThreshold << accuracyLog) >> tableLog;
    return ((minNbBits + 1) << accuracyLog) - normalizedDelta;
}

// ---------------------------------------------------------------------------------------------------- histogram (hist.c)
ZE_FN_NOINLINE u32 hist(u32* count, u32* maxSymPtr, const u8* src, u32 n)    // returns largest count, sets highest present symbol
{
    u32 ms = *maxSymPtr;
    for (u32 i = 0; i <= ms; ++i) count[i] = 0;
    if (n == 0) { *maxSymPtr = 0; return 0; }
    for (u32 i = 0; i < n; ++i) count[src[i]]++;
    while (!count[ms]) ms--;
    *maxSymPtr = ms;
    u32 largest = 0;
    for (u32 i = 0; i <= ms; ++i) if (count[i] > largest) largest = count[i];
    return largest;
}

// the same over a warp: lanes count into the shared-memory table with atomics, the counters are then copied to `count`
ZE_FN_NOINLINE u32 hist_w(Work& w, u32* count, u32* maxSymPtr, const u8* src, u32 n)
{
#if defined(__CUDA_ARCH__)
    if (w.histTab && n >= 64) {
        u32 ms = *maxSymPtr;
        u32* T = w.histTab;
        for (u32 i = ZE_LANE; i < 256; i += ZE_LANES) T[i] = 0;
        ze_sync();
        for (u32 i = ZE_LANE; i < n; i += ZE_LANES) atomicAdd(&T[src[i]], 1u);
        ze_sync();
        u32 top = 0, largest = 0;
        for (u32 i = ZE_LANE; i <= ms; i += ZE_LANES) { const u32 c = T[i]; count[i] = c; if (c) top = i; if (c > largest) largest = c; }
        ze_sync();
        *maxSymPtr = ze_reduce_max(top);
        return ze_reduce_max(largest);
    }
#endif
    return hist(count, maxSymPtr, src, n);
}

// ---------------------------------------------------------------------------------------------------- Huffman (huf_compress.c)
ZE_FN u32 huf_get_index(u32 c) { return c < 166 ? c : highbit(c) + 158; }       // HUF_getIndex (:497)
ZE_FN_NOINLINE void huf_insertion_sort(HufNode* a, i32 low, i32 high)
{
    i32 size = high - low + 1; a += low;
    for (i32 i = 1; i < size; ++i) { HufNode key = a[i]; i32 j = i - 1; while (j >= 0 && a[j].count < key.count) { a[j + 1] = a[j]; j--; } a[j + 1] = key; }
}
ZE_FN_NOINLINE i32 huf_partition(HufNode* arr, i32 low, i32 high)
{
    u32 pivot = arr[high].count; i32 i = low - 1;
    for (i32 j = low; j < high; j++) if (arr[j].count > pivot) { i++; HufNode t = arr[i]; arr[i] = arr[j]; arr[j] = t; }
    HufNode t = arr[i + 1]; arr[i + 1] = arr[high]; arr[high] = t;
    return i + 1;
}
// HUF_simpleQuickSort (:548) with its recursion unrolled onto an explicit stack (same visiting order => same result)
ZE_FN_NOINLINE void huf_quick_sort(HufNode* arr, i32 low0, i32 high0)
{
    i32 stk[160][2]; i32 sp = 0;
    stk[sp][0] = low0; stk[sp][1] = high0; sp++;
    while (sp) {
        sp--; i32 low = stk[sp][0], high = stk[sp][1];
        if (high - low < 8) { huf_insertion_sort(arr, low, high); continue; }     // threshold only at function entry, as in the reference
        // the reference recurses into the smaller side and keeps partitioning the larger one in place; the two sub-ranges
        // are disjoint, so handling the deferred one later gives the same array
        while (low < high) {
            i32 idx = huf_partition(arr, low, high);
            if (idx - low < high - idx) { if (idx - 1 > low) { stk[sp][0] = low; stk[sp][1] = idx - 1; sp++; } low = idx + 1; }
            else { if (high > idx + 1) { stk[sp][0] = idx + 1; stk[sp][1] = high; sp++; } high = idx - 1; }
        }
    }
}

// HUF_sort (:627): bucket sort by descending count (counts >= 166 share log2 buckets that are quick-sorted)
ZE_FN_NOINLINE void huf_sort(HufNode* huffNode, const u32* count, u32 maxSym, u32* rankPos /* 192*2: base, curr */)
{
    u32 n1 = maxSym + 1;
    for (u32 i = 0; i < 192 * 2; ++i) rankPos[i] = 0;
    for (u32 n = 0; n < n1; ++n) rankPos[2 * huf_get_index(count[n])]++;
    for (u32 n = 191; n > 0; --n) { rankPos[2 * (n - 1)] += rankPos[2 * n]; rankPos[2 * (n - 1) + 1] = rankPos[2 * (n - 1)]; }
    for (u32 n = 0; n < n1; ++n) {
        u32 c = count[n], r = huf_get_index(c) + 1, pos = rankPos[2 * r + 1]++;
        huffNode[pos].count = c; huffNode[pos].byte = (u8)n;
    }
    for (u32 n = 166; n < 191; ++n) {
        i32 bucketSize = (i32)rankPos[2 * n + 1] - (i32)rankPos[2 * n];
        u32 start = rankPos[2 * n];
        if (bucketSize > 1) huf_quick_sort(huffNode + start, 0, bucketSize - 1);
    }
}
// HUF_buildTree (:680). huffNode[-1] must be addressable (huffNode = table + 1)
ZE_FN_NOINLINE i32 huf_build_tree(HufNode* huffNode, u32 maxSym)
{
    const i32 STARTNODE = 256;
    HufNode* huffNode0 = huffNode - 1;
    i32 nonNullRank = (i32)maxSym, lowS, lowN, nodeNb = STARTNODE, n, nodeRoot;
    while (huffNode[nonNullRank].count == 0) nonNullRank--;
    lowS = nonNullRank; nodeRoot = nodeNb + lowS - 1; lowN = nodeNb;
    huffNode[nodeNb].count = huffNode[lowS].count + huffNode[lowS - 1].count;
    huffNode[lowS].parent = huffNode[lowS - 1].parent = (u16)nodeNb;
    nodeNb++; lowS -= 2;
    for (n = nodeNb; n <= nodeRoot; n++) huffNode[n].count = 1u << 30;
    huffNode0[0].count = 1u << 31;
    while (nodeNb <= nodeRoot) {
        i32 n1 = (huffNode[lowS].count < huffNode[lowN].count) ? lowS-- : lowN++;
        i32 n2 = (huffNode[lowS].count < huffNode[lowN].count) ? lowS-- : lowN++;
        huffNode[nodeNb].count = huffNode[n1].count + huffNode[n2].count;
        huffNode[n1].parent = huffNode[n2].parent = (u16)nodeNb;
        nodeNb++;
    }
    huffNode[nodeRoot].nbBits = 0;
    for (n = nodeRoot - 1; n >= STARTNODE; n--) huffNode[n].nbBits = huffNode[huffNode[n].parent].nbBits + 1;
    for (n = 0; n <= nonNullRank; n++) huffNode[n].nbBits = huffNode[huffNode[n].parent].nbBits + 1;
    return nonNullRank;
}
// HUF_setMaxHeight (:340)
ZE_FN_NOINLINE u32 huf_set_max_height(HufNode* huffNode, u32 lastNonNull, u32 targetNbBits)
{
    u32 largestBits = huffNode[lastNonNull].nbBits;
    if (largestBits <= targetNbBits) return largestBits;
    {   i32 totalCost = 0; u32 baseCost = 1u << (largestBits - targetNbBits); i32 n = (i32)lastNonNull;
        while (huffNode[n].nbBits > targetNbBits) { totalCost += (i32)(baseCost - (1u << (largestBits - huffNode[n].nbBits))); huffNode[n].nbBits = (u8)targetNbBits; n--; }
        while (huffNode[n].nbBits == targetNbBits) --n;
        totalCost >>= (largestBits - targetNbBits);
        {   const u32 noSymbol = 0xF0F0F0F0;
            u32 rankLast[14];
            for (u32 i = 0; i < 14; ++i) rankLast[i] = noSymbol;
            {   u32 currentNbBits = targetNbBits;
                for (i32 pos = n; pos >= 0; pos--) {
                    if (huffNode[pos].nbBits >= currentNbBits) continue;
                    currentNbBits = huffNode[pos].nbBits;
                    rankLast[targetNbBits - currentNbBits] = (u32)pos;
                }
            }
            while (totalCost > 0) {
                u32 nBitsToDecrease = highbit((u32)totalCost) + 1;
                for (; nBitsToDecrease > 1; nBitsToDecrease--) {
                    u32 highPos = rankLast[nBitsToDecrease], lowPos = rankLast[nBitsToDecrease - 1];
                    if (highPos == noSymbol) continue;
                    if (lowPos == noSymbol) break;
                    {   u32 highTotal = huffNode[highPos].count, lowTotal = 2 * huffNode[lowPos].count;
                        if (highTotal <= lowTotal) break; }
                }
                while (nBitsToDecrease <= 12 && rankLast[nBitsToDecrease] == noSymbol) nBitsToDecrease++;
                totalCost -= 1 << (nBitsToDecrease - 1);
                huffNode[rankLast[nBitsToDecrease]].nbBits++;
                if (rankLast[nBitsToDecrease - 1] == noSymbol) rankLast[nBitsToDecrease - 1] = rankLast[nBitsToDecrease];
                if (rankLast[nBitsToDecrease] == 0) rankLast[nBitsToDecrease] = noSymbol;
                else {
                    rankLast[nBitsToDecrease]--;
                    if (huffNode[rankLast[nBitsToDecrease]].nbBits != targetNbBits - nBitsToDecrease) rankLast[nBitsToDecrease] = noSymbol;
                }
            }
            while (totalCost < 0) {
                if (rankLast[1] == noSymbol) {
                    while (huffNode[n].nbBits == targetNbBits) n--;
                    huffNode[n + 1].nbBits--;
                    rankLast[1] = (u32)(n + 1);
                    totalCost++;
                    continue;
                }
                huffNode[rankLast[1] + 1].nbBits--;
                rankLast[1]++;
                totalCost++;
            }
        }
    }
    return targetNbBits;
}
// HUF_buildCTable_wksp (:693) + HUF_buildCTableFromTree (:660). returns maxNbBits
ZE_FN_NOINLINE u32 huf_build_ctable(Work& w, HufCT& ct, const u32* count, u32 maxSym, u32 maxNbBits)
{
    HufNode* huffNode0 = w.huffNode; HufNode* huffNode = huffNode0 + 1;
    if (maxNbBits == 0) maxNbBits = 11;
    for (u32 i = 0; i < 513; ++i) { huffNode0[i].count = 0; huffNode0[i].parent = 0; huffNode0[i].byte = 0; huffNode0[i].nbBits = 0; }
    huf_sort(huffNode, count, maxSym, w.rankPos);
    i32 nonNullRank = huf_build_tree(huffNode, maxSym);
    maxNbBits = huf_set_max_height(huffNode, (u32)nonNullRank, maxNbBits);
    {   u16 nbPerRank[13], valPerRank[13];
        for (u32 i = 0; i < 13; ++i) { nbPerRank[i] = 0; valPerRank[i] = 0; }
        for (i32 n = 0; n <= nonNullRank; n++) nbPerRank[huffNode[n].nbBits]++;
        {   u16 mn = 0; for (i32 n = (i32)maxNbBits; n > 0; n--) { valPerRank[n] = mn; mn += nbPerRank[n]; mn >>= 1; } }
        for (u32 i = 0; i < 256; ++i) { ct.nbBits[i] = 0; ct.val[i] = 0; }
        for (u32 n = 0; n <= maxSym; n++) ct.nbBits[huffNode[n].byte] = huffNode[n].nbBits;
        for (u32 n = 0; n <= maxSym; n++) ct.val[n] = valPerRank[ct.nbBits[n]]++;       // (value kept even for 0-bit symbols: never read)
        ct.tableLog = maxNbBits; ct.maxSym = maxSym;
    }
    return maxNbBits;
}
ZE_FN_NOINLINE u32 huf_estimate_size(const HufCT& ct, const u32* count, u32 maxSym)       // HUF_estimateCompressedSize (:722)
{
    u64 nbBits = 0;
    for (u32 s = 0; s <= maxSym; ++s) nbBits += (u64)ct.nbBits[s] * count[s];
    return (u32)(nbBits >> 3);
}
ZE_FN_NOINLINE int huf_validate(const HufCT& ct, const u32* count, u32 maxSym)            // HUF_validateCTable (:733)
{
    if (ct.maxSym < maxSym) return 0;
    int bad = 0;
    for (u32 s = 0; s <= maxSym; ++s) bad |= (count[s] != 0) & (ct.nbBits[s] == 0);
    return !bad;
}
// FSE_compress_usingCTable_generic (fse_compress.c:556) -- only used for the Huffman weights.  returns 0 if not compressible
ZE_FN_NOINLINE u32 fse_compress_weights(u8* dst, u32 dstSize, const u8* src, u32 srcSize, const FseCT& ct)
{
    const u8* ip = src + srcSize;
    BitW b; FseState s1, s2;
    if (srcSize <= 2) return 0;
    if (!bit_init(b, dst, dstSize)) return 0;
    if (srcSize & 1) { fse_init2(s1, ct, *--ip); fse_init2(s2, ct, *--ip); fse_encode(b, s1, *--ip); bit_flush(b); }
    else { fse_init2(s2, ct, *--ip); fse_init2(s1, ct, *--ip); }
    srcSize -= 2;
    if (srcSize & 2) { fse_encode(b, s2, *--ip); fse_encode(b, s1, *--ip); bit_flush(b); }
    while (ip > src) {
        fse_encode(b, s2, *--ip); fse_encode(b, s1, *--ip);
        fse_encode(b, s2, *--ip); fse_encode(b, s1, *--ip);
        bit_flush(b);
    }
    fse_flush_state(b, s2); fse_flush_state(b, s1);
    return bit_close(b);
}
// HUF_compressWeights (:128).  returns 0 = not compressible, 1 = rle, else size
ZE_FN_NOINLINE u32 huf_compress_weights(Work& w, u8* dst, u32 dstSize, const u8* wt, u32 wtSize)
{
    u32 maxSym = 12, tableLog = 6;
    u32 cnt[13]; i16 norm[13];
    if (wtSize <= 1) return 0;
    {   u32 mc = hist(cnt, &maxSym, wt, wtSize);
        if (mc == wtSize) return 1;
        if (mc == 1) return 0; }
    tableLog = fse_optimal_table_log(tableLog, wtSize, maxSym, 2);
    if (fse_normalize(norm, tableLog, cnt, wtSize, maxSym, 0) == ~0u) return 0;      // (cannot fail for valid weights)
    u8* op = dst;
    {   u32 h = fse_write_ncount(op, norm, maxSym, tableLog); if (!h) return 0; op += h; }
    fse_build_ctable(*w.tmpCT, norm, maxSym, tableLog, w.scratch);
    {   u32 c = fse_compress_weights(op, (u32)(dst + dstSize - op), wt, wtSize, *w.tmpCT); if (c == 0) return 0; op += c; }
    return (u32)(op - dst);
}
// HUF_writeCTable_wksp (:226). dst must hold 129+ bytes. returns header size (0 on failure)
ZE_FN_NOINLINE u32 huf_write_ctable(Work& w, u8* dst, u32 maxDst, const HufCT& ct, u32 maxSym, u32 huffLog)
{
    u8 bitsToWeight[13]; u8* huffWeight = w.scratch + 768;     // 256 bytes
    bitsToWeight[0] = 0;
    for (u32 n = 1; n < huffLog + 1; n++) bitsToWeight[n] = (u8)(huffLog + 1 - n);
    for (u32 n = 0; n < maxSym; n++) huffWeight[n] = bitsToWeight[ct.nbBits[n]];
    {   u32 hSize = huf_compress_weights(w, dst + 1, maxDst - 1, huffWeight, maxSym);
        if ((hSize > 1) & (hSize < maxSym / 2)) { dst[0] = (u8)hSize; return hSize + 1; } }
    if (maxSym > 128) return 0;
    dst[0] = (u8)(128 + (maxSym - 1));
    huffWeight[maxSym] = 0;
    for (u32 n = 0; n < maxSym; n += 2) dst[(n / 2) + 1] = (u8)((huffWeight[n] << 4) + huffWeight[n + 1]);
    return ((maxSym + 1) / 2) + 1;
}
// HUF_optimalTableLog (:1233)
ZE_FN_NOINLINE u32 huf_optimal_table_log(Work& w, u32 maxTableLog, u32 srcSize, u32 maxSym, HufCT& table, const u32* count, int optimalDepth)
{
    if (!optimalDepth) return fse_optimal_table_log(maxTableLog, srcSize, maxSym, 1);
    u8* dst = w.scratch + 256;           // 512 bytes
    u32 card = 0; for (u32 i = 0; i <= maxSym; ++i) card += count[i] != 0;
    u32 minTableLog = highbit(card) + 1;
    u64 optSize = ~0ull - 1; u32 optLog = maxTableLog;
    for (u32 guess = minTableLog; guess <= maxTableLog; guess++) {
        u32 maxBits = huf_build_ctable(w, table, count, maxSym, guess);
        if (maxBits < guess && guess > minTableLog) break;
        u32 hSize = huf_write_ctable(w, dst, 500, table, maxSym, maxBits);
        if (hSize == 0) continue;
        u64 newSize = (u64)huf_estimate_size(table, count, maxSym) + hSize;
        if (newSize > optSize + 1) break;
        if (newSize < optSize) { optSize = newSize; optLog = guess; }
    }
    return optLog;
}
// HUF_compress1X_usingCTable_internal_body (:1003): the bit stream is the codes of src[n-1] .. src[0], then the end mark
ZE_FN_NOINLINE u32 huf_compress1x(u8* dst, u32 dstSize, const u8* src, u32 srcSize, const HufCT& ct)
{
    if (dstSize < 8) return 0;
    BitW b; if (!bit_init(b, dst, dstSize)) return 0;
    for (u32 n = srcSize; n > 0; --n) {
        u8 sy = src[n - 1];
        bit_add(b, ct.val[sy], ct.nbBits[sy]);
        if (b.pos > 48) bit_flush(b);
    }
    return bit_close(b);
}
// HUF_compress4X_usingCTable_internal (:1068)
ZE_FN_NOINLINE u32 huf_compress4x(u8* dst, u32 dstSize, const u8* src, u32 srcSize, const HufCT& ct)
{
    u32 segmentSize = (srcSize + 3) / 4;
    const u8* ip = src; const u8* iend = src + srcSize;
    u8* op = dst; u8* oend = dst + dstSize;
    if (dstSize < 6 + 1 + 1 + 1 + 8) return 0;
    if (srcSize < 12) return 0;
    op += 6;
    for (int i = 0; i < 3; ++i) {
        u32 c = huf_compress1x(op, (u32)(oend - op), ip, segmentSize, ct);
        if (c == 0 || c > 65535) return 0;
        wr16(dst + 2 * i, c); op += c; ip += segmentSize;
    }
    {   u32 c = huf_compress1x(op, (u32)(oend - op), ip, (u32)(iend - ip), ct);
        if (c == 0 || c > 65535) return 0;
        op += c; }
    return (u32)(op - dst);
}
ZE_FN_NOINLINE u32 huf_compress_ctable(u8* ostart, u8* op, u8* oend, const u8* src, u32 srcSize, int single, const HufCT& ct)   // HUF_compressCTable_internal (:1126)
{
    u32 c = single ? huf_compress1x(op, (u32)(oend - op), src, srcSize, ct) : huf_compress4x(op, (u32)(oend - op), src, srcSize, ct);
    if (c == 0) return 0;
    op += c;
    if ((u32)(op - ostart) >= srcSize - 1) return 0;
    return (u32)(op - ostart);
}
// HUF_compress_internal (:1285) as called by ZSTD_compressLiterals (no preferRepeat: strategy >= lazy).
// returns 0 = not compressible, 1 = single symbol (dst[0] = symbol), else size.  *repeat updated as in the reference.
ZE_FN_NOINLINE u32 huf_compress(Work& w, u8* dst, u32 dstSize, const u8* src, u32 srcSize, int single, HufCT& oldTable, i32* repeat,
                       int optimalDepth, int suspectUncompressible)
{
    u8* ostart = dst; u8* oend = dst + dstSize; u8* op = ostart;
    u32 maxSym = 255, huffLog = 11;
    if (!srcSize || !dstSize) return 0;
    if (suspectUncompressible && srcSize >= 4096 * 10) {
        u32 lt = 0, m1 = 255, m2 = 255;
        lt += hist_w(w, w.count, &m1, src, 4096);
        lt += hist_w(w, w.count, &m2, src + srcSize - 4096, 4096);
        if (lt <= ((2 * 4096) >> 7) + 4) return 0;
    }
    {   u32 largest = hist_w(w, w.count, &maxSym, src, srcSize);
        if (largest == srcSize) { *ostart = src[0]; return 1; }
        if (largest <= (srcSize >> 7) + 4) return 0; }
    if (*repeat == REP_CHECK && !huf_validate(oldTable, w.count, maxSym)) *repeat = REP_NONE;
    HufCT& table = *w.tmpHuf;
    huffLog = huf_optimal_table_log(w, huffLog, srcSize, maxSym, table, w.count, optimalDepth);
    huffLog = huf_build_ctable(w, table, w.count, maxSym, huffLog);
    {   u32 hSize = huf_write_ctable(w, op, dstSize, table, maxSym, huffLog);
        if (hSize == 0) return 0;                                  // (error paths of the reference end as "not compressible")
        if (*repeat != REP_NONE) {
            u32 oldSize = huf_estimate_size(oldTable, w.count, maxSym), newSize = huf_estimate_size(table, w.count, maxSym);
            if (oldSize <= hSize + newSize || hSize + 12 >= srcSize)
                return huf_compress_ctable(ostart, op, oend, src, srcSize, single, oldTable);
        }
        if (hSize + 12ul >= srcSize) return 0;
        op += hSize;
        *repeat = REP_NONE;
        oldTable = table;
    }
    return huf_compress_ctable(ostart, op, oend, src, srcSize, single, table);
}
ZE_FN u32 min_gain(u32 srcSize, u32 strat) { u32 minlog = strat >= ST_BTULTRA ? strat - 1 : 6; return (srcSize >> minlog) + 2; }   // ZSTD_minGain

// ---------------------------------------------------------------------------------------------------- literals (zstd_compress_literals.c)
ZE_FN_NOINLINE u32 no_compress_literals(u8* dst, const u8* src, u32 srcSize)
{
    u32 fl = 1 + (srcSize > 31) + (srcSize > 4095);
    if (fl == 1) dst[0] = (u8)(SET_BASIC + (srcSize << 3));
    else if (fl == 2) wr16(dst, SET_BASIC + (1 << 2) + (srcSize << 4));
    else wr32(dst, SET_BASIC + (3 << 2) + (srcSize << 4));
    for (u32 i = ZE_LANE; i < srcSize; i += ZE_LANES) dst[fl + i] = src[i];
    ze_sync();
    return srcSize + fl;
}
ZE_FN_NOINLINE u32 rle_literals(u8* dst, const u8* src, u32 srcSize)
{
    u32 fl = 1 + (srcSize > 31) + (srcSize > 4095);
    if (fl == 1) dst[0] = (u8)(SET_RLE + (srcSize << 3));
    else if (fl == 2) wr16(dst, SET_RLE + (1 << 2) + (srcSize << 4));
    else wr32(dst, SET_RLE + (3 << 2) + (srcSize << 4));
    dst[fl] = src[0];
    return fl + 1;
}
// ZSTD_compressLiterals (:129)
ZE_FN_NOINLINE u32 compress_literals(Work& w, u8* dst, u32 dstCap, const u8* src, u32 srcSize, const Entropy& prev, Entropy& next, int suspectUncompressible)
{
    u32 strategy = w.cp.strategy;
    u32 lhSize = 3 + (srcSize >= 1024) + (srcSize >= 16384);
    u32 singleStream = srcSize < 256;
    int hType = SET_COMPRESSED;
    next.huf = prev.huf; next.hufRepeat = prev.hufRepeat;
    {   i32 shift = 9 - (i32)strategy; if (shift > 3) shift = 3;
        u32 mintc = (prev.hufRepeat == REP_VALID) ? 6 : (8u << shift);
        if (srcSize < mintc) return no_compress_literals(dst, src, srcSize); }
    i32 repeat = prev.hufRepeat;
    if (repeat == REP_VALID && lhSize == 3) singleStream = 1;
    u32 cLitSize = huf_compress(w, dst + lhSize, dstCap - lhSize, src, srcSize, singleStream, next.huf, &repeat,
                                strategy >= ST_BTULTRA, suspectUncompressible);
    if (repeat != REP_NONE) hType = SET_REPEAT;
    {   u32 mg = min_gain(srcSize, strategy);
        if (cLitSize == 0 || cLitSize >= srcSize - mg) { next.huf = prev.huf; next.hufRepeat = prev.hufRepeat; return no_compress_literals(dst, src, srcSize); } }
    if (cLitSize == 1) {
        bool same = true;
        if (srcSize < 8) for (u32 p = 1; p < srcSize; ++p) if (src[p] != src[0]) same = false;
        if (srcSize >= 8 || same) { next.huf = prev.huf; next.hufRepeat = prev.hufRepeat; return rle_literals(dst, src, srcSize); }
    }
    if (hType == SET_COMPRESSED) next.hufRepeat = REP_CHECK;
    if (lhSize == 3) wr24(dst, (u32)hType + ((u32)(!singleStream) << 2) + (srcSize << 4) + (cLitSize << 14));
    else if (lhSize == 4) wr32(dst, (u32)hType + (2 << 2) + (srcSize << 4) + (cLitSize << 18));
    else { wr32(dst, (u32)hType + (3 << 2) + (srcSize << 4) + (cLitSize << 22)); dst[4] = (u8)(cLitSize >> 10); }
    return lhSize + cLitSize;
}


// ---------------------------------------------------------------------------------------------------- sequences (zstd_compress_sequences.c)
ZE_FN_NOINLINE u32 entropy_cost(const u32* count, u32 max, u32 total)               // ZSTD_entropyCost (:85)
{
    u32 cost = 0;
    for (u32 s = 0; s <= max; ++s) {
        u32 norm = (u32)((256 * (u64)count[s]) / total);
        if (count[s] != 0 && norm == 0) norm = 1;
        cost += count[s] * kInvProbLog256[norm];
    }
    return cost >> 8;
}
ZE_FN_NOINLINE u64 fse_bit_cost_total(const FseCT& ct, const u32* count, u32 max)    // ZSTD_fseBitCost (:107) ; ~0 = error
{
    const u32 kAcc = 8; u64 cost = 0;
    if (ct.maxSym < max) return ~0ull;
    for (u32 s = 0; s <= max; ++s) {
        u32 tableLog = ct.tableLog, badCost = (tableLog + 1) << kAcc;
        u32 bc = fse_bit_cost(ct, s, kAcc);
        if (count[s] == 0) continue;
        if (bc >= badCost) return ~0ull;
        cost += (u64)count[s] * bc;
    }
    return cost >> kAcc;
}
ZE_FN_NOINLINE u64 cross_entropy_cost(const i16* norm, u32 accuracyLog, const u32* count, u32 max)   // ZSTD_crossEntropyCost (:141)
{
    u32 shift = 8 - accuracyLog; u64 cost = 0;
    for (u32 s = 0; s <= max; ++s) {
        u32 normAcc = (norm[s] != -1) ? (u32)norm[s] : 1, norm256 = normAcc << shift;
        cost += (u64)count[s] * kInvProbLog256[norm256];
    }
    return cost >> 8;
}
ZE_FN_NOINLINE u64 ncount_cost(Work& w, const u32* count, u32 max, u32 nbSeq, u32 FSELog)              // ZSTD_NCountCost (:72)
{
    i16 norm[53];
    u32 tableLog = fse_optimal_table_log(FSELog, nbSeq, max, 2);
    if (fse_normalize(norm, tableLog, count, nbSeq, max, nbSeq >= 2048) == ~0u) return ~0ull;
    u32 r = fse_write_ncount(w.scratch + 256, norm, max, tableLog);
    return r ? r : ~0ull;
}
// ZSTD_selectEncodingType (:157) for strategy >= lazy
ZE_FN_NOINLINE int select_encoding_type(Work& w, i32* repeatMode, const u32* count, u32 max, u32 mostFrequent, u32 nbSeq, u32 FSELog,
                               const FseCT& prevCT, const i16* defaultNorm, u32 defaultNormLog, int isDefaultAllowed)
{
    if (mostFrequent == nbSeq) {
        *repeatMode = REP_NONE;
        if (isDefaultAllowed && nbSeq <= 2) return SET_BASIC;
        return SET_RLE;
    }
    u64 basicCost = isDefaultAllowed ? cross_entropy_cost(defaultNorm, defaultNormLog, count, max) : ~0ull;
    u64 repeatCost = *repeatMode != REP_NONE ? fse_bit_cost_total(prevCT, count, max) : ~0ull;
    u64 NCountCost = ncount_cost(w, count, max, nbSeq, FSELog);
    u64 compressedCost = (NCountCost << 3) + entropy_cost(count, max, nbSeq);
    // (error codes of the reference are huge size_t values: (size_t)-N; ~0ull keeps every comparison identical)
    if (basicCost <= repeatCost && basicCost <= compressedCost) { *repeatMode = REP_NONE; return SET_BASIC; }
    if (repeatCost <= compressedCost) return SET_REPEAT;
    *repeatMode = REP_CHECK;
    return SET_COMPRESSED;
}
// ZSTD_buildCTable (:243).  returns bytes written to op (NCount header / rle symbol); ~0u on error
ZE_FN_NOINLINE u32 build_ctable(Work& w, u8* op, FseCT& next, u32 FSELog, int type, u32* count, u32 max, const u8* codeTable, u32 nbSeq,
                       const i16* defaultNorm, u32 defaultNormLog, u32 defaultMax, const FseCT& prev)
{
    switch (type) {
    case SET_RLE: fse_build_rle(next, (u8)max); *op = codeTable[0]; return 1;
    case SET_REPEAT: next = prev; return 0;
    case SET_BASIC: fse_build_ctable(next, defaultNorm, defaultMax, defaultNormLog, w.scratch); return 0;
    default: {
        i16 norm[53];
        u32 nbSeq_1 = nbSeq;
        u32 tableLog = fse_optimal_table_log(FSELog, nbSeq, max, 2);
        if (count[codeTable[nbSeq - 1]] > 1) { count[codeTable[nbSeq - 1]]--; nbSeq_1--; }
        if (fse_normalize(norm, tableLog, count, nbSeq_1, max, nbSeq_1 >= 2048) == ~0u) return ~0u;
        u32 nc = fse_write_ncount(op, norm, max, tableLog);
        if (!nc) return ~0u;
        fse_build_ctable(next, norm, max, tableLog, w.scratch);
        return nc; }
    }
}
// ZSTD_seqToCodes (zstd_compress.c:2679)
ZE_FN_NOINLINE void seq_to_codes(const SeqStore& ss)
{
    u32 nbSeq = (u32)(ss.seq - ss.seqStart);
    for (u32 u = ZE_LANE; u < nbSeq; u += ZE_LANES) {
        ss.llCode[u] = (u8)LLcode(ss.seqStart[u].litLength);
        ss.ofCode[u] = (u8)highbit(ss.seqStart[u].offBase);
        ss.mlCode[u] = (u8)MLcode(ss.seqStart[u].mlBase);
    }
    ze_sync();
    if (ss.longType == 1) ss.llCode[ss.longPos] = MaxLL;
    if (ss.longType == 2) ss.mlCode[ss.longPos] = MaxML;
}
struct SeqStats { int LLtype, Offtype, MLtype; u32 size, lastCountSize; int err; };
// ZSTD_buildSequencesStatistics (zstd_compress.c:2749)
ZE_FN_NOINLINE SeqStats build_seq_stats(Work& w, const SeqStore& ss, u32 nbSeq, const Entropy& prev, Entropy& next, u8* dst)
{
    SeqStats st; st.lastCountSize = 0; st.err = 0; st.size = 0;
    u8* op = dst; u32* count = w.count;
    seq_to_codes(ss);
    {   u32 max = MaxLL; u32 mf = hist_w(w, count, &max, ss.llCode, nbSeq);
        next.llRep = prev.llRep;
        st.LLtype = select_encoding_type(w, &next.llRep, count, max, mf, nbSeq, 9, prev.ll, kLLnorm, 6, 1);
        u32 cs = build_ctable(w, op, next.ll, 9, st.LLtype, count, max, ss.llCode, nbSeq, kLLnorm, 6, MaxLL, prev.ll);
        if (cs == ~0u) { st.err = 1; return st; }
        if (st.LLtype == SET_COMPRESSED) st.lastCountSize = cs;
        op += cs; }
    {   u32 max = MaxOff; u32 mf = hist_w(w, count, &max, ss.ofCode, nbSeq);
        int defaultAllowed = max <= DefaultMaxOff;
        next.ofRep = prev.ofRep;
        st.Offtype = select_encoding_type(w, &next.ofRep, count, max, mf, nbSeq, 8, prev.of, kOFnorm, 5, defaultAllowed);
        u32 cs = build_ctable(w, op, next.of, 8, st.Offtype, count, max, ss.ofCode, nbSeq, kOFnorm, 5, DefaultMaxOff, prev.of);
        if (cs == ~0u) { st.err = 1; return st; }
        if (st.Offtype == SET_COMPRESSED) st.lastCountSize = cs;
        op += cs; }
    {   u32 max = MaxML; u32 mf = hist_w(w, count, &max, ss.mlCode, nbSeq);
        next.mlRep = prev.mlRep;
        st.MLtype = select_encoding_type(w, &next.mlRep, count, max, mf, nbSeq, 9, prev.ml, kMLnorm, 6, 1);
        u32 cs = build_ctable(w, op, next.ml, 9, st.MLtype, count, max, ss.mlCode, nbSeq, kMLnorm, 6, MaxML, prev.ml);
        if (cs == ~0u) { st.err = 1; return st; }
        if (st.MLtype == SET_COMPRESSED) st.lastCountSize = cs;
        op += cs; }
    st.size = (u32)(op - dst);
    return st;
}
// ZSTD_encodeSequences_body (zstd_compress_sequences.c:292), 64-bit, no long offsets.  returns 0 if dst too small
ZE_FN_NOINLINE u32 encode_sequences(u8* dst, u64 cap, const Entropy& e, const SeqStore& ss, u32 nbSeq)
{
    BitW b; FseState sML, sOF, sLL;
    if (!bit_init(b, dst, cap)) return 0;
    const Seq* sq = ss.seqStart;
    fse_init2(sML, e.ml, ss.mlCode[nbSeq - 1]);
    fse_init2(sOF, e.of, ss.ofCode[nbSeq - 1]);
    fse_init2(sLL, e.ll, ss.llCode[nbSeq - 1]);
    bit_add(b, sq[nbSeq - 1].litLength, kLLbits[ss.llCode[nbSeq - 1]]);
    bit_add(b, sq[nbSeq - 1].mlBase, kMLbits[ss.mlCode[nbSeq - 1]]);
    bit_add(b, sq[nbSeq - 1].offBase, ss.ofCode[nbSeq - 1]);
    bit_flush(b);
    for (u32 n = nbSeq - 2; n < nbSeq; n--) {
        u32 llCode = ss.llCode[n], ofCode = ss.ofCode[n], mlCode = ss.mlCode[n];
        u32 llBits = kLLbits[llCode], ofBits = ofCode, mlBits = kMLbits[mlCode];
        fse_encode(b, sOF, ofCode);
        fse_encode(b, sML, mlCode);
        fse_encode(b, sLL, llCode);
        if (ofBits + mlBits + llBits >= 64 - 7 - (9 + 9 + 8)) bit_flush(b);
        bit_add(b, sq[n].litLength, llBits);
        bit_add(b, sq[n].mlBase, mlBits);
        if (ofBits + mlBits + llBits > 56) bit_flush(b);
        bit_add(b, sq[n].offBase, ofBits);
        bit_flush(b);
    }
    fse_flush_state(b, sML); fse_flush_state(b, sOF); fse_flush_state(b, sLL);
    return bit_close(b);
}

// ZSTD_entropyCompressSeqStore (zstd_compress.c:2874-3027). returns compressed block body size, 0 = "not compressed"
ZE_FN_NOINLINE u32 entropy_compress_seqstore(Work& w, const SeqStore& ss, const Entropy& prev, Entropy& next, u8* dst, u64 dstCap, u32 srcSize)
{
    u8* op = dst;
    u32 nbSeq = (u32)(ss.seq - ss.seqStart);
    {   u32 numLiterals = (u32)(ss.lit - ss.litStart);
        int suspect = (nbSeq == 0) || (numLiterals / nbSeq >= 20);
        op += compress_literals(w, op, (u32)(dstCap > 0xffffffffu ? 0xffffffffu : dstCap), ss.litStart, numLiterals, prev, next, suspect); }
    if (nbSeq < 128) *op++ = (u8)nbSeq;
    else if (nbSeq < 0x7F00) { op[0] = (u8)((nbSeq >> 8) + 0x80); op[1] = (u8)nbSeq; op += 2; }
    else { op[0] = 0xFF; wr16(op + 1, nbSeq - 0x7F00); op += 3; }
    u32 cSize;
    if (nbSeq == 0) {
        next.ll = prev.ll; next.of = prev.of; next.ml = prev.ml; next.llRep = prev.llRep; next.ofRep = prev.ofRep; next.mlRep = prev.mlRep;
        cSize = (u32)(op - dst);
    } else {
        u8* seqHead = op++;
        SeqStats st = build_seq_stats(w, ss, nbSeq, prev, next, op);
        if (st.err) { w.error = 2; return 0; }
        *seqHead = (u8)((st.LLtype << 6) + (st.Offtype << 4) + (st.MLtype << 2));
        op += st.size;
        u32 bs = encode_sequences(op, dstCap - (u64)(op - dst), next, ss, nbSeq);
        if (bs == 0) return 0;                         // dstSize_tooSmall with srcSize <= dstCapacity => "not compressed"
        op += bs;
        if (st.lastCountSize && (st.lastCountSize + bs) < 4) return 0;
        cSize = (u32)(op - dst);
    }
    {   u32 maxCSize = srcSize - min_gain(srcSize, w.cp.strategy);
        if (cSize >= maxCSize) return 0; }
    return cSize;
}

// ---------------------------------------------------------------------------------------------------- block-split estimation (zstd_compress.c:3522-3860)
// ZSTD_buildBlockEntropyStats_literals: returns desSize, sets hType
ZE_FN_NOINLINE u32 block_stats_literals(Work& w, const u8* src, u32 srcSize, const Entropy& prev, Entropy& next, HufMeta& hm, int optimalDepth)
{
    u32 maxSym = 255, huffLog = 11;
    i32 repeat = prev.hufRepeat;
    next.huf = prev.huf; next.hufRepeat = prev.hufRepeat;
    {   u32 minLit = (prev.hufRepeat == REP_VALID) ? 6 : 63;
        if (srcSize <= minLit) { hm.hType = SET_BASIC; return 0; } }
    {   u32 largest = hist_w(w, w.count, &maxSym, src, srcSize);
        if (largest == srcSize) { hm.hType = SET_RLE; return 0; }
        if (largest <= (srcSize >> 7) + 4) { hm.hType = SET_BASIC; return 0; } }
    if (repeat == REP_CHECK && !huf_validate(prev.huf, w.count, maxSym)) repeat = REP_NONE;
    for (u32 i = 0; i < 256; ++i) { next.huf.nbBits[i] = 0; next.huf.val[i] = 0; }
    huffLog = huf_optimal_table_log(w, huffLog, srcSize, maxSym, next.huf, w.count, optimalDepth);
    huffLog = huf_build_ctable(w, next.huf, w.count, maxSym, huffLog);
    {   u32 newCSize = huf_estimate_size(next.huf, w.count, maxSym);
        u32 hSize = huf_write_ctable(w, hm.des, 128, next.huf, maxSym, huffLog);
        if (repeat != REP_NONE) {
            u32 oldCSize = huf_estimate_size(prev.huf, w.count, maxSym);
            if (oldCSize < srcSize && (oldCSize <= hSize + newCSize || hSize + 12 >= srcSize)) {
                next.huf = prev.huf; next.hufRepeat = prev.hufRepeat; hm.hType = SET_REPEAT; return 0; }
        }
        if (newCSize + hSize >= srcSize) { next.huf = prev.huf; next.hufRepeat = prev.hufRepeat; hm.hType = SET_BASIC; return 0; }
        hm.hType = SET_COMPRESSED; next.hufRepeat = REP_CHECK;
        return hSize;
    }
}
ZE_FN_NOINLINE u32 estimate_literal(Work& w, const u8* lits, u32 litSize, const HufCT& huf, const HufMeta& hm, int writeEntropy)
{
    u32 maxSym = 255, hdr = 3 + (litSize >= 1024) + (litSize >= 16384), single = litSize < 256;
    if (hm.hType == SET_BASIC) return litSize;
    if (hm.hType == SET_RLE) return 1;
    hist_w(w, w.count, &maxSym, lits, litSize);
    u32 est = huf_estimate_size(huf, w.count, maxSym);
    if (writeEntropy) est += hm.desSize;
    if (!single) est += 6;
    return est + hdr;
}
ZE_FN_NOINLINE u64 estimate_symbol_type(Work& w, int type, const u8* codeTable, u32 nbSeq, u32 maxCode, const FseCT& ct, const u8* addBits,
                               const i16* defaultNorm, u32 defaultNormLog)
{
    u32 max = maxCode; u64 bits = 0;
    hist_w(w, w.count, &max, codeTable, nbSeq);
    if (type == SET_BASIC) bits = cross_entropy_cost(defaultNorm, defaultNormLog, w.count, max);
    else if (type == SET_RLE) bits = 0;
    else bits = fse_bit_cost_total(ct, w.count, max);
    if (bits == ~0ull) return (u64)nbSeq * 10;
    {   u32 part = 0;                                         // nbSeq <= 128 K/3 sequences x <= 31 bits: fits 32 bits per lane
        for (u32 i = ZE_LANE; i < nbSeq; i += ZE_LANES) part += addBits ? addBits[codeTable[i]] : codeTable[i];
        bits += ze_reduce_add(part); }
    return bits >> 3;
}
// ZSTD_buildEntropyStatisticsAndEstimateSubBlockSize (:3845). ~0 = error
ZE_FN_NOINLINE u64 estimate_subblock(Work& w, const SeqStore& ss)
{
    const Entropy& prev = w.bs[w.prevIdx]->e; Entropy& next = w.bs[w.prevIdx ^ 1]->e;
    HufMeta& hm = *w.hufMeta; FseMeta& fm = *w.fseMeta;
    u32 litSize = (u32)(ss.lit - ss.litStart), nbSeq = (u32)(ss.seq - ss.seqStart);
    hm.desSize = block_stats_literals(w, ss.litStart, litSize, prev, next, hm, w.cp.strategy >= ST_BTULTRA);
    if (nbSeq) {
        SeqStats st = build_seq_stats(w, ss, nbSeq, prev, next, fm.buf);
        if (st.err) return ~0ull;
        fm.llType = st.LLtype; fm.ofType = st.Offtype; fm.mlType = st.MLtype; fm.tablesSize = st.size; fm.lastCountSize = st.lastCountSize;
    } else {
        fm.llType = fm.ofType = fm.mlType = SET_BASIC; fm.tablesSize = 0; fm.lastCountSize = 0;
        next.llRep = next.ofRep = next.mlRep = REP_NONE;
    }
    u64 literalsSize = estimate_literal(w, ss.litStart, litSize, next.huf, hm, hm.hType == SET_COMPRESSED);
    u64 seqSize = 1 + 1 + (nbSeq >= 128) + (nbSeq >= 0x7F00);
    seqSize += estimate_symbol_type(w, fm.ofType, ss.ofCode, nbSeq, MaxOff, next.of, nullptr, kOFnorm, 5);
    seqSize += estimate_symbol_type(w, fm.llType, ss.llCode, nbSeq, MaxLL, next.ll, kLLbits, kLLnorm, 6);
    seqSize += estimate_symbol_type(w, fm.mlType, ss.mlCode, nbSeq, MaxML, next.ml, kMLbits, kMLnorm, 6);
    seqSize += fm.tablesSize;
    return seqSize + literalsSize + 3;
}
ZE_FN_NOINLINE u32 count_lit_bytes(const SeqStore& ss)          // ZSTD_countSeqStoreLiteralsBytes (:3862)
{
    u32 n = (u32)(ss.seq - ss.seqStart), b = 0;
    for (u32 i = 0; i < n; ++i) { b += ss.seqStart[i].litLength; if (i == ss.longPos && ss.longType == 1) b += 0x10000; }
    return b;
}
ZE_FN_NOINLINE u32 count_match_bytes(const SeqStore& ss)        // ZSTD_countSeqStoreMatchBytes (:3877)
{
    u32 n = (u32)(ss.seq - ss.seqStart), b = 0;
    for (u32 i = 0; i < n; ++i) { b += ss.seqStart[i].mlBase + 3; if (i == ss.longPos && ss.longType == 2) b += 0x10000; }
    return b;
}
// ZSTD_deriveSeqStoreChunk (:3894)
ZE_FN_NOINLINE void derive_chunk(SeqStore& r, const SeqStore& o, u32 startIdx, u32 endIdx)
{
    r = o;
    if (startIdx > 0) { r.seq = o.seqStart + startIdx; r.litStart += count_lit_bytes(r); }
    if (o.longType != 0) {
        if (o.longPos < startIdx || o.longPos > endIdx) r.longType = 0; else r.longPos -= startIdx;
    }
    r.seqStart = o.seqStart + startIdx;
    r.seq = o.seqStart + endIdx;
    if (endIdx != (u32)(o.seq - o.seqStart)) { u32 lb = count_lit_bytes(r); r.lit = r.litStart + lb; }
    r.llCode += startIdx; r.mlCode += startIdx; r.ofCode += startIdx;
}
// ZSTD_deriveBlockSplitsHelper (:4092) with its recursion on an explicit stack (in-order traversal keeps split order)
ZE_FN_NOINLINE u32 derive_block_splits(Work& w, u32* partitions, u32 nbSeq)
{
    if (nbSeq <= 4) return 0;
    u32 nsplits = 0;
    struct Fr { u32 a, b; int stage; };
    Fr stk[24]; int sp = 0;
    stk[sp].a = 0; stk[sp].b = nbSeq; stk[sp].stage = 0; sp++;
    while (sp) {
        Fr& f = stk[sp - 1];
        u32 mid = (f.a + f.b) / 2;
        if (f.stage == 0) {
            if (f.b - f.a < 300 || nsplits >= 196) { sp--; continue; }
            SeqStore full, h1, h2;
            derive_chunk(full, w.ss, f.a, f.b); derive_chunk(h1, w.ss, f.a, mid); derive_chunk(h2, w.ss, mid, f.b);
            u64 eo = estimate_subblock(w, full), e1 = estimate_subblock(w, h1), e2 = estimate_subblock(w, h2);
            if (eo == ~0ull || e1 == ~0ull || e2 == ~0ull) { sp--; continue; }
            if (e1 + e2 < eo) { f.stage = 1; stk[sp].a = f.a; stk[sp].b = mid; stk[sp].stage = 0; sp++; }
            else sp--;
        } else if (f.stage == 1) {
            partitions[nsplits++] = mid;
            f.stage = 2; u32 b = f.b; stk[sp].a = mid; stk[sp].b = b; stk[sp].stage = 0; sp++;
        } else sp--;
    }
    partitions[nsplits] = nbSeq;
    return nsplits;
}


// ---------------------------------------------------------------------------------------------------- blocks and frame (zstd_compress.c)
ZE_FN void reset_seqstore(Work& w) { w.ss.seq = w.ss.seqStart; w.ss.lit = w.ss.litStart; w.ss.longType = 0; w.ss.longPos = 0; }

// ZSTD_buildSeqStore (:3200) + the block compressors of zstd_opt.c:1441-1541.  returns false for ZSTDbss_noCompress
ZE_FN_NOINLINE bool build_seqstore(Work& w, const u8* src, u32 srcSize)
{
    if (srcSize < 2 + 3 + 1 + 1) return false;
    reset_seqstore(w);
    {   u32 curr = (u32)(src - w.src) + w.baseOff;
        if (curr > w.nextToUpdate + 384) { u32 d = curr - w.nextToUpdate - 384; w.nextToUpdate = curr - (d < 192 ? d : 192); } }
    BlockState& prev = *w.bs[w.prevIdx]; BlockState& next = *w.bs[w.prevIdx ^ 1];
    next.rep[0] = prev.rep[0]; next.rep[1] = prev.rep[1]; next.rep[2] = prev.rep[2];
    u32 lastLL;
    if (w.cp.strategy == ST_BTLAZY2) lastLL = compress_block_btlazy2(w, next.rep, src, srcSize);
    else if (w.cp.strategy == ST_BTOPT) lastLL = compress_block_opt(w, next.rep, src, srcSize, 0);
    else if (w.cp.strategy == ST_BTULTRA) lastLL = compress_block_opt(w, next.rep, src, srcSize, 2);
    else {
        // ZSTD_compressBlock_btultra2 (zstd_opt.c:1513): first block is parsed twice, the first pass only seeds the statistics
        u32 curr = (u32)(src - w.src) + w.baseOff;
        if (w.llSum == 0 && w.ss.seq == w.ss.seqStart && w.dictLimit == w.lowLimit && curr == w.dictLimit && srcSize > 8) {
            u32 tmpRep[3] = { next.rep[0], next.rep[1], next.rep[2] };
            compress_block_opt(w, tmpRep, src, srcSize, 2);
            reset_seqstore(w);
            w.baseOff += srcSize; w.dictLimit += srcSize; w.lowLimit = w.dictLimit; w.nextToUpdate = w.dictLimit;
        }
        lastLL = compress_block_opt(w, next.rep, src, srcSize, 2);
    }
    {   const u8* ll = src + srcSize - lastLL;                    // ZSTD_storeLastLiterals
        for (u32 i = ZE_LANE; i < lastLL; i += ZE_LANES) w.ss.lit[i] = ll[i];
        ze_sync();
        w.ss.lit += lastLL; }
    return true;
}
ZE_FN bool is_rle(const u8* src, u32 n) { for (u32 i = 1; i < n; ++i) if (src[i] != src[0]) return false; return true; }   // ZSTD_isRLE (:3469)
ZE_FN_NOINLINE u32 no_compress_block(u8* dst, const u8* src, u32 srcSize, u32 last)
{
    wr24(dst, last + (0u << 1) + (srcSize << 3));
    for (u32 i = ZE_LANE; i < srcSize; i += ZE_LANES) dst[3 + i] = src[i];
    ze_sync();
    return 3 + srcSize;
}
ZE_FN_NOINLINE u32 rle_compress_block(u8* dst, u8 b, u32 srcSize, u32 last) { wr24(dst, last + (1u << 1) + (srcSize << 3)); dst[3] = b; return 4; }
ZE_FN void confirm(Work& w) { w.prevIdx ^= 1; }              // ZSTD_blockState_confirmRepcodesAndEntropyTables

// ZSTD_resolveRepcodeToRawOffset (:3927) / ZSTD_seqStore_resolveOffCodes (:3959)
ZE_FN u32 resolve_rep_raw(const u32* rep, u32 offBase, u32 ll0)
{
    u32 adj = offBase - 1 + ll0;
    if (adj == 3) return rep[0] - 1;
    return rep[adj];
}
ZE_FN_NOINLINE void resolve_off_codes(u32* dRep, u32* cRep, const SeqStore& ss, u32 nbSeq)
{
    u32 longLit = ss.longType == 1 ? ss.longPos : nbSeq;
    for (u32 idx = 0; idx < nbSeq; ++idx) {
        Seq* sq = ss.seqStart + idx;
        u32 ll0 = (sq->litLength == 0) && (idx != longLit);
        u32 offBase = sq->offBase;
        if (offBase >= 1 && offBase <= 3) {
            u32 dRaw = resolve_rep_raw(dRep, offBase, ll0), cRaw = resolve_rep_raw(cRep, offBase, ll0);
            if (dRaw != cRaw) sq->offBase = cRaw + 3;
        }
        update_rep(dRep, sq->offBase, ll0);
        update_rep(cRep, offBase, ll0);
    }
}
// ZSTD_compressSeqStore_singleBlock (:4002)
ZE_FN_NOINLINE u32 compress_seqstore_single(Work& w, const SeqStore& ss, u32* dRep, u32* cRep, u8* dst, u64 dstCap, const u8* src, u32 srcSize, u32 last, int isPartition)
{
    u32 dRepOrig[3] = { dRep[0], dRep[1], dRep[2] };
    if (isPartition) resolve_off_codes(dRep, cRep, ss, (u32)(ss.seq - ss.seqStart));
    u32 cSeqs = entropy_compress_seqstore(w, ss, w.bs[w.prevIdx]->e, w.bs[w.prevIdx ^ 1]->e, dst + 3, dstCap - 3, srcSize);
    if (!w.isFirstBlock && cSeqs < 25 && is_rle(src, srcSize)) cSeqs = 1;
    u32 cSize;
    if (cSeqs == 0) { cSize = no_compress_block(dst, src, srcSize, last); dRep[0] = dRepOrig[0]; dRep[1] = dRepOrig[1]; dRep[2] = dRepOrig[2]; }
    else if (cSeqs == 1) { cSize = rle_compress_block(dst, src[0], srcSize, last); dRep[0] = dRepOrig[0]; dRep[1] = dRepOrig[1]; dRep[2] = dRepOrig[2]; }
    else { confirm(w); wr24(dst, last + (2u << 1) + (cSeqs << 3)); cSize = 3 + cSeqs; }
    Entropy& pe = w.bs[w.prevIdx]->e;
    if (pe.ofRep == REP_VALID) pe.ofRep = REP_CHECK;
    return cSize;
}
// ZSTD_compressBlock_splitBlock (:4250) + _internal (:4165)
ZE_FN_NOINLINE u32 compress_block_split(Work& w, u8* dst, u64 dstCap, const u8* src, u32 blockSize, u32 lastBlock)
{
    if (!build_seqstore(w, src, blockSize)) {
        Entropy& pe = w.bs[w.prevIdx]->e;
        if (pe.ofRep == REP_VALID) pe.ofRep = REP_CHECK;
        return no_compress_block(dst, src, blockSize, lastBlock);
    }
    u32 nbSeq = (u32)(w.ss.seq - w.ss.seqStart);
    u32* partitions = w.partitions;
    u32 numSplits = derive_block_splits(w, partitions, nbSeq);
    u32 dRep[3], cRep[3];
    for (int i = 0; i < 3; ++i) dRep[i] = cRep[i] = w.bs[w.prevIdx]->rep[i];
    if (numSplits == 0) return compress_seqstore_single(w, w.ss, dRep, cRep, dst, dstCap, src, blockSize, lastBlock, 0);
    u32 cSize = 0, srcBytesTotal = 0;
    const u8* ip = src; u8* op = dst;
    SeqStore curr, nextS; nextS = w.ss;
    derive_chunk(curr, w.ss, 0, partitions[0]);
    for (u32 i = 0; i <= numSplits; ++i) {
        u32 lastPartition = (i == numSplits), lastBlockEntireSrc = 0;
        u32 srcBytes = count_lit_bytes(curr) + count_match_bytes(curr);
        srcBytesTotal += srcBytes;
        if (lastPartition) { srcBytes += blockSize - srcBytesTotal; lastBlockEntireSrc = lastBlock; }
        else derive_chunk(nextS, w.ss, partitions[i], partitions[i + 1]);
        u32 cs = compress_seqstore_single(w, curr, dRep, cRep, op, dstCap, ip, srcBytes, lastBlockEntireSrc, 1);
        ip += srcBytes; op += cs; dstCap -= cs; cSize += cs;
        curr = nextS;
    }
    for (int i = 0; i < 3; ++i) w.bs[w.prevIdx]->rep[i] = dRep[i];
    return cSize;
}
// ZSTD_compressBlock_internal (:4277) + the header logic of ZSTD_compress_frameChunk (:4496-4512)
ZE_FN_NOINLINE u32 compress_block_plain(Work& w, u8* dst, u64 dstCap, const u8* src, u32 blockSize, u32 lastBlock)
{
    u32 cSize = 0;
    if (build_seqstore(w, src, blockSize)) {
        cSize = entropy_compress_seqstore(w, w.ss, w.bs[w.prevIdx]->e, w.bs[w.prevIdx ^ 1]->e, dst + 3, dstCap - 3, blockSize);
        if (!w.isFirstBlock && cSize < 25 && is_rle(src, blockSize)) { cSize = 1; dst[3] = src[0]; }
    }
    if (cSize > 1) confirm(w);
    Entropy& pe = w.bs[w.prevIdx]->e;
    if (pe.ofRep == REP_VALID) pe.ofRep = REP_CHECK;
    if (cSize == 0) return no_compress_block(dst, src, blockSize, lastBlock);
    u32 hdr = cSize == 1 ? lastBlock + (1u << 1) + (blockSize << 3) : lastBlock + (2u << 1) + (cSize << 3);
    wr24(dst, hdr);
    return cSize + 3;
}

// workspace carving -----------------------------------------------------------------------------------------------
ZE_FN u64 align_up(u64 x, u64 a) { return (x + a - 1) / a * a; }
struct WorkSizes { u64 hash, chain, hash3, opt, matches, freqs, seqs, lits, codes, bstates, misc, total; };
ZE_FN WorkSizes work_sizes(const Params& cp)
{
    WorkSizes z;
    u32 maxNbSeq = cp.blockSize / (cp.minMatch == 3 ? 3 : 4);
    u32 hashLog3 = cp.minMatch == 3 ? (cp.windowLog < 17 ? cp.windowLog : 17) : 0;
    z.hash = align_up((u64)4 << cp.hashLog, 256);
    z.chain = align_up((u64)4 << cp.chainLog, 256);
    z.hash3 = hashLog3 ? align_up((u64)4 << hashLog3, 256) : 256;
    z.opt = align_up((u64)sizeof(Opt) * OPT_SIZE, 256);
    z.matches = align_up((u64)sizeof(Match) * OPT_SIZE, 256);
    z.freqs = align_up(4 * (256 + 36 + 53 + 32 + 16), 256);
    z.seqs = align_up((u64)sizeof(Seq) * (maxNbSeq + 2), 256);
    z.lits = align_up((u64)cp.blockSize + 64, 256);
    z.codes = align_up((u64)3 * (maxNbSeq + 2), 256);
    z.bstates = align_up(2 * sizeof(BlockState), 256);
    z.misc = align_up(4 * 256 + sizeof(HufNode) * 520 + 4 * 192 * 2 + 1024 + sizeof(FseCT) + sizeof(HufCT) + sizeof(FseMeta) + sizeof(HufMeta) + 4 * 200 + 256 + 16 + 4 * 384, 256)
           + align_up((u64)sizeof(Win), 256);
    z.total = z.hash + z.chain + z.hash3 + z.opt + z.matches + z.freqs + z.seqs + z.lits + z.codes + z.bstates + z.misc;
    return z;
}
// low-latency scratch layout: match-finder window, match list (<= 3 repcodes + 1 hash3 + 2^searchLog tree matches), frequency tables
struct FastSizes { u32 win, matches, freqs, prices, hist, total; };
ZE_FN FastSizes fast_sizes()
{
    FastSizes f;
    f.win = (u32)align_up((u64)sizeof(Win), 16);                  // first: the helper threads of the CTA find it at offset 0
    f.matches = (u32)align_up((u64)sizeof(Match) * (256 + 8), 16);
    f.freqs = (u32)align_up(4 * (256 + 36 + 53 + 32 + 16), 16);
    f.prices = 4 * 384;
    f.hist = 4 * 256;
    f.total = f.win + f.matches + f.freqs + f.prices + f.hist;
    return f;
}
ZE_FN u64 compress_bound(u64 n) { return n + (n >> 8) + (n < (128u << 10) ? (((128u << 10) - n) >> 11) : 0); }    // ZSTD_COMPRESSBOUND

// ZSTD_compressCCtx (:5317) for one input.  `mem` = zero-initialised workspace of work_sizes(cp).total bytes.
// returns the frame size, 0 on failure (w.error says why)
// `fast` / `fastBytes`: optional low-latency scratch (shared memory on the device) for the parser's hot state, see fast_sizes()
ZE_FN_NOINLINE u64 compress_frame(const u8* src, u64 srcSize64, int level, u8* dst, u64 dstCap, u8* mem, int* err, u64* prof_out = nullptr,
                                  u8* fast = nullptr, u32 fastBytes = 0)
{
    Work w; *err = 0;
    w.histTab = nullptr;
#if defined(ZE_PROF) && defined(__CUDA_ARCH__)
    for (int i = 0; i < ZE_PROF_N; ++i) w.prof[i] = 0;
    long long t_frame = clock64();
#endif
    Params cp = get_params(level, srcSize64);
    if (!cp.supported) { *err = 1; return 0; }
    u32 srcSize = (u32)srcSize64;
    w.src = src; w.srcSize = srcSize; w.cp = cp; w.error = 0;
    WorkSizes z = work_sizes(cp);
    u8* p = mem;
    w.hashTable = (u32*)p; p += z.hash; w.chainTable = (u32*)p; p += z.chain; w.hashTable3 = (u32*)p; p += z.hash3;
    w.opt = (Opt*)p; p += z.opt; w.matches = (Match*)p; p += z.matches;
    w.litFreq = (u32*)p; w.llFreq = w.litFreq + 256; w.mlFreq = w.llFreq + 36; w.ofFreq = w.mlFreq + 53; p += z.freqs;
    if (fast && fastBytes >= fast_sizes().total) {               // parser tables in the low-latency scratch
        FastSizes fz = fast_sizes(); u8* f = fast;
        w.win = (Win*)f; f += fz.win;
        w.matches = (Match*)f; f += fz.matches;
        w.litFreq = (u32*)f; w.llFreq = w.litFreq + 256; w.mlFreq = w.llFreq + 36; w.ofFreq = w.mlFreq + 53; f += fz.freqs;
        w.priceTab = (u32*)f; f += fz.prices;
        w.histTab = (u32*)f;
    }
    w.ss.seqStart = (Seq*)p; p += z.seqs; w.ss.litStart = p; p += z.lits;
    w.maxNbSeq = cp.blockSize / (cp.minMatch == 3 ? 3 : 4);
    w.ss.llCode = p; w.ss.mlCode = p + (w.maxNbSeq + 2); w.ss.ofCode = p + 2 * (w.maxNbSeq + 2); p += z.codes;
    w.bs[0] = (BlockState*)p; w.bs[1] = w.bs[0] + 1; p += z.bstates;
    w.count = (u32*)p; p += 4 * 256;
    w.huffNode = (HufNode*)p; p += sizeof(HufNode) * 520;
    w.rankPos = (u32*)p; p += 4 * 192 * 2;
    w.scratch = p; p += 1024;
    w.tmpCT = (FseCT*)p; p += sizeof(FseCT);
    w.tmpHuf = (HufCT*)p; p += sizeof(HufCT);
    w.fseMeta = (FseMeta*)p; p += sizeof(FseMeta);
    p = (u8*)align_up((u64)p, 8);
    w.hufMeta = (HufMeta*)p; p += sizeof(HufMeta);
    p = (u8*)align_up((u64)p, 8);
    w.partitions = (u32*)p; p += 4 * 200;
    w.dummySlot = (u32*)p; p += 256;
    if (!(fast && fastBytes >= fast_sizes().total)) w.priceTab = (u32*)p;
    p += 4 * 384;
    p = (u8*)align_up((u64)p, 8);
    if (!(fast && fastBytes >= fast_sizes().total)) w.win = (Win*)p;
    if (ZE_LANE == 0) { w.win->count = 0; w.win->next = 0; w.win->baseOff = 0; w.win->pending = 0; w.win->built = 0; for (int i = 0; i < 8; ++i) w.win->phase[i] = 0; }
    ze_sync();
    reset_seqstore(w);
    w.hashLog3 = cp.minMatch == 3 ? (cp.windowLog < 17 ? cp.windowLog : 17) : 0;
    w.baseOff = 2; w.lowLimit = 2; w.dictLimit = 2; w.nextToUpdate = 2;
    w.litSum = w.llSum = w.mlSum = w.ofSum = 0; w.pricePredef = 0;
    w.prevIdx = 0; w.isFirstBlock = 1;
    {   BlockState& b = *w.bs[0];                             // ZSTD_reset_compressedBlockState
        b.rep[0] = 1; b.rep[1] = 4; b.rep[2] = 8;
        b.e.hufRepeat = REP_NONE; b.e.llRep = b.e.ofRep = b.e.mlRep = REP_NONE; }

    // frame header (ZSTD_writeFrameHeader :4530): magic, FHD (single segment, FCS code), content size
    u8* op = dst;
    wr32(op, 0xFD2FB528u); op += 4;
    {   u32 fcs = (srcSize >= 256) + (srcSize >= 65536 + 256);
        *op++ = (u8)((1u << 5) + (fcs << 6));
        if (fcs == 0) *op++ = (u8)srcSize;
        else if (fcs == 1) { wr16(op, srcSize - 256); op += 2; }
        else { wr32(op, srcSize); op += 4; } }
    // ZSTD_compress_frameChunk (:4448)
    u32 remaining = srcSize, blockSize = cp.blockSize;
    const u8* ip = src;
    u32 maxDist = 1u << cp.windowLog;
    while (remaining) {
        u32 lastBlock = blockSize >= remaining;
        if (remaining < blockSize) blockSize = remaining;
        {   u32 blockEndIdx = (u32)(ip - src) + w.baseOff;          // ZSTD_window_enforceMaxDist(window, ip, ...)
            if (blockEndIdx > maxDist) {
                u32 nl = blockEndIdx - maxDist;
                if (w.lowLimit < nl) w.lowLimit = nl;
                if (w.dictLimit < w.lowLimit) w.dictLimit = w.lowLimit;
            } }
        if (w.nextToUpdate < w.lowLimit) w.nextToUpdate = w.lowLimit;
        u64 cap = dstCap - (u64)(op - dst);
        u32 cs = cp.splitter ? compress_block_split(w, op, cap, ip, blockSize, lastBlock) : compress_block_plain(w, op, cap, ip, blockSize, lastBlock);
        if (w.error) { *err = w.error; return 0; }
        ip += blockSize; remaining -= blockSize; op += cs;
        w.isFirstBlock = 0;
    }
    if (srcSize == 0) { wr24(op, 1); op += 3; }                     // ZSTD_writeEpilogue: empty last raw block
#if defined(ZE_PROF) && defined(__CUDA_ARCH__)
    w.prof[0] = (u64)(clock64() - t_frame);
    for (int i = 0; i < 6; ++i) w.prof[19 + i] = w.win->phase[i];
    if (prof_out && ZE_LANE == 0) for (int i = 0; i < ZE_PROF_N; ++i) prof_out[i] = w.prof[i];
#endif
    return (u64)(op - dst);
}

}  // namespace ZE_NS
