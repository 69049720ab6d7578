// kernels_lz_chunk.cu -- CLZDiff_V2::Encode (src/common/lz_diff.cpp:669-798) as a chunk-parallel parse (sm_100a).
//
// The greedy parse of one segment is sequential, but its state between two tokens is tiny -- (i, pred_pos, no_prev_literals) --
// and after every match it is (i, pred_pos, 0).  A segment is therefore cut into chunks of LZC_CHUNK text positions and ONE THREAD
// parses one chunk speculatively from its first position with no_prev_literals = 0:
//   * k_lzc_parse : thread per (segment, chunk).  A CTA works on chunks of ONE group; the group's 2-bit packed reference and its
//     hash table are staged in shared memory by two TMA bulk copies (as in kernels_lz.cu).  The text is read straight from the
//     packed contig store (8-byte loads; a thread walks forward through its chunk, so its sectors stay in L1).  Match extension is
//     a 64-bit XOR + clz per 32 bases.  A single-candidate match that reaches the end of the chunk is left OPEN (its extension is
//     the next chunk's business), so every chunk costs O(chunk) whatever the match lengths are.  Each chunk writes the bytes of its
//     tokens after its first match (their pred_pos chain is self-contained) and a small record describing its first match, the
//     literals before it and its final state.
//   * k_lzc_stitch : thread per segment.  Walks the chunk records in order carrying the TRUE state, emits every chunk's first
//     match (the only token whose bytes depend on the incoming state: dif_pos, the '!' rewrite of lz_diff.cpp:769-779, a backward
//     extension that crosses the chunk boundary, the merge of an open match with its continuation) and copies the rest.
//     Whatever the records cannot prove identical to the sequential parse (a failed candidate whose outcome depends on the number
//     of previous literals, several candidates for a first match, a continuation on another diagonal) marks the segment for the
//     sequential kernel (k_lz_packed<0>), which then produces it; the result is always the reference's byte string.
//
// Algorithmic bytes per segment: ceil(n/4) + ceil(m/4) + e.
#include "internal.cuh"
#include "lz_chunk_core.cuh"
#include <cstdlib>

__device__ __forceinline__ uint32_t lzc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void lzc_mbar_init(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(lzc_smem_u32(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void lzc_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
:: "r"(lzc_smem_u32(dst)), "l"(src), "r"(bytes), "r"(lzc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void lzc_mbar_wait(uint64_t* bar, uint32_t phase)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" :: "r"(lzc_smem_u32(bar)), "r"(phase) : "memory");
}

// hands the chunks of a unit to the lanes: item -> (request, chunk) through the running chunk count the requests carry (unit_base)
template <bool COSTS>
struct DevFetch {
    const LzcReq* reqs; LzcUnit u; uint32_t* next; uint8_t* cslab; LzcRec* recs; uint32_t* costv;
    template <bool STAGED>
    __device__ __forceinline__ bool operator()(LzcView<STAGED>& a, LzcItem& it)
    {
        const uint32_t k = atomicAdd(next, 1u);
        if (k >= u.n_items) return false;
        uint32_t lo = 0, hi = u.count - 1;
        while (lo < hi) { const uint32_t mid = (lo + hi + 1) >> 1; if (reqs[u.first + mid].unit_base <= k + u.item0) lo = mid; else hi = mid - 1; }
        const LzcReq q = reqs[u.first + lo];
        const uint32_t ch = k + u.item0 - q.unit_base;
        a.gs = (int64_t)q.gstart; a.n = q.n; a.rc = q.is_rc;
        it.c0 = ch * q.chunk; it.c1 = lzc_min(q.n, it.c0 + q.chunk);
        it.out = cslab + (uint64_t)(q.chunk_first + ch) * LZC_CSLAB;
        it.rec = recs + q.chunk_first + ch;
        it.v = COSTS ? costv + q.out_off : nullptr; it.prefix = q.out_cap;
        return true;
    }
};

// COSTS: cost vectors instead of deltas (GetCodingCostVector): request q's vector starts at costv + q.out_off, q.out_cap = prefix_costs
template <bool COSTS>
__global__ void __launch_bounds__(LZC_THREADS, 2) k_lzc_parse(
    const uint64_t* __restrict__ P, const GroupRefDev* __restrict__ groups, const LzcReq* __restrict__ reqs,
    const LzcUnit* __restrict__ units, uint32_t mml, uint32_t stage_limit, uint8_t* __restrict__ cslab, LzcRec* __restrict__ recs,
    uint32_t* __restrict__ costv)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t s_next;
    if (threadIdx.x == 0) s_next = 0;
    const LzcUnit u = units[blockIdx.x];
    const GroupRefDev g = groups[u.group];
    const uint32_t ht_bytes = g.ht_size * ((g.flags & GRF_SHORT) ? 2u : 4u);
    const bool stage = (g.packed_bytes + ht_bytes <= stage_limit);
    if (stage) {
        if (threadIdx.x == 0) lzc_mbar_init(&bar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(lzc_smem_u32(&bar)), "r"(g.packed_bytes + ht_bytes) : "memory");
            lzc_bulk_g2s(smem, g.packed, g.packed_bytes, &bar);
            lzc_bulk_g2s(smem + g.packed_bytes, g.ht, ht_bytes, &bar);
        }
        lzc_mbar_wait(&bar, 0);
    }
    // every lane pulls chunks from the unit's counter until none is left (lzc_parse_stream)
    __syncthreads();
    DevFetch<COSTS> fetch; fetch.reqs = reqs; fetch.u = u; fetch.next = &s_next; fetch.cslab = cslab; fetch.recs = recs; fetch.costv = costv;
    if (stage) {
        LzcView<true> a; a.T = P; a.gs = 0; a.n = 0; a.rc = 0; a.R = nullptr; a.r_s = lzc_smem_u32(smem);
        a.ht = nullptr; a.ht_s = lzc_smem_u32(smem + g.packed_bytes); a.mask = g.ht_size - 1; a.is_short = g.flags & GRF_SHORT; a.m = g.m;
        lzc_parse_stream<true, COSTS>(a, mml, fetch);
    } else {
        LzcView<false> a; a.T = P; a.gs = 0; a.n = 0; a.rc = 0; a.R = (const uint64_t*)g.packed; a.r_s = 0;
        a.ht = g.ht; a.ht_s = 0; a.mask = g.ht_size - 1; a.is_short = g.flags & GRF_SHORT; a.m = g.m;
        lzc_parse_stream<false, COSTS>(a, mml, fetch);
    }
}

// ------------------------------------------------------------------------------------------------ phase 2: one thread, one segment
template <bool COSTS>
__global__ void __launch_bounds__(128) k_lzc_stitch(
    const uint64_t* __restrict__ P, const GroupRefDev* __restrict__ groups, const LzcReq* __restrict__ reqs, uint32_t n_req,
    uint32_t mml, const uint8_t* __restrict__ cslab, const LzcRec* __restrict__ recs, uint8_t* __restrict__ slab,
    uint32_t* __restrict__ res, uint32_t* __restrict__ fb, uint32_t* __restrict__ counters, uint32_t* __restrict__ costv, uint32_t per_warp)
{
    // per_warp: one segment per WARP (lane 0 works).  The stitcher is a serial, latency-bound walk; 32 of them in one warp run
    // their divergent paths one after the other, so small batches (one sample of a collection) spread over warps instead.
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (per_warp && (threadIdx.x & 31u)) return;
    const uint32_t r = per_warp ? t >> 5 : t;
    if (r >= n_req) return;
    const LzcReq q = reqs[r];
    const GroupRefDev g = groups[q.group];
    LzcView<false> a; a.T = P; a.gs = (int64_t)q.gstart; a.n = q.n; a.rc = q.is_rc; a.R = (const uint64_t*)g.packed; a.r_s = 0;
    a.ht = g.ht; a.ht_s = 0; a.mask = g.ht_size - 1; a.is_short = g.flags & GRF_SHORT; a.m = g.m;
    const int64_t o = COSTS ? lzc_stitch_segment<LzcView<false>, true>(a, q, mml, recs + q.chunk_first, nullptr, nullptr, 0, costv + q.out_off, q.out_cap)
                            : lzc_stitch_segment<LzcView<false>, false>(a, q, mml, recs + q.chunk_first, cslab + (uint64_t)q.chunk_first * LZC_CSLAB, slab + q.out_off, q.out_cap);
    if (o <= -10) { fb[r] = 1; atomicAdd(counters + 0, 1u); res[q.orig] = 0; return; }
    fb[r] = 0;
    if (o == -2) { atomicOr(counters + 1, 1u); res[q.orig] = 0; return; }
    res[q.orig] = (uint32_t)o;
}

// ------------------------------------------------------------------------------------------------ host side
int agc_lzc_launch(agcgpu_ctx* ctx, const LzcReq* d_reqs, uint32_t n_req, const LzcUnit* d_units, uint32_t n_units, size_t smem,
                   uint8_t* cslab, LzcRec* recs, uint8_t* slab, uint32_t* res, uint32_t* fb, uint32_t* counters, uint32_t* costv)
{
    static bool attr_set = false;
    if (!attr_set) {
        cudaFuncSetAttribute(k_lzc_parse<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LZC_STAGE_LIMIT);
        cudaFuncSetAttribute(k_lzc_parse<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LZC_STAGE_LIMIT);
        attr_set = true;
    }
    static const bool no_stage = getenv("AGCGPU_LZC_NOSTAGE") != nullptr;        // diagnostics: read reference and table from global memory
    const GroupRefDev* groups = (const GroupRefDev*)ctx->d_groups.p;
    const uint64_t* P = (const uint64_t*)ctx->packed.p;
    const uint32_t mml = ctx->prm.min_match_len, sl = no_stage ? 0u : (uint32_t)smem;
    if (costv) k_lzc_parse<true><<<n_units, LZC_THREADS, smem, ctx->st>>>(P, groups, d_reqs, d_units, mml, sl, cslab, recs, costv);
    else k_lzc_parse<false><<<n_units, LZC_THREADS, smem, ctx->st>>>(P, groups, d_reqs, d_units, mml, sl, cslab, recs, nullptr);
    CKL();
    const uint32_t per_warp = n_req <= 32u * (uint32_t)ctx->n_sm ? 1u : 0u;
    const uint32_t n_thr = per_warp ? n_req * 32u : n_req;
    if (costv) k_lzc_stitch<true><<<(n_thr + 127) / 128, 128, 0, ctx->st>>>(P, groups, d_reqs, n_req, mml, cslab, recs, slab, res, fb, counters, costv, per_warp);
    else k_lzc_stitch<false><<<(n_thr + 127) / 128, 128, 0, ctx->st>>>(P, groups, d_reqs, n_req, mml, cslab, recs, slab, res, fb, counters, nullptr, per_warp);
    CKL();
    return 0;
}
