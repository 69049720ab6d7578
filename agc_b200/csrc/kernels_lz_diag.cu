// kernels_lz_diag.cu -- CLZDiff_V2::Encode (src/common/lz_diff.cpp:669-798): one warp per segment, streaming along the current
// diagonal (lz_diag_core.cuh has the algorithm and the argument why the bytes are the sequential parse's).
//
// A CTA (LZD_THREADS / 32 warps) works on requests of ONE group: the group's 2-bit packed reference and its hash table are
// staged in shared memory by two TMA bulk copies (cp.async.bulk + mbarrier), every warp owns an LzdScratch (window bitmap,
// mismatch list, per-state results) behind them and pulls requests from the unit's counter.  The text is read from the packed
// contig store in HBM exactly once per window (coalesced: consecutive lanes read consecutive 8-byte words); the reference never
// leaves shared memory.  Algorithmic bytes per segment: ceil(n/4) + ceil(m/4) + e.
#include "internal.cuh"
#include "lz_chunk.cuh"
#include "lz_diag_core.cuh"


__device__ __forceinline__ uint32_t lzd_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// HT_STAGED: the hash table is staged as well (32 warps, one CTA per SM); otherwise only the reference is (16 warps, up to two
// CTAs per SM; the few index probes per mismatch go to L2) -- the launch for groups whose table does not fit
template <int LZD_THREADS, bool HT_STAGED>
__global__ void __launch_bounds__(LZD_THREADS, HT_STAGED ? 1 : 2) k_lz_diag(
    const uint64_t* __restrict__ P, const GroupRefDev* __restrict__ groups, const LzReqDev* __restrict__ reqs,
    const LzUnit* __restrict__ units, uint32_t mml, uint32_t stage_bytes, uint8_t* __restrict__ slab,
    uint32_t* __restrict__ res, uint32_t* __restrict__ err)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t next_req;
    const LzUnit u = units[blockIdx.x];
    const GroupRefDev g = groups[u.group];
    const uint32_t ht_bytes = g.ht_size * ((g.flags & GRF_SHORT) ? 2u : 4u);
    const uint32_t st_bytes = g.packed_bytes + (HT_STAGED ? ht_bytes : 0u);
    const bool stage = (st_bytes <= stage_bytes);
    if (threadIdx.x == 0) {
        next_req = 0;
        if (stage) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(lzd_smem_u32(&bar)), "r"(1));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
    }
    __syncthreads();
    if (stage) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(lzd_smem_u32(&bar)), "r"(st_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         :: "r"(lzd_smem_u32(smem)), "l"(g.packed), "r"(g.packed_bytes), "r"(lzd_smem_u32(&bar)) : "memory");
            if (HT_STAGED)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(lzd_smem_u32(smem + g.packed_bytes)), "l"(g.ht), "r"(ht_bytes), "r"(lzd_smem_u32(&bar)) : "memory");
        }
        if (threadIdx.x == 0)                    // one thread polls the barrier; the others wait for it at the CTA barrier below
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "WAIT_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra DONE_%=;\n\t"
                "bra WAIT_%=;\n\t"
                "DONE_%=:\n\t}" :: "r"(lzd_smem_u32(&bar)), "r"(0) : "memory");
        __syncthreads();
    }
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    LzdScratch& S = *reinterpret_cast<LzdScratch*>(smem + ((stage_bytes + 127u) & ~127u) + (size_t)warp * ((sizeof(LzdScratch) + 15u) & ~15u));
    for (;;) {
        uint32_t r = 0;
        if (lane == 0) r = atomicAdd(&next_req, 1u);
        r = __shfl_sync(0xffffffffu, r, 0);
        if (r >= u.count) break;
        const LzReqDev q = reqs[u.first + r];
        if (lane == 0) {                                 // pull the whole packed segment into L2 while the parse works on its head
            const uint64_t b0 = (q.gstart >> 2) & ~(uint64_t)15;
            const uint32_t nb = (uint32_t)(((q.gstart + q.n + 3) >> 2) - b0 + 15) & ~15u;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"((const uint8_t*)P + b0), "r"(nb) : "memory");
        }
        int64_t len;
        if (stage) {
            LzcView<true, HT_STAGED> a; a.T = P; a.gs = (int64_t)q.gstart; a.n = q.n; a.rc = q.is_rc; a.R = nullptr; a.r_s = lzd_smem_u32(smem);
            a.ht = g.ht; a.ht_s = lzd_smem_u32(smem + g.packed_bytes); a.mask = g.ht_size - 1; a.is_short = g.flags & GRF_SHORT; a.m = g.m;
            len = lzd_encode_segment(a, S, mml, slab + q.out_off, q.out_cap);
        } else {
            LzcView<false> a; a.T = P; a.gs = (int64_t)q.gstart; a.n = q.n; a.rc = q.is_rc; a.R = (const uint64_t*)g.packed; a.r_s = 0;
            a.ht = g.ht; a.ht_s = 0; a.mask = g.ht_size - 1; a.is_short = g.flags & GRF_SHORT; a.m = g.m;
            len = lzd_encode_segment(a, S, mml, slab + q.out_off, q.out_cap);
        }
        if (lane == 0) {
            if (len < 0) { atomicOr(err, 1u); res[q.orig] = 0; }
            else res[q.orig] = (uint32_t)len;
        }
        __syncwarp();
    }
}

// stage_bytes: bytes of shared memory set aside for the staged reference (+ table when ht_staged); 0: groups are read from global memory
int agc_lzd_launch(agcgpu_ctx* ctx, const LzReqDev* d_reqs, const LzUnit* d_units, uint32_t n_units, size_t stage_bytes, int ht_staged,
                   uint8_t* slab, uint32_t* res, uint32_t* err)
{
    const uint32_t warps = ht_staged ? 32u : 16u;
    const size_t smem = ((stage_bytes + 127u) & ~(size_t)127u) + (size_t)warps * ((sizeof(LzdScratch) + 15u) & ~15u);
    if (ht_staged) {
        CK(cudaFuncSetAttribute(k_lz_diag<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_lz_diag<1024, true><<<n_units, 1024, smem, ctx->st>>>((const uint64_t*)ctx->packed.p, (const GroupRefDev*)ctx->d_groups.p, d_reqs, d_units,
                                                                ctx->prm.min_match_len, (uint32_t)stage_bytes, slab, res, err);
    } else {
        CK(cudaFuncSetAttribute(k_lz_diag<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_lz_diag<512, false><<<n_units, 512, smem, ctx->st>>>((const uint64_t*)ctx->packed.p, (const GroupRefDev*)ctx->d_groups.p, d_reqs, d_units,
                                                               ctx->prm.min_match_len, (uint32_t)stage_bytes, slab, res, err);
    }
    CKL();
    return 0;
}
size_t agc_lzd_scratch_bytes(int ht_staged) { return (size_t)(ht_staged ? 32 : 16) * ((sizeof(LzdScratch) + 15u) & ~15u); }
