// kernels_zstd.cu -- residual coder on the device: agcgpu_zstd_compress_batch = ZSTD_compressCCtx(level) for a batch of
// independent inputs.  The frame producer itself is zstd_enc.cuh (bit-identical to the reference's
// vendored libzstd; see the header for the function-by-function citations).
//
// Mapping: one CTA per input; inputs are independent so the batch is the parallel dimension (HPP scale: ~10^5 frames per
// batch).  The optimal parser is a sequential dynamic program -- one warp runs it with warp-uniform control flow -- over a
// binary-tree match finder, whose dependent tree walks are taken out of the parser's critical path by the window engine of
// zstd_enc.cuh: the wide kernel (inputs > 32 KB) gives the engine 15 more warps and 221 KB of shared memory, the narrow one
// (one warp, 19 KB) keeps ~1600 small frames resident.  zstd's tables live in a per-input workspace in HBM (L2 resident).
// Largest inputs are scheduled first.
#include "internal.cuh"
#define ZE_NS ze                 // wide coder: 512-slot match-finder window, parser warp + 15 warps of tree walks
#define ZE_WN_W 512
#include "zstd_enc.cuh"
#undef ZE_NS
#undef ZE_WN_W
#define ZE_NS zen                // narrow coder: 32-slot window, one warp and ~20 KB of shared memory per frame
#define ZE_WN_W 32
#include "zstd_enc.cuh"
#undef ZE_NS
#undef ZE_WN_W
#include <algorithm>
#include <numeric>
#include <cstdio>
#include <cstdlib>

struct ZTaskDev {
    const uint8_t* src; uint8_t* dst; uint8_t* mem;
    uint64_t n, dst_cap;
    int32_t level; int32_t err;
    uint64_t out_size;
    uint64_t t_start, t_end;          // %globaltimer at CTA start / end (trace output only)
#ifdef ZE_PROF
    uint64_t prof[32];
#endif
};

static const uint32_t ZS_THREADS = 512;
__device__ __forceinline__ uint64_t zs_globaltimer() { uint64_t t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__global__ void __launch_bounds__(ZS_THREADS, 1) k_zstd(ZTaskDev* __restrict__ tasks, uint32_t n_tasks, uint32_t smem_bytes)
{
    uint32_t t = blockIdx.x;
    if (t >= n_tasks) return;
    extern __shared__ __align__(16) uint8_t zs_smem[];
    // warp 0 runs the coder: its 32 lanes carry identical scalar state (see zstd_enc.cuh) and split array-wide steps; the other
    // warps serve the match finder's window jobs (one tree walk per thread) between two named barriers
    if (threadIdx.x >= 32) {
        ze::Win& W = *reinterpret_cast<ze::Win*>(zs_smem);
        for (;;) {
            ze::ze_bar_sync(1, ZS_THREADS);                      // a job was posted
            const uint32_t job = *reinterpret_cast<volatile uint32_t*>(&W.job);
            if (job == ze::WJ_EXIT) break;
            ze::win_run(W, job);
            ze::ze_bar_arrive(2, ZS_THREADS);                    // done (the parser waits on barrier 2 when it needs the result)
        }
        return;
    }
    ZTaskDev k = tasks[t];
    int err = 0;
    if (threadIdx.x == 0) tasks[t].t_start = zs_globaltimer();
#ifdef ZE_PROF
    uint64_t r = ze::compress_frame(k.src, k.n, k.level, k.dst, k.dst_cap, k.mem, &err, tasks[t].prof, zs_smem, smem_bytes);
#else
    uint64_t r = ze::compress_frame(k.src, k.n, k.level, k.dst, k.dst_cap, k.mem, &err, nullptr, zs_smem, smem_bytes);
#endif
    if (threadIdx.x == 0) { tasks[t].err = err; tasks[t].out_size = r; tasks[t].t_end = zs_globaltimer(); reinterpret_cast<ze::Win*>(zs_smem)->job = ze::WJ_EXIT; }
    __syncwarp();
    ze::ze_bar_arrive(1, ZS_THREADS);
}

// small inputs: one warp per frame, many frames per SM
__global__ void __launch_bounds__(32) k_zstd_narrow(ZTaskDev* __restrict__ tasks, uint32_t n_tasks, uint32_t smem_bytes)
{
    uint32_t t = blockIdx.x;
    if (t >= n_tasks) return;
    extern __shared__ __align__(16) uint8_t zs_smem[];
    ZTaskDev k = tasks[t];
    int err = 0;
    if (threadIdx.x == 0) tasks[t].t_start = zs_globaltimer();
#ifdef ZE_PROF
    uint64_t r = zen::compress_frame(k.src, k.n, k.level, k.dst, k.dst_cap, k.mem, &err, tasks[t].prof, zs_smem, smem_bytes);
#else
    uint64_t r = zen::compress_frame(k.src, k.n, k.level, k.dst, k.dst_cap, k.mem, &err, nullptr, zs_smem, smem_bytes);
#endif
    if (threadIdx.x == 0) { tasks[t].err = err; tasks[t].out_size = r; tasks[t].t_end = zs_globaltimer(); }
}
// inputs up to this size take the narrow coder (AGCGPU_ZSTD_NARROW_MAX overrides it: diagnostics)
static uint64_t zs_narrow_max()
{
    static const uint64_t v = getenv("AGCGPU_ZSTD_NARROW_MAX") ? strtoull(getenv("AGCGPU_ZSTD_NARROW_MAX"), nullptr, 10) : (32u << 10);
    return v;
}

// keep_lead == ~0: frames go to the host (dst / dst_offsets).  Otherwise (sharded coder) they stay on the device, back to back in input
// order in ctx->scr_zkeep behind keep_lead bytes the caller fills in; dst_offsets still receives their offsets.
static int zstd_compress_batch_impl(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* src_offsets, const int32_t* levels,
                                    uint32_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* dst_offsets, uint64_t keep_lead)
{
    const bool keep = keep_lead != ~0ull;
    dst_offsets[0] = 0;
    if (n == 0) return 0;
    // per-input parameters, workspace and output sizes
    std::vector<uint64_t> ws(n), ob(n);
    uint64_t total_src = src_offsets[n];
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t len = src_offsets[i + 1] - src_offsets[i];
        ze::Params cp = ze::get_params(levels[i], len);
        if (!cp.supported)
            return agc_fail(ctx, AGCGPU_EUNSUPPORTED, "zstd: input %u (level %d, %llu bytes) is outside the implemented envelope "
                            "(levels 13/17/18/19, inputs below 1 GiB)", i, levels[i], (unsigned long long)len);
        ws[i] = (ze::work_sizes(cp).total + 255) / 256 * 256;
        ob[i] = (ze::compress_bound(len) + 64 + 255) / 256 * 256;
    }
    // Which coder.  Inputs above 32 KB take the wide one (a whole SM per frame: 15 warps of tree walks feed the parser) as long as
    // every such frame gets an SM of its own.  When there are more of them than SMs, the frames of LZ-diff delta text
    // ("0,85.C0,33.A...": the match-finder windows are cut after a few dozen positions, the helper warps idle, and the one-warp
    // coder is as fast per frame -- 4.9 vs 5.8 us/B on C3's packs) go to the narrow coder, 11 of which fit on an SM, instead of
    // waiting for a second round of SMs (C3: 223 frames, wave 2.74 s -> 2.20 s).  Packs of raw sequences (1 byte per base; C2's
    // 1.35 MB raw-group pack: 0.71 us/B wide) always stay wide.  Both coders produce the same bytes; this is scheduling only.
    std::vector<uint8_t> is_wide(n, 0), is_text(n, 0);
    uint32_t n_big = 0;
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t len = src_offsets[i + 1] - src_offsets[i];
        if (len <= zs_narrow_max()) continue;
        ++n_big;
        const uint8_t* p = src + src_offsets[i];
        uint32_t sym = 0; const uint32_t probe = 64;
        for (uint32_t k = 0; k < probe; ++k) { const uint8_t c = p[(uint64_t)k * (len / probe)]; sym += c < 32u || c == 0xffu; }
        is_text[i] = sym * 2 <= probe;               // mostly printable: digits, ',', '.', letters
        is_wide[i] = 1;
    }
    if (n_big > (uint32_t)ctx->n_sm && !getenv("AGCGPU_ZSTD_WIDE_ALL"))
        for (uint32_t i = 0; i < n; ++i) if (is_text[i]) is_wide[i] = 0;
    // schedule: wide frames first, biggest inputs first; waves bounded by a workspace budget
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        if (is_wide[a] != is_wide[b]) return is_wide[a] > is_wide[b];
        return src_offsets[a + 1] - src_offsets[a] > src_offsets[b + 1] - src_offsets[b]; });
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t budget = std::max<uint64_t>((uint64_t)(free_b * 0.6), 512ull << 20);
    if (int r = agc_reserve(ctx, ctx->scr_bytes, total_src + 64)) return r;
    CK(cudaMemcpyAsync(ctx->scr_bytes.p, src, total_src, cudaMemcpyHostToDevice, ctx->st));
    ctx->stats.h2d_bytes += total_src;
    ctx->stats.zstd_input_mb += (float)(total_src * 1e-6);
    std::vector<uint64_t> out_size(n, 0);
    std::vector<std::vector<uint8_t>> frames(n);
    std::vector<const uint8_t*> keep_src(n, nullptr);
    size_t pos = 0;
    while (pos < n) {
        size_t end = pos; uint64_t wsum = 0, osum = 0;
        while (end < n && (end == pos || wsum + ws[order[end]] + osum + ob[order[end]] <= budget)) { wsum += ws[order[end]]; osum += ob[order[end]]; ++end; }
        if (ws[order[pos]] + ob[order[pos]] > (uint64_t)free_b)
            return agc_fail(ctx, AGCGPU_ENOMEM, "zstd: not enough device memory for one %llu-byte workspace", (unsigned long long)ws[order[pos]]);
        uint32_t cnt = (uint32_t)(end - pos);
        if (int r = agc_reserve(ctx, ctx->scr_out, wsum + 256)) return r;
        if (int r = agc_reserve(ctx, ctx->scr_dense, osum + 256)) return r;
        if (int r = agc_reserve(ctx, ctx->scr_req, cnt * sizeof(ZTaskDev))) return r;
        CK(cudaMemsetAsync(ctx->scr_out.p, 0, wsum, ctx->st));
        std::vector<ZTaskDev> tasks(cnt);
        uint64_t wo = 0, oo = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
            uint32_t i = order[pos + j];
            ZTaskDev& k = tasks[j];
            k.src = (const uint8_t*)ctx->scr_bytes.p + src_offsets[i]; k.n = src_offsets[i + 1] - src_offsets[i];
            k.dst = (uint8_t*)ctx->scr_dense.p + oo; k.dst_cap = ob[i]; k.mem = (uint8_t*)ctx->scr_out.p + wo;
            k.level = levels[i]; k.err = 0; k.out_size = 0; k.t_start = k.t_end = 0;
            wo += ws[i]; oo += ob[i];
        }
        CK(cudaMemcpyAsync(ctx->scr_req.p, tasks.data(), cnt * sizeof(ZTaskDev), cudaMemcpyHostToDevice, ctx->st));
        CK(cudaDeviceSetLimit(cudaLimitStackSize, 16384));
        CK(cudaEventRecord(ctx->ev0, ctx->st));
        // inputs are sorted by size: the first n_wide take the wide coder, the rest the narrow one on a second stream so that
        // the two kernels share the device
        uint32_t n_wide = 0;
        while (n_wide < cnt && is_wide[order[pos + n_wide]]) ++n_wide;
        if (n_wide) {
            const uint32_t smem_bytes = ze::fast_sizes().total;
            CK(cudaFuncSetAttribute(k_zstd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
            k_zstd<<<n_wide, ZS_THREADS, smem_bytes, ctx->st>>>((ZTaskDev*)ctx->scr_req.p, n_wide, smem_bytes);
            CKL();
        }
        if (cnt > n_wide) {
            if (!ctx->st2) { CK(cudaStreamCreateWithFlags(&ctx->st2, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&ctx->ev2, cudaEventDisableTiming)); }
            const uint32_t smem_n = zen::fast_sizes().total;
            CK(cudaStreamWaitEvent(ctx->st2, ctx->ev0, 0));
            k_zstd_narrow<<<cnt - n_wide, 32, smem_n, ctx->st2>>>((ZTaskDev*)ctx->scr_req.p + n_wide, cnt - n_wide, smem_n);
            CKL();
            CK(cudaEventRecord(ctx->ev2, ctx->st2));
            CK(cudaStreamWaitEvent(ctx->st, ctx->ev2, 0));
        }
        CK(cudaEventRecord(ctx->ev1, ctx->st));
        CK(cudaMemcpyAsync(tasks.data(), ctx->scr_req.p, cnt * sizeof(ZTaskDev), cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        {   float ms = 0; cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1);
            ctx->stats.zstd_kernel_ms += ms;
            if (getenv("AGCGPU_TRACE")) {
                uint64_t tot = 0; for (uint32_t j = 0; j < cnt; ++j) tot += tasks[j].n;
                fprintf(stderr, "[agcgpu] zstd wave: %u inputs, %llu bytes, largest %llu (level %d), kernel %.1f ms\n", cnt,
                        (unsigned long long)tot, (unsigned long long)tasks[0].n, tasks[0].level, ms);
                uint64_t t0 = ~0ull, tw = 0, tn = 0, endw = 0, endn = 0;
                for (uint32_t j = 0; j < cnt; ++j) t0 = std::min<uint64_t>(t0, tasks[j].t_start);
                for (uint32_t j = 0; j < cnt; ++j) { uint64_t d = tasks[j].t_end - tasks[j].t_start; if (j < n_wide) { tw += d; endw = std::max(endw, tasks[j].t_end - t0); } else { tn += d; endn = std::max(endn, tasks[j].t_end - t0); } }
                fprintf(stderr, "[agcgpu]   wide: %u frames, sum of frame times %.1f ms, last ends at %.1f ms | narrow: %u frames, sum %.1f ms, last ends at %.1f ms\n",
                        n_wide, tw * 1e-6, endw * 1e-6, cnt - n_wide, tn * 1e-6, endn * 1e-6);
                for (uint32_t j = 0; j < cnt; j += (j < 8 ? 1 : std::max<uint32_t>(1, cnt / 24)))
                    fprintf(stderr, "[agcgpu]   frame %u: %llu B L%d start %.1f ms dur %.1f ms (%.2f us/B)\n", j, (unsigned long long)tasks[j].n, tasks[j].level,
                            (tasks[j].t_start - t0) * 1e-6, (tasks[j].t_end - tasks[j].t_start) * 1e-6, (tasks[j].t_end - tasks[j].t_start) * 1e-3 / std::max<uint64_t>(1, tasks[j].n));
#ifdef ZE_PROF
                for (uint32_t j = 0; j < cnt && j < 4; ++j) {
                    const uint64_t* p = tasks[j].prof; const double us = 1.0 / 1965.0;     // ticks at the max SM clock
                    fprintf(stderr, "[agcgpu]  frame %u (%llu B, L%d): total %.0f us | parse %.0f  matches %.0f (update_tree %.0f, window build %.0f, commit %.0f, seq query %.0f, replay %.0f)\n"
                                    "[agcgpu]    counts: get_all_matches %llu, windows %llu, commits %llu, seq inserts %llu, seq queries %llu, replayed queries %llu, window inserts %llu, cuts: unusable slot %llu, skipped positions %llu, re-resolves %llu (%.0f us)\n[agcgpu]    build phases (thread 0): walk %.0f +wait %.0f, pairs %.0f +wait %.0f, resolve %.0f +wait %.0f us\n",
                            j, (unsigned long long)tasks[j].n, tasks[j].level, p[0] * us, p[1] * us, p[2] * us, p[3] * us, p[4] * us, p[5] * us, p[6] * us, p[7] * us,
                            (unsigned long long)p[12], (unsigned long long)p[8], (unsigned long long)p[14], (unsigned long long)p[9], (unsigned long long)p[10], (unsigned long long)p[11],
                            (unsigned long long)p[13], (unsigned long long)p[15], (unsigned long long)p[16], (unsigned long long)p[17], p[18] * us, p[19] * us, p[20] * us, p[21] * us, p[22] * us, p[23] * us, p[24] * us);
                }
#endif
            } }
        // frames back: one copy when the batch is small or mostly incompressible, one copy per frame otherwise
        uint64_t out_total = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
            uint32_t i = order[pos + j];
            if (tasks[j].err) return agc_fail(ctx, AGCGPU_EUNSUPPORTED, "zstd: device coder failed on input %u (code %d)", i, tasks[j].err);
            out_size[i] = tasks[j].out_size; out_total += out_size[i];
        }
        if (keep) {
            // sharded coder: note where each frame lies; they are compacted on the device once every wave is done
            for (uint32_t j = 0; j < cnt; ++j) { uint32_t i = order[pos + j]; keep_src[i] = tasks[j].dst; }
            if (end < n) {     // more waves will reuse scr_dense: park this wave's frames
                for (uint32_t j = 0; j < cnt; ++j) {
                    uint32_t i = order[pos + j];
                    frames[i].resize(out_size[i]);
                    CK(cudaMemcpyAsync(frames[i].data(), tasks[j].dst, out_size[i], cudaMemcpyDeviceToHost, ctx->st));
                    keep_src[i] = nullptr;
                }
                CK(cudaStreamSynchronize(ctx->st));
            }
        } else if (osum <= (64ull << 20) || out_total * 2 >= osum) {
            std::vector<uint8_t> host(osum);
            CK(cudaMemcpyAsync(host.data(), ctx->scr_dense.p, osum, cudaMemcpyDeviceToHost, ctx->st));
            CK(cudaStreamSynchronize(ctx->st));
            ctx->stats.d2h_bytes += osum;
            for (uint32_t j = 0; j < cnt; ++j) {
                uint32_t i = order[pos + j];
                const uint8_t* p = host.data() + (tasks[j].dst - (uint8_t*)ctx->scr_dense.p);
                frames[i].assign(p, p + out_size[i]);
            }
        } else {
            for (uint32_t j = 0; j < cnt; ++j) {
                uint32_t i = order[pos + j];
                frames[i].resize(out_size[i]);
                CK(cudaMemcpyAsync(frames[i].data(), tasks[j].dst, out_size[i], cudaMemcpyDeviceToHost, ctx->st));
                ctx->stats.d2h_bytes += out_size[i];
            }
            CK(cudaStreamSynchronize(ctx->st));
        }
        pos = end;
    }
    for (uint32_t i = 0; i < n; ++i) dst_offsets[i + 1] = dst_offsets[i] + out_size[i];
    if (keep) {
        if (int r = agc_reserve(ctx, ctx->scr_zkeep, keep_lead + dst_offsets[n] + 64)) return r;
        uint8_t* base = (uint8_t*)ctx->scr_zkeep.p + keep_lead;
        for (uint32_t i = 0; i < n; ++i) {
            if (!out_size[i]) continue;
            if (keep_src[i]) CK(cudaMemcpyAsync(base + dst_offsets[i], keep_src[i], out_size[i], cudaMemcpyDeviceToDevice, ctx->st));
            else CK(cudaMemcpyAsync(base + dst_offsets[i], frames[i].data(), out_size[i], cudaMemcpyHostToDevice, ctx->st));
        }
        CK(cudaStreamSynchronize(ctx->st));
        return 0;
    }
    if (dst_offsets[n] > dst_cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "zstd: need %llu output bytes", (unsigned long long)dst_offsets[n]);
    for (uint32_t i = 0; i < n; ++i) if (out_size[i]) memcpy(dst + dst_offsets[i], frames[i].data(), out_size[i]);
    return 0;
}

extern "C" int agcgpu_zstd_compress_batch(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* src_offsets, const int32_t* levels,
                                          uint32_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* dst_offsets)
{
    if (!ctx || !src_offsets || !dst_offsets || (n && (!src || !levels || !dst))) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    return zstd_compress_batch_impl(ctx, src, src_offsets, levels, n, dst, dst_cap, dst_offsets, ~0ull);
}

// The residual coder over the ranks of the NCCL communicator: frames are dealt out by size (largest first, each to the least
// loaded rank; every rank computes the same assignment), coded where they land, and all-gathered between device buffers:
// block of rank r = [u64 offsets of its frames (count+1)][frames back to back].
extern "C" int agcgpu_zstd_compress_batch_sharded(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* src_offsets, const int32_t* levels,
                                                  uint32_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* dst_offsets)
{
    if (!ctx || !src_offsets || !dst_offsets || (n && (!src || !levels || !dst))) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    if (!agc_comm_active() || n == 0) return zstd_compress_batch_impl(ctx, src, src_offsets, levels, n, dst, dst_cap, dst_offsets, ~0ull);
    const uint32_t W = agc_comm_world(), me = agc_comm_rank();
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return src_offsets[a + 1] - src_offsets[a] > src_offsets[b + 1] - src_offsets[b]; });
    std::vector<uint64_t> load(W, 0);
    std::vector<std::vector<uint32_t>> of_rank(W);
    for (uint32_t i : order) {
        uint32_t r = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
        of_rank[r].push_back(i); load[r] += src_offsets[i + 1] - src_offsets[i] + 512;      // + a per-frame constant: tiny frames are not free
    }
    const std::vector<uint32_t>& mine = of_rank[me];
    const uint32_t cnt = (uint32_t)mine.size();
    std::vector<uint64_t> so((size_t)cnt + 1, 0), fo((size_t)cnt + 1, 0);
    std::vector<int32_t> lv(cnt);
    for (uint32_t j = 0; j < cnt; ++j) { so[j + 1] = so[j] + (src_offsets[mine[j] + 1] - src_offsets[mine[j]]); lv[j] = levels[mine[j]]; }
    std::vector<uint8_t> sub(so[cnt] + 1);
    for (uint32_t j = 0; j < cnt; ++j) if (so[j + 1] > so[j]) memcpy(sub.data() + so[j], src + src_offsets[mine[j]], so[j + 1] - so[j]);
    const uint64_t hdr = ((uint64_t)(cnt + 1) * 8 + 15) / 16 * 16;
    int local = 0;                                       // a local failure still enters the collective (all ranks fail together)
    if (cnt) local = zstd_compress_batch_impl(ctx, sub.data(), so.data(), lv.data(), cnt, nullptr, 0, fo.data(), hdr);
    else local = agc_reserve(ctx, ctx->scr_zkeep, hdr + 64);
    if (!local && cudaMemcpyAsync(ctx->scr_zkeep.p, fo.data(), (size_t)(cnt + 1) * 8, cudaMemcpyHostToDevice, ctx->st) != cudaSuccess) local = AGCGPU_ECUDA;
    std::vector<uint64_t> sizes; uint64_t stride = 0;
    if (int r = agc_comm_allgatherv(ctx, ctx->scr_zkeep.p, hdr + fo[cnt], local, sizes, &stride)) return r;
    std::vector<uint8_t> host((size_t)stride * W);
    if (stride) CK(cudaMemcpyAsync(host.data(), ctx->scr_gather.p, (size_t)stride * W, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.d2h_bytes += (size_t)stride * W;
    std::vector<uint64_t> fsize(n, 0);
    std::vector<const uint8_t*> fptr(n, nullptr);
    for (uint32_t r = 0; r < W; ++r) {
        const uint32_t c = (uint32_t)of_rank[r].size();
        const uint64_t h = ((uint64_t)(c + 1) * 8 + 15) / 16 * 16;
        if (sizes[r] < h) return agc_fail(ctx, AGCGPU_ECUDA, "sharded coder: truncated block from rank %u", r);
        const uint64_t* offs = (const uint64_t*)(host.data() + (size_t)stride * r);
        if (sizes[r] != h + offs[c]) return agc_fail(ctx, AGCGPU_ECUDA, "sharded coder: block of rank %u has the wrong size", r);
        for (uint32_t j = 0; j < c; ++j) { fsize[of_rank[r][j]] = offs[j + 1] - offs[j]; fptr[of_rank[r][j]] = host.data() + (size_t)stride * r + h + offs[j]; }
    }
    dst_offsets[0] = 0;
    for (uint32_t i = 0; i < n; ++i) dst_offsets[i + 1] = dst_offsets[i] + fsize[i];
    if (dst_offsets[n] > dst_cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "zstd: need %llu output bytes", (unsigned long long)dst_offsets[n]);
    for (uint32_t i = 0; i < n; ++i) if (fsize[i]) memcpy(dst + dst_offsets[i], fptr[i], fsize[i]);
    return 0;
}
