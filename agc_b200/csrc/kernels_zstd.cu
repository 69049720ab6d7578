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
#include <chrono>
#include <cstring>

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

// the coder's call depth needs more than the default local-memory stack; cudaDeviceSetLimit waits for the device to go idle, so it
// is only called when the limit is not there yet (never while asynchronous waves are in flight: submit sets it first)
static int zs_stack_limit(agcgpu_ctx* ctx)
{
    size_t cur = 0;
    CK(cudaDeviceGetLimit(&cur, cudaLimitStackSize));
    if (cur < 16384) CK(cudaDeviceSetLimit(cudaLimitStackSize, 16384));
    return 0;
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
        if (int r = zs_stack_limit(ctx)) return r;
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


// ---------------------------------------------------------------------------------------------------------------- asynchronous waves
// agcgpu_zstd_submit / agcgpu_zstd_collect: the same coder, but a batch ("wave") is only QUEUED -- inputs copied to the device, its
// kernels launched on streams of its own -- and the call returns; the frames of every wave submitted so far come back from one
// agcgpu_zstd_collect.  A pack is coded while the host and the library stream go on with the next samples (the reference overlaps
// the same way: its compression threads run behind the segment queue, agc_compressor.cpp:1093-1272), so at Close only the last
// packs are still to be coded.  With a communicator every rank submits the same waves and codes its share of each (largest first
// to the least loaded rank, the same assignment on every rank); collect all-gathers the frames once.
struct ZWave {
    uint32_t n_in = 0;                               // inputs of the submit call
    std::vector<std::vector<uint32_t>> of_rank;      // with a communicator: who codes which input
    std::vector<uint32_t> mine;                      // inputs coded here, in launch order (wide first, largest first)
    std::vector<uint64_t> in_size;                   // per input
    uint32_t n_wide = 0;
    void* d_src = nullptr; size_t src_cap = 0;
    void* d_ws = nullptr; size_t ws_cap = 0;
    void* d_out = nullptr; size_t out_cap = 0;
    void* d_tasks = nullptr; size_t tasks_cap = 0;
    uint64_t osum = 0;
    std::vector<ZTaskDev> tasks;
    cudaStream_t st_w = nullptr, st_n = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr, en = nullptr;
    bool launched = false;
    // no room on the device at submit time: the inputs wait on the host and are coded synchronously at collect
    std::vector<uint8_t> h_src; std::vector<uint64_t> h_offs; std::vector<int32_t> h_levels;
    std::vector<std::vector<uint8_t>> frames;        // results of `mine` (index = position in mine)
};

static void zwave_release(agcgpu_ctx* ctx, ZWave* w)
{
    if (w->st_w) { cudaStreamSynchronize(w->st_w); cudaStreamDestroy(w->st_w); }
    if (w->st_n) { cudaStreamSynchronize(w->st_n); cudaStreamDestroy(w->st_n); }
    if (w->e0) cudaEventDestroy(w->e0);
    if (w->e1) cudaEventDestroy(w->e1);
    if (w->en) cudaEventDestroy(w->en);
    agc_dev_free(ctx->dev, w->d_src, w->src_cap); agc_dev_free(ctx->dev, w->d_ws, w->ws_cap);
    agc_dev_free(ctx->dev, w->d_out, w->out_cap); agc_dev_free(ctx->dev, w->d_tasks, w->tasks_cap);
    delete w;
}
void agc_zstd_waves_drop(agcgpu_ctx* ctx)            // destroy(): waves nobody collected
{
    for (void* p : ctx->zwaves) zwave_release(ctx, (ZWave*)p);
    ctx->zwaves.clear();
}

static bool zs_is_text(const uint8_t* p, uint64_t len)
{
    uint32_t sym = 0; const uint32_t probe = 64;
    for (uint32_t k = 0; k < probe; ++k) { const uint8_t c = p[(uint64_t)k * (len / probe)]; sym += c < 32u || c == 0xffu; }
    return sym * 2 <= probe;                         // mostly printable: digits, ',', '.', letters
}

extern "C" void* agcgpu_host_alloc(uint64_t bytes, uint64_t* out_cap);
extern "C" void agcgpu_host_free(void* p, uint64_t cap);

// input i = ptrs[i][0 .. sizes[i])
static int zstd_submit_impl(agcgpu_ctx* ctx, const uint8_t* const* ptrs, const uint64_t* sizes, const int32_t* levels, uint32_t n)
{
    cudaSetDevice(ctx->dev);
    if (n == 0) return 0;
    std::vector<uint64_t> ws(n), ob(n);
    for (uint32_t i = 0; i < n; ++i) {
        const uint64_t len = sizes[i];
        ze::Params cp = ze::get_params(levels[i], len);
        if (!cp.supported)
            return agc_fail(ctx, AGCGPU_EUNSUPPORTED, "zstd: input %u (level %d, %llu bytes) is outside the implemented envelope "
                            "(levels 13/17/18/19, inputs below 1 GiB)", i, levels[i], (unsigned long long)len);
        ws[i] = (ze::work_sizes(cp).total + 255) / 256 * 256;
        ob[i] = (ze::compress_bound(len) + 64 + 255) / 256 * 256;
    }
    ZWave* w = new ZWave;
    w->n_in = n; w->in_size.resize(n);
    for (uint32_t i = 0; i < n; ++i) w->in_size[i] = sizes[i];
    std::vector<uint32_t> order(n);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return w->in_size[a] > w->in_size[b]; });
    if (agc_comm_active()) {
        const uint32_t W = agc_comm_world();
        std::vector<uint64_t> load(W, 0);
        w->of_rank.assign(W, {});
        for (uint32_t i : order) {
            const uint32_t r = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
            w->of_rank[r].push_back(i); load[r] += w->in_size[i] + 512;
        }
        w->mine = w->of_rank[agc_comm_rank()];
    } else w->mine = order;
    // Delta text ("0,85.C0,33.A!!!!...": runs and a tiny alphabet cut the match-finder windows after a few dozen positions, the
    // helper warps of the wide coder idle and the one-warp coder is faster per frame, 4.9 vs 5.8 us/B) always takes the narrow
    // coder here; other inputs above 32 KB (packs of raw sequences) take the wide one.  Scheduling only: same bytes either way.
    std::vector<uint8_t> wide(n, 0);
    for (uint32_t i : w->mine)
        if (w->in_size[i] > zs_narrow_max() && (getenv("AGCGPU_ZSTD_WIDE_ALL") || !zs_is_text(ptrs[i], w->in_size[i]))) wide[i] = 1;
    std::stable_sort(w->mine.begin(), w->mine.end(), [&](uint32_t a, uint32_t b) {
        if (wide[a] != wide[b]) return wide[a] > wide[b];
        return w->in_size[a] > w->in_size[b]; });
    const uint32_t cnt = (uint32_t)w->mine.size();
    ctx->zwaves.push_back(w);
    if (cnt == 0) return 0;
    uint64_t total_src = 0, wsum = 0, osum = 0;
    for (uint32_t i : w->mine) { total_src += w->in_size[i]; wsum += ws[i]; osum += ob[i]; w->n_wide += wide[i]; }
    ctx->stats.zstd_input_mb += (float)(total_src * 1e-6);
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    auto park = [&]() {                              // keep the inputs on the host; collect codes them with the synchronous call
        w->h_offs.assign(1, 0); w->h_levels.clear(); w->h_src.resize(total_src);
        for (uint32_t i : w->mine) {
            if (w->in_size[i]) memcpy(w->h_src.data() + w->h_offs.back(), ptrs[i], w->in_size[i]);
            w->h_offs.push_back(w->h_offs.back() + w->in_size[i]); w->h_levels.push_back(levels[i]);
        }
        return 0;
    };
    if (total_src + wsum + osum + (64ull << 20) > (uint64_t)(free_b * 0.5) || getenv("AGCGPU_ZSTD_SYNC")) return park();
    w->d_src = agc_dev_alloc(ctx->dev, total_src + 256, &w->src_cap);
    w->d_ws = agc_dev_alloc(ctx->dev, wsum + 256, &w->ws_cap);
    w->d_out = agc_dev_alloc(ctx->dev, osum + 256, &w->out_cap);
    w->d_tasks = agc_dev_alloc(ctx->dev, cnt * sizeof(ZTaskDev) + 256, &w->tasks_cap);
    if (!w->d_src || !w->d_ws || !w->d_out || !w->d_tasks) {
        agc_dev_free(ctx->dev, w->d_src, w->src_cap); agc_dev_free(ctx->dev, w->d_ws, w->ws_cap);
        agc_dev_free(ctx->dev, w->d_out, w->out_cap); agc_dev_free(ctx->dev, w->d_tasks, w->tasks_cap);
        w->d_src = w->d_ws = w->d_out = w->d_tasks = nullptr;
        return park();
    }
    w->osum = osum;
    if (int r = zs_stack_limit(ctx)) return r;
    CK(cudaStreamCreateWithFlags(&w->st_w, cudaStreamNonBlocking));
    CK(cudaEventCreate(&w->e0)); CK(cudaEventCreate(&w->e1));
    w->tasks.resize(cnt);
    {   // inputs back to back in launch order, gathered into a page-locked staging buffer (pooled) and sent by DMA
        uint64_t stage_cap = 0;
        uint8_t* stage = (uint8_t*)agcgpu_host_alloc(total_src + 64, &stage_cap);
        std::vector<uint8_t> pageable;
        if (!stage) { pageable.resize(total_src + 1); stage = pageable.data(); }
        uint64_t so = 0, wo = 0, oo = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t i = w->mine[j];
            if (w->in_size[i]) memcpy(stage + so, ptrs[i], w->in_size[i]);
            ZTaskDev& k = w->tasks[j];
            k.src = (const uint8_t*)w->d_src + so; k.n = w->in_size[i];
            k.dst = (uint8_t*)w->d_out + oo; k.dst_cap = ob[i]; k.mem = (uint8_t*)w->d_ws + wo;
            k.level = levels[i]; k.err = 0; k.out_size = 0; k.t_start = k.t_end = 0;
            so += w->in_size[i]; wo += ws[i]; oo += ob[i];
        }
        cudaError_t e = cudaMemcpyAsync(w->d_src, stage, total_src, cudaMemcpyHostToDevice, w->st_w);
        if (e == cudaSuccess) e = cudaStreamSynchronize(w->st_w);          // the staging buffer goes back to the pool
        if (pageable.empty()) agcgpu_host_free(stage, stage_cap);
        if (e != cudaSuccess) return agc_fail(ctx, AGCGPU_ECUDA, "zstd submit: upload failed: %s", cudaGetErrorString(e));
        ctx->stats.h2d_bytes += total_src;
    }
    CK(cudaMemsetAsync(w->d_ws, 0, wsum, w->st_w));
    CK(cudaMemcpyAsync(w->d_tasks, w->tasks.data(), cnt * sizeof(ZTaskDev), cudaMemcpyHostToDevice, w->st_w));
    CK(cudaEventRecord(w->e0, w->st_w));
    if (w->n_wide) {
        const uint32_t smem_bytes = ze::fast_sizes().total;
        CK(cudaFuncSetAttribute(k_zstd, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        k_zstd<<<w->n_wide, ZS_THREADS, smem_bytes, w->st_w>>>((ZTaskDev*)w->d_tasks, w->n_wide, smem_bytes);
        CKL();
    }
    if (cnt > w->n_wide) {
        const uint32_t smem_n = zen::fast_sizes().total;
        cudaStream_t sn = w->st_w;
        if (w->n_wide) {                             // beside the wide kernel, not behind it
            CK(cudaStreamCreateWithFlags(&w->st_n, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&w->en, cudaEventDisableTiming));
            CK(cudaStreamWaitEvent(w->st_n, w->e0, 0));
            sn = w->st_n;
        }
        k_zstd_narrow<<<cnt - w->n_wide, 32, smem_n, sn>>>((ZTaskDev*)w->d_tasks + w->n_wide, cnt - w->n_wide, smem_n);
        CKL();
        if (w->st_n) { CK(cudaEventRecord(w->en, w->st_n)); CK(cudaStreamWaitEvent(w->st_w, w->en, 0)); }
    }
    CK(cudaEventRecord(w->e1, w->st_w));
    w->launched = true;
    return 0;
}

extern "C" int agcgpu_zstd_submit(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* src_offsets, const int32_t* levels, uint32_t n)
{
    if (!ctx || !src_offsets || (n && (!src || !levels))) return AGCGPU_EINVAL;
    std::vector<const uint8_t*> ptrs(n); std::vector<uint64_t> sizes(n);
    for (uint32_t i = 0; i < n; ++i) { ptrs[i] = src + src_offsets[i]; sizes[i] = src_offsets[i + 1] - src_offsets[i]; }
    return zstd_submit_impl(ctx, ptrs.data(), sizes.data(), levels, n);
}
// the same with one pointer per input (the parts of a pack queue are separate buffers: no concatenation on the caller's side)
extern "C" int agcgpu_zstd_submit_parts(agcgpu_ctx* ctx, const uint8_t* const* ptrs, const uint64_t* sizes, const int32_t* levels, uint32_t n)
{
    if (!ctx || (n && (!ptrs || !sizes || !levels))) return AGCGPU_EINVAL;
    for (uint32_t i = 0; i < n; ++i) if (sizes[i] && !ptrs[i]) return AGCGPU_EINVAL;
    return zstd_submit_impl(ctx, ptrs, sizes, levels, n);
}

// frames of every input submitted since the last collect, in submission order: frame i = dst[dst_offsets[i] .. dst_offsets[i+1]);
// n_expected = total number of inputs (checked)
extern "C" int agcgpu_zstd_collect(agcgpu_ctx* ctx, uint32_t n_expected, uint8_t* dst, uint64_t dst_cap, uint64_t* dst_offsets)
{
    if (!ctx || !dst_offsets || (n_expected && !dst)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    std::vector<ZWave*> waves;
    for (void* p : ctx->zwaves) waves.push_back((ZWave*)p);
    ctx->zwaves.clear();
    struct Guard { agcgpu_ctx* c; std::vector<ZWave*>& v; ~Guard() { for (ZWave* w : v) zwave_release(c, w); } } guard{ ctx, waves };
    uint64_t n_total = 0;
    for (ZWave* w : waves) n_total += w->n_in;
    dst_offsets[0] = 0;
    int local = 0;
    if (n_total != n_expected) local = agc_fail(ctx, AGCGPU_EINVAL, "zstd collect: %llu inputs were submitted, the caller expects %u", (unsigned long long)n_total, n_expected);
    const bool trace = getenv("AGCGPU_TRACE") != nullptr;
    const auto t_wait0 = std::chrono::steady_clock::now();
    // 1. this rank's frames of every wave, on the host
    for (size_t wi = 0; wi < waves.size() && !local; ++wi) {
        ZWave* w = waves[wi];
        const uint32_t cnt = (uint32_t)w->mine.size();
        w->frames.resize(cnt);
        if (!cnt) continue;
        if (!w->launched) {
            std::vector<uint64_t> fo((size_t)cnt + 1, 0);
            const uint64_t cap = w->h_offs.back() + w->h_offs.back() / 128 + 1024ull * (cnt + 1);
            std::vector<uint8_t> out(cap);
            local = zstd_compress_batch_impl(ctx, w->h_src.data(), w->h_offs.data(), w->h_levels.data(), cnt, out.data(), cap, fo.data(), ~0ull);
            if (!local) for (uint32_t j = 0; j < cnt; ++j) w->frames[j].assign(out.begin() + fo[j], out.begin() + fo[j + 1]);
            continue;
        }
        if (cudaEventSynchronize(w->e1) != cudaSuccess) { local = agc_fail(ctx, AGCGPU_ECUDA, "zstd wave %zu: %s", wi, cudaGetErrorString(cudaGetLastError())); break; }
        float ms = 0; cudaEventElapsedTime(&ms, w->e0, w->e1);
        ctx->stats.zstd_kernel_ms += ms;
        if (cudaMemcpyAsync(w->tasks.data(), w->d_tasks, cnt * sizeof(ZTaskDev), cudaMemcpyDeviceToHost, w->st_w) != cudaSuccess ||
            cudaStreamSynchronize(w->st_w) != cudaSuccess) { local = agc_fail(ctx, AGCGPU_ECUDA, "zstd wave %zu: task copy failed", wi); break; }
        uint64_t out_total = 0;
        for (uint32_t j = 0; j < cnt && !local; ++j) {
            if (w->tasks[j].err) local = agc_fail(ctx, AGCGPU_EUNSUPPORTED, "zstd: device coder failed on input %u of wave %zu (code %d)", w->mine[j], wi, w->tasks[j].err);
            out_total += w->tasks[j].out_size;
        }
        if (local) break;
        if (trace) {
            uint64_t tot = 0, big = 0; for (uint32_t j = 0; j < cnt; ++j) { tot += w->tasks[j].n; big = std::max<uint64_t>(big, w->tasks[j].n); }
            fprintf(stderr, "[agcgpu] zstd wave %zu (async): %u inputs (%u wide), %llu bytes, largest %llu, device %.1f ms\n", wi, cnt, w->n_wide,
                    (unsigned long long)tot, (unsigned long long)big, ms);
        }
        if (w->osum <= (64ull << 20) || out_total * 2 >= w->osum) {
            std::vector<uint8_t> host(w->osum);
            if (cudaMemcpyAsync(host.data(), w->d_out, w->osum, cudaMemcpyDeviceToHost, w->st_w) != cudaSuccess || cudaStreamSynchronize(w->st_w) != cudaSuccess) {
                local = agc_fail(ctx, AGCGPU_ECUDA, "zstd wave %zu: frame copy failed", wi); break; }
            ctx->stats.d2h_bytes += w->osum;
            for (uint32_t j = 0; j < cnt; ++j) { const uint8_t* p = host.data() + (w->tasks[j].dst - (uint8_t*)w->d_out); w->frames[j].assign(p, p + w->tasks[j].out_size); }
        } else {
            for (uint32_t j = 0; j < cnt; ++j) {
                w->frames[j].resize(w->tasks[j].out_size);
                if (cudaMemcpyAsync(w->frames[j].data(), w->tasks[j].dst, w->tasks[j].out_size, cudaMemcpyDeviceToHost, w->st_w) != cudaSuccess) { local = AGCGPU_ECUDA; break; }
                ctx->stats.d2h_bytes += w->tasks[j].out_size;
            }
            if (cudaStreamSynchronize(w->st_w) != cudaSuccess || local) { local = agc_fail(ctx, AGCGPU_ECUDA, "zstd wave %zu: frame copy failed", wi); break; }
        }
    }
    ctx->stats.zstd_wait_ms += (float)std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_wait0).count();
    if (!agc_comm_active()) {
        if (local) return local;
        uint64_t i = 0;
        for (ZWave* w : waves) {
            std::vector<uint64_t> sz(w->n_in, 0);
            for (size_t j = 0; j < w->mine.size(); ++j) sz[w->mine[j]] = w->frames[j].size();
            for (uint32_t k = 0; k < w->n_in; ++k, ++i) dst_offsets[i + 1] = dst_offsets[i] + sz[k];
        }
        if (dst_offsets[n_total] > dst_cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "zstd: need %llu output bytes", (unsigned long long)dst_offsets[n_total]);
        uint64_t base = 0;
        for (ZWave* w : waves) {
            for (size_t j = 0; j < w->mine.size(); ++j) if (!w->frames[j].empty()) memcpy(dst + dst_offsets[base + w->mine[j]], w->frames[j].data(), w->frames[j].size());
            base += w->n_in;
        }
        return 0;
    }
    // 2. all-gather: block of a rank = [u64 size of each of its frames, wave by wave in its launch order][the frames back to back]
    const uint32_t W = agc_comm_world(), me = agc_comm_rank();
    uint64_t my_cnt = 0, my_bytes = 0;
    for (ZWave* w : waves) { my_cnt += w->mine.size(); if (!local) for (auto& f : w->frames) my_bytes += f.size(); }
    const uint64_t hdr = (my_cnt * 8 + 15) / 16 * 16;
    std::vector<uint8_t> blk(hdr + my_bytes + 16, 0);
    if (!local) {
        // frames in the order of the assignment (of_rank), which every rank can rebuild -- `mine` is in launch order
        uint64_t* sz = (uint64_t*)blk.data(); uint64_t o = hdr, k = 0;
        for (ZWave* w : waves) {
            std::vector<uint32_t> at(w->n_in, 0);
            for (size_t j = 0; j < w->mine.size(); ++j) at[w->mine[j]] = (uint32_t)j;
            for (uint32_t i : w->of_rank[me]) { auto& f = w->frames[at[i]]; sz[k++] = f.size(); if (!f.empty()) memcpy(blk.data() + o, f.data(), f.size()); o += f.size(); }
        }
    }
    if (!local) local = agc_reserve(ctx, ctx->scr_zkeep, hdr + my_bytes + 64);
    if (!local && cudaMemcpyAsync(ctx->scr_zkeep.p, blk.data(), hdr + my_bytes, cudaMemcpyHostToDevice, ctx->st) != cudaSuccess) local = AGCGPU_ECUDA;
    std::vector<uint64_t> sizes; uint64_t stride = 0;
    if (int r = agc_comm_allgatherv(ctx, ctx->scr_zkeep.p, hdr + my_bytes, local, sizes, &stride)) return r;
    std::vector<uint8_t> host((size_t)stride * W);
    if (stride) CK(cudaMemcpyAsync(host.data(), ctx->scr_gather.p, (size_t)stride * W, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.d2h_bytes += (size_t)stride * W;
    std::vector<const uint8_t*> fptr(n_total, nullptr);
    std::vector<uint64_t> fsize(n_total, 0);
    for (uint32_t r = 0; r < W; ++r) {
        uint64_t c = 0; for (ZWave* w : waves) c += w->of_rank[r].size();
        const uint64_t h = (c * 8 + 15) / 16 * 16;
        if (sizes[r] < h) return agc_fail(ctx, AGCGPU_ECUDA, "sharded coder: truncated block from rank %u", r);
        const uint64_t* sz = (const uint64_t*)(host.data() + (size_t)stride * r);
        uint64_t o = h, k = 0, base = 0;
        for (ZWave* w : waves) {
            for (uint32_t i : w->of_rank[r]) {
                if (o + sz[k] > sizes[r]) return agc_fail(ctx, AGCGPU_ECUDA, "sharded coder: block of rank %u has the wrong size", r);
                fptr[base + i] = host.data() + (size_t)stride * r + o; fsize[base + i] = sz[k]; o += sz[k]; ++k;
            }
            base += w->n_in;
        }
        if (o != sizes[r]) return agc_fail(ctx, AGCGPU_ECUDA, "sharded coder: block of rank %u has the wrong size", r);
    }
    for (uint64_t i = 0; i < n_total; ++i) dst_offsets[i + 1] = dst_offsets[i] + fsize[i];
    if (dst_offsets[n_total] > dst_cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "zstd: need %llu output bytes", (unsigned long long)dst_offsets[n_total]);
    for (uint64_t i = 0; i < n_total; ++i) if (fsize[i]) memcpy(dst + dst_offsets[i], fptr[i], fsize[i]);
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
