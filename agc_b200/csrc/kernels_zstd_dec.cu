// kernels_zstd_dec.cu -- the decoding side of the residual coder on the device: ZSTD_decompressDCtx as CSegment::unpack / get
// (src/common/segment.cpp:500-577, 220-399) and CCollection_V3 use it, for a batch of independent frames.  One CTA per frame,
// its first thread runs the (inherently sequential) frame decoder of zstd_dec.cuh; the frames of a batch run side by side.
// Used by the decode-and-compare self check of `create` (CAGCCompressor::SetVerify); `append` will need it to reload packs.
#include "internal.cuh"
#include "zstd_dec.cuh"
#include <algorithm>

struct ZDTaskDev { const uint8_t* src; uint64_t n; uint8_t* dst; uint64_t cap; long long result; };

__global__ void __launch_bounds__(32) k_zstd_decode(ZDTaskDev* __restrict__ tasks, uint32_t n, zd::Work* __restrict__ work)
{
    const uint32_t i = blockIdx.x;
    if (i >= n || threadIdx.x != 0) return;
    tasks[i].result = zd::decompress_frame(tasks[i].src, tasks[i].n, tasks[i].dst, tasks[i].cap, work[i]);
}

extern "C" int agcgpu_zstd_decompress_batch(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* src_offsets, uint32_t n, uint8_t* dst,
                                            uint64_t dst_cap, uint64_t* dst_offsets)
{
    if (!ctx || !src_offsets || !dst_offsets || (n && !src) || (dst_cap && !dst)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    dst_offsets[0] = 0;
    if (n == 0) return 0;
    for (uint32_t i = 0; i < n; ++i) {                      // output sizes come from the frame headers
        int64_t fcs = zd::frame_content_size(src + src_offsets[i], src_offsets[i + 1] - src_offsets[i]);
        if (fcs < 0) return agc_fail(ctx, AGCGPU_EUNSUPPORTED, "zstd decode: frame %u has no content size in its header", i);
        dst_offsets[i + 1] = dst_offsets[i] + (uint64_t)fcs;
    }
    if (dst_offsets[n] > dst_cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "zstd decode: need %llu output bytes", (unsigned long long)dst_offsets[n]);
    const uint64_t total_src = src_offsets[n], total_dst = dst_offsets[n];
    if (int r = agc_reserve(ctx, ctx->scr_bytes, total_src + 64)) return r;
    if (int r = agc_reserve(ctx, ctx->scr_dense, total_dst + 64)) return r;
    CK(cudaMemcpyAsync(ctx->scr_bytes.p, src, total_src, cudaMemcpyHostToDevice, ctx->st));
    ctx->stats.h2d_bytes += total_src;
    const uint32_t wave = 1024;                             // frames in flight: 1024 x sizeof(zd::Work) ~ 150 MB of tables and literal buffers
    if (int r = agc_reserve(ctx, ctx->scr_out, (size_t)std::min<uint32_t>(n, wave) * sizeof(zd::Work) + 256)) return r;
    if (int r = agc_reserve(ctx, ctx->scr_req, (size_t)std::min<uint32_t>(n, wave) * sizeof(ZDTaskDev))) return r;
    {   size_t cur = 0;                                   // (setting the limit waits for an idle device: only when it is not there yet)
        CK(cudaDeviceGetLimit(&cur, cudaLimitStackSize));
        if (cur < 16384) CK(cudaDeviceSetLimit(cudaLimitStackSize, 16384)); }
    for (uint32_t pos = 0; pos < n; pos += wave) {
        const uint32_t cnt = std::min<uint32_t>(wave, n - pos);
        std::vector<ZDTaskDev> tasks(cnt);
        for (uint32_t j = 0; j < cnt; ++j) {
            const uint32_t i = pos + j;
            tasks[j].src = (const uint8_t*)ctx->scr_bytes.p + src_offsets[i]; tasks[j].n = src_offsets[i + 1] - src_offsets[i];
            tasks[j].dst = (uint8_t*)ctx->scr_dense.p + dst_offsets[i]; tasks[j].cap = dst_offsets[i + 1] - dst_offsets[i];
            tasks[j].result = 0;
        }
        CK(cudaMemcpyAsync(ctx->scr_req.p, tasks.data(), cnt * sizeof(ZDTaskDev), cudaMemcpyHostToDevice, ctx->st));
        k_zstd_decode<<<cnt, 32, 0, ctx->st>>>((ZDTaskDev*)ctx->scr_req.p, cnt, (zd::Work*)ctx->scr_out.p);
        CKL();
        CK(cudaMemcpyAsync(tasks.data(), ctx->scr_req.p, cnt * sizeof(ZDTaskDev), cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        for (uint32_t j = 0; j < cnt; ++j)
            if (tasks[j].result < 0 || (uint64_t)tasks[j].result != tasks[j].cap)
                return agc_fail(ctx, AGCGPU_EINVAL, "zstd decode: frame %u is malformed or truncated (code %lld)", pos + j, tasks[j].result);
    }
    if (total_dst) { CK(cudaMemcpy(dst, ctx->scr_dense.p, total_dst, cudaMemcpyDeviceToHost)); ctx->stats.d2h_bytes += total_dst; }
    return 0;
}
