// kernels_lz_dec.cu -- CLZDiff_V2::Decode (src/common/lz_diff.cpp:801-836) on the device, as CSegment::get calls it
// (src/common/segment.cpp:220-399): a delta + its group's resident reference -> the segment's symbols, 1 byte each.
// The parse of a delta is sequential (every token moves pred_pos), so one thread walks one delta; the deltas of a batch run
// side by side.  Used by the decode-and-compare self check of `create`; the decompression side proper is out of scope.
#include "internal.cuh"
#include <algorithm>

struct LzDecTask { const uint8_t* enc; uint64_t en; uint8_t* out; uint64_t cap; uint32_t group; int32_t err; uint64_t produced; };

__device__ __forceinline__ uint8_t ref_sym(const GroupRefDev& g, uint32_t pos)
{
    if (g.flags & GRF_DIRTY) return g.codes[pos];
    return (g.packed[pos >> 2] >> (6 - 2 * (pos & 3))) & 3;
}

// write != 0: symbols go to t.out; write == 0: only the size is computed
__global__ void __launch_bounds__(32) k_lz_decode(LzDecTask* __restrict__ tasks, uint32_t n, const GroupRefDev* __restrict__ groups,
                                                 uint32_t min_match_len, int write)
{
    const uint32_t i = blockIdx.x;
    if (i >= n || threadIdx.x != 0) return;
    LzDecTask t = tasks[i];
    const GroupRefDev g = groups[t.group];
    uint64_t o = 0, p = 0;
    uint32_t pred_pos = 0;
    int err = 0;
    while (p < t.en && !err) {
        const uint8_t c = t.enc[p];
        if ((c >= 'A' && c <= 'A' + 20) || c == '!') {                 // literal (is_literal / decode_literal, lz_diff.h)
            if (c == '!' && pred_pos >= g.m) { err = 1; break; }
            if (write) { if (o >= t.cap) { err = 2; break; } t.out[o] = c == '!' ? ref_sym(g, pred_pos) : (uint8_t)(c - 'A'); }
            ++o; ++pred_pos; ++p;
        } else if (c == 30) {                                           // N run: 30 <len-4> 4
            ++p; uint64_t v = 0;
            while (p < t.en && t.enc[p] >= '0' && t.enc[p] <= '9') v = v * 10 + (t.enc[p++] - '0');
            ++p;
            if (write) { if (o + v + 4 > t.cap) { err = 2; break; } for (uint64_t q = 0; q < v + 4; ++q) t.out[o + q] = 4; }
            o += v + 4;
        } else {                                                        // match: [-]<ref_pos - pred_pos>[,<len - min_match_len>].
            bool neg = false; int64_t v = 0;
            if (c == '-') { neg = true; ++p; }
            while (p < t.en && t.enc[p] >= '0' && t.enc[p] <= '9') v = v * 10 + (t.enc[p++] - '0');
            if (neg) v = -v;
            const uint32_t rp = (uint32_t)(v + (int64_t)pred_pos);
            uint32_t len;
            if (p < t.en && t.enc[p] == ',') {
                ++p; int64_t l = 0;
                while (p < t.en && t.enc[p] >= '0' && t.enc[p] <= '9') l = l * 10 + (t.enc[p++] - '0');
                len = (uint32_t)(l + min_match_len); ++p;
            } else { if (rp > g.m) { err = 1; break; } len = g.m - rp; ++p; }      // to the end of the reference
            if ((uint64_t)rp + len > g.m) { err = 1; break; }
            if (write) { if (o + len > t.cap) { err = 2; break; } for (uint32_t q = 0; q < len; ++q) t.out[o + q] = ref_sym(g, rp + q); }
            o += len; pred_pos = rp + len;
        }
    }
    tasks[i].err = err; tasks[i].produced = o;
}

extern "C" int agcgpu_lz_decode_batch(agcgpu_ctx* ctx, const uint32_t* group_ids, const uint8_t* deltas, const uint64_t* delta_offsets,
                                      uint32_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_offsets)
{
    if (!ctx || !out_offsets || !delta_offsets || (n && (!group_ids || !deltas)) || (out_cap && !out)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    out_offsets[0] = 0;
    if (n == 0) return 0;
    for (uint32_t i = 0; i < n; ++i)
        if (group_ids[i] >= ctx->h_groups.size() || !(ctx->h_groups[group_ids[i]].flags & GRF_PRESENT))
            return agc_fail(ctx, AGCGPU_EINVAL, "lz_decode: group %u has no reference", group_ids[i]);
    const uint64_t total_in = delta_offsets[n];
    if (int r = agc_reserve(ctx, ctx->scr_bytes, total_in + 64)) return r;
    if (int r = agc_reserve(ctx, ctx->scr_req, (size_t)n * sizeof(LzDecTask))) return r;
    if (total_in) { CK(cudaMemcpyAsync(ctx->scr_bytes.p, deltas, total_in, cudaMemcpyHostToDevice, ctx->st)); ctx->stats.h2d_bytes += total_in; }
    std::vector<LzDecTask> tasks(n);
    for (uint32_t i = 0; i < n; ++i) {
        tasks[i].enc = (const uint8_t*)ctx->scr_bytes.p + delta_offsets[i]; tasks[i].en = delta_offsets[i + 1] - delta_offsets[i];
        tasks[i].out = nullptr; tasks[i].cap = 0; tasks[i].group = group_ids[i]; tasks[i].err = 0; tasks[i].produced = 0;
    }
    // pass 1: sizes
    CK(cudaMemcpyAsync(ctx->scr_req.p, tasks.data(), (size_t)n * sizeof(LzDecTask), cudaMemcpyHostToDevice, ctx->st));
    k_lz_decode<<<n, 32, 0, ctx->st>>>((LzDecTask*)ctx->scr_req.p, n, (const GroupRefDev*)ctx->d_groups.p, ctx->prm.min_match_len, 0);
    CKL();
    CK(cudaMemcpyAsync(tasks.data(), ctx->scr_req.p, (size_t)n * sizeof(LzDecTask), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    for (uint32_t i = 0; i < n; ++i) {
        if (tasks[i].err) return agc_fail(ctx, AGCGPU_EINVAL, "lz_decode: delta %u does not fit its reference (group %u)", i, group_ids[i]);
        out_offsets[i + 1] = out_offsets[i] + tasks[i].produced;
    }
    if (out_offsets[n] > out_cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "lz_decode: need %llu output bytes", (unsigned long long)out_offsets[n]);
    if (out_offsets[n] == 0) return 0;
    // pass 2: symbols
    if (int r = agc_reserve(ctx, ctx->scr_dense, out_offsets[n] + 64)) return r;
    for (uint32_t i = 0; i < n; ++i) { tasks[i].out = (uint8_t*)ctx->scr_dense.p + out_offsets[i]; tasks[i].cap = tasks[i].produced; }
    CK(cudaMemcpyAsync(ctx->scr_req.p, tasks.data(), (size_t)n * sizeof(LzDecTask), cudaMemcpyHostToDevice, ctx->st));
    k_lz_decode<<<n, 32, 0, ctx->st>>>((LzDecTask*)ctx->scr_req.p, n, (const GroupRefDev*)ctx->d_groups.p, ctx->prm.min_match_len, 1);
    CKL();
    CK(cudaMemcpyAsync(tasks.data(), ctx->scr_req.p, (size_t)n * sizeof(LzDecTask), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(out, ctx->scr_dense.p, out_offsets[n], cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.d2h_bytes += out_offsets[n];
    for (uint32_t i = 0; i < n; ++i) if (tasks[i].err) return agc_fail(ctx, AGCGPU_EINVAL, "lz_decode: delta %u failed in the write pass", i);
    return 0;
}
