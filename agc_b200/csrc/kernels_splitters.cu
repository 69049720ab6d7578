// kernels_splitters.cu -- determine_splitters on the device (src/core/agc_compressor.cpp:428-563).
//
//   k_enum_kmers      : start_kmer_collecting_threads (707-759): every canonical k-mer of the reference sample
//   cub radix sort    : raduls::RadixSortMSD (490) -- a third-party sort library in the reference too
//   k_singleton_flags : remove_non_singletons (664-704)
//   k_find_splitters  : find_splitters_in_contig (762-825): one warp walks one contig; membership in the sorted
//                       singleton list is tested lazily (binary search) for 32 consecutive positions per step.
#include "internal.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_select.cuh>
#include <algorithm>

#define FULL 0xffffffffu

__device__ __forceinline__ uint64_t canon_at(const uint64_t* __restrict__ P, uint64_t gend, uint32_t k)
{
    const uint32_t shift = 64 - 2 * k;
    uint64_t dir = agc_win(P, gend - (k - 1)) & ((~0ULL) << shift);
    uint64_t rc = (~agc_rev2(dir)) << shift;
    return dir < rc ? dir : rc;
}
// last exception position <= hi that is >= lo, or ~0
__device__ __forceinline__ uint64_t last_exc_in(const uint64_t* __restrict__ exc_pos, uint64_t n_exc, uint64_t lo, uint64_t hi)
{
    if (n_exc == 0) return ~0ULL;
    uint64_t a = 0, b = n_exc;                       // first index with pos > hi
    while (a < b) { uint64_t mid = (a + b) >> 1; if (exc_pos[mid] <= hi) a = mid + 1; else b = mid; }
    if (a == 0) return ~0ULL;
    uint64_t e = exc_pos[a - 1];
    return e >= lo ? e : ~0ULL;
}

__global__ void __launch_bounds__(256) k_enum_kmers(const uint64_t* __restrict__ P, const uint64_t* __restrict__ cstart,
                                                    uint32_t n_contigs, const uint32_t* __restrict__ chunk_prefix, uint32_t total_chunks,
                                                    uint32_t k, const uint64_t* __restrict__ exc_pos, uint64_t n_exc,
                                                    uint64_t* __restrict__ out, uint64_t out_base)
{
    out -= out_base;                                 // out[0] belongs to the first base of the first listed contig
    const uint32_t shift = 64 - 2 * k;
    const uint64_t kmask = (~0ULL) << shift;
    for (uint32_t u = blockIdx.x; u < total_chunks; u += gridDim.x) {
        uint32_t lo = 0, hi = n_contigs;
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (chunk_prefix[mid] <= u) lo = mid; else hi = mid; }
        uint32_t c = lo;
        uint64_t base = cstart[c], len = cstart[c + 1] - base;
        uint64_t p0 = (uint64_t)(u - chunk_prefix[c]) * AGC_SCAN_CHUNK + (uint64_t)threadIdx.x * 32;
        uint64_t pend = p0 + 32 < len ? p0 + 32 : len;
        for (uint64_t p = p0; p < pend && p < (uint64_t)(k - 1); ++p) out[base + p] = ~0ULL;
        uint64_t ps = p0 > (uint64_t)(k - 1) ? p0 : (uint64_t)(k - 1);
        if (ps >= pend) continue;
        bool any_exc = last_exc_in(exc_pos, n_exc, base + ps - (k - 1), base + pend - 1) != ~0ULL;
        uint64_t dir = agc_win(P, base + ps - (k - 1)) & kmask;
        uint64_t rc = (~agc_rev2(dir)) << shift;
        uint64_t nxt = agc_win(P, base + ps + 1);
        for (uint64_t p = ps; p < pend; ++p) {
            uint64_t canon = dir < rc ? dir : rc;
            if (any_exc && last_exc_in(exc_pos, n_exc, base + p - (k - 1), base + p) != ~0ULL) canon = ~0ULL;
            out[base + p] = canon;
            uint64_t s = nxt >> 62; nxt <<= 2;
            dir = ((dir << 2) | (s << shift)) & kmask;
            rc = ((rc >> 2) | ((3 - s) << 62)) & kmask;
        }
    }
}

__global__ void k_singleton_flags(const uint64_t* __restrict__ sorted, uint64_t n, uint8_t* __restrict__ flags)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t v = sorted[i];
    bool single = v != ~0ULL && (i == 0 || sorted[i - 1] != v) && (i + 1 == n || sorted[i + 1] != v);
    flags[i] = single;
}

__device__ __forceinline__ bool in_sorted(const uint64_t* __restrict__ v, uint64_t n, uint64_t x)
{
    uint64_t a = 0, b = n;
    while (a < b) { uint64_t mid = (a + b) >> 1; if (v[mid] < x) a = mid + 1; else b = mid; }
    return a < n && v[a] == x;
}

// find_new_splitters' two set_difference passes (agc_compressor.cpp:2066-2076) as one membership test per candidate
__global__ void k_absent_flags(const uint64_t* __restrict__ cand, uint64_t n, const uint64_t* __restrict__ ref, uint64_t n_ref,
                               uint8_t* __restrict__ flags)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = !in_sorted(ref, n_ref, cand[i]);
}

// -f mode: k-mers of one resident range that pass kmer_filter_t (agc_compressor.h:570-599); one thread per position,
// appended in any order (the host sorts by position)
__global__ void k_filter_kmers(const uint64_t* __restrict__ P, uint64_t gstart, uint64_t len, uint32_t k, uint64_t thr,
                               const uint64_t* __restrict__ exc_pos, uint64_t n_exc, agcgpu_fkmer* __restrict__ out,
                               uint32_t* __restrict__ out_count, uint32_t out_cap)
{
    const uint32_t shift = 64 - 2 * k;
    for (uint64_t p = (uint64_t)(k - 1) + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; p < len; p += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t g = gstart + p;
        if (last_exc_in(exc_pos, n_exc, g - (k - 1), g) != ~0ULL) continue;
        uint64_t dir = agc_win(P, g - (k - 1)) & ((~0ULL) << shift);
        uint64_t rc = (~agc_rev2(dir)) << shift;
        uint64_t canon = dir < rc ? dir : rc;
        if ((agc_murmur64(canon) ^ 0xD73F8BF11046C40EULL) >= thr) continue;
        uint32_t idx = atomicAdd(out_count, 1u);
        if (idx < out_cap) { agcgpu_fkmer f; f.pos = p; f.kmer = canon; f.is_dir_oriented = dir <= rc; f.is_symmetric = dir == rc; out[idx] = f; }
    }
}

int agc_filtered_kmers(agcgpu_ctx* ctx, uint64_t gstart, uint64_t len, uint64_t thr, std::vector<agcgpu_fkmer>& out)
{
    out.clear();
    const uint32_t k = ctx->prm.kmer_length;
    if (len < k) return 0;
    // expected count = len * thr / 2^64; room for 2x + slack, retried once with the exact count if that was not enough
    uint64_t cap = (uint64_t)((double)len * ((double)thr / 18446744073709551616.0) * 2.0) + 1024;
    if (cap > len) cap = len;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (cap > 0x7fffffffull) return agc_fail(ctx, AGCGPU_EUNSUPPORTED, "filtered_kmers: range too large");
        if (int r = agc_reserve(ctx, ctx->scr_misc, cap * sizeof(agcgpu_fkmer) + 64)) return r;
        if (int r = agc_reserve(ctx, ctx->counters, 64)) return r;
        CK(cudaMemsetAsync(ctx->counters.p, 0, 64, ctx->st));
        uint32_t grid = (uint32_t)std::min<uint64_t>((len + 255) / 256, (uint64_t)ctx->n_sm * 8);
        k_filter_kmers<<<grid, 256, 0, ctx->st>>>((const uint64_t*)ctx->packed.p, gstart, len, k, thr, (const uint64_t*)ctx->exc_pos.p,
                                                  ctx->n_exc, (agcgpu_fkmer*)ctx->scr_misc.p, (uint32_t*)ctx->counters.p, (uint32_t)cap);
        CKL();
        uint32_t cnt = 0;
        CK(cudaMemcpyAsync(&cnt, ctx->counters.p, 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        if (cnt > cap) { cap = cnt; continue; }
        out.resize(cnt);
        if (cnt) { CK(cudaMemcpy(out.data(), ctx->scr_misc.p, (size_t)cnt * sizeof(agcgpu_fkmer), cudaMemcpyDeviceToHost)); ctx->stats.d2h_bytes += (size_t)cnt * sizeof(agcgpu_fkmer); }
        std::sort(out.begin(), out.end(), [](const agcgpu_fkmer& a, const agcgpu_fkmer& b) { return a.pos < b.pos; });
        return 0;
    }
    return agc_fail(ctx, AGCGPU_EUNSUPPORTED, "filtered_kmers: count changed between two passes");
}

// one warp per contig
__global__ void __launch_bounds__(128) k_find_splitters(const uint64_t* __restrict__ P, const uint64_t* __restrict__ cstart,
                                                        uint32_t n_contigs, uint32_t k, uint64_t segment_size,
                                                        const uint64_t* __restrict__ singles, uint64_t n_singles,
                                                        const uint64_t* __restrict__ exc_pos, uint64_t n_exc,
                                                        uint64_t* __restrict__ out, uint32_t* __restrict__ out_count, uint32_t out_cap,
                                                        SplPos* __restrict__ out_where)
{
    uint32_t c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (c >= n_contigs) return;
    const uint64_t base = cstart[c], len = cstart[c + 1] - base;
    if (len < k) return;
    uint64_t q = k - 1;                 // next position to test (k-mer end, contig coordinates)
    int64_t last_acc = -1;              // position of the last accepted splitter
    // forward walk: first singleton at a position >= need
    while (q < len) {
        uint64_t p = q + lane;
        bool ok = false; uint64_t canon = 0; uint64_t skip_to = 0;
        if (p < len) {
            uint64_t e = last_exc_in(exc_pos, n_exc, base + p - (k - 1), base + p);
            if (e != ~0ULL) skip_to = e - base + k;      // no valid k-mer ends before e + k
            else { canon = canon_at(P, base + p, k); ok = in_sorted(singles, n_singles, canon); }
        }
        uint32_t mk = __ballot_sync(FULL, ok);
        if (mk) {
            uint32_t L = __ffs(mk) - 1;
            uint64_t pa = q + L;
            uint64_t d = __shfl_sync(FULL, canon, L);
            if (lane == 0) { uint32_t idx = atomicAdd(out_count, 1u); if (idx < out_cap) { out[idx] = d; out_where[idx] = SplPos{ pa, c, 0u }; } }
            last_acc = (int64_t)pa;
            q = pa + segment_size;
            continue;
        }
        // nothing in these 32 positions: advance; when every lane was blocked by an exception jump past the farthest one
        uint64_t nq = q + 32;
        uint32_t blocked = __ballot_sync(FULL, skip_to != 0 || p >= len);
        if (blocked == FULL) {
            uint64_t mx = skip_to;
#pragma unroll
            for (int o = 16; o; o >>= 1) { uint64_t t = __shfl_xor_sync(FULL, mx, o); mx = t > mx ? t : mx; }
            if (mx > nq) nq = mx;
        }
        q = nq;
    }
    // right-most singleton among the k-mers seen after the last accepted splitter (817-824)
    uint64_t lo_pos = last_acc < 0 ? (uint64_t)(k - 1) : (uint64_t)last_acc + k;
    if (lo_pos >= len) return;
    uint64_t hi_pos = len;              // exclusive
    while (hi_pos > lo_pos) {
        uint64_t span = hi_pos - lo_pos < 32 ? hi_pos - lo_pos : 32;
        bool ok = false; uint64_t canon = 0;
        if (lane < span) {
            uint64_t p = hi_pos - 1 - lane;
            if (last_exc_in(exc_pos, n_exc, base + p - (k - 1), base + p) == ~0ULL) {
                canon = canon_at(P, base + p, k); ok = in_sorted(singles, n_singles, canon);
            }
        }
        uint32_t mk = __ballot_sync(FULL, ok);
        if (mk) {
            uint32_t L = __ffs(mk) - 1;
            uint64_t d = __shfl_sync(FULL, canon, L);
            if (lane == 0) { uint32_t idx = atomicAdd(out_count, 1u); if (idx < out_cap) { out[idx] = d; out_where[idx] = SplPos{ hi_pos - 1 - L, c, 1u }; } }
            return;
        }
        hi_pos -= span;
    }
}

int agc_enumerate_splitters(agcgpu_ctx* ctx, uint32_t c0, uint32_t n_contigs, bool exclude_ref, bool keep_kmers, std::vector<uint64_t>& out_sorted)
{
    out_sorted.clear();
    const uint32_t k = ctx->prm.kmer_length;
    if (c0 + n_contigs > ctx->n_contigs) return agc_fail(ctx, AGCGPU_EINVAL, "splitters: contig range outside the resident batch");
    const uint64_t base0 = n_contigs ? ctx->h_cstart[c0] : 0;
    const uint64_t total = n_contigs ? ctx->h_cstart[c0 + n_contigs] - base0 : 0;
    if (keep_kmers && exclude_ref) return agc_fail(ctx, AGCGPU_EINVAL, "splitters: keep_kmers and exclude_ref are exclusive");
    if (keep_kmers) ctx->n_ref_kmers = 0;
    if (total == 0 || n_contigs == 0) return 0;
    if (exclude_ref && ctx->n_ref_kmers == 0 && !(ctx->prm.flags & AGCGPU_F_ADAPTIVE))
        return agc_fail(ctx, AGCGPU_EINVAL, "find_new_splitters needs AGCGPU_F_ADAPTIVE and a reference sample (agcgpu_determine_splitters)");
    std::vector<uint32_t> cp(n_contigs + 1);
    uint32_t total_chunks = 0;
    for (uint32_t c = 0; c < n_contigs; ++c) {
        cp[c] = total_chunks;
        total_chunks += (uint32_t)((ctx->h_cstart[c0 + c + 1] - ctx->h_cstart[c0 + c] + AGC_SCAN_CHUNK - 1) / AGC_SCAN_CHUNK);
    }
    cp[n_contigs] = total_chunks;
    if (int r = agc_reserve(ctx, ctx->chunk_prefix, (n_contigs + 1) * 4)) return r;
    CK(cudaMemcpyAsync(ctx->chunk_prefix.p, cp.data(), (n_contigs + 1) * 4, cudaMemcpyHostToDevice, ctx->st));
    const uint64_t* d_cstart = (const uint64_t*)ctx->d_cstart.p + c0;
    DevBuf keys_a, keys_b, flags, tmp, nsel, outb, whereb;
    int rc = 0;
    auto cleanup = [&]() { for (DevBuf* b : { &keys_a, &keys_b, &flags, &tmp, &nsel, &outb, &whereb }) if (b->p) { agc_dev_free(ctx->dev, b->p, b->cap + 64); ctx->device_bytes -= b->cap; b->p = nullptr; b->cap = 0; } };
    uint32_t out_cap = (uint32_t)std::min<uint64_t>(0x7fffffff, total / std::max<uint32_t>(ctx->prm.segment_size, 1) + 2ull * n_contigs + 16);
    if ((rc = agc_reserve(ctx, keys_a, total * 8)) || (rc = agc_reserve(ctx, keys_b, total * 8)) || (rc = agc_reserve(ctx, flags, total)) ||
        (rc = agc_reserve(ctx, nsel, 64)) || (rc = agc_reserve(ctx, outb, (size_t)out_cap * 8)) ||
        (rc = agc_reserve(ctx, whereb, (size_t)out_cap * sizeof(SplPos)))) { cleanup(); return rc; }
    uint32_t grid = std::min<uint32_t>(total_chunks, (uint32_t)ctx->n_sm * 8);
    k_enum_kmers<<<grid, 256, 0, ctx->st>>>((const uint64_t*)ctx->packed.p, d_cstart, n_contigs,
        (const uint32_t*)ctx->chunk_prefix.p, total_chunks, k, (const uint64_t*)ctx->exc_pos.p, ctx->n_exc, (uint64_t*)keys_a.p, base0);
    ctx->stats.kernel_launches++;
    size_t tb = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb, (const uint64_t*)keys_a.p, (uint64_t*)keys_b.p, (uint64_t)total, 0, 64, ctx->st);
    if ((rc = agc_reserve(ctx, tmp, tb + 256))) { cleanup(); return rc; }
    cub::DeviceRadixSort::SortKeys(tmp.p, tb, (const uint64_t*)keys_a.p, (uint64_t*)keys_b.p, (uint64_t)total, 0, 64, ctx->st);
    ctx->stats.kernel_launches++;
    k_singleton_flags<<<(uint32_t)((total + 255) / 256), 256, 0, ctx->st>>>((const uint64_t*)keys_b.p, total, (uint8_t*)flags.p);
    ctx->stats.kernel_launches++;
    size_t tb2 = 0;
    cub::DeviceSelect::Flagged(nullptr, tb2, (const uint64_t*)keys_b.p, (const uint8_t*)flags.p, (uint64_t*)keys_a.p, (uint64_t*)nsel.p, (int64_t)total, ctx->st);
    if ((rc = agc_reserve(ctx, tmp, tb2 + 256))) { cleanup(); return rc; }
    cub::DeviceSelect::Flagged(tmp.p, tb2, (const uint64_t*)keys_b.p, (const uint8_t*)flags.p, (uint64_t*)keys_a.p, (uint64_t*)nsel.p, (int64_t)total, ctx->st);
    ctx->stats.kernel_launches++;
    uint64_t n_singles = 0;
    cudaError_t e = cudaMemcpyAsync(&n_singles, nsel.p, 8, cudaMemcpyDeviceToHost, ctx->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
    if (e != cudaSuccess) { cleanup(); return agc_fail(ctx, AGCGPU_ECUDA, "determine_splitters: %s", cudaGetErrorString(e)); }
    const uint64_t* d_cand = (const uint64_t*)keys_a.p;
    if (exclude_ref && n_singles) {
        // candidates := singletons of these contigs that are not k-mers of the reference sample (sorted order is kept)
        k_absent_flags<<<(uint32_t)((n_singles + 255) / 256), 256, 0, ctx->st>>>((const uint64_t*)keys_a.p, n_singles,
            (const uint64_t*)ctx->ref_kmers.p, ctx->n_ref_kmers, (uint8_t*)flags.p);
        ctx->stats.kernel_launches++;
        size_t tb3 = 0;
        cub::DeviceSelect::Flagged(nullptr, tb3, (const uint64_t*)keys_a.p, (const uint8_t*)flags.p, (uint64_t*)keys_b.p, (uint64_t*)nsel.p, (int64_t)n_singles, ctx->st);
        if ((rc = agc_reserve(ctx, tmp, tb3 + 256))) { cleanup(); return rc; }
        cub::DeviceSelect::Flagged(tmp.p, tb3, (const uint64_t*)keys_a.p, (const uint8_t*)flags.p, (uint64_t*)keys_b.p, (uint64_t*)nsel.p, (int64_t)n_singles, ctx->st);
        ctx->stats.kernel_launches++;
        e = cudaMemcpyAsync(&n_singles, nsel.p, 8, cudaMemcpyDeviceToHost, ctx->st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
        if (e != cudaSuccess) { cleanup(); return agc_fail(ctx, AGCGPU_ECUDA, "find_new_splitters: %s", cudaGetErrorString(e)); }
        d_cand = (const uint64_t*)keys_b.p;
    }
    uint64_t* d_out = (uint64_t*)outb.p;
    uint32_t* d_cnt = (uint32_t*)nsel.p + 4;
    cudaMemsetAsync(nsel.p, 0, 64, ctx->st);
    k_find_splitters<<<(n_contigs + 3) / 4, 128, 0, ctx->st>>>((const uint64_t*)ctx->packed.p, d_cstart, n_contigs, k,
        ctx->prm.segment_size, d_cand, n_singles, (const uint64_t*)ctx->exc_pos.p, ctx->n_exc, d_out, d_cnt, out_cap, (SplPos*)whereb.p);
    ctx->stats.kernel_launches++;
    uint32_t cnt = 0;
    e = cudaMemcpyAsync(&cnt, d_cnt, 4, cudaMemcpyDeviceToHost, ctx->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->st);
    if (e == cudaSuccess && cnt <= out_cap && cnt) {
        out_sorted.resize(cnt); e = cudaMemcpy(out_sorted.data(), d_out, (size_t)cnt * 8, cudaMemcpyDeviceToHost);
        std::vector<SplPos> where(cnt);
        if (e == cudaSuccess) e = cudaMemcpy(where.data(), whereb.p, (size_t)cnt * sizeof(SplPos), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess) for (uint32_t i = 0; i < cnt; ++i)       // where each splitter was found (agcgpu_last_splitter_positions)
            ctx->h_last_spl.push_back(agcgpu_ctx::SplFound{ c0 + where[i].contig, where[i].pos, out_sorted[i], (uint8_t)where[i].is_last });
    }
    if (keep_kmers && e == cudaSuccess) {                  // the sorted k-mer list of the reference sample stays on the device
        if (ctx->ref_kmers.p) { agc_dev_free(ctx->dev, ctx->ref_kmers.p, ctx->ref_kmers.cap + 64); ctx->device_bytes -= ctx->ref_kmers.cap; }
        ctx->ref_kmers = keys_b; keys_b.p = nullptr; keys_b.cap = 0;
        ctx->n_ref_kmers = total;
    }
    cleanup();
    if (e != cudaSuccess) return agc_fail(ctx, AGCGPU_ECUDA, "determine_splitters: %s", cudaGetErrorString(e));
    if (cnt > out_cap) return agc_fail(ctx, AGCGPU_EUNSUPPORTED, "determine_splitters: too many splitters");
    std::sort(out_sorted.begin(), out_sorted.end());
    out_sorted.erase(std::unique(out_sorted.begin(), out_sorted.end()), out_sorted.end());
    ctx->stats.d2h_bytes += (size_t)cnt * 8;
    return 0;
}
