// api.cu -- the C ABI of libagcgpu (include/agcgpu.h): context, memory plumbing and the entry points.
#include "internal.cuh"
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>

static thread_local std::string g_create_err;

// Device-memory pool shared by all contexts of the process: a context returns its buffers here when it is destroyed and the next
// one (another `create` in the same process, e.g. a server compressing collection after collection) takes them back instead of
// paying cudaMalloc / cudaFree -- both synchronise the device and cost tens to hundreds of milliseconds for the large tables.
#include <mutex>
namespace {
struct PoolBlock { void* p; size_t cap; };
std::mutex g_pool_mu;
std::vector<PoolBlock> g_pool[32];
const size_t kPoolMaxBytes = (size_t)64 << 30;
size_t g_pool_bytes[32];
std::vector<PoolBlock> g_pin_pool;
}
void* agc_dev_alloc(int dev, size_t bytes, size_t* cap_out)
{
    {   std::lock_guard<std::mutex> lk(g_pool_mu);
        auto& v = g_pool[dev & 31];
        int best = -1;
        for (int i = 0; i < (int)v.size(); ++i)
            if (v[i].cap >= bytes && v[i].cap <= 4 * bytes + ((size_t)1 << 20) && (best < 0 || v[i].cap < v[best].cap)) best = i;
        if (best >= 0) { PoolBlock b = v[best]; v.erase(v.begin() + best); g_pool_bytes[dev & 31] -= b.cap; *cap_out = b.cap; return b.p; }
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        agc_dev_trim(dev);                                       // give the pooled blocks back to the driver and retry once
        if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    }
    *cap_out = bytes;
    return p;
}
void agc_dev_free(int dev, void* p, size_t cap)
{
    if (!p) return;
    {   std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pool_bytes[dev & 31] + cap <= kPoolMaxBytes) { g_pool[dev & 31].push_back(PoolBlock{ p, cap }); g_pool_bytes[dev & 31] += cap; return; }
    }
    cudaFree(p);
}
void agc_dev_trim(int dev)
{
    std::lock_guard<std::mutex> lk(g_pool_mu);
    for (auto& b : g_pool[dev & 31]) cudaFree(b.p);
    g_pool[dev & 31].clear(); g_pool_bytes[dev & 31] = 0;
}

int agc_fail(agcgpu_ctx* c, int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
    if (c) c->err = buf; else g_create_err = buf;
    return code;
}

int agc_reserve(agcgpu_ctx* ctx, DevBuf& b, size_t bytes, bool keep)
{
    if (bytes <= b.cap) return 0;
    size_t ncap = std::max(bytes + bytes / 4, (size_t)4096);
    ncap = (ncap + 255) / 256 * 256;
    size_t got = 0;
    void* np = agc_dev_alloc(ctx->dev, ncap + 64, &got);
    if (!np) return agc_fail(ctx, AGCGPU_ENOMEM, "cudaMalloc(%zu) failed", ncap);
    ncap = got - 64;
    if (keep && b.p && b.cap) {
        if (cudaMemcpyAsync(np, b.p, b.cap, cudaMemcpyDeviceToDevice, ctx->st) != cudaSuccess || cudaStreamSynchronize(ctx->st) != cudaSuccess)
            return agc_fail(ctx, AGCGPU_ECUDA, "device realloc copy failed");
    } else if (b.p) cudaStreamSynchronize(ctx->st);
    if (b.p) { agc_dev_free(ctx->dev, b.p, b.cap + 64); ctx->device_bytes -= b.cap; }
    b.p = np; b.cap = ncap; ctx->device_bytes += ncap;
    return 0;
}

void* agc_arena_alloc(agcgpu_ctx* ctx, size_t bytes)
{
    bytes = (bytes + 255) / 256 * 256;
    if (bytes > ctx->arena_left) {
        size_t chunk = std::max(bytes, (size_t)256 << 20);
        size_t got = 0;
        void* p = agc_dev_alloc(ctx->dev, chunk, &got);
        if (!p) return nullptr;
        chunk = got;
        ctx->arena_chunks.emplace_back(p, chunk);
        ctx->arena_cur = (uint8_t*)p; ctx->arena_left = chunk; ctx->device_bytes += chunk;
    }
    void* r = ctx->arena_cur;
    ctx->arena_cur += bytes; ctx->arena_left -= bytes;
    return r;
}

// page-locked host memory for the caller's ingest buffers (include/agcgpu.h), pooled like the contexts' staging buffers: a file read
// straight into such a buffer goes to the device by DMA without the driver's staging copy
extern "C" void* agcgpu_host_alloc(uint64_t bytes, uint64_t* out_cap)
{
    if (!out_cap) return nullptr;
    {   std::lock_guard<std::mutex> lk(g_pool_mu);
        int best = -1;
        for (int i = 0; i < (int)g_pin_pool.size(); ++i)
            if (g_pin_pool[i].cap >= bytes && (best < 0 || g_pin_pool[i].cap < g_pin_pool[best].cap)) best = i;
        if (best >= 0) { PoolBlock b = g_pin_pool[best]; g_pin_pool.erase(g_pin_pool.begin() + best); *out_cap = b.cap; return b.p; }
    }
    void* p = nullptr;
    const size_t ncap = (size_t)bytes + (size_t)bytes / 4 + 4096;
    if (cudaMallocHost(&p, ncap) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    *out_cap = ncap;
    return p;
}
extern "C" void agcgpu_host_free(void* p, uint64_t cap)
{
    if (!p) return;
    {   std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pin_pool.size() < 6) { g_pin_pool.push_back(PoolBlock{ p, (size_t)cap }); return; }
    }
    cudaFreeHost(p);
}

int agc_pin_reserve(agcgpu_ctx* ctx, size_t bytes)
{
    if (bytes <= ctx->pin_cap) return 0;
    if (ctx->pin) cudaFreeHost(ctx->pin);
    ctx->pin = nullptr; ctx->pin_cap = 0;
    {   std::lock_guard<std::mutex> lk(g_pool_mu);              // pinned staging of a destroyed context
        for (size_t i = 0; i < g_pin_pool.size(); ++i)
            if (g_pin_pool[i].cap >= bytes) { ctx->pin = g_pin_pool[i].p; ctx->pin_cap = g_pin_pool[i].cap; g_pin_pool.erase(g_pin_pool.begin() + i); return 0; }
    }
    size_t ncap = bytes + bytes / 4 + 4096;
    if (cudaMallocHost(&ctx->pin, ncap) != cudaSuccess) { cudaGetLastError(); return agc_fail(ctx, AGCGPU_ENOMEM, "cudaMallocHost(%zu) failed", ncap); }
    ctx->pin_cap = ncap;
    return 0;
}

extern "C" {

int agcgpu_create(const agcgpu_params* p, agcgpu_ctx** out)
{
    if (!p || !out) return agc_fail(nullptr, AGCGPU_EINVAL, "null argument");
    if (p->kmer_length < 17 || p->kmer_length > 32 || p->min_match_len < 15 || p->min_match_len > 32)
        return agc_fail(nullptr, AGCGPU_EINVAL, "k must be in [17,32] and min_match_len in [15,32] (src/app/application.h:24-84)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return agc_fail(nullptr, AGCGPU_ENODEV, "no CUDA device: libagcgpu has no CPU fallback");
    }
    if (p->device < 0 || p->device >= ndev) return agc_fail(nullptr, AGCGPU_EINVAL, "device %d out of range (%d devices)", p->device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, p->device) != cudaSuccess) return agc_fail(nullptr, AGCGPU_ECUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return agc_fail(nullptr, AGCGPU_ENODEV, "device %d is sm_%d%d; this library is built for sm_100a only", p->device, prop.major, prop.minor);
    if (cudaSetDevice(p->device) != cudaSuccess) return agc_fail(nullptr, AGCGPU_ECUDA, "cudaSetDevice failed");
    agcgpu_ctx* ctx = new agcgpu_ctx();
    ctx->prm = *p; ctx->dev = p->device; ctx->n_sm = prop.multiProcessorCount;
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    if (cudaStreamCreateWithFlags(&ctx->st, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
        delete ctx; return agc_fail(nullptr, AGCGPU_ECUDA, "stream/event creation failed");
    }
    // map_segments[(~0,~0)] = 0 (CAGCCompressor::Create, agc_compressor.cpp:2307)
    ctx->h_map_k1.push_back(~0ULL); ctx->h_map_k2.push_back(~0ULL); ctx->h_map_val.push_back(0);
    *out = ctx;
    return 0;
}

void agcgpu_destroy(agcgpu_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->dev);
    cudaStreamSynchronize(ctx->st);
    agc_zstd_waves_drop(ctx);
    DevBuf* bufs[] = { &ctx->spl_keys, &ctx->spl_filter, &ctx->raw, &ctx->packed, &ctx->exc_pos, &ctx->exc_code, &ctx->tile_desc,
                       &ctx->tile_cnt, &ctx->tile_base, &ctx->d_cstart, &ctx->chunk_prefix, &ctx->hits, &ctx->counters, &ctx->map_k1,
                       &ctx->map_k2, &ctx->map_val, &ctx->d_groups, &ctx->scr_req, &ctx->scr_units, &ctx->scr_out, &ctx->scr_sizes,
                       &ctx->scr_offs, &ctx->scr_dense, &ctx->scr_misc, &ctx->scr_bytes, &ctx->scr_chunk, &ctx->scr_rec, &ctx->scr_gsz, &ctx->scr_gather, &ctx->scr_zkeep, &ctx->scr_cost, &ctx->ref_kmers };
    for (DevBuf* b : bufs) if (b->p) agc_dev_free(ctx->dev, b->p, b->cap + 64);
    for (auto& c : ctx->arena_chunks) agc_dev_free(ctx->dev, c.first, c.second);
    if (ctx->pin) {
        std::lock_guard<std::mutex> lk(g_pool_mu);
        if (g_pin_pool.size() < 4) { g_pin_pool.push_back(PoolBlock{ ctx->pin, ctx->pin_cap }); ctx->pin = nullptr; }
    }
    if (ctx->pin) cudaFreeHost(ctx->pin);
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    if (ctx->st2) { cudaStreamSynchronize(ctx->st2); cudaStreamDestroy(ctx->st2); cudaEventDestroy(ctx->ev2); }
    cudaStreamDestroy(ctx->st);
    delete ctx;
}

const char* agcgpu_last_error(const agcgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int agcgpu_sync(agcgpu_ctx* ctx)
{
    if (!ctx) return AGCGPU_EINVAL;
    CK(cudaStreamSynchronize(ctx->st));
    return 0;
}

void* agcgpu_stream(agcgpu_ctx* ctx) { return ctx ? (void*)ctx->st : nullptr; }

int agcgpu_get_stats(agcgpu_ctx* ctx, agcgpu_stats* out)
{
    if (!ctx || !out) return AGCGPU_EINVAL;
    ctx->stats.device_bytes_in_use = ctx->device_bytes;
    *out = ctx->stats;
    return 0;
}

int agcgpu_set_splitters(agcgpu_ctx* ctx, const uint64_t* s, uint64_t n)
{
    if (!ctx || (n && !s)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    return agc_upload_splitters(ctx, s, n);
}

// compress_contig's sequential part on the sparse hit list (agc_compressor.cpp:2019-2048): a hit is taken only if the
// rolling k-mer was refilled since the previous accepted hit (kmer.Reset() at 2031 => next full k-mer ends >= k later).
static void resolve_cuts(agcgpu_ctx* ctx, const std::vector<ScanHit>& hits, std::vector<agcgpu_cut>& cuts)
{
    const uint32_t k = ctx->prm.kmer_length;
    size_t h = 0;
    for (uint32_t c = 0; c < ctx->n_contigs; ++c) {
        uint64_t len = ctx->h_cstart[c + 1] - ctx->h_cstart[c];
        uint64_t split_pos = 0; bool have_front = false; uint64_t fdir = 0, frc = 0;
        int64_t last = -(int64_t)k;
        for (; h < hits.size() && hits[h].contig == c; ++h) {
            const ScanHit& x = hits[h];
            if ((int64_t)x.pos - last < (int64_t)k) continue;
            agcgpu_cut cut; memset(&cut, 0, sizeof cut);
            cut.contig = c; cut.start = split_pos; cut.len = x.pos + 1 - split_pos;
            cut.has_front = have_front; cut.front_dir = fdir; cut.front_rc = frc;
            cut.has_back = 1; cut.back_dir = x.dir; cut.back_rc = x.rc;
            cuts.push_back(cut);
            split_pos = x.pos + 1 - k; have_front = true; fdir = x.dir; frc = x.rc; last = (int64_t)x.pos;
        }
        if (split_pos < len) {
            agcgpu_cut cut; memset(&cut, 0, sizeof cut);
            cut.contig = c; cut.start = split_pos; cut.len = len - split_pos;
            cut.has_front = have_front; cut.front_dir = fdir; cut.front_rc = frc;
            cuts.push_back(cut);
        }
    }
}

static int scan_common(agcgpu_ctx* ctx, const uint8_t* raw_dev, uint64_t raw_bytes, const uint64_t* raw_offsets, uint32_t n_contigs,
                       uint64_t* out_contig_len, agcgpu_cut* out_cuts, uint64_t cap_cuts, uint64_t* out_n_cuts)
{
    std::vector<ScanHit> hits;
    if (int r = agc_prep_and_scan(ctx, raw_dev, raw_bytes, raw_offsets, n_contigs, true, &hits)) return r;
    if (out_contig_len) for (uint32_t c = 0; c < n_contigs; ++c) out_contig_len[c] = ctx->h_cstart[c + 1] - ctx->h_cstart[c];
    std::vector<agcgpu_cut> cuts;
    resolve_cuts(ctx, hits, cuts);
    if (out_n_cuts) *out_n_cuts = cuts.size();
    if (cuts.size() > cap_cuts) return agc_fail(ctx, AGCGPU_EOVERFLOW, "scan: %zu cuts, caller buffer holds %llu", cuts.size(), (unsigned long long)cap_cuts);
    if (!cuts.empty()) memcpy(out_cuts, cuts.data(), cuts.size() * sizeof(agcgpu_cut));
    return 0;
}

static int upload_raw(agcgpu_ctx* ctx, const uint8_t* raw, uint64_t raw_bytes)
{
    if (int r = agc_reserve(ctx, ctx->raw, raw_bytes + 64)) return r;
    if (raw_bytes) {
        CK(cudaMemcpyAsync(ctx->raw.p, raw, raw_bytes, cudaMemcpyHostToDevice, ctx->st));
        ctx->stats.h2d_bytes += raw_bytes;
    }
    return 0;
}

int agcgpu_scan_contigs(agcgpu_ctx* ctx, const uint8_t* raw, const uint64_t* raw_offsets, uint32_t n_contigs,
                        uint64_t* out_contig_len, agcgpu_cut* out_cuts, uint64_t cap_cuts, uint64_t* out_n_cuts)
{
    if (!ctx || !raw_offsets || (n_contigs && !raw && raw_offsets[n_contigs])) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    uint64_t raw_bytes = raw_offsets[n_contigs];
    if (int r = upload_raw(ctx, raw, raw_bytes)) return r;
    return scan_common(ctx, (const uint8_t*)ctx->raw.p, raw_bytes, raw_offsets, n_contigs, out_contig_len, out_cuts, cap_cuts, out_n_cuts);
}

int agcgpu_scan_contigs_dev(agcgpu_ctx* ctx, const void* raw_dev, uint64_t raw_bytes, const uint64_t* raw_offsets, uint32_t n_contigs,
                            uint64_t* out_contig_len, agcgpu_cut* out_cuts, uint64_t cap_cuts, uint64_t* out_n_cuts)
{
    if (!ctx || !raw_offsets || !raw_dev) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    return scan_common(ctx, (const uint8_t*)raw_dev, raw_bytes, raw_offsets, n_contigs, out_contig_len, out_cuts, cap_cuts, out_n_cuts);
}

int agcgpu_determine_splitters(agcgpu_ctx* ctx, const uint8_t* raw, const uint64_t* raw_offsets, uint32_t n_contigs,
                               uint64_t* out_splitters, uint64_t cap, uint64_t* out_n)
{
    if (!ctx || !raw_offsets || !out_n) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    uint64_t raw_bytes = raw_offsets[n_contigs];
    if (int r = upload_raw(ctx, raw, raw_bytes)) return r;
    if (int r = agc_prep_and_scan(ctx, (const uint8_t*)ctx->raw.p, raw_bytes, raw_offsets, n_contigs, false, nullptr)) return r;
    std::vector<uint64_t> spl;
    ctx->h_last_spl.clear();
    if (int r = agc_enumerate_splitters(ctx, 0, ctx->n_contigs, false, (ctx->prm.flags & AGCGPU_F_ADAPTIVE) != 0, spl)) return r;
    *out_n = spl.size();
    if (spl.size() > cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "determine_splitters: %zu splitters, buffer holds %llu", spl.size(), (unsigned long long)cap);
    if (!spl.empty()) memcpy(out_splitters, spl.data(), spl.size() * 8);
    return agc_upload_splitters(ctx, spl.data(), spl.size());
}

int agcgpu_find_new_splitters(agcgpu_ctx* ctx, const uint32_t* contigs, uint32_t n, uint64_t* out_splitters, uint64_t cap, uint64_t* out_n)
{
    if (!ctx || !out_n || (n && !contigs)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    if (!(ctx->prm.flags & AGCGPU_F_ADAPTIVE)) return agc_fail(ctx, AGCGPU_EINVAL, "find_new_splitters needs AGCGPU_F_ADAPTIVE");
    std::vector<uint64_t> all, one;
    ctx->h_last_spl.clear();
    for (uint32_t i = 0; i < n; ++i) {                   // candidates are per contig (its own singletons), so one pass each
        if (contigs[i] >= ctx->n_contigs) return agc_fail(ctx, AGCGPU_EINVAL, "find_new_splitters: contig %u is not resident", contigs[i]);
        if (int r = agc_enumerate_splitters(ctx, contigs[i], 1, true, false, one)) return r;
        all.insert(all.end(), one.begin(), one.end());
    }
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    *out_n = all.size();
    if (all.size() > cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "find_new_splitters: %zu splitters, buffer holds %llu", all.size(), (unsigned long long)cap);
    if (!all.empty()) memcpy(out_splitters, all.data(), all.size() * 8);
    return 0;
}

int agcgpu_filtered_kmers(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint64_t threshold, agcgpu_fkmer* out, uint64_t cap,
                          uint64_t* out_offsets)
{
    if (!ctx || !out_offsets || (n && !reqs) || (cap && !out)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    out_offsets[0] = 0;
    std::vector<agcgpu_fkmer> one;
    bool overflow = false;
    for (uint32_t i = 0; i < n; ++i) {
        const agcgpu_seg_req& q = reqs[i];
        if (q.contig >= ctx->n_contigs || q.start > ctx->h_cstart[q.contig + 1] - ctx->h_cstart[q.contig])
            return agc_fail(ctx, AGCGPU_EINVAL, "filtered_kmers: range %u is outside the resident batch", i);
        const uint64_t len = std::min<uint64_t>(q.len, ctx->h_cstart[q.contig + 1] - ctx->h_cstart[q.contig] - q.start);   // clipped like get_part
        if (int r = agc_filtered_kmers(ctx, ctx->h_cstart[q.contig] + q.start, len, threshold, one)) return r;
        uint64_t o = out_offsets[i];
        out_offsets[i + 1] = o + one.size();
        if (out_offsets[i + 1] > cap) overflow = true;
        else if (!one.empty()) memcpy(out + o, one.data(), one.size() * sizeof(agcgpu_fkmer));
    }
    if (overflow) return agc_fail(ctx, AGCGPU_EOVERFLOW, "filtered_kmers: %llu k-mers, buffer holds %llu", (unsigned long long)out_offsets[n], (unsigned long long)cap);
    return 0;
}

int agcgpu_last_splitter_positions(agcgpu_ctx* ctx, uint32_t* out_contig, uint64_t* out_pos, uint64_t* out_kmer, uint8_t* out_is_last,
                                   uint64_t cap, uint64_t* out_n)
{
    if (!ctx || !out_n) return AGCGPU_EINVAL;
    auto v = ctx->h_last_spl;
    std::sort(v.begin(), v.end(), [](const agcgpu_ctx::SplFound& a, const agcgpu_ctx::SplFound& b) { return a.contig != b.contig ? a.contig < b.contig : a.pos < b.pos; });
    *out_n = v.size();
    if (v.size() > cap) return agc_fail(ctx, AGCGPU_EOVERFLOW, "last_splitter_positions: %zu entries, buffer holds %llu", v.size(), (unsigned long long)cap);
    for (size_t i = 0; i < v.size(); ++i) { out_contig[i] = v[i].contig; out_pos[i] = v[i].pos; out_kmer[i] = v[i].kmer; out_is_last[i] = v[i].is_last; }
    return 0;
}

int agcgpu_rescan_contigs(agcgpu_ctx* ctx, agcgpu_cut* out_cuts, uint64_t cap_cuts, uint64_t* out_n_cuts)
{
    if (!ctx || (cap_cuts && !out_cuts)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    std::vector<ScanHit> hits;
    if (ctx->total_bases) { if (int r = agc_scan_resident(ctx, &hits)) return r; }
    std::vector<agcgpu_cut> cuts;
    resolve_cuts(ctx, hits, cuts);
    if (out_n_cuts) *out_n_cuts = cuts.size();
    if (cuts.size() > cap_cuts) return agc_fail(ctx, AGCGPU_EOVERFLOW, "rescan: %zu cuts, caller buffer holds %llu", cuts.size(), (unsigned long long)cap_cuts);
    if (!cuts.empty()) memcpy(out_cuts, cuts.data(), cuts.size() * sizeof(agcgpu_cut));
    return 0;
}

int agcgpu_get_segment(agcgpu_ctx* ctx, uint32_t contig, uint64_t start, uint32_t len, uint32_t is_rc, uint8_t* out)
{
    if (!ctx || (len && !out)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    if (contig >= ctx->n_contigs) return agc_fail(ctx, AGCGPU_EINVAL, "get_segment: contig %u not resident", contig);
    if (start + len > ctx->h_cstart[contig + 1] - ctx->h_cstart[contig]) return agc_fail(ctx, AGCGPU_EINVAL, "get_segment: range outside contig");
    if (len == 0) return 0;
    if (int r = agc_reserve(ctx, ctx->scr_bytes, (size_t)len + 64)) return r;
    if (int r = agc_expand_segment(ctx, ctx->h_cstart[contig] + start, len, is_rc, (uint8_t*)ctx->scr_bytes.p, 0)) return r;
    CK(cudaMemcpyAsync(out, ctx->scr_bytes.p, len, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.d2h_bytes += len;
    return 0;
}

int agcgpu_map_insert(agcgpu_ctx* ctx, const uint64_t* k1, const uint64_t* k2, const int32_t* gid, uint64_t n)
{
    if (!ctx || (n && (!k1 || !k2 || !gid))) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    for (uint64_t i = 0; i < n; ++i) {
        if (gid[i] < 0) return agc_fail(ctx, AGCGPU_EINVAL, "map_insert: negative group id");
        ctx->h_map_k1.push_back(k1[i]); ctx->h_map_k2.push_back(k2[i]); ctx->h_map_val.push_back(gid[i]);
    }
    return agc_map_update(ctx);
}

int agcgpu_assign_cuts(agcgpu_ctx* ctx, const agcgpu_cut* cuts, uint64_t n, agcgpu_assign* out)
{
    if (!ctx || (n && (!cuts || !out))) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    return agc_assign_launch(ctx, cuts, n, out);
}

int agcgpu_group_put_reference_batch(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n)
{
    if (!ctx || (n && !reqs)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    return agc_refs_from_segments(ctx, reqs, n);
}

int agcgpu_group_put_reference(agcgpu_ctx* ctx, uint32_t group_id, const uint8_t* symbols, uint32_t len)
{
    if (!ctx || (len && !symbols)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    return agc_ref_from_host(ctx, group_id, symbols, len);
}

int agcgpu_group_get_index(agcgpu_ctx* ctx, uint32_t group_id, uint32_t* out_slots, uint64_t cap, uint64_t* out_ht_size)
{
    if (!ctx || !out_ht_size) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    if (group_id >= ctx->h_groups.size() || !(ctx->h_groups[group_id].flags & GRF_PRESENT))
        return agc_fail(ctx, AGCGPU_EINVAL, "get_index: group %u has no reference", group_id);
    const GroupRefDev& g = ctx->h_groups[group_id];
    *out_ht_size = g.ht_size;
    if (!out_slots) return 0;
    if (cap < g.ht_size) return agc_fail(ctx, AGCGPU_EOVERFLOW, "get_index: need %u slots", g.ht_size);
    if (g.flags & GRF_SHORT) {
        std::vector<uint16_t> t(g.ht_size);
        CK(cudaMemcpy(t.data(), g.ht, (size_t)g.ht_size * 2, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < g.ht_size; ++i) out_slots[i] = t[i] == 0xffffu ? 0xffffffffu : t[i];
    } else CK(cudaMemcpy(out_slots, g.ht, (size_t)g.ht_size * 4, cudaMemcpyDeviceToHost));
    return 0;
}

int agcgpu_lz_encode_batch(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_offsets)
{
    if (!ctx || !out_offsets || (n && (!reqs || !out))) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    return agc_lz_run(ctx, 0, reqs, n, 0, out, out_cap, out_offsets, nullptr);
}

int agcgpu_lz_encode_batch_sharded(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_offsets)
{
    if (!ctx || !out_offsets || (n && (!reqs || !out))) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    if (!agc_comm_active() || n == 0) return agc_lz_run(ctx, 0, reqs, n, 0, out, out_cap, out_offsets, nullptr);
    return agc_lz_encode_sharded(ctx, reqs, n, out, out_cap, out_offsets);
}

int agcgpu_debug_lz_chunk_records(agcgpu_ctx* ctx, void* out, uint64_t cap_bytes, uint64_t* out_n_records)
{
    if (!ctx || !out_n_records) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    *out_n_records = ctx->last_lzc_chunks;
    const uint64_t need = ctx->last_lzc_chunks * 64;
    if (!out || cap_bytes < need) return need ? AGCGPU_EOVERFLOW : 0;
    if (need) CK(cudaMemcpy(out, ctx->scr_rec.p, need, cudaMemcpyDeviceToHost));
    return 0;
}

int agcgpu_comm_world(void) { return agc_comm_active() ? (int)agc_comm_world() : 1; }
int agcgpu_comm_rank(void) { return agc_comm_active() ? (int)agc_comm_rank() : 0; }

int agcgpu_lz_estimate_batch(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint32_t* out)
{
    if (!ctx || (n && (!reqs || !out))) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    return agc_lz_run(ctx, 1, reqs, n, 0, nullptr, 0, nullptr, out);
}

int agcgpu_lz_cost_vector(agcgpu_ctx* ctx, const agcgpu_seg_req* req, int prefix_costs, uint32_t* out)
{
    if (!ctx || !req || (req->len && !out)) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    if (req->len == 0) return 0;
    agcgpu_seg_req q = *req;
    q.bound = prefix_costs ? 1u : 0u;                    // the cost-vector kernels take prefix_costs per request
    return agc_lz_run(ctx, 2, &q, 1, prefix_costs, nullptr, 0, nullptr, out);
}

int agcgpu_lz_cost_split_batch(agcgpu_ctx* ctx, const agcgpu_split_req* reqs, uint32_t n, uint32_t* out_best_pos, uint32_t* out_best_sum)
{
    if (!ctx || (n && (!reqs || !out_best_pos || !out_best_sum))) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    if (n == 0) return 0;
    return agc_lz_cost_split(ctx, reqs, n, out_best_pos, out_best_sum);
}

int agcgpu_pack_ref_batch(agcgpu_ctx* ctx, const uint32_t* group_ids, uint32_t n, uint8_t* out, uint64_t out_cap,
                          uint64_t* out_offsets, uint8_t* out_use_tuples)
{
    if (!ctx || !out_offsets || (n && (!group_ids || !out || !out_use_tuples))) return AGCGPU_EINVAL;
    cudaSetDevice(ctx->dev);
    return agc_pack_refs(ctx, group_ids, n, out, out_cap, out_offsets, out_use_tuples);
}

}  // extern "C"
