// kernels_prep.cu -- ingest + splitter scan + hash-assign kernels (sm_100a).
//
//   k_tile_count / k_tile_scan / k_tile_compact : preprocess_raw_contig (src/core/agc_compressor.cpp:907-951) for a whole
//        batch: drop bytes < 64, map ASCII -> symbol code (cnv_num, src/common/agc_basic.h:39-49), write the symbols
//        2-bit packed (4 per byte, first base in the top bits) + a sorted exception list for every non-ACGT symbol.
//   k_scan : compress_contig's per-base loop (agc_compressor.cpp:2009-2034): rolling canonical k-mer (src/core/kmer.h
//        284-301) + splitter-set membership (bloom_set_t/hash_set_lp in the reference; a shared-memory one-hash bitmap
//        in front of an exact open-addressing set here -- the set layout is not observable).
//   k_assign : add_segment's key construction + map_segments.find (agc_compressor.cpp:1287-1313,1363).
//
// All integer/byte work, HBM-bound: no tensor cores.  Loads of the raw FASTA are 128-bit and coalesced.
#include "internal.cuh"
#include <algorithm>
#include <cstring>

// cnv_num[64..127] (agc_basic.h:39-49); bytes >= 128 are outside the reference's table (mapped to 30)
__constant__ uint8_t c_cnv[64] = {
    32,  0, 11,  1, 12, 30, 30,  2, 13, 30, 30,  9, 30, 10,  4, 30,
    30, 30,  5,  7,  3, 15, 14,  8, 30,  6, 30, 30, 30, 30, 30, 30,
    32,  0, 11,  1, 12, 30, 30,  2, 13, 30, 30,  9, 30, 10,  4, 30,
    30, 30,  5,  7,  3, 15, 14,  8, 30,  6, 30, 30, 30, 30, 30, 30 };

struct TileDesc {          // one preprocessing tile: 512 aligned 16-byte chunks, masked to [lo, hi) raw bytes
    uint64_t chunk0;       // index of the first 16-byte chunk (raw_dev + 16*chunk0)
    uint64_t lo, hi;       // raw byte range of the owning contig that this tile may touch
};

__device__ __forceinline__ uint32_t sym_of(uint32_t c)
{
    // ACGT/acgt fast path: x = (c>>1)&3 gives A0 C1 G3 T2; x ^ (x>>1) -> 0 1 2 3
    uint32_t u = c & 0xDFu;
    if (u == 'A' || u == 'C' || u == 'G' || u == 'T') { uint32_t x = (c >> 1) & 3u; return x ^ (x >> 1); }
    return c < 128u ? c_cnv[c - 64u] : 30u;
}

// per-thread: 32 consecutive raw bytes (two aligned 16-byte chunks); returns kept count (low 16) | exception count (high 16)
__device__ __forceinline__ uint32_t tile_thread_load(const uint8_t* __restrict__ raw, const TileDesc& td, uint32_t t,
                                                     uint32_t (&w)[8], uint64_t& byte0)
{
    byte0 = (td.chunk0 + 2ull * t) * 16ull;
    const uint4* p = reinterpret_cast<const uint4*>(raw + byte0);
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (byte0 < td.hi && byte0 + 16 > td.lo) a = __ldg(p);
    if (byte0 + 16 < td.hi && byte0 + 32 > td.lo) b = __ldg(p + 1);
    w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        uint32_t c = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
        uint64_t pos = byte0 + j;
        bool in = pos >= td.lo && pos < td.hi;
        if (in && c >= 64u) {
            uint32_t u = c & 0xDFu;
            bool acgt = (u == 'A' || u == 'C' || u == 'G' || u == 'T');
            cnt += acgt ? 1u : 0x10001u;
        }
    }
    return cnt;
}

__global__ void __launch_bounds__(256) k_tile_count(const uint8_t* __restrict__ raw, const TileDesc* __restrict__ tiles,
                                                    uint32_t n_tiles, uint32_t* __restrict__ tile_cnt)
{
    __shared__ uint32_t s_part[8];
    uint32_t tile = blockIdx.x;
    if (tile >= n_tiles) return;
    TileDesc td = tiles[tile];
    uint32_t w[8]; uint64_t b0;
    uint32_t cnt = tile_thread_load(raw, td, threadIdx.x, w, b0);
#pragma unroll
    for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t s = 0;
        for (int i = 0; i < 8; ++i) s += s_part[i];
        tile_cnt[tile] = s;
    }
}

// single-block exclusive scan of the (kept | exc<<16) tile counters into two u64 base arrays; totals in base[n]
__global__ void __launch_bounds__(1024) k_tile_scan(const uint32_t* __restrict__ tile_cnt, uint32_t n_tiles,
                                                    uint64_t* __restrict__ base_kept, uint64_t* __restrict__ base_exc)
{
    __shared__ uint64_t s_k[32], s_e[32];
    __shared__ uint64_t carry_k, carry_e;
    if (threadIdx.x == 0) { carry_k = 0; carry_e = 0; }
    __syncthreads();
    for (uint32_t start = 0; start < n_tiles; start += 1024) {
        uint32_t i = start + threadIdx.x;
        uint32_t v = i < n_tiles ? tile_cnt[i] : 0;
        uint64_t k = v & 0xffffu, e = v >> 16;
        uint64_t ik = k, ie = e;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint64_t tk = __shfl_up_sync(0xffffffffu, ik, o), te = __shfl_up_sync(0xffffffffu, ie, o);
            if ((threadIdx.x & 31) >= o) { ik += tk; ie += te; }
        }
        if ((threadIdx.x & 31) == 31) { s_k[threadIdx.x >> 5] = ik; s_e[threadIdx.x >> 5] = ie; }
        __syncthreads();
        if (threadIdx.x < 32) {
            uint64_t wk = s_k[threadIdx.x], we = s_e[threadIdx.x];
            uint64_t xk = wk, xe = we;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint64_t tk = __shfl_up_sync(0xffffffffu, xk, o), te = __shfl_up_sync(0xffffffffu, xe, o);
                if (threadIdx.x >= o) { xk += tk; xe += te; }
            }
            s_k[threadIdx.x] = xk - wk; s_e[threadIdx.x] = xe - we;   // exclusive warp offsets
        }
        __syncthreads();
        uint64_t ok = carry_k + s_k[threadIdx.x >> 5] + ik - k;
        uint64_t oe = carry_e + s_e[threadIdx.x >> 5] + ie - e;
        if (i < n_tiles) { base_kept[i] = ok; base_exc[i] = oe; }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_k = ok + k; carry_e = oe + e; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { base_kept[n_tiles] = carry_k; base_exc[n_tiles] = carry_e; }
}

__global__ void __launch_bounds__(256) k_tile_compact(const uint8_t* __restrict__ raw, const TileDesc* __restrict__ tiles,
                                                      uint32_t n_tiles, const uint64_t* __restrict__ base_kept,
                                                      const uint64_t* __restrict__ base_exc, uint32_t* __restrict__ packed,
                                                      uint64_t* __restrict__ exc_pos, uint8_t* __restrict__ exc_code)
{
    __shared__ uint8_t s_codes[AGC_TILE_BYTES + 32];
    __shared__ uint32_t s_warp[8];
    uint32_t tile = blockIdx.x;
    if (tile >= n_tiles) return;
    TileDesc td = tiles[tile];
    uint64_t ob = base_kept[tile], eb = base_exc[tile];
    uint32_t total = (uint32_t)(base_kept[tile + 1] - ob);
    uint32_t w[8]; uint64_t b0;
    uint32_t cnt = tile_thread_load(raw, td, threadIdx.x, w, b0);
    // block exclusive scan of cnt (both 16-bit fields at once; totals <= 8192 each)
    uint32_t inc = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, o); if ((threadIdx.x & 31) >= o) inc += t; }
    if ((threadIdx.x & 31) == 31) s_warp[threadIdx.x >> 5] = inc;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t i = 0; i < (threadIdx.x >> 5); ++i) woff += s_warp[i];
    uint32_t excl = woff + inc - cnt;
    uint32_t k = excl & 0xffffu, e = excl >> 16;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        uint32_t c = (w[j >> 2] >> (8 * (j & 3))) & 0xffu;
        uint64_t pos = b0 + j;
        if (pos >= td.lo && pos < td.hi && c >= 64u) {
            uint32_t s = sym_of(c);
            s_codes[k] = (uint8_t)s;
            if (s > 3u) { exc_pos[eb + e] = ob + k; exc_code[eb + e] = (uint8_t)s; ++e; }
            ++k;
        }
    }
    __syncthreads();
    if (total == 0) return;
    // pack: output words (16 bases each); boundary words are shared with neighbouring tiles -> atomicOr into zeroed memory
    uint64_t w0 = ob >> 4, w1 = (ob + total - 1) >> 4;
    for (uint64_t wi = w0 + threadIdx.x; wi <= w1; wi += 256) {
        uint32_t v = 0;
        bool partial = false;
#pragma unroll
        for (int b = 0; b < 16; ++b) {
            uint64_t g = wi * 16 + b;
            if (g >= ob && g < ob + total) {
                uint32_t s = s_codes[(uint32_t)(g - ob)];
                if (s > 3u) s = 0;
                v |= s << (8 * (b >> 2) + 6 - 2 * (b & 3));
            } else partial = true;
        }
        if (partial) atomicOr(&packed[wi], v); else packed[wi] = v;
    }
}

// ------------------------------------------------------------------------------------------------ k-mer scan
__device__ __forceinline__ bool exc_in_range(const uint64_t* __restrict__ exc_pos, uint64_t n_exc, uint64_t lo, uint64_t hi)
{   // any exception position in [lo, hi] ?
    if (n_exc == 0) return false;
    uint64_t a = 0, b = n_exc;
    while (a < b) { uint64_t mid = (a + b) >> 1; if (exc_pos[mid] < lo) a = mid + 1; else b = mid; }
    return a < n_exc && exc_pos[a] <= hi;
}

__device__ __forceinline__ bool splitter_lookup(const uint64_t* __restrict__ keys, uint64_t mask, uint64_t canon, uint64_t h)
{
    uint64_t slot = (h >> 24) & mask;
    while (true) {
        uint64_t kx = __ldg(&keys[slot]);
        if (kx == canon) return true;
        if (kx == ~0ULL) return false;
        slot = (slot + 1) & mask;
    }
}

__global__ void __launch_bounds__(AGC_SCAN_THREADS) k_scan(
    const uint64_t* __restrict__ P, const uint64_t* __restrict__ cstart, uint32_t n_contigs,
    const uint32_t* __restrict__ chunk_prefix, uint32_t total_chunks, uint32_t k,
    const uint64_t* __restrict__ spl_keys, uint64_t spl_mask, const uint32_t* __restrict__ filter, uint32_t filter_log2,
    const uint64_t* __restrict__ exc_pos, uint64_t n_exc, ScanHit* __restrict__ hits, uint32_t* __restrict__ hit_count,
    uint32_t hit_cap)
{
    extern __shared__ uint32_t s_filter[];
    const uint32_t fwords = 1u << (filter_log2 - 5);
    for (uint32_t i = threadIdx.x; i < fwords; i += blockDim.x) s_filter[i] = filter[i];
    __syncthreads();
    const uint64_t fmask = (1ull << filter_log2) - 1;
    const uint32_t shift = 64 - 2 * k;
    const uint64_t kmask = (~0ULL) << shift;

    for (uint32_t u = blockIdx.x; u < total_chunks; u += gridDim.x) {
        // contig of this chunk: last c with chunk_prefix[c] <= u
        uint32_t lo = 0, hi = n_contigs;
        while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (chunk_prefix[mid] <= u) lo = mid; else hi = mid; }
        uint32_t c = lo;
        uint64_t base = cstart[c], len = cstart[c + 1] - base;
        uint64_t p0 = (uint64_t)(u - chunk_prefix[c]) * AGC_SCAN_CHUNK + (uint64_t)threadIdx.x * 32;
        uint64_t pend = p0 + 32 < len ? p0 + 32 : len;
        uint64_t ps = p0 > (uint64_t)(k - 1) ? p0 : (uint64_t)(k - 1);
        if (ps >= pend) continue;
        uint64_t dir = agc_win(P, base + ps - (k - 1)) & kmask;
        uint64_t rc = (~agc_rev2(dir)) << shift;
        uint64_t nxt = agc_win(P, base + ps + 1);
        for (uint64_t p = ps; p < pend; ++p) {
            uint64_t canon = dir < rc ? dir : rc;
            uint64_t h = agc_murmur64(canon);
            uint64_t bit = h & fmask;
            if ((s_filter[bit >> 5] >> (bit & 31)) & 1u) {
                if (splitter_lookup(spl_keys, spl_mask, canon, h) &&
                    !exc_in_range(exc_pos, n_exc, base + p - (k - 1), base + p)) {
                    uint32_t idx = atomicAdd(hit_count, 1u);
                    if (idx < hit_cap) { ScanHit hh; hh.pos = p; hh.dir = dir; hh.rc = rc; hh.contig = c; hh.pad = 0; hits[idx] = hh; }
                }
            }
            uint64_t s = nxt >> 62; nxt <<= 2;
            dir = ((dir << 2) | (s << shift)) & kmask;
            rc = ((rc >> 2) | ((3 - s) << 62)) & kmask;
        }
    }
}

// ------------------------------------------------------------------------------------------------ segment expansion
// symbols (1 byte each) of segment [gstart, gstart+n) (optionally reverse complemented) + pad bytes of 31
__global__ void k_expand(const uint64_t* __restrict__ P, uint64_t gstart, uint32_t n, uint32_t is_rc,
                         uint8_t* __restrict__ dst, uint32_t pad)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        uint64_t g = is_rc ? gstart + (n - 1 - i) : gstart + i;
        uint32_t s = (uint32_t)(agc_win(P, g) >> 62);
        dst[i] = (uint8_t)(is_rc ? 3 - s : s);
    } else if (i < n + pad) dst[i] = 31;
}
__global__ void k_patch_exc(const uint64_t* __restrict__ exc_pos, const uint8_t* __restrict__ exc_code, uint64_t e0, uint64_t e1,
                            uint64_t gstart, uint32_t n, uint32_t is_rc, uint8_t* __restrict__ dst)
{
    uint64_t i = e0 + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= e1) return;
    uint64_t q = exc_pos[i] - gstart;
    dst[is_rc ? (n - 1 - q) : q] = exc_code[i];      // non-ACGT symbols are their own complement (agc_basic.cpp:280-316)
}

// ------------------------------------------------------------------------------------------------ hash-assign
__global__ void k_assign(const agcgpu_cut* __restrict__ cuts, uint64_t n, const uint64_t* __restrict__ mk1,
                         const uint64_t* __restrict__ mk2, const int32_t* __restrict__ mval, uint64_t mmask,
                         agcgpu_assign* __restrict__ out)
{
    uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    agcgpu_cut c = cuts[i];
    agcgpu_assign a; a.reserved = 0;
    uint64_t fc = c.front_dir < c.front_rc ? c.front_dir : c.front_rc;
    uint64_t bc = c.back_dir < c.back_rc ? c.back_dir : c.back_rc;
    a.is_rc = 0; a.group_id = -1;
    if (c.has_front && c.has_back) {
        a.klass = 0;
        if (fc < bc) { a.key1 = fc; a.key2 = bc; } else { a.key1 = bc; a.key2 = fc; a.is_rc = 1; }
    } else if (c.has_front) { a.klass = 1; a.key1 = fc; a.key2 = ~0ULL; }
    else if (c.has_back) { a.klass = 2; a.key1 = ~0ULL; a.key2 = bc; }
    else { a.klass = 3; a.key1 = a.key2 = ~0ULL; }
    if (a.klass == 0 || a.klass == 3) {
        uint64_t slot = agc_murmur_pair(a.key1, a.key2) & mmask;
        while (true) {
            int32_t v = mval[slot];
            if (v < 0) break;
            if (mk1[slot] == a.key1 && mk2[slot] == a.key2) { a.group_id = v; break; }
            slot = (slot + 1) & mmask;
        }
    }
    out[i] = a;
}

// ================================================================================================ host side
int agc_upload_splitters(agcgpu_ctx* ctx, const uint64_t* s, uint64_t n)
{
    uint64_t cap = 64;
    while (cap < 4 * n + 16) cap <<= 1;
    uint32_t fl = 10;
    while (fl < 19 && (1ull << fl) < 16 * n) ++fl;
    std::vector<uint64_t> keys(cap, ~0ULL);
    std::vector<uint32_t> filt((size_t)1 << (fl - 5), 0u);
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t h = agc_murmur64(s[i]);
        uint64_t bit = h & ((1ull << fl) - 1);
        filt[bit >> 5] |= 1u << (bit & 31);
        uint64_t slot = (h >> 24) & (cap - 1);
        while (keys[slot] != ~0ULL && keys[slot] != s[i]) slot = (slot + 1) & (cap - 1);
        keys[slot] = s[i];
    }
    if (int r = agc_reserve(ctx, ctx->spl_keys, cap * 8)) return r;
    if (int r = agc_reserve(ctx, ctx->spl_filter, filt.size() * 4)) return r;
    CK(cudaMemcpyAsync(ctx->spl_keys.p, keys.data(), cap * 8, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->spl_filter.p, filt.data(), filt.size() * 4, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.h2d_bytes += cap * 8 + filt.size() * 4;
    ctx->spl_mask = cap - 1; ctx->filter_log2 = fl; ctx->n_spl = n;
    return 0;
}

// Preprocess a batch that is already on the device (raw_dev readable up to raw_bytes rounded up to 16) and optionally scan it.
int agc_prep_and_scan(agcgpu_ctx* ctx, const uint8_t* raw_dev, uint64_t raw_bytes, const uint64_t* raw_offsets,
                      uint32_t n_contigs, bool do_scan, std::vector<ScanHit>* hits_out)
{
    if ((uintptr_t)raw_dev & 15) return agc_fail(ctx, AGCGPU_EINVAL, "raw device buffer must be 16-byte aligned");
    // tiles never straddle contigs: contig c owns ceil((end - align_down(start)) / 8192) tiles
    std::vector<TileDesc> tiles;
    std::vector<uint32_t> first_tile(n_contigs + 1);
    for (uint32_t c = 0; c < n_contigs; ++c) {
        uint64_t lo = raw_offsets[c], hi = raw_offsets[c + 1];
        first_tile[c] = (uint32_t)tiles.size();
        if (hi < lo || hi > raw_bytes) return agc_fail(ctx, AGCGPU_EINVAL, "bad raw_offsets at contig %u", c);
        for (uint64_t ch = lo >> 4; ch * 16 < hi; ch += AGC_TILE_CHUNKS) { TileDesc t; t.chunk0 = ch; t.lo = lo; t.hi = hi; tiles.push_back(t); }
    }
    first_tile[n_contigs] = (uint32_t)tiles.size();
    uint32_t n_tiles = (uint32_t)tiles.size();
    ctx->n_contigs = n_contigs;
    ctx->h_cstart.assign(n_contigs + 1, 0);
    ctx->n_exc = 0; ctx->h_exc_pos.clear(); ctx->total_bases = 0;
    if (n_tiles == 0) {
        if (int r = agc_reserve(ctx, ctx->d_cstart, (n_contigs + 1) * 8)) return r;
        CK(cudaMemsetAsync(ctx->d_cstart.p, 0, (n_contigs + 1) * 8, ctx->st));
        if (int r = agc_reserve(ctx, ctx->packed, 256)) return r;
        CK(cudaMemsetAsync(ctx->packed.p, 0, 256, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        return 0;
    }
    if (int r = agc_reserve(ctx, ctx->tile_desc, n_tiles * sizeof(TileDesc))) return r;
    if (int r = agc_reserve(ctx, ctx->tile_cnt, n_tiles * 4)) return r;
    if (int r = agc_reserve(ctx, ctx->tile_base, (size_t)(n_tiles + 1) * 16)) return r;
    uint64_t* base_kept = (uint64_t*)ctx->tile_base.p;
    uint64_t* base_exc = base_kept + (n_tiles + 1);
    CK(cudaMemcpyAsync(ctx->tile_desc.p, tiles.data(), n_tiles * sizeof(TileDesc), cudaMemcpyHostToDevice, ctx->st));
    ctx->stats.h2d_bytes += n_tiles * sizeof(TileDesc);
    k_tile_count<<<n_tiles, 256, 0, ctx->st>>>(raw_dev, (const TileDesc*)ctx->tile_desc.p, n_tiles, (uint32_t*)ctx->tile_cnt.p);
    CKL();
    k_tile_scan<<<1, 1024, 0, ctx->st>>>((const uint32_t*)ctx->tile_cnt.p, n_tiles, base_kept, base_exc);
    CKL();
    // need totals + per-contig starts on the host (allocation sizes)
    std::vector<uint64_t> h_base(n_tiles + 1);
    uint64_t totals[2];
    CK(cudaMemcpyAsync(h_base.data(), base_kept, (size_t)(n_tiles + 1) * 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaMemcpyAsync(&totals[1], base_exc + n_tiles, 8, cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.d2h_bytes += (size_t)(n_tiles + 1) * 8 + 8;
    totals[0] = h_base[n_tiles];
    for (uint32_t c = 0; c <= n_contigs; ++c) ctx->h_cstart[c] = h_base[first_tile[c]];
    ctx->total_bases = totals[0]; ctx->n_exc = totals[1];
    size_t packed_bytes = ((totals[0] + 63) / 64) * 16 + 256;
    if (int r = agc_reserve(ctx, ctx->packed, packed_bytes)) return r;
    if (int r = agc_reserve(ctx, ctx->exc_pos, (totals[1] + 1) * 8)) return r;
    if (int r = agc_reserve(ctx, ctx->exc_code, totals[1] + 1)) return r;
    if (int r = agc_reserve(ctx, ctx->d_cstart, (n_contigs + 1) * 8)) return r;
    CK(cudaMemsetAsync(ctx->packed.p, 0, packed_bytes, ctx->st));
    CK(cudaMemcpyAsync(ctx->d_cstart.p, ctx->h_cstart.data(), (n_contigs + 1) * 8, cudaMemcpyHostToDevice, ctx->st));
    k_tile_compact<<<n_tiles, 256, 0, ctx->st>>>(raw_dev, (const TileDesc*)ctx->tile_desc.p, n_tiles, base_kept, base_exc,
                                               (uint32_t*)ctx->packed.p, (uint64_t*)ctx->exc_pos.p, (uint8_t*)ctx->exc_code.p);
    CKL();
    if (totals[1]) {
        ctx->h_exc_pos.resize(totals[1]);
        CK(cudaMemcpyAsync(ctx->h_exc_pos.data(), ctx->exc_pos.p, totals[1] * 8, cudaMemcpyDeviceToHost, ctx->st));
        ctx->stats.d2h_bytes += totals[1] * 8;
    }
    if (!do_scan) { CK(cudaStreamSynchronize(ctx->st)); return 0; }
    return agc_scan_resident(ctx, hits_out);
}

// compress_contig's scan loop over the resident batch under the current splitter set (also the hard_contigs stage of -a mode)
int agc_scan_resident(agcgpu_ctx* ctx, std::vector<ScanHit>* hits_out)
{
    const uint32_t n_contigs = ctx->n_contigs;
    const uint64_t totals[1] = { ctx->total_bases };
    hits_out->clear();
    if (ctx->spl_keys.p == nullptr) return agc_fail(ctx, AGCGPU_EINVAL, "scan requested before agcgpu_set_splitters");
    std::vector<uint32_t> cp(n_contigs + 1);
    uint32_t total_chunks = 0;
    for (uint32_t c = 0; c < n_contigs; ++c) {
        cp[c] = total_chunks;
        uint64_t len = ctx->h_cstart[c + 1] - ctx->h_cstart[c];
        uint64_t nch = (len + AGC_SCAN_CHUNK - 1) / AGC_SCAN_CHUNK;
        if (total_chunks + nch > 0xfffffff0ull) return agc_fail(ctx, AGCGPU_EUNSUPPORTED, "batch too large for one scan launch");
        total_chunks += (uint32_t)nch;
    }
    cp[n_contigs] = total_chunks;
    if (int r = agc_reserve(ctx, ctx->chunk_prefix, (n_contigs + 1) * 4)) return r;
    CK(cudaMemcpyAsync(ctx->chunk_prefix.p, cp.data(), (n_contigs + 1) * 4, cudaMemcpyHostToDevice, ctx->st));
    uint32_t hit_cap = (uint32_t)std::min<uint64_t>(0x7fffffffull, totals[0] / 64 + 4ull * n_contigs + 1024);
    if (int r = agc_reserve(ctx, ctx->hits, (size_t)hit_cap * sizeof(ScanHit))) return r;
    if (int r = agc_reserve(ctx, ctx->counters, 64)) return r;
    CK(cudaMemsetAsync(ctx->counters.p, 0, 64, ctx->st));
    uint32_t h_count = 0;
    if (total_chunks) {
        size_t smem = (size_t)4 << (ctx->filter_log2 - 5);
        CK(cudaFuncSetAttribute(k_scan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        uint32_t grid = std::min<uint32_t>(total_chunks, (uint32_t)ctx->n_sm * 3);
        CK(cudaEventRecord(ctx->ev0, ctx->st));
        k_scan<<<grid, AGC_SCAN_THREADS, smem, ctx->st>>>((const uint64_t*)ctx->packed.p, (const uint64_t*)ctx->d_cstart.p, n_contigs,
            (const uint32_t*)ctx->chunk_prefix.p, total_chunks, ctx->prm.kmer_length, (const uint64_t*)ctx->spl_keys.p, ctx->spl_mask,
            (const uint32_t*)ctx->spl_filter.p, ctx->filter_log2, (const uint64_t*)ctx->exc_pos.p, ctx->n_exc,
            (ScanHit*)ctx->hits.p, (uint32_t*)ctx->counters.p, hit_cap);
        CKL();
        CK(cudaEventRecord(ctx->ev1, ctx->st));
        CK(cudaMemcpyAsync(&h_count, ctx->counters.p, 4, cudaMemcpyDeviceToHost, ctx->st));
        CK(cudaStreamSynchronize(ctx->st));
        cudaEventElapsedTime(&ctx->stats.last_scan_kernel_ms, ctx->ev0, ctx->ev1);
        ctx->stats.scan_kernel_ms_total += ctx->stats.last_scan_kernel_ms; ctx->stats.scan_bytes_total += totals[0] / 4; ctx->stats.scan_launches++;
    } else CK(cudaStreamSynchronize(ctx->st));
    if (h_count > hit_cap) return agc_fail(ctx, AGCGPU_EUNSUPPORTED, "splitter hits (%u) exceed capacity (%u)", h_count, hit_cap);
    hits_out->resize(h_count);
    if (h_count) {
        CK(cudaMemcpy(hits_out->data(), ctx->hits.p, (size_t)h_count * sizeof(ScanHit), cudaMemcpyDeviceToHost));
        ctx->stats.d2h_bytes += (size_t)h_count * sizeof(ScanHit);
        std::sort(hits_out->begin(), hits_out->end(), [](const ScanHit& a, const ScanHit& b) {
            return a.contig != b.contig ? a.contig < b.contig : a.pos < b.pos; });
    }
    return 0;
}

bool agc_segment_dirty(agcgpu_ctx* ctx, uint64_t gstart, uint32_t n)
{
    if (ctx->h_exc_pos.empty() || n == 0) return false;
    auto it = std::lower_bound(ctx->h_exc_pos.begin(), ctx->h_exc_pos.end(), gstart);
    return it != ctx->h_exc_pos.end() && *it < gstart + n;
}

int agc_expand_segment(agcgpu_ctx* ctx, uint64_t gstart, uint32_t n, uint32_t is_rc, uint8_t* dst_dev, uint32_t pad_bytes)
{
    uint32_t tot = n + pad_bytes;
    if (tot == 0) return 0;
    k_expand<<<(tot + 255) / 256, 256, 0, ctx->st>>>((const uint64_t*)ctx->packed.p, gstart, n, is_rc, dst_dev, pad_bytes);
    CKL();
    if (!ctx->h_exc_pos.empty() && n) {
        auto lo = std::lower_bound(ctx->h_exc_pos.begin(), ctx->h_exc_pos.end(), gstart);
        auto hi = std::lower_bound(ctx->h_exc_pos.begin(), ctx->h_exc_pos.end(), gstart + n);
        uint64_t e0 = lo - ctx->h_exc_pos.begin(), e1 = hi - ctx->h_exc_pos.begin();
        if (e1 > e0) {
            k_patch_exc<<<(uint32_t)((e1 - e0 + 255) / 256), 256, 0, ctx->st>>>((const uint64_t*)ctx->exc_pos.p,
                (const uint8_t*)ctx->exc_code.p, e0, e1, gstart, n, is_rc, dst_dev);
            CKL();
        }
    }
    return 0;
}

// The (k1, k2) -> group table lives twice: the device copy k_assign reads and a host mirror that decides where a new key goes.
// A rebuild (first use, or the load factor reached 1/4) sizes the table for twice the keys and uploads it whole; every other
// insertion only touches the slots of the new keys: they are placed in the mirror and sent as a short update list that
// k_map_apply writes into the device copy (one registration unit adds a handful of groups to a table of up to millions).
struct MapUpd { uint64_t slot, k1, k2; int32_t val; int32_t pad; };
__global__ void k_map_apply(const MapUpd* __restrict__ u, uint32_t n, uint64_t* __restrict__ mk1, uint64_t* __restrict__ mk2, int32_t* __restrict__ mval)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const MapUpd x = u[i];
    mk1[x.slot] = x.k1; mk2[x.slot] = x.k2; mval[x.slot] = x.val;
}
// place key i of the insertion log in the mirror; returns the slot when the device copy has to change, ~0 otherwise
static uint64_t map_place(agcgpu_ctx* ctx, uint64_t i)
{
    const uint64_t cap = ctx->t_val.size();
    const uint64_t a = ctx->h_map_k1[i], b = ctx->h_map_k2[i];
    const int32_t v = ctx->h_map_val[i];
    uint64_t slot = agc_murmur_pair(a, b) & (cap - 1);
    while (ctx->t_val[slot] >= 0 && !(ctx->t_k1[slot] == a && ctx->t_k2[slot] == b)) slot = (slot + 1) & (cap - 1);
    if (ctx->t_val[slot] < 0) { ctx->t_k1[slot] = a; ctx->t_k2[slot] = b; ctx->t_val[slot] = v; ++ctx->map_count; return slot; }
    if (ctx->t_val[slot] > v) { ctx->t_val[slot] = v; return slot; }      // keep the smallest id (agc_compressor.cpp:1010-1012)
    return ~0ull;
}
int agc_map_rebuild(agcgpu_ctx* ctx)
{
    const uint64_t n = ctx->h_map_k1.size();
    uint64_t cap = 1024;
    while (cap < 8 * n + 16) cap <<= 1;                          // load <= 1/8 now, rebuilt again at 1/4
    ctx->t_k1.assign(cap, 0); ctx->t_k2.assign(cap, 0); ctx->t_val.assign(cap, -1);
    ctx->map_count = 0;
    for (uint64_t i = 0; i < n; ++i) map_place(ctx, i);
    ctx->map_placed = n;
    if (int r = agc_reserve(ctx, ctx->map_k1, cap * 8)) return r;
    if (int r = agc_reserve(ctx, ctx->map_k2, cap * 8)) return r;
    if (int r = agc_reserve(ctx, ctx->map_val, cap * 4)) return r;
    CK(cudaMemcpyAsync(ctx->map_k1.p, ctx->t_k1.data(), cap * 8, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->map_k2.p, ctx->t_k2.data(), cap * 8, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaMemcpyAsync(ctx->map_val.p, ctx->t_val.data(), cap * 4, cudaMemcpyHostToDevice, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.h2d_bytes += cap * 20;
    ctx->map_mask = cap - 1;
    return 0;
}
// the keys appended to the insertion log since the last call
int agc_map_update(agcgpu_ctx* ctx)
{
    const uint64_t n = ctx->h_map_k1.size();
    if (ctx->map_val.p == nullptr || ctx->t_val.empty() || 4 * n + 16 > ctx->t_val.size()) return agc_map_rebuild(ctx);
    std::vector<MapUpd> upd;
    for (uint64_t i = ctx->map_placed; i < n; ++i) {
        const uint64_t slot = map_place(ctx, i);
        if (slot != ~0ull) { MapUpd u; u.slot = slot; u.k1 = ctx->t_k1[slot]; u.k2 = ctx->t_k2[slot]; u.val = ctx->t_val[slot]; u.pad = 0; upd.push_back(u); }
    }
    ctx->map_placed = n;
    if (upd.empty()) return 0;
    if (int r = agc_reserve(ctx, ctx->scr_gsz, upd.size() * sizeof(MapUpd) + 64)) return r;
    CK(cudaMemcpyAsync(ctx->scr_gsz.p, upd.data(), upd.size() * sizeof(MapUpd), cudaMemcpyHostToDevice, ctx->st));
    k_map_apply<<<(uint32_t)((upd.size() + 127) / 128), 128, 0, ctx->st>>>((const MapUpd*)ctx->scr_gsz.p, (uint32_t)upd.size(),
        (uint64_t*)ctx->map_k1.p, (uint64_t*)ctx->map_k2.p, (int32_t*)ctx->map_val.p);
    CKL();
    CK(cudaStreamSynchronize(ctx->st));                          // `upd` goes out of scope
    ctx->stats.h2d_bytes += upd.size() * sizeof(MapUpd);
    return 0;
}

int agc_assign_launch(agcgpu_ctx* ctx, const agcgpu_cut* cuts, uint64_t n, agcgpu_assign* out)
{
    if (n == 0) return 0;
    if (ctx->map_val.p == nullptr) if (int r = agc_map_rebuild(ctx)) return r;
    if (int r = agc_reserve(ctx, ctx->scr_misc, n * (sizeof(agcgpu_cut) + sizeof(agcgpu_assign)))) return r;
    agcgpu_cut* d_c = (agcgpu_cut*)ctx->scr_misc.p;
    agcgpu_assign* d_a = (agcgpu_assign*)(d_c + n);
    CK(cudaMemcpyAsync(d_c, cuts, n * sizeof(agcgpu_cut), cudaMemcpyHostToDevice, ctx->st));
    k_assign<<<(uint32_t)((n + 127) / 128), 128, 0, ctx->st>>>(d_c, n, (const uint64_t*)ctx->map_k1.p, (const uint64_t*)ctx->map_k2.p,
                                                              (const int32_t*)ctx->map_val.p, ctx->map_mask, d_a);
    CKL();
    CK(cudaMemcpyAsync(out, d_a, n * sizeof(agcgpu_assign), cudaMemcpyDeviceToHost, ctx->st));
    CK(cudaStreamSynchronize(ctx->st));
    ctx->stats.h2d_bytes += n * sizeof(agcgpu_cut); ctx->stats.d2h_bytes += n * sizeof(agcgpu_assign);
    return 0;
}
