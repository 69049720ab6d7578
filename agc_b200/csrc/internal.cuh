// internal.cuh -- shared declarations of libagcgpu (sm_100a). Not part of the public ABI (include/agcgpu.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <vector>
#include <unordered_map>
#include "../../include/agcgpu.h"

#define AGC_EMPTY32 0xffffffffu
#define AGC_TILE_CHUNKS 512u          // 16-byte chunks of raw FASTA per preprocessing tile
#define AGC_TILE_BYTES (AGC_TILE_CHUNKS * 16u)
#define AGC_SCAN_THREADS 256u
#define AGC_SCAN_CHUNK (AGC_SCAN_THREADS * 32u)   // k-mer end positions per scan work unit

// ------------------------------------------------------------------------------------------------ device helpers
#ifdef __CUDACC__
// murmur3 fmix64 -- reference: src/common/utils.h:164-176 (MurMur64Hash)
__host__ __device__ __forceinline__ uint64_t agc_murmur64(uint64_t h)
{
    h ^= h >> 33; h *= 0xff51afd7ed558ccdULL;
    h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL;
    h ^= h >> 33;
    return h;
}
__host__ __device__ __forceinline__ uint64_t agc_murmur_pair(uint64_t a, uint64_t b)   // utils.h:203-225
{
    return agc_murmur64(agc_murmur64(a) ^ b);
}

// byte-swap a 64-bit little-endian load so that the first base (first byte, top two bits) lands in bits 63:62
__device__ __forceinline__ uint64_t agc_be64(uint64_t x)
{
    uint32_t lo = __byte_perm((uint32_t)x, 0, 0x0123);
    uint32_t hi = __byte_perm((uint32_t)(x >> 32), 0, 0x0123);
    return ((uint64_t)lo << 32) | hi;
}
// 32 bases starting at base index g of a 2-bit packed sequence (first base at the MSB). Reads words g/32 and g/32+1.
__device__ __forceinline__ uint64_t agc_win(const uint64_t* __restrict__ P, uint64_t g)
{
    uint64_t i = g >> 5;
    uint32_t sh = (uint32_t)(g & 31u) * 2u;
    uint64_t a = agc_be64(P[i]);
    uint64_t b = agc_be64(P[i + 1]);
    return (a << sh) | ((b >> 1) >> (63u - sh));
}
// same, g may be negative (> -32): slots before base 0 are zero-filled garbage the callers mask out
__device__ __forceinline__ uint64_t agc_win_s(const uint64_t* __restrict__ P, int64_t g)
{
    if (g >= 0) return agc_win(P, (uint64_t)g);
    if (g <= -32) return 0;
    return agc_win(P, 0) >> (uint32_t)(2 * (-g));
}
// reverse the order of the 32 2-bit groups of x
__device__ __forceinline__ uint64_t agc_rev2(uint64_t x)
{
    uint64_t r = __brevll(x);
    return ((r & 0xAAAAAAAAAAAAAAAAULL) >> 1) | ((r & 0x5555555555555555ULL) << 1);
}
#endif

// ------------------------------------------------------------------------------------------------ host structures
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

// device-side descriptor of one group's reference segment + LZ index
struct GroupRefDev {
    const uint8_t* packed;     // 2-bit packed reference, base 0 at byte 0; 16-byte aligned, zero padded (+16 B slack)
    const void* ht;            // u16 (short) or u32 slots, value = ref_pos / 4, all-ones = empty
    const uint8_t* codes;      // 1 byte / symbol + key_len bytes of 31 (only when the reference has non-ACGT symbols)
    uint32_t m;                // reference length in symbols
    uint32_t ht_size;          // power of two
    uint32_t flags;            // bit0 short (u16) table, bit1 dirty (non-ACGT present), bit2 present
    uint32_t packed_bytes;     // bytes staged to shared memory (multiple of 16)
};
#define GRF_SHORT 1u
#define GRF_DIRTY 2u
#define GRF_PRESENT 4u

struct ScanHit {               // one k-mer occurrence that is a splitter
    uint64_t pos;              // end position of the k-mer, contig coordinates
    uint64_t dir, rc;          // CKmer words
    uint32_t contig;
    uint32_t pad;
};

struct SplPos {                // where k_find_splitters found a splitter
    uint64_t pos;              // end position of the k-mer, contig coordinates
    uint32_t contig;           // index within the launch
    uint32_t is_last;          // the right-most candidate added after the walk
};

struct LzReqDev {              // device form of agcgpu_seg_req
    uint64_t gstart;           // global base index of the segment's first base in the packed contig store
    uint32_t n;
    uint32_t is_rc;
    uint32_t group;
    uint32_t bound;
    uint64_t out_off;          // byte offset into the output slab (encode) / u32 index (cost vector)
    uint32_t out_cap;
    uint32_t orig;             // index in the caller's request array
};

struct LzUnit {                // one CTA's work: requests [first, first+count) all of group `group`
    uint32_t group, first, count, pad;
};

struct agcgpu_ctx {
    agcgpu_params prm;
    int dev = 0;
    int n_sm = 0;
    cudaStream_t st = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaStream_t st2 = nullptr; cudaEvent_t ev2 = nullptr;      // second stream (residual coder: narrow kernel beside the wide one)
    std::string err;
    agcgpu_stats stats;

    // splitter set (any exact set is observationally equivalent)
    DevBuf spl_keys;           // open addressing, ~0 = empty
    uint64_t spl_mask = 0;
    DevBuf spl_filter;         // 2^filter_log2 bits, one hash: first-level reject in shared memory
    uint32_t filter_log2 = 10;
    uint64_t n_spl = 0;
    // -a mode (AGCGPU_F_ADAPTIVE): every k-mer of the reference sample, sorted (invalid positions = ~0 at the end) --
    // v_candidate_kmers and v_duplicated_kmers of the reference in one list (agc_compressor.cpp:493-494, 2066-2076)
    DevBuf ref_kmers;
    uint64_t n_ref_kmers = 0;
    struct SplFound { uint32_t contig; uint64_t pos, kmer; uint8_t is_last; };
    std::vector<SplFound> h_last_spl;      // splitters of the last determine / find_new call with their positions (-f mode)

    // resident contig batch
    DevBuf raw;                // raw FASTA bytes (only when uploaded through agcgpu_scan_contigs)
    DevBuf packed;             // 2-bit packed symbols of all contigs, back to back
    DevBuf exc_pos, exc_code;  // sorted exception list (global base index, code)
    DevBuf tile_desc, tile_cnt, tile_base;
    DevBuf d_cstart;           // n_contigs+1 global base offsets
    DevBuf chunk_prefix;
    DevBuf hits; DevBuf counters;
    std::vector<uint64_t> h_cstart;
    std::vector<uint64_t> h_exc_pos;     // host mirror (dirty-segment classification)
    uint32_t n_contigs = 0;
    uint64_t n_exc = 0;
    uint64_t total_bases = 0;

    // segment map
    DevBuf map_k1, map_k2, map_val;
    uint64_t map_mask = 0, map_count = 0;
    std::vector<uint64_t> h_map_k1, h_map_k2; std::vector<int32_t> h_map_val;   // insertion log for rebuilds
    std::vector<uint64_t> t_k1, t_k2; std::vector<int32_t> t_val;               // host mirror of the table (places new keys)
    uint64_t map_placed = 0;                                                    // log entries already in the mirror

    // reference store
    std::vector<GroupRefDev> h_groups;
    DevBuf d_groups;
    std::vector<std::pair<void*, size_t>> arena_chunks;
    uint8_t* arena_cur = nullptr; size_t arena_left = 0;
    size_t device_bytes = 0;

    // scratch
    DevBuf scr_req, scr_units, scr_out, scr_sizes, scr_offs, scr_dense, scr_misc, scr_bytes, scr_chunk, scr_rec, scr_gsz, scr_gather, scr_zkeep, scr_cost;
    void* pin = nullptr; size_t pin_cap = 0;   // pinned host staging
    std::vector<uint64_t> last_slab_off;       // device-only encode: slab offset of every delta (request order)
    uint64_t last_lzc_chunks = 0;              // chunk records of the last chunk-parallel encode (diagnostics)
    std::vector<void*> zwaves;                 // residual-coder batches in flight (agcgpu_zstd_submit), in submission order
};

// ------------------------------------------------------------------------------------------------ internal API
int agc_fail(agcgpu_ctx* c, int code, const char* fmt, ...);
int agc_reserve(agcgpu_ctx* c, DevBuf& b, size_t bytes, bool keep = false);
void* agc_arena_alloc(agcgpu_ctx* c, size_t bytes);    // 256-byte aligned, lives until destroy
int agc_pin_reserve(agcgpu_ctx* c, size_t bytes);
void agc_zstd_waves_drop(agcgpu_ctx* c);      // kernels_zstd.cu: waits for and frees the waves nobody collected
void* agc_dev_alloc(int dev, size_t bytes, size_t* cap_out);   // process-wide device-memory pool (api.cu)
void agc_dev_free(int dev, void* p, size_t cap);
void agc_dev_trim(int dev);

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return agc_fail(ctx, AGCGPU_ECUDA, "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CKL() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) \
    return agc_fail(ctx, AGCGPU_ECUDA, "%s:%d launch: %s", __FILE__, __LINE__, cudaGetErrorString(e_)); \
    ctx->stats.kernel_launches++; } while (0)

// kernels_prep.cu
int agc_prep_and_scan(agcgpu_ctx* ctx, const uint8_t* raw_dev, uint64_t raw_bytes, const uint64_t* raw_offsets,
                      uint32_t n_contigs, bool do_scan, std::vector<ScanHit>* hits_out);
int agc_scan_resident(agcgpu_ctx* ctx, std::vector<ScanHit>* hits_out);
int agc_filtered_kmers(agcgpu_ctx* ctx, uint64_t gstart, uint64_t len, uint64_t thr, std::vector<agcgpu_fkmer>& out);
// splitters of resident contigs [c0, c0+nc): candidates = singletons among the k-mers of those contigs, minus (exclude_ref)
// the k-mers of the reference sample; keep_kmers moves the sorted k-mer list into ctx->ref_kmers
int agc_enumerate_splitters(agcgpu_ctx* ctx, uint32_t c0, uint32_t nc, bool exclude_ref, bool keep_kmers, std::vector<uint64_t>& out_sorted);
int agc_expand_segment(agcgpu_ctx* ctx, uint64_t gstart, uint32_t n, uint32_t is_rc, uint8_t* dst_dev, uint32_t pad_bytes);
int agc_upload_splitters(agcgpu_ctx* ctx, const uint64_t* s, uint64_t n);
int agc_map_rebuild(agcgpu_ctx* ctx);
int agc_map_update(agcgpu_ctx* ctx);      // places the keys logged since the last call (rebuilds when the table is too full)
int agc_assign_launch(agcgpu_ctx* ctx, const agcgpu_cut* cuts, uint64_t n, agcgpu_assign* out);

// comm.cu (NCCL exchange)
bool agc_comm_active();
uint32_t agc_comm_rank();
uint32_t agc_comm_world();
int agc_comm_allgather(agcgpu_ctx* ctx, const void* d_send, void* d_recv, size_t bytes, cudaStream_t st);
int agc_comm_allgatherv(agcgpu_ctx* ctx, const void* d_mine, uint64_t my_bytes, int local_status, std::vector<uint64_t>& sizes, uint64_t* stride_out);

// kernels_lz.cu
int agc_refs_from_segments(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n);
int agc_ref_from_host(agcgpu_ctx* ctx, uint32_t group, const uint8_t* symbols, uint32_t len);
int agc_lz_run(agcgpu_ctx* ctx, int mode, const agcgpu_seg_req* reqs, uint32_t n, int prefix_costs,
               uint8_t* out_bytes, uint64_t out_cap, uint64_t* out_offsets, uint32_t* out_u32);
int agc_lz_encode_sharded(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_offsets);
int agc_lz_cost_split(agcgpu_ctx* ctx, const agcgpu_split_req* reqs, uint32_t n, uint32_t* out_pos, uint32_t* out_sum);
int agc_pack_refs(agcgpu_ctx* ctx, const uint32_t* group_ids, uint32_t n, uint8_t* out, uint64_t out_cap,
                  uint64_t* out_offsets, uint8_t* out_use_tuples);
bool agc_segment_dirty(agcgpu_ctx* ctx, uint64_t gstart, uint32_t n);
