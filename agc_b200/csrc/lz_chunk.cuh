// lz_chunk.cuh -- shared declarations of the chunk-parallel LZ-diff encoder (kernels_lz_chunk.cu, driven from kernels_lz.cu)
#pragma once
#include <stdint.h>
#include <stddef.h>

#define LZC_CHUNK 2048u            // text positions per chunk of a large batch (a multiple of 4: diagonal-0 matches start on an indexed position)
#define LZC_CHUNK_MIN 512u         // ... of a small batch: more lanes, fewer tokens per lane (the chunk size travels in LzcReq::chunk)
#define LZC_CSLAB 2560u            // bytes of token output a chunk can produce: (LZC_CHUNK + min_match_len) positions, <= 23 bytes per 21
#define LZC_THREADS 512
#define LZC_STAGE_LIMIT (110u * 1024u)

// chunk record flags
#define LZC_HAS_FIRST   1u         // the chunk found a match
#define LZC_FIRST_OPEN  2u         // ... and it was still matching at the end of the chunk (nothing follows it)
#define LZC_FIRST_MULTI 4u         // ... chosen among several candidates
#define LZC_FIRST_BLIM  8u         // ... whose backward extension stopped at the chunk's first position
#define LZC_SENS       16u         // a probe before the first match failed only because of the local no_prev_literals
#define LZC_END_OPEN   32u         // a later match was still matching at the end of the chunk
#define LZC_EQ         64u         // chunk == reference at the same positions (one diagonal-0 match over the whole chunk)
#define LZC_NEQ       128u         // a position of the chunk differs from the reference at the same position

struct LzcRec {                    // what a chunk's parse leaves for the stitcher
    uint32_t flags;
    uint32_t lit0;                 // literals before the first match (bytes [0, lit0) of the chunk's slab; all bytes if there is no match)
    uint32_t first_p;              // text position whose probe found the first match
    uint32_t first_ts, first_mp, first_len;     // first match after the local rewind: text start, reference start, length (so far, if open)
    uint32_t bytes;                // bytes in the chunk's slab: lit0 raw literals, then the tokens after the first match
    uint32_t end_i, end_np, end_pred;           // state after the chunk's last token
    int32_t  end_diag;             // reference - text position of the chunk's last match
    uint32_t open_ts, open_mp, open_predb;      // LZC_END_OPEN: the open match and pred_pos before it
    uint32_t pad[2];
};

struct LzcReq {                    // one segment to encode
    uint64_t gstart;               // global base index of its first base
    uint32_t n, is_rc, group;
    uint32_t chunk_first;          // index of its first chunk record / chunk slab
    uint32_t nch;
    uint32_t unit_base;            // chunks of the requests before it in its unit
    uint64_t out_off;              // byte offset of its delta in the output slab (cost vectors: u32 index of its vector)
    uint32_t out_cap;              // (cost vectors: prefix_costs)
    uint32_t orig;                 // index in the caller's request array
    uint32_t chunk;                // text positions per chunk (LZC_CHUNK_MIN .. LZC_CHUNK, a multiple of 4)
    uint32_t pad;
};

struct LzcUnit { uint32_t group, first, count, item0, n_items, pad; };   // one CTA: chunks [item0, item0 + n_items) of requests [first, first+count)

struct agcgpu_ctx;
int agc_lzc_launch(agcgpu_ctx* ctx, const LzcReq* d_reqs, uint32_t n_req, const LzcUnit* d_units, uint32_t n_units, size_t smem,
                   uint8_t* cslab, LzcRec* recs, uint8_t* slab, uint32_t* res, uint32_t* fb, uint32_t* counters, uint32_t* costv = nullptr);
// kernels_lz_diag.cu: warp per segment, streaming along the current diagonal (mode 0 only)
struct LzReqDev; struct LzUnit;
int agc_lzd_launch(agcgpu_ctx* ctx, const LzReqDev* d_reqs, const LzUnit* d_units, uint32_t n_units, size_t stage_bytes, int ht_staged,
                   uint8_t* slab, uint32_t* res, uint32_t* err);
size_t agc_lzd_scratch_bytes(int ht_staged);
#define LZD_MAX_N (1u << 17)      // longer segments go to the chunk-parallel kernels (one warp would walk them alone)
