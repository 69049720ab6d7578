/*
 * agcgpu.h -- C ABI of libagcgpu.so: the B200 (sm_100a) implementation of AGC's compression hot path.
 *
 * The reference (refresh-bio/agc v3.2.2) has no plugin/FFI seam on the compress side; the seam is the set of
 * internal calls that CAGCCompressor / CSegment make per contig and per segment.  Each entry point below names the
 * reference call site it replaces (paths relative to the reference tree).  INTEGRATION.md shows the binding a
 * maintainer adds on the reference side.
 *
 * Conventions: plain pointers and sizes only (no C++/torch types); every function returns 0 on success and a
 * negative AGCGPU_E* code on failure, agcgpu_last_error() gives the message; the caller owns host buffers, the
 * context owns device memory and streams; no exception crosses the boundary.  There is NO CPU fallback:
 * without a usable sm_100 device agcgpu_create() fails.
 *
 * Sequence representation on the device: bases are kept 2-bit packed (4 bases / byte, first base in the two most
 * significant bits -- the same layout as the reference's 4-per-byte "tuples", src/common/segment.h:73-138) plus a
 * sorted exception list (position, code) for every non-ACGT symbol, so the store is lossless.
 */
#ifndef AGCGPU_H
#define AGCGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */
#endif

#define AGCGPU_OK            0
#define AGCGPU_ECUDA        -1   /* CUDA runtime error */
#define AGCGPU_EINVAL       -2   /* bad argument */
#define AGCGPU_ENOMEM       -3   /* host or device allocation failed / capacity exceeded */
#define AGCGPU_EOVERFLOW    -4   /* caller-provided output buffer too small */
#define AGCGPU_ENODEV       -5   /* no sm_100 device */
#define AGCGPU_EUNSUPPORTED -6   /* input outside the implemented envelope (fails loudly, never falls back) */

typedef struct agcgpu_ctx agcgpu_ctx;

/* Compression parameters: CAGCCompressor::Create arguments (src/core/agc_compressor.cpp:2273-2288) */
typedef struct {
    uint32_t kmer_length;        /* -k, 17..32 */
    uint32_t min_match_len;      /* -l, 15..32 */
    uint32_t segment_size;       /* -s */
    uint32_t pack_cardinality;   /* -b */
    int32_t  device;             /* CUDA device ordinal */
    uint32_t flags;              /* AGCGPU_F_* */
} agcgpu_params;
/* -a (adaptive_compression, agc_compressor.cpp:493-494,536-541): agcgpu_determine_splitters keeps the sorted k-mer
 * list of the reference sample on the device (v_candidate_kmers + v_duplicated_kmers) for agcgpu_find_new_splitters */
#define AGCGPU_F_ADAPTIVE 1u

/* One segment cut out of a contig: what compress_contig hands to add_segment
 * (src/core/agc_compressor.cpp:2019-2048).  front/back are the CKmer words (src/core/kmer.h:21-31) of the
 * terminal splitters, left-aligned; has_* = CKmer::is_full(). */
typedef struct {
    uint32_t contig;             /* index into the resident contig batch */
    uint32_t has_front;
    uint32_t has_back;
    uint32_t reserved;
    uint64_t start;              /* first base, in preprocessed contig coordinates */
    uint64_t len;
    uint64_t front_dir, front_rc;
    uint64_t back_dir, back_rc;
} agcgpu_cut;

/* A segment of a resident contig, optionally reverse-complemented (reverse_complement_copy,
 * src/common/agc_basic.cpp:280-316) -- the `contig_t` argument of CSegment::add / estimate / get_coding_cost. */
typedef struct {
    uint32_t contig;
    uint32_t is_rc;
    uint64_t start;
    uint32_t len;
    uint32_t group_id;           /* reference-segment group to code against */
    uint32_t bound;              /* estimate only: CLZDiff_V2::Estimate bound (src/common/lz_diff.cpp:868) */
    uint32_t reserved;
} agcgpu_seg_req;

/* Result of the hash-assign kernel for one cut (the common branches of CAGCCompressor::add_segment,
 * src/core/agc_compressor.cpp:1287-1313 + map_segments.find 1363). */
typedef struct {
    uint64_t key1, key2;         /* (min,max) canonical splitter pair; ~0 where a splitter is missing */
    int32_t  group_id;           /* >= 0: known group; -1: pair not in the map (host decides: new group / split) */
    uint32_t is_rc;              /* store reverse-complemented */
    uint32_t klass;              /* 0 both splitters, 1 front only, 2 back only, 3 none */
    uint32_t reserved;
} agcgpu_assign;

typedef struct {
    uint64_t kernel_launches;    /* kernels of this library launched since agcgpu_create */
    uint64_t h2d_bytes, d2h_bytes;
    uint64_t device_bytes_in_use;
    uint64_t lz_alg_bytes;       /* sum over LZ requests of ceil(n/4)+ceil(m/4)+e of the last LZ batch */
    float    last_lz_kernel_ms;  /* device time of the last LZ kernel (CUDA events on the library stream) */
    float    last_scan_kernel_ms;
    float    zstd_kernel_ms;     /* device time spent in the residual coder since agcgpu_create (sum over batches) */
    float    zstd_input_mb;      /* bytes handed to the residual coder since agcgpu_create, in 10^6 bytes */
    uint64_t lz_chunk_segments;       /* segments encoded by the chunk-parallel kernels since agcgpu_create ... */
    uint64_t lz_sequential_segments;  /* ... of which the stitcher handed to the sequential kernel */
    /* sums since agcgpu_create (a step of the bench = one create): LZ-diff encode launches and splitter-scan launches */
    uint64_t lz_alg_bytes_total;      /* sum of ceil(n/4)+ceil(m/4)+e over all encode requests */
    uint64_t scan_bytes_total;        /* 2-bit packed bytes read by k_scan (bases / 4) */
    float    lz_kernel_ms_total;      /* device time of the encode kernels (CUDA events) */
    float    scan_kernel_ms_total;
    uint32_t lz_encode_launches, scan_launches;
    float    zstd_wait_ms;            /* host time blocked in agcgpu_zstd_collect: the part of the residual coder that did NOT overlap */
    uint32_t lz_diag_segments;        /* segments encoded by the warp-per-segment diagonal kernel since agcgpu_create */
} agcgpu_stats;

/* ---- lifetime ------------------------------------------------------------------------------------------------- */
/* ctor/dtor of the device side of CAGCCompressor (agc_compressor.h:540-764). */
int agcgpu_create(const agcgpu_params* params, agcgpu_ctx** out_ctx);
void agcgpu_destroy(agcgpu_ctx* ctx);
const char* agcgpu_last_error(const agcgpu_ctx* ctx);     /* ctx may be NULL: error of the last failed create */
int agcgpu_sync(agcgpu_ctx* ctx);                         /* wait for all work queued on the library stream */
void* agcgpu_stream(agcgpu_ctx* ctx);                     /* cudaStream_t the kernels are launched on (for event timing) */
int agcgpu_get_stats(agcgpu_ctx* ctx, agcgpu_stats* out);
/* Page-locked host memory for ingest buffers (CGenomeIO::ReadContigRaw's destination, src/core/genome_io.cpp:206-250): raw FASTA
 * read straight into such a buffer is uploaded by agcgpu_scan_contigs without a staging copy.  Pooled per process; *out_cap = the
 * capacity actually handed out (>= bytes), to be passed back to agcgpu_host_free.  NULL when the allocation fails. */
void* agcgpu_host_alloc(uint64_t bytes, uint64_t* out_cap);
void agcgpu_host_free(void* p, uint64_t cap);

/* ---- splitters -------------------------------------------------------------------------------------------------- */
/* determine_splitters (src/core/agc_compressor.cpp:428-563): raw FASTA bodies of the reference sample in
 * (newlines included, as CGenomeIO::ReadContigRaw returns them) -> sorted splitter list out.  The contigs are
 * uploaded by this call; they do not stay resident.  Also installs the set (as agcgpu_set_splitters). */
int agcgpu_determine_splitters(agcgpu_ctx* ctx, const uint8_t* raw, const uint64_t* raw_offsets, uint32_t n_contigs,
                               uint64_t* out_splitters, uint64_t cap, uint64_t* out_n);
/* hs_splitters / bloom_splitters fill (agc_compressor.cpp:543-555; append: 339-351; -a: 1191-1209) */
int agcgpu_set_splitters(agcgpu_ctx* ctx, const uint64_t* splitters, uint64_t n);

/* -a mode, CAGCCompressor::find_new_splitters (agc_compressor.cpp:2054-2082) for resident contigs that the scan left
 * without a single splitter (compress_contig 2038-2044): per contig, its singleton k-mers that do not occur in the
 * reference sample (neither as singletons nor duplicated) are the candidates of find_splitters_in_contig (762-825).
 * Needs AGCGPU_F_ADAPTIVE.  out = the new splitters of all listed contigs, sorted, de-duplicated (the reference
 * inserts them into a hash set, 1191-1209).  The splitter set itself is NOT changed: the caller merges and calls
 * agcgpu_set_splitters. */
int agcgpu_find_new_splitters(agcgpu_ctx* ctx, const uint32_t* contigs, uint32_t n, uint64_t* out_splitters, uint64_t cap,
                              uint64_t* out_n);

/* -f mode (fallback minimizers).  One k-mer that passes CAGCCompressor::kmer_filter_t (agc_compressor.h:570-599):
 * (MurMur64Hash(canonical k-mer) ^ 0xD73F8BF11046C40E) < threshold. */
typedef struct {
    uint64_t pos;                /* position of the k-mer's last base, relative to the range's first base */
    uint64_t kmer;               /* canonical CKmer word */
    uint32_t is_dir_oriented;    /* CKmer::is_dir_oriented(): kmer_dir <= kmer_rc */
    uint32_t is_symmetric;       /* kmer_dir == kmer_rc (find_splitters_in_contig skips those, agc_compressor.cpp:788) */
} agcgpu_fkmer;
/* The filtered k-mers of resident ranges, in position order: the loops of find_cand_segment_using_fallback_minimizers
 * (agc_compressor.cpp:1826-1855) and find_splitters_in_contig (776-789).  Range i = reqs[i] (contig, start, len; is_rc and
 * group_id are ignored: the canonical k-mers of a reverse complement are the same ones, orientation flipped; len is
 * clipped to the contig's end, so 0xFFFFFFFF means "to the end");
 * k-mers that overlap a non-ACGT symbol do not exist.  Range i's k-mers = out[out_offsets[i] .. out_offsets[i+1]). */
int agcgpu_filtered_kmers(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint64_t threshold, agcgpu_fkmer* out,
                          uint64_t cap, uint64_t* out_offsets);
/* Where the splitters of the last agcgpu_determine_splitters / agcgpu_find_new_splitters call were found:
 * (contig index in that call's batch, position of the k-mer's last base, canonical k-mer), ordered by (contig, position);
 * is_last[i] = 1 for the right-most-candidate splitter that find_splitters_in_contig adds after its walk (816-824).
 * The contigs of agcgpu_determine_splitters stay resident until the next scan call, so agcgpu_filtered_kmers can be
 * applied to them: together these give the v_fallbacks of find_splitters_in_contig (797-802, 821-822). */
int agcgpu_last_splitter_positions(agcgpu_ctx* ctx, uint32_t* out_contig, uint64_t* out_pos, uint64_t* out_kmer, uint8_t* out_is_last,
                                   uint64_t cap, uint64_t* out_n);

/* ---- contig batch: preprocess + scan ---------------------------------------------------------------------------- */
/* preprocess_raw_contig (agc_compressor.cpp:907-951) + compress_contig's scan loop (1997-2051) for a batch of
 * contigs.  raw = concatenated raw bodies, contig i = raw[raw_offsets[i] .. raw_offsets[i+1]).  The preprocessed
 * contigs replace the previous resident batch.  out_contig_len[i] = number of symbols after preprocessing.
 * cuts come out ordered by (contig, start). */
int agcgpu_scan_contigs(agcgpu_ctx* ctx, const uint8_t* raw, const uint64_t* raw_offsets, uint32_t n_contigs,
                        uint64_t* out_contig_len, agcgpu_cut* out_cuts, uint64_t cap_cuts, uint64_t* out_n_cuts);
/* Same, input already resident on the device (raw_dev = device pointer); used by bench.py for the HBM-resident
 * timing and by pipelines that ingest through their own pinned staging. */
int agcgpu_scan_contigs_dev(agcgpu_ctx* ctx, const void* raw_dev, uint64_t raw_bytes, const uint64_t* raw_offsets,
                            uint32_t n_contigs, uint64_t* out_contig_len, agcgpu_cut* out_cuts, uint64_t cap_cuts,
                            uint64_t* out_n_cuts);
/* compress_contig's scan loop again over the RESIDENT batch under the current splitter set: the hard_contigs stage of
 * -a mode (agc_compressor.cpp:1211-1227), after agcgpu_set_splitters added the new splitters.  Cuts of every resident
 * contig come out ordered by (contig, start); the caller keeps those of the contigs it re-queued. */
int agcgpu_rescan_contigs(agcgpu_ctx* ctx, agcgpu_cut* out_cuts, uint64_t cap_cuts, uint64_t* out_n_cuts);
/* get_part (agc_compressor.cpp:2085-2091) / reverse_complement_copy: download symbols (1 byte each) of a segment */
int agcgpu_get_segment(agcgpu_ctx* ctx, uint32_t contig, uint64_t start, uint32_t len, uint32_t is_rc, uint8_t* out);

/* ---- hash-assign ------------------------------------------------------------------------------------------------ */
/* map_segments updates (store_segments, agc_compressor.cpp:1007-1012) */
int agcgpu_map_insert(agcgpu_ctx* ctx, const uint64_t* key1, const uint64_t* key2, const int32_t* group_id, uint64_t n);
/* add_segment's key construction + map_segments.find for a batch of cuts (agc_compressor.cpp:1287-1313,1363) */
int agcgpu_assign_cuts(agcgpu_ctx* ctx, const agcgpu_cut* cuts, uint64_t n, agcgpu_assign* out);

/* ---- reference segments and LZ-diff ----------------------------------------------------------------------------- */
/* CLZDiffBase::Prepare + prepare_index (src/common/lz_diff.cpp:48-149,375-428), called from CSegment::add for the
 * first sequence of a group (src/common/segment.cpp:41-47).  Batch form: reference i of group req[i].group_id is
 * the (possibly reverse-complemented) resident segment req[i]. */
int agcgpu_group_put_reference_batch(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n);
/* Same from host symbols (append path: CSegment::unpack, src/common/segment.cpp:528) */
int agcgpu_group_put_reference(agcgpu_ctx* ctx, uint32_t group_id, const uint8_t* symbols, uint32_t len);
/* hash-table slots of a group's index, widened to u32 (0xFFFFFFFF = empty): for layout-parity tests */
int agcgpu_group_get_index(agcgpu_ctx* ctx, uint32_t group_id, uint32_t* out_slots, uint64_t cap, uint64_t* out_ht_size);

/* CLZDiff_V2::Encode (lz_diff.cpp:669-798) as called from CSegment::add (segment.cpp:59).
 * out_offsets has n+1 entries; delta i = out[out_offsets[i] .. out_offsets[i+1]). */
int agcgpu_lz_encode_batch(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n,
                           uint8_t* out, uint64_t out_cap, uint64_t* out_offsets);
/* diagnostics: the chunk records (agc_b200/csrc/lz_chunk.cuh, LzcRec, 64 bytes each) the last agcgpu_lz_encode_batch left on the
 * device, in request order as sorted by group; tests compare them with the host build of the same source */
int agcgpu_debug_lz_chunk_records(agcgpu_ctx* ctx, void* out, uint64_t cap_bytes, uint64_t* out_n_records);

/* CLZDiff_V2::Estimate (lz_diff.cpp:839-946) as called from CSegment::estimate (agc_compressor.cpp:1705,1733,1757) */
int agcgpu_lz_estimate_batch(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint32_t* out);
/* CLZDiffBase::GetCodingCostVector (lz_diff.cpp:159-284) as called from CSegment::get_coding_cost
 * (agc_compressor.cpp:1540-1571).  out has req->len entries. */
int agcgpu_lz_cost_vector(agcgpu_ctx* ctx, const agcgpu_seg_req* req, int prefix_costs, uint32_t* out);

/* One decision of find_cand_segment_with_missing_middle_splitter (agc_compressor.cpp:1502-1627), everything between the two
 * get_coding_cost calls (1540-1571) and the argmin loop (1605-1616): v1 = cost vector of the resident segment against group1,
 * v2 = against group2; each is taken in the orientation / prefix_costs mode the reference picks from the order of the splitters,
 * reversed when it was computed on the reverse complement; v1 is cumulated from the left (partial_sum), v2 from the right, and
 * best_pos = the FIRST i with the smallest v1cum[i] + v2cum[i] (u32 arithmetic), best_sum = that value (~0 when len == 0).
 * The host applies the k+1 snapping (1621-1624) itself.  Only 8 bytes per decision come back instead of 2 x len x 4. */
typedef struct {
    uint32_t contig;
    uint32_t len;
    uint64_t start;
    uint32_t group1, group2;
    uint32_t flags;              /* AGCGPU_SPLIT_* */
    uint32_t reserved;
} agcgpu_split_req;
#define AGCGPU_SPLIT_RC1      1u   /* v1: code the reverse complement of the segment */
#define AGCGPU_SPLIT_PREFIX1  2u   /* v1: get_coding_cost(..., prefix_costs = true) */
#define AGCGPU_SPLIT_REV1     4u   /* v1: reverse the vector before cumulating (1551-1552) */
#define AGCGPU_SPLIT_RC2      8u
#define AGCGPU_SPLIT_PREFIX2 16u
#define AGCGPU_SPLIT_REV2    32u
int agcgpu_lz_cost_split_batch(agcgpu_ctx* ctx, const agcgpu_split_req* reqs, uint32_t n, uint32_t* out_best_pos, uint32_t* out_best_sum);

/* CLZDiff_V2::Decode (lz_diff.cpp:801-836) as CSegment::get calls it (segment.cpp:220-399): delta i = deltas[delta_offsets[i] ..
 * delta_offsets[i+1]) is decoded against the resident reference of group_ids[i]; symbols (1 byte each) of segment i =
 * out[out_offsets[i] .. out_offsets[i+1]).  out_offsets is filled even when out_cap is too small (AGCGPU_EOVERFLOW).
 * An empty delta decodes to nothing (CSegment stores "equal to the reference" that way, segment.cpp:61-64). */
int agcgpu_lz_decode_batch(agcgpu_ctx* ctx, const uint32_t* group_ids, const uint8_t* deltas, const uint64_t* delta_offsets,
                           uint32_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_offsets);

/* ---- packing of reference segments ------------------------------------------------------------------------------ */
/* CSegment::store_in_archive(ref) up to the zstd call (src/common/segment.h:218-255): periodicity probe and
 * bytes2tuples (73-138).  For group i: out_use_tuples[i] = 1 -> payload = tuples (zstd level 13, marker 1),
 * 0 -> payload = raw symbols (zstd level 19, marker 0).  Payload i = out[out_offsets[i] .. out_offsets[i+1]). */
int agcgpu_pack_ref_batch(agcgpu_ctx* ctx, const uint32_t* group_ids, uint32_t n, uint8_t* out, uint64_t out_cap,
                          uint64_t* out_offsets, uint8_t* out_use_tuples);

/* ---- residual coder ---------------------------------------------------------------------------------------------- */
/* ZSTD_compressCCtx(cctx, dst, cap, src, n, level) of the vendored zstd 1.5.5 (3rd_party/zstd/lib/compress/
 * zstd_compress.c:5317) for a batch of independent inputs, as called from add_to_archive / add_to_archive_tuples
 * (src/common/segment.h:176,201) and CCollection_V3::zstd_compress (src/common/collection_v3.cpp:139).
 * Input i = src[src_offsets[i] .. src_offsets[i+1]), compressed at levels[i]; frame i = dst[dst_offsets[i] ..
 * dst_offsets[i+1]).  Frames are byte-identical to libzstd 1.5.5's. */
int agcgpu_zstd_compress_batch(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* src_offsets, const int32_t* levels,
                               uint32_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* dst_offsets);
/* The same coder, asynchronous -- the reference codes its packs on worker threads behind the segment queue
 * (src/core/agc_compressor.cpp:1093-1272: store_segments -> CSegment::add -> add_to_archive) while the next contigs are read;
 * here a batch is queued on streams of its own and the call returns once the inputs are on the device, so a pack is coded
 * while the host and the library stream go on with the following samples.  agcgpu_zstd_collect waits for every batch
 * submitted since the last collect and returns their frames in submission order (n_expected = their total number of
 * inputs; frame i = dst[dst_offsets[i] .. dst_offsets[i+1])).  With a communicator (agcgpu_comm_init) every rank submits
 * the same batches and codes its share of each; collect all-gathers the frames once. */
int agcgpu_zstd_submit(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* src_offsets, const int32_t* levels, uint32_t n);
/* agcgpu_zstd_submit with one pointer per input: input i = ptrs[i][0 .. sizes[i]) (a pack queue holds separate buffers) */
int agcgpu_zstd_submit_parts(agcgpu_ctx* ctx, const uint8_t* const* ptrs, const uint64_t* sizes, const int32_t* levels, uint32_t n);
int agcgpu_zstd_collect(agcgpu_ctx* ctx, uint32_t n_expected, uint8_t* dst, uint64_t dst_cap, uint64_t* dst_offsets);

/* ZSTD_decompressDCtx (3rd_party/zstd/lib/decompress) for a batch of independent frames, as CSegment::unpack / get
 * (src/common/segment.cpp:500-577, 220-399) and CCollection_V3 call it: the decode side of the residual coder, used by the
 * decode-and-compare self check (agcgpu_compressor_set_verify).  Frame i = src[src_offsets[i] .. src_offsets[i+1]); every frame
 * header must carry its content size (all frames of ZSTD_compressCCtx do); output i = dst[dst_offsets[i] .. dst_offsets[i+1]).
 * dst_offsets is filled even when dst_cap is too small (AGCGPU_EOVERFLOW), so a caller can size the buffer with one call. */
int agcgpu_zstd_decompress_batch(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* src_offsets, uint32_t n, uint8_t* dst,
                                 uint64_t dst_cap, uint64_t* dst_offsets);

/* ---- CAGCCompressor facade (src/core/agc_compressor.h:754-763) for non-C++ callers -------------------------------- */
typedef struct agcgpu_compressor agcgpu_compressor;
/* CAGCCompressor::Create; dump_parts_path (may be NULL) = test hook writing every part's pre-zstd content */
int agcgpu_compressor_create(const char* out_file, uint32_t pack_cardinality, uint32_t kmer_length, const char* reference_file,
                             uint32_t segment_size, uint32_t min_match_len, int concatenated_genomes, int adaptive_compression,
                             uint32_t verbosity, uint32_t no_threads, double fallback_frac, int device,
                             const char* dump_parts_path, agcgpu_compressor** out);
/* CAGCCompressor::Append (src/core/agc_compressor.cpp:2330-2374): continue the archive in_archive into out_file; afterwards
 * add_sample_files / close as after create. */
int agcgpu_compressor_append(const char* in_archive, const char* out_file, uint32_t verbosity, int prefetch_archive, int concatenated_genomes,
                             int adaptive_compression, uint32_t no_threads, double fallback_frac, int device, agcgpu_compressor** out);
/* CAGCCompressor::AddSampleFiles */
int agcgpu_compressor_add_sample_files(agcgpu_compressor* c, const char* const* sample_names, const char* const* file_names,
                                       uint32_t n, uint32_t no_threads);
/* AddSampleFiles for contigs already in memory: contig i = raw[offsets[i] .. offsets[i+1]) (raw FASTA body, as
 * CGenomeIO::ReadContigRaw returns it) of sample sample_of_contig[i] (samples contiguous, in order).  raw is a host
 * pointer, or a device pointer when raw_is_device != 0 (16-byte aligned, readable up to offsets[n] rounded up to 16). */
int agcgpu_compressor_add_samples_memory(agcgpu_compressor* c, const char* const* sample_names, uint32_t n_samples,
                                         const uint32_t* sample_of_contig, const char* const* contig_ids, uint32_t n_contigs,
                                         const void* raw, const uint64_t* offsets, int raw_is_device);
/* ---- multi-GPU: one process per GPU -----------------------------------------------------------------------
 * Every rank makes the same Create / AddSampleFiles / Close calls on the same inputs.  The host bookkeeping (O(#segments))
 * and the cheap device passes (ingest, scan, hash-assign, reference indexes) are replicated, so every rank holds the same
 * splitter set, segment map and reference store without an exchange; the per-base work whose results do not depend on where
 * they are computed -- LZ-diff encoding of the segments (CSegment::add -> CLZDiff_V2::Encode) and the residual coding of the
 * parts (ZSTD_compressCCtx) -- is split across the ranks and the results are all-gathered, the one exchange step of the path.
 * Rank 0 writes the archive (byte-identical to the single-GPU one); the other ranks write nothing.
 * `allgather` must gather `bytes` bytes from every rank into recv (rank r's block at recv + r*bytes) on all ranks -- e.g.
 * ncclAllGather / torch.distributed.all_gather_into_tensor (agc_b200/dist.py).  Process-wide; applies to compressors created
 * afterwards; world <= 1 or allgather == NULL switches it off. */
typedef int (*agcgpu_allgather_fn)(void* user, const void* send, void* recv, uint64_t bytes);
int agcgpu_set_exchange(uint32_t rank, uint32_t world, agcgpu_allgather_fn allgather, void* user);

/* The exchange over NCCL, inside the library (agc_b200/csrc/comm.cu): one communicator per process, device-to-device
 * ncclAllGather on the library's stream.  One rank calls agcgpu_comm_unique_id and hands the 128 bytes to the others over any
 * side channel (torch.distributed broadcast, MPI, a file); every rank then calls agcgpu_comm_init, after which compressors
 * created in this process split LZ-diff encoding and residual coding across the ranks: the deltas / frames a rank produced stay
 * in HBM, are all-gathered there and come to the host once.  Takes precedence over agcgpu_set_exchange. */
#define AGCGPU_UNIQUE_ID_BYTES 128
int agcgpu_comm_unique_id(uint8_t* out_id);
int agcgpu_comm_init(uint32_t rank, uint32_t world, const uint8_t* id, int device);
int agcgpu_comm_destroy(void);
typedef struct {
    uint32_t nranks;             /* ncclCommCount of the live communicator (1 without one) */
    uint32_t rank;
    uint64_t collectives;        /* ncclAllGather calls issued by the data path since agcgpu_comm_init */
    uint64_t bytes_gathered;     /* bytes received through them on this rank */
} agcgpu_comm_stats;
int agcgpu_comm_get_stats(agcgpu_comm_stats* out);
int agcgpu_comm_world(void);     /* ranks of the live communicator, 1 without one */
int agcgpu_comm_rank(void);
/* agcgpu_lz_encode_batch / agcgpu_zstd_compress_batch with the work split across the ranks of the communicator: every rank passes
 * the SAME arguments, works on its share (requests: contiguous runs balanced by bases; frames: largest first to the least loaded
 * rank) and receives everything; the results are all-gathered between device buffers.  Without a communicator they are the
 * plain calls. */
int agcgpu_lz_encode_batch_sharded(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_offsets);
int agcgpu_zstd_compress_batch_sharded(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* src_offsets, const int32_t* levels,
                                       uint32_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* dst_offsets);
const char* agcgpu_comm_last_error(void);

/* measurement hook: build every archive part but skip the residual coder and the file writes */
int agcgpu_compressor_set_discard_parts(agcgpu_compressor* c, int discard);
/* self check: every frame the residual coder writes is decoded again on the device and compared with its input before the
 * part goes to the archive; a mismatch fails the call (AGC's own check is a later `agc getset`) */
int agcgpu_compressor_set_verify(agcgpu_compressor* c, int verify);
/* CAGCCompressor::AddCmdLine */
int agcgpu_compressor_add_cmd_line(agcgpu_compressor* c, const char* cmd_line);
/* CAGCCompressor::Close, then frees the object */
int agcgpu_compressor_close(agcgpu_compressor* c, uint32_t no_threads);
const char* agcgpu_compressor_last_error(const agcgpu_compressor* c);   /* c may be NULL: last failed create */
uint64_t agcgpu_compressor_total_bases(const agcgpu_compressor* c);
agcgpu_ctx* agcgpu_compressor_ctx(agcgpu_compressor* c);
/* counters of the compressor most recently closed in this process (its context is gone by then); measurement hook */
int agcgpu_compressor_last_stats(agcgpu_stats* out);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
