// compressor.h -- host side of the B200 compression path: mirrors the public surface of the reference's
// CAGCCompressor (src/core/agc_compressor.h:754-763) so that `agc create` can call it unchanged:
//   Create(file, pack_cardinality, k, ref_file, segment_size, min_match_len, concatenated, adaptive, verbosity, threads, fallback_frac)
//   AddSampleFiles(vector<pair<sample,file>>, threads) ; AddCmdLine(string) ; Close(threads)
// bool success, no exceptions, diagnostics to stderr when is_app_mode -- as in the reference.
//
// Everything per-base runs on the GPU through the C ABI (include/agcgpu.h): ingest/2-bit packing, splitter
// determination, splitter scan, hash-assign, LZ index build, LZ-diff encode/estimate/cost vectors, reference tuple
// packing and the zstd-format residual coder.  The host keeps what is O(#segments): canonical ordering
// (CBufferedSegPart), group / in-group id assignment, pack bookkeeping (CSegment), metadata (CCollection_V3) and the
// container (CArchive).  The written .agc is byte-identical to the reference's.
#pragma once
#include <cstdint>
#include <cstdio>
#include <map>
#include <memory>
#include <string>
#include <array>
#include <unordered_map>
#include <unordered_set>
#include <vector>
#include <type_traits>
#include "../../../include/agcgpu.h"

namespace agc_b200 {

// ---- container: src/common/archive.{h,cpp} (output side) ------------------------------------------------------------
class CArchive {
public:
    struct part_t { uint64_t offset, size; };
    struct stream_t { std::string stream_name; uint64_t raw_size = 0; std::vector<part_t> parts; };

    bool Open(const std::string& file_name);
    bool Close();                                             // flush + footer (archive.cpp:68-85,142-169)
    int RegisterStream(const std::string& name);              // archive.cpp:228-262
    bool AddPart(int stream_id, const std::vector<uint8_t>& data, uint64_t metadata);          // immediate (280-293)
    bool AddPartBuffered(int stream_id, std::vector<uint8_t>&& data, uint64_t metadata);       // 332-339
    bool FlushOutBuffers();                                   // ascending stream id, then call order (342-351)
    bool IsOpen() const { return f != nullptr; }
    size_t NoStreams() const { return v_streams.size(); }
    const std::string& StreamName(size_t i) const { return v_streams[i].stream_name; }
private:
    size_t write_varint(uint64_t x);                          // archive.h:110-125: [n][big-endian bytes]
    FILE* f = nullptr;
    bool use_stdout = false;
    uint64_t f_offset = 0;
    std::vector<stream_t> v_streams;
    std::unordered_map<std::string, int> rm_streams;
    std::map<int, std::vector<std::pair<std::vector<uint8_t>, uint64_t>>> m_buffer;
};

// ---- metadata: src/common/collection_v3.{h,cpp} (compression side) -----------------------------------------------
struct segment_desc_t { uint32_t group_id = ~0u, in_group_id = ~0u; bool is_rev_comp = false; uint32_t raw_length = 0; };
struct contig_desc_t { std::string name; std::vector<segment_desc_t> segments; };
struct sample_desc_t { std::string name; std::vector<contig_desc_t> contigs; };

class CCollection_V3 {
public:
    void set_params(uint32_t batch_size, uint32_t segment_size, uint32_t kmer_length);
    bool register_sample_contig(const std::string& sample_name, const std::string& contig_name);   // collection_v3.cpp:706-731
    std::unordered_set<std::string> cur_contig_names, dup_warned;   // contig names of the sample being registered / samples already warned about
    void reset_prev_sample_name() { prev_sample_name.clear(); }
    void add_segment_placed(uint32_t sample_id, uint32_t contig_idx, uint32_t place, const segment_desc_t& d);
    size_t get_no_samples() const { return sample_desc.size(); }
    // raw (pre-zstd) serializations; the caller compresses them (levels 19 / 18 / 5x19) and assembles the parts
    void serialize_sample_names(std::vector<uint8_t>& v) const;                                     // 325-331
    void serialize_contig_names(std::vector<uint8_t>& v, uint32_t id_from, uint32_t id_to) const;   // 468-497
    void serialize_contig_details(std::vector<uint8_t> (&v)[5], uint32_t id_from, uint32_t id_to);  // 539-586
    void clear_batch(uint32_t id_from, uint32_t id_to);
    // append side (collection_v3.cpp:332-345, 498-536, 589-680)
    bool deserialize_sample_names(const std::vector<uint8_t>& v);
    bool deserialize_contig_names(const std::vector<uint8_t>& v, size_t i_sample, uint32_t& n_in_batch);
    bool deserialize_contig_details(const std::vector<uint8_t> (&v)[5], size_t i_sample);
    std::vector<sample_desc_t> sample_desc;
    static void append(std::vector<uint8_t>& data, uint32_t num);                                   // collection.h:126-160
    static void append(std::vector<uint8_t>& data, const std::string& s);
private:
    static std::vector<std::string> split_string(const std::string& s);
    static std::string encode_split(const std::vector<std::string>& prev, const std::vector<std::string>& curr);
    std::unordered_map<std::string, uint32_t> sample_ids;
    std::string prev_sample_name;
    uint32_t batch_size = 50, segment_size = 60000, kmer_length = 31;
    std::vector<int> v_in_group_ids;
};

// ---- FASTA input: src/core/genome_io.cpp:206-250 (plain or gzip through zlib) ----------------------------------------
class CGenomeIO {
public:
    ~CGenomeIO() { Close(); }
    bool Open(const std::string& file_name);
    void Close();
    bool ReadContigRaw(std::string& id, std::vector<uint8_t>& contig);
private:
    bool fill();
    void* gz = nullptr;
    std::vector<uint8_t> buf;
    size_t pos = 0, filled = 0;
    bool at_eof = false;
};

// one zstd job: raw bytes in, frame out (level as the reference call site uses it)
struct ZTask { std::vector<uint8_t> raw; int level; std::vector<uint8_t> packed; };

// a part waiting for its zstd frames; written in (epoch, stream id, seq) order == the reference's flush order
struct PartJob {
    uint64_t epoch; int stream_id; uint64_t seq;
    int kind;                // 0 plain (marker 0), 1 tuples (marker 1), 2 collection part (frame only), 3 collection details (5 frames)
    uint64_t raw_size;       // metadata when the packed form is kept
    uint64_t fallback_size = 0;          // size of the un-coded form (the part's metadata when the coded form is kept)
    std::vector<uint8_t> fallback_raw;   // kinds 0/1: stored as is with metadata 0 when packed+1 >= raw (segment.h:180-187)
    std::vector<ZTask> tasks;
};
static_assert(std::is_nothrow_move_constructible<PartJob>::value, "ZTask pointers stay valid while the job queue grows and is sorted");

class CAGCCompressor {
public:
    CAGCCompressor();
    ~CAGCCompressor();
    bool Create(const std::string& file_name, uint32_t pack_cardinality, uint32_t kmer_length, const std::string& reference_file_name,
                uint32_t segment_size, uint32_t min_match_len, bool concatenated_genomes, bool adaptive_compression,
                uint32_t verbosity, uint32_t no_threads, double fallback_frac);
    // CAGCCompressor::Append (agc_compressor.cpp:2330-2374): continue an existing archive (host/append.cpp)
    bool Append(const std::string& in_archive_fn, const std::string& out_archive_fn, uint32_t verbosity, bool prefetch_archive,
                bool concatenated_genomes, bool adaptive_compression, uint32_t no_threads, double fallback_frac);
    bool AddSampleFiles(std::vector<std::pair<std::string, std::string>> v_sample_file_name, uint32_t no_threads);
    // Same as AddSampleFiles for contigs that are already in memory (raw FASTA bodies as ReadContigRaw returns them):
    // contig i = raw[offsets[i] .. offsets[i+1]) belongs to sample_of_contig[i]; `raw` is a host pointer, or a device
    // pointer when raw_is_device (the HBM-resident entry used by bench.py).  Samples must be contiguous and ordered.
    bool AddSamplesFromMemory(const std::vector<std::string>& sample_names, const std::vector<uint32_t>& sample_of_contig,
                              const std::vector<std::string>& contig_ids, const uint8_t* raw, const uint64_t* offsets, bool raw_is_device);
    void SetDiscardParts(bool d) { discard_parts = d; }              // bench hook: build every part, skip residual coder + file
    void AddCmdLine(const std::string& cmd_line);
    bool Close(uint32_t no_threads = 1);

    // extras of this implementation
    void SetDevice(int dev) { device = dev; }
    void SetAppMode(bool m) { is_app_mode = m; }
    void SetDumpParts(const std::string& path) { dump_path = path; }   // test hook: write every part's pre-zstd content
    void SetBatchBases(uint64_t b) { batch_bases = b; }
    void SetVerify(bool v) { verify = v; }                             // decode every coded frame on the device and compare with its input
    // multi-GPU (include/agcgpu.h, agcgpu_set_exchange): this object is rank `rank` of `world` identical ones
    void SetExchange(uint32_t rank, uint32_t world, agcgpu_allgather_fn fn, void* user) { xrank = rank; xworld = fn && world > 1 ? world : 1; xfn = fn; xuser = user; }
    const std::string& LastError() const { return last_error; }
    void SetLastError(const std::string& e) { last_error = e; }
    uint64_t TotalBases() const { return total_bases; }
    agcgpu_ctx* Ctx() { return ctx; }

private:
    struct Item {            // one buffered segment (CBufferedSegPart::seg_part_t, agc_compressor.h:29-118)
        uint32_t sample_id, contig_idx, seg_part_no;
        uint32_t batch_contig;           // contig index inside the resident device batch
        uint64_t start; uint32_t len; bool is_rc;
        uint64_t k1, k2; int32_t group;  // group < 0: new
        const std::string* contig_name;
    };
    struct GroupState {      // CSegment (src/common/segment.{h,cpp}) minus the sequence data
        uint32_t no_seqs = 0;
        std::vector<std::vector<uint8_t>> pack;      // v_lzp or v_raw
        int stream_ref = -1, stream_delta = -1;
        bool exists = false;
        // append mode: the last pack of the input archive, written back verbatim unless the group is touched (segment.h:283-292)
        std::vector<uint8_t> packed_delta; uint64_t packed_meta = 0; bool packed_pending = false;
        // append mode: loaded but not unpacked yet.  CSegment::estimate / get_coding_cost test ref_size == 0 BEFORE they unpack
        // (segment.cpp:84-86, 101-103), so until the first add() such a group estimates to 0 and returns no cost vector
        bool lazy = false;
        uint32_t ref_size = 0;                       // symbols + 1 (segment.cpp:47)
    };
    struct BatchContig { uint32_t sample_id, contig_idx; uint32_t unit = 0; };   // unit: contigs registered at the same synchronisation token
    void account_registration();

    bool fail(const std::string& msg);
    bool gpu_ok(int rc, const char* what);
    bool process_batch(std::vector<std::vector<uint8_t>>& raws, std::vector<BatchContig>& owners);
    bool process_batch_raw(const uint8_t* cat, bool is_device, const std::vector<uint64_t>& offs, std::vector<BatchContig>& owners);
    bool flush_jobs(bool final_flush);
    bool compress_tasks(std::vector<ZTask*>& tasks);
    bool compress_tasks_local(std::vector<ZTask*>& tasks);
    // -f mode (fallback minimizers)
    bool collect_fallbacks(const std::vector<uint32_t>& contigs);
    void apply_pending_fallbacks();
    bool find_cand_segment_using_fallback_minimizers(uint32_t bc, uint64_t start, uint32_t len, bool rc_view, uint64_t max_val,
                                                     std::pair<uint64_t, uint64_t>& pk, bool& store_rc);
    bool exchange(const std::vector<uint8_t>& mine, std::vector<std::vector<uint8_t>>& all);
    bool lz_encode(std::vector<agcgpu_seg_req>& lz, std::vector<uint8_t>& deltas, std::vector<uint64_t>& doffs);
    bool lz_encode_local(const agcgpu_seg_req* lz, size_t n, std::vector<uint8_t>& deltas, std::vector<uint64_t>& doffs);
    void add_job(PartJob&& j);
    void store_pack(uint32_t group_id, GroupState& g, uint64_t epoch);
    void store_contig_batch(uint32_t id_from, uint32_t id_to, uint64_t epoch);
    std::string ss_base(uint32_t n) const;

    // parameters
    uint32_t pack_cardinality = 50, kmer_length = 31, min_match_len = 20, segment_size = 60000, verbosity = 0;
    bool concatenated_genomes = false, adaptive_compression = false, is_app_mode = true;
    int device = 0;
    uint32_t xrank = 0, xworld = 1; agcgpu_allgather_fn xfn = nullptr; void* xuser = nullptr;
    uint64_t batch_bases = 1ull << 30;
    std::string dump_path, last_error, oversize_error;
    bool discard_parts = false, verify = false;
    FILE* dump_f = nullptr;

    agcgpu_ctx* ctx = nullptr;
    CArchive out_archive;
    CCollection_V3 collection;
    bool working = false, appending = false;
    std::map<std::string, std::string> file_type_info;          // append: as loaded from the input archive (agc_basic.cpp:68-88)

    std::vector<uint64_t> splitters;                          // sorted
    std::map<std::pair<uint64_t, uint64_t>, int32_t> map_segments;            // agc_compressor.h:628
    std::unordered_map<uint64_t, std::vector<uint64_t>> map_segments_terminators;   // 629
    uint64_t fallback_thr = 0;                                                // kmer_filter_t::thr (agc_compressor.h:570-599); 0 = off
    std::unordered_map<uint64_t, std::vector<std::pair<uint64_t, uint64_t>>> map_fallback_minimizers;   // 632
    std::vector<std::array<uint64_t, 4>> pending_fallbacks;                   // vv_fallback_minimizers: applied at the next registration
    std::vector<GroupState> v_segments;
    uint32_t no_segments = 0;
    uint32_t processed_samples = 0;
    uint64_t epoch = 0, job_seq = 0, total_bases = 0;
    std::vector<PartJob> jobs;
    uint8_t* arena = nullptr; uint64_t arena_cap = 0, arena_used = 0;   // page-locked ingest buffer: the raw contigs of the device batch
    bool arena_reserve(uint64_t need);
    bool add_sample_files_arena(std::vector<std::pair<std::string, std::string>>& files);
    uint64_t submitted_job_bytes = 0;                            // pending_job_bytes at the last asynchronous submit
    size_t jobs_submitted = 0;                                   // jobs[0 .. jobs_submitted) are with the device already
    std::vector<ZTask*> inflight;                                // their tasks, in submission order
    bool async_coder() const;
    bool submit_pending(bool with_extra);
    bool collect_inflight();
    std::vector<ZTask*> extra_tasks;                             // coded with the next drain, written by their owner
    uint64_t pending_job_bytes = 0, flush_threshold = 1ull << 30;
    std::vector<std::pair<std::string, std::string>> cmd_lines;
    int collection_samples_id = -1, collection_contig_id = -1, collection_details_id = -1;
};

}  // namespace agc_b200
