// compressor.cpp -- see compressor.h.  Reference citations are relative to /root/reference.
#include "compressor.h"
#include <fcntl.h>
#include <unistd.h>
#include <sys/stat.h>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <zlib.h>
#include <algorithm>
#include <cstring>
#include <iostream>
#include <numeric>
#include <set>

// AGCGPU_TRACE=1: wall time of the host-visible phases on stderr (diagnostics only)
struct PhaseTimer {
    const char* name; std::chrono::steady_clock::time_point t0; bool on;
    explicit PhaseTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()), on(getenv("AGCGPU_TRACE") != nullptr) {}
    ~PhaseTimer() { if (on) fprintf(stderr, "[agcgpu] phase %-22s %8.1f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); }
};
// accumulated timers (AGCGPU_TRACE): total ms and call count per label, printed by Close
struct AccTimers {
    std::map<std::string, std::pair<double, uint64_t>> acc;
    bool on = getenv("AGCGPU_TRACE") != nullptr;
    void report() { if (!on) return; for (auto& kv : acc) fprintf(stderr, "[agcgpu] total %-26s %9.1f ms in %llu calls\n", kv.first.c_str(), kv.second.first, (unsigned long long)kv.second.second); acc.clear(); }
};
static AccTimers g_acc;
struct AccTimer {
    const char* name; std::chrono::steady_clock::time_point t0;
    explicit AccTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()) {}
    ~AccTimer() { if (!g_acc.on) return; auto& a = g_acc.acc[name]; a.first += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); ++a.second; }
};
namespace agc_b200 {

// =====================================================================================================================
// CArchive (output side), src/common/archive.cpp
// =====================================================================================================================
bool CArchive::Open(const std::string& file_name)
{
    if (f) return false;
    use_stdout = file_name.empty();                          // COutFile::Open (src/common/io.h:281-300): no name = stdout
    f = use_stdout ? stdout : fopen(file_name.c_str(), "wb");
    if (!f) return false;
    if (!use_stdout) setvbuf(f, nullptr, _IOFBF, 32 << 20);
    f_offset = 0;
    return true;
}

size_t CArchive::write_varint(uint64_t x)
{
    int no_bytes = 0;
    for (uint64_t tmp = x; tmp; tmp >>= 8) ++no_bytes;
    putc(no_bytes, f);
    for (int i = no_bytes; i; --i) putc((int)((x >> ((i - 1) * 8)) & 0xff), f);
    return (size_t)no_bytes + 1;
}

int CArchive::RegisterStream(const std::string& name)
{
    auto p = rm_streams.find(name);
    if (p != rm_streams.end()) return p->second;
    int id = (int)v_streams.size();
    v_streams.emplace_back();
    v_streams[id].stream_name = name;
    rm_streams[name] = id;
    return id;
}

bool CArchive::AddPart(int stream_id, const std::vector<uint8_t>& data, uint64_t metadata)
{
    v_streams[stream_id].parts.push_back(part_t{ f_offset, data.size() });
    f_offset += write_varint(metadata);
    if (!data.empty()) fwrite(data.data(), 1, data.size(), f);
    f_offset += data.size();
    return true;
}

bool CArchive::AddPartBuffered(int stream_id, std::vector<uint8_t>&& data, uint64_t metadata)
{
    m_buffer[stream_id].emplace_back(std::move(data), metadata);
    return true;
}

bool CArchive::FlushOutBuffers()
{
    for (auto& x : m_buffer)
        for (auto& y : x.second) AddPart(x.first, y.first, y.second);
    m_buffer.clear();
    return true;
}

bool CArchive::Close()
{
    if (!f) return false;
    FlushOutBuffers();
    // footer: archive.cpp:142-169 (raw_size is never updated on the write path: always 0)
    size_t footer_size = 0;
    footer_size += write_varint(v_streams.size());
    for (auto& s : v_streams) {
        fwrite(s.stream_name.data(), 1, s.stream_name.size(), f); putc(0, f);
        footer_size += s.stream_name.size() + 1;
        footer_size += write_varint(s.parts.size());
        footer_size += write_varint(s.raw_size);
        for (auto& p : s.parts) { footer_size += write_varint(p.offset); footer_size += write_varint(p.size); }
    }
    for (int i = 0; i < 8; ++i) putc((int)((footer_size >> (8 * i)) & 0xff), f);     // io.h:371-380 WriteUInt little endian
    bool ok = use_stdout ? fflush(f) == 0 : fclose(f) == 0;
    f = nullptr;
    return ok;
}

// =====================================================================================================================
// CCollection_V3 (compression side), src/common/collection_v3.cpp + collection.h
// =====================================================================================================================
void CCollection_V3::set_params(uint32_t b, uint32_t s, uint32_t k) { batch_size = b; segment_size = s; kmer_length = k; }

void CCollection_V3::append(std::vector<uint8_t>& data, uint32_t num)
{
    const uint32_t thr_1 = 1u << 7, thr_2 = thr_1 + (1u << 14), thr_3 = thr_2 + (1u << 21), thr_4 = thr_3 + (1u << 28);
    if (num < thr_1) data.push_back((uint8_t)num);
    else if (num < thr_2) { num -= thr_1; data.push_back((uint8_t)(0x80u + (num >> 8))); data.push_back((uint8_t)(num & 0xff)); }
    else if (num < thr_3) { num -= thr_2; data.push_back((uint8_t)(0xC0u + (num >> 16))); data.push_back((uint8_t)((num >> 8) & 0xff)); data.push_back((uint8_t)(num & 0xff)); }
    else if (num < thr_4) { num -= thr_3; data.push_back((uint8_t)(0xE0u + (num >> 24))); data.push_back((uint8_t)((num >> 16) & 0xff)); data.push_back((uint8_t)((num >> 8) & 0xff)); data.push_back((uint8_t)(num & 0xff)); }
    else { num -= thr_4; data.push_back(0xF0u); data.push_back((uint8_t)((num >> 24) & 0xff)); data.push_back((uint8_t)((num >> 16) & 0xff)); data.push_back((uint8_t)((num >> 8) & 0xff)); data.push_back((uint8_t)(num & 0xff)); }
}
void CCollection_V3::append(std::vector<uint8_t>& data, const std::string& s)
{
    data.insert(data.end(), s.begin(), s.end());
    data.push_back(0);
}

static std::string extract_contig_name(const std::string& s)      // collection.cpp:17-27
{
    auto p = s.begin();
    for (; p != s.end(); ++p)
        if ((*p < '0') && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) break;
    return std::string(s.begin(), p);
}

bool CCollection_V3::register_sample_contig(const std::string& sample_name, const std::string& contig_name)
{
    std::string stored = sample_name;
    if (sample_name.empty()) stored = extract_contig_name(contig_name);
    if (stored != prev_sample_name) {
        if (sample_ids.count(stored)) return false;
        uint32_t id = (uint32_t)sample_ids.size();
        sample_ids[stored] = id;
        sample_desc.emplace_back();
        sample_desc.back().name = stored;
        prev_sample_name = stored;
        cur_contig_names.clear();
    }
    // Two contigs of one sample with the same name: the reference accepts them, but its segment placement looks contigs up by name
    // (first match) and its result then depends on how std::sort leaves equal keys -- data of the second contig is lost and the
    // bytes are not reproducible.  Same treatment here (so no failure), but said out loud once per sample.
    if (!cur_contig_names.insert(contig_name).second && !dup_warned.count(stored)) {
        dup_warned.insert(stored);
        std::cerr << "Warning: sample " << stored << " holds more than one contig named \"" << contig_name
                  << "\": the archive will not contain all of them (rename the contigs)\n";
    }
    sample_desc.back().contigs.emplace_back();
    sample_desc.back().contigs.back().name = contig_name;
    return true;
}

void CCollection_V3::add_segment_placed(uint32_t sample_id, uint32_t contig_idx, uint32_t place, const segment_desc_t& d)
{
    // add_segments_placed looks the contig up by name and takes the first match (collection_v3.cpp:790-803)
    auto& contigs = sample_desc[sample_id].contigs;
    const std::string& nm = contigs[contig_idx].name;
    for (auto& x : contigs)
        if (x.name == nm) {
            if (place >= x.segments.size()) x.segments.resize((size_t)place + 1);
            x.segments[place] = d;
            break;
        }
}

void CCollection_V3::serialize_sample_names(std::vector<uint8_t>& v) const
{
    append(v, (uint32_t)sample_desc.size());
    for (auto& x : sample_desc) append(v, x.name);
}

std::vector<std::string> CCollection_V3::split_string(const std::string& s)
{
    std::vector<std::string> comp;
    auto p = s.begin();
    while (true) {
        auto q = std::find(p, s.end(), ' ');
        comp.emplace_back(p, q);
        if (q == s.end()) break;
        p = q + 1;
    }
    return comp;
}

std::string CCollection_V3::encode_split(const std::vector<std::string>& prev, const std::vector<std::string>& curr)
{
    std::string enc;
    for (size_t i = 0; i < curr.size(); ++i) {
        if (prev[i] == curr[i]) enc.push_back((char)-127);
        else if (prev[i].size() != curr[i].size()) enc.append(curr[i]);
        else {
            signed char cnt = 0;
            for (size_t j = 0; j < curr[i].size(); ++j) {
                if (prev[i][j] == curr[i][j]) {
                    if (cnt == 100) { enc.push_back((char)-cnt); cnt = 1; } else ++cnt;
                } else {
                    if (cnt) { enc.push_back((char)-cnt); cnt = 0; }
                    enc.push_back(curr[i][j]);
                }
            }
            if (cnt) enc.push_back((char)-cnt);
        }
        enc.push_back(' ');
    }
    enc.pop_back();
    return enc;
}

void CCollection_V3::serialize_contig_names(std::vector<uint8_t>& v, uint32_t id_from, uint32_t id_to) const
{
    append(v, id_to - id_from);
    for (uint32_t s = id_from; s < id_to; ++s) {
        const auto& sd = sample_desc[s];
        append(v, (uint32_t)sd.contigs.size());
        std::vector<std::string> prev_split, curr_split;
        for (auto& x : sd.contigs) {
            curr_split = split_string(x.name);
            if (curr_split.size() != prev_split.size()) append(v, x.name);
            else append(v, encode_split(prev_split, curr_split));
            prev_split = std::move(curr_split);
        }
    }
}

static uint64_t zigzag_encode(uint64_t x_curr, uint64_t x_prev)      // utils.h:123-132
{
    if (x_curr < x_prev) return 2 * (x_prev - x_curr) - 1u;
    if (x_curr < 2 * x_prev) return 2 * (x_curr - x_prev);
    return x_curr;
}

void CCollection_V3::serialize_contig_details(std::vector<uint8_t> (&v)[5], uint32_t id_from, uint32_t id_to)
{
    append(v[0], id_to - id_from);
    v_in_group_ids.clear();
    auto get_igi = [&](uint32_t pos) -> int { return pos >= v_in_group_ids.size() ? -1 : v_in_group_ids[pos]; };
    auto set_igi = [&](uint32_t pos, int val) {
        if (pos >= v_in_group_ids.size()) v_in_group_ids.resize((size_t)((int)(pos * 1.2) + 1), -1);
        v_in_group_ids[pos] = val;
    };
    for (uint32_t s = id_from; s < id_to; ++s) {
        auto& sd = sample_desc[s];
        append(v[0], (uint32_t)sd.contigs.size());
        uint32_t pred_raw_length = segment_size + kmer_length;
        for (auto& x : sd.contigs) {
            append(v[0], (uint32_t)x.segments.size());
            for (auto& seg : x.segments) {
                int prev = get_igi(seg.group_id);
                uint32_t e_in;
                if (prev == -1) e_in = (uint32_t)(int)seg.in_group_id;
                else if (seg.in_group_id == 0) e_in = 0;
                else if ((int)seg.in_group_id == prev + 1) e_in = 1;
                else e_in = (uint32_t)zigzag_encode(seg.in_group_id, (uint64_t)(prev + 1)) + 1u;
                uint32_t e_raw = (uint32_t)zigzag_encode(seg.raw_length, pred_raw_length);
                append(v[1], seg.group_id);
                append(v[2], e_in);
                append(v[3], e_raw);
                append(v[4], (uint32_t)seg.is_rev_comp);
                if ((int)seg.in_group_id > prev && seg.in_group_id > 0) set_igi(seg.group_id, (int)seg.in_group_id);
            }
        }
    }
}

void CCollection_V3::clear_batch(uint32_t id_from, uint32_t id_to)
{
    for (uint32_t s = id_from; s < id_to; ++s) { sample_desc[s].contigs.clear(); sample_desc[s].contigs.shrink_to_fit(); }
}

// =====================================================================================================================
// CGenomeIO: read_contig_raw (src/core/genome_io.cpp:206-250)
// =====================================================================================================================
bool CGenomeIO::Open(const std::string& file_name)
{
    Close();
    gz = gzopen(file_name.c_str(), "rb");
    if (!gz) return false;
    gzbuffer((gzFile)gz, 1 << 20);
    buf.resize(8 << 20);
    pos = filled = 0; at_eof = false;
    return true;
}
void CGenomeIO::Close() { if (gz) { gzclose((gzFile)gz); gz = nullptr; } }
bool CGenomeIO::fill()
{
    if (pos < filled) { memmove(buf.data(), buf.data() + pos, filled - pos); filled -= pos; pos = 0; }
    else pos = filled = 0;
    if (at_eof) return filled != 0;
    int r = gzread((gzFile)gz, buf.data() + filled, (unsigned)(buf.size() - filled));
    if (r <= 0) { at_eof = true; return filled != 0; }
    filled += (size_t)r;
    return filled != 0;
}
bool CGenomeIO::ReadContigRaw(std::string& id, std::vector<uint8_t>& contig)
{
    if (!gz) return false;
    id.clear(); contig.clear();
    while (true) {
        if (pos >= filled) if (!fill()) return false;
        uint8_t c = buf[pos++];
        if (c == '\n' || c == '\r') break;
        id.push_back((char)c);
    }
    if (!id.empty()) id.erase(id.begin());
    while (true) {
        uint8_t* b = buf.data() + pos; uint8_t* e = buf.data() + filled;
        uint8_t* p = (uint8_t*)memchr(b, '>', (size_t)(e - b));
        if (p) { contig.insert(contig.end(), b, p); pos = (size_t)(p - buf.data()); break; }
        contig.insert(contig.end(), b, e);
        pos = filled;
        if (!fill()) break;
    }
    return !id.empty() && !contig.empty();
}

// =====================================================================================================================
// CAGCCompressor
// =====================================================================================================================
static const uint32_t NO_RAW_GROUPS = 16;       // agc_basic.h:79
static const uint64_t EMPTY = ~0ull;

CAGCCompressor::CAGCCompressor() {}
CAGCCompressor::~CAGCCompressor()
{
    if (working) Close(1);
    if (arena) agcgpu_host_free(arena, arena_cap);
    if (ctx) agcgpu_destroy(ctx);
    if (dump_f) fclose(dump_f);
}

bool CAGCCompressor::fail(const std::string& msg)
{
    last_error = msg;
    if (is_app_mode) std::cerr << msg << std::endl;
    return false;
}
bool CAGCCompressor::gpu_ok(int rc, const char* what)
{
    if (rc == 0) return true;
    return fail(std::string(what) + ": " + agcgpu_last_error(ctx));
}

std::string CAGCCompressor::ss_base(uint32_t n) const       // utils.cpp:30-66 (file version 3: "x" + base64)
{
    static const char dig[] = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz_#";
    std::string res = "x";
    do { res.push_back(dig[n & 0x3fu]); n /= 64; } while (n);
    return res;
}

void CAGCCompressor::add_job(PartJob&& j)
{
    j.seq = job_seq++;
    // The device residual coder restates libzstd's index handling for inputs below 1 GiB only (include/agcgpu.h): a part beyond that
    // (a pack of pack_cardinality splitter-less contigs of a large genome) is refused HERE, when it is queued, with a message that
    // says what to do -- not at the drain after all the device work has been done.  flush_jobs reports it.
    for (auto& t : j.tasks)
        if (t.raw.size() >= (1ull << 30) && oversize_error.empty())
            oversize_error = "a part of " + std::to_string(t.raw.size()) + " bytes (stream " + std::to_string(j.stream_id) + ") exceeds the residual coder's 1 GiB input limit: "
                             "lower the pack cardinality (-b) so that a pack of raw contigs stays below 1 GiB";
    for (auto& t : j.tasks) pending_job_bytes += t.raw.size();
    jobs.emplace_back(std::move(j));
}

bool CAGCCompressor::Create(const std::string& file_name, uint32_t _pack_cardinality, uint32_t _kmer_length,
                            const std::string& reference_file_name, uint32_t _segment_size, uint32_t _min_match_len,
                            bool _concatenated_genomes, bool _adaptive_compression, uint32_t _verbosity, uint32_t, double fallback_frac)
{
    if (working) return false;
    pack_cardinality = _pack_cardinality; kmer_length = _kmer_length; min_match_len = _min_match_len;
    segment_size = _segment_size; verbosity = _verbosity;
    concatenated_genomes = _concatenated_genomes; adaptive_compression = _adaptive_compression;
    fallback_thr = fallback_frac == 0.0 ? 0ull : (uint64_t)(((double)~0ull) * fallback_frac);       // kmer_filter_t::reset
    map_fallback_minimizers.clear(); pending_fallbacks.clear(); appending = false;
    agcgpu_params prm; memset(&prm, 0, sizeof prm);
    prm.kmer_length = kmer_length; prm.min_match_len = min_match_len; prm.segment_size = segment_size;
    prm.pack_cardinality = pack_cardinality; prm.device = device;
    prm.flags = adaptive_compression ? AGCGPU_F_ADAPTIVE : 0;      // keeps v_candidate_kmers / v_duplicated_kmers on the device (493-494)
    PhaseTimer pt_create("create+splitters");
    int rc = agcgpu_create(&prm, &ctx);
    if (rc) return fail(std::string("agcgpu_create: ") + agcgpu_last_error(nullptr));

    // determine_splitters (agc_compressor.cpp:428-563)
    {
        CGenomeIO gio;
        if (!gio.Open(reference_file_name)) return fail("Cannot open file: " + reference_file_name);
        std::vector<uint8_t> raw, contig; std::vector<uint64_t> offs{ 0 };
        std::string id;
        while (gio.ReadContigRaw(id, contig)) { raw.insert(raw.end(), contig.begin(), contig.end()); offs.push_back(raw.size()); }
        if (verbosity > 0 && is_app_mode) std::cerr << "Determination of splitters\n";
        uint64_t cap = raw.size() / std::max<uint32_t>(segment_size, 1) + 2 * offs.size() + 64, n = 0;
        splitters.resize(cap);
        if (raw.empty()) raw.push_back(0);
        if (!gpu_ok(agcgpu_determine_splitters(ctx, raw.data(), offs.data(), (uint32_t)offs.size() - 1, splitters.data(), cap, &n), "determine_splitters"))
            return false;
        splitters.resize(n);
        if (verbosity > 1 && is_app_mode) std::cerr << "No. of splitters: " << n << std::endl;
        if (fallback_thr) {                                  // v_fallbacks of find_splitters_in_contig for the reference contigs (797-802, 821-822)
            std::vector<uint32_t> all_ctg(offs.size() - 1);
            std::iota(all_ctg.begin(), all_ctg.end(), 0u);
            if (!collect_fallbacks(all_ctg)) return false;
        }
    }
    if (!dump_path.empty()) {
        dump_f = fopen(dump_path.c_str(), "wb");
        if (!dump_f) return fail("cannot open dump file " + dump_path);
    }
    // multi-GPU: every rank keeps the identical bookkeeping, rank 0 alone owns the output file
    if (agcgpu_comm_world() > 1) xrank = (uint32_t)agcgpu_comm_rank();
    if (!out_archive.Open(xrank == 0 ? file_name : std::string("/dev/null"))) return fail("Cannot create archive " + file_name);
    working = true;
    collection.set_params(pack_cardinality, segment_size, kmer_length);
    collection_samples_id = out_archive.RegisterStream("collection-samples");      // collection_v3.cpp:38-45
    collection_contig_id = out_archive.RegisterStream("collection-contigs");
    collection_details_id = out_archive.RegisterStream("collection-details");
    map_segments[std::make_pair(EMPTY, EMPTY)] = 0;                               // agc_compressor.cpp:2307
    v_segments.resize(NO_RAW_GROUPS);
    for (no_segments = 0; no_segments < NO_RAW_GROUPS; ++no_segments) {             // 2311-2321
        GroupState& g = v_segments[no_segments];
        g.exists = true;
        g.stream_delta = out_archive.RegisterStream(ss_base(no_segments) + "d");
        g.no_seqs = 1;
        g.pack.emplace_back(std::vector<uint8_t>{ 0x7f });
    }
    collection.reset_prev_sample_name();
    epoch = 0; processed_samples = 0;
    return true;
}

void CAGCCompressor::AddCmdLine(const std::string& cmd_line) { cmd_lines.emplace_back(cmd_line, ""); }   // never serialized in v3

// CSegment::store_in_archive(pack) (segment.h:258-280): deltas (or raw sequences) each followed by 0xFF, zstd level 17
void CAGCCompressor::store_pack(uint32_t group_id, GroupState& g, uint64_t ep)
{
    PartJob j; j.epoch = ep; j.kind = 0;
    if (g.stream_delta < 0) g.stream_delta = out_archive.RegisterStream(ss_base(group_id) + "d");
    j.stream_id = g.stream_delta;
    std::vector<uint8_t> pack;
    size_t sz = 0; for (auto& x : g.pack) sz += x.size() + 1;
    pack.reserve(sz);
    for (auto& x : g.pack) { pack.insert(pack.end(), x.begin(), x.end()); pack.push_back(0xff); }
    j.raw_size = pack.size(); j.fallback_size = pack.size();
    j.tasks.emplace_back(); j.tasks[0].level = 17; j.tasks[0].raw = pack;
    j.fallback_raw = std::move(pack);
    add_job(std::move(j));
    g.pack.clear();
}

void CAGCCompressor::store_contig_batch(uint32_t id_from, uint32_t id_to, uint64_t ep)      // collection_v3.cpp:682-703
{
    PartJob jn; jn.epoch = ep; jn.kind = 2; jn.stream_id = collection_contig_id;
    jn.tasks.emplace_back(); jn.tasks[0].level = 18;
    collection.serialize_contig_names(jn.tasks[0].raw, id_from, id_to);
    jn.raw_size = jn.tasks[0].raw.size();
    PartJob jd; jd.epoch = ep; jd.kind = 3; jd.stream_id = collection_details_id; jd.raw_size = 0;
    std::vector<uint8_t> v[5];
    collection.serialize_contig_details(v, id_from, id_to);
    for (int i = 0; i < 5; ++i) { jd.tasks.emplace_back(); jd.tasks[i].level = 19; jd.tasks[i].raw = std::move(v[i]); }
    add_job(std::move(jn));
    add_job(std::move(jd));
    collection.clear_batch(id_from, id_to);
}

// variable-size all-gather over the fixed-size primitive: sizes first, then blocks padded to the largest
bool CAGCCompressor::exchange(const std::vector<uint8_t>& mine, std::vector<std::vector<uint8_t>>& all)
{
    all.assign(xworld, {});
    uint64_t sz = mine.size();
    std::vector<uint64_t> sizes(xworld, 0);
    if (xfn(xuser, &sz, sizes.data(), 8)) return fail("exchange: all-gather of the block sizes failed");
    if (sizes[xrank] != sz) return fail("exchange: all-gather returned a wrong block for this rank");
    uint64_t mx = *std::max_element(sizes.begin(), sizes.end());
    if (mx == 0) return true;
    std::vector<uint8_t> send(mx, 0), recv(mx * xworld);
    if (sz) memcpy(send.data(), mine.data(), sz);
    if (xfn(xuser, send.data(), recv.data(), mx)) return fail("exchange: all-gather failed");
    for (uint32_t r = 0; r < xworld; ++r) all[r].assign(recv.begin() + r * mx, recv.begin() + r * mx + sizes[r]);
    return true;
}

// residual coder over all ranks: the frames of a drain are dealt out by size (largest first, to the least loaded rank), coded
// where they land and all-gathered; every rank ends up with every frame (only rank 0 writes them)
bool CAGCCompressor::compress_tasks(std::vector<ZTask*>& tasks)
{
    if (xworld <= 1 || tasks.empty() || agcgpu_comm_world() > 1) return compress_tasks_local(tasks);
    // longest-processing-time-first: frames sorted by size, each to the rank with the least bytes so far (ties: lowest rank);
    // every rank computes the same assignment from the same task list
    std::vector<size_t> order(tasks.size());
    std::iota(order.begin(), order.end(), (size_t)0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return tasks[a]->raw.size() > tasks[b]->raw.size(); });
    std::vector<uint64_t> load(xworld, 0);
    std::vector<std::vector<size_t>> of_rank(xworld);
    for (size_t i : order) {
        uint32_t r = (uint32_t)(std::min_element(load.begin(), load.end()) - load.begin());
        of_rank[r].push_back(i); load[r] += tasks[i]->raw.size() + 512;     // + a per-frame constant: empty and tiny frames are not free
    }
    std::vector<ZTask*> mine;
    for (size_t i : of_rank[xrank]) mine.push_back(tasks[i]);
    if (!compress_tasks_local(mine)) return false;
    std::vector<uint8_t> blob;
    for (auto* t : mine) { uint64_t n = t->packed.size(); const uint8_t* p = (const uint8_t*)&n; blob.insert(blob.end(), p, p + 8); blob.insert(blob.end(), t->packed.begin(), t->packed.end()); }
    std::vector<std::vector<uint8_t>> all;
    if (!exchange(blob, all)) return false;
    for (uint32_t r = 0; r < xworld; ++r) {
        size_t o = 0;
        for (size_t i : of_rank[r]) {
            uint64_t n;
            if (o + 8 > all[r].size()) return fail("exchange: truncated frame block");
            memcpy(&n, all[r].data() + o, 8); o += 8;
            if (o + n > all[r].size()) return fail("exchange: truncated frame block");
            if (r != xrank) tasks[i]->packed.assign(all[r].begin() + o, all[r].begin() + o + n);
            o += n;
        }
    }
    return true;
}

bool CAGCCompressor::lz_encode_local(const agcgpu_seg_req* lz, size_t n, std::vector<uint8_t>& deltas, std::vector<uint64_t>& doffs)
{
    doffs.assign(n + 1, 0);
    deltas.clear();
    if (!n) return true;
    uint64_t cap = 64; for (size_t i = 0; i < n; ++i) cap += (uint64_t)lz[i].len * 3 / 2 + 32;
    // the ABI wants a buffer large enough for the worst case; typical deltas are ~1% of that, so try small first
    uint64_t try_cap = std::max<uint64_t>(cap / 16, 1 << 20);
    deltas.resize(try_cap);
    int rc2 = agcgpu_lz_encode_batch(ctx, lz, (uint32_t)n, deltas.data(), try_cap, doffs.data());
    if (rc2 == AGCGPU_EOVERFLOW) { deltas.resize(cap); rc2 = agcgpu_lz_encode_batch(ctx, lz, (uint32_t)n, deltas.data(), cap, doffs.data()); }
    return gpu_ok(rc2, "lz_encode_batch");
}

// LZ-diff encoding over all ranks: contiguous runs of requests (they are grouped by reference), balanced by bases
bool CAGCCompressor::lz_encode(std::vector<agcgpu_seg_req>& lz, std::vector<uint8_t>& deltas, std::vector<uint64_t>& doffs)
{
    if (agcgpu_comm_world() > 1 && !lz.empty()) {
        // NCCL communicator inside the library: the split, the device-to-device all-gather and the one copy to the host are
        // agcgpu_lz_encode_batch_sharded's business; every rank gets every delta
        const size_t n = lz.size();
        doffs.assign(n + 1, 0);
        uint64_t cap = 64; for (size_t i = 0; i < n; ++i) cap += (uint64_t)lz[i].len * 3 / 2 + 32;
        uint64_t try_cap = std::max<uint64_t>(cap / 16, 1 << 20);
        deltas.resize(try_cap);
        int rc2 = agcgpu_lz_encode_batch_sharded(ctx, lz.data(), (uint32_t)n, deltas.data(), try_cap, doffs.data());
        if (rc2 == AGCGPU_EOVERFLOW) { deltas.resize(cap); rc2 = agcgpu_lz_encode_batch_sharded(ctx, lz.data(), (uint32_t)n, deltas.data(), cap, doffs.data()); }
        return gpu_ok(rc2, "lz_encode_batch_sharded");
    }
    if (xworld <= 1 || lz.empty()) return lz_encode_local(lz.data(), lz.size(), deltas, doffs);
    std::vector<uint64_t> cum(lz.size() + 1, 0);
    for (size_t i = 0; i < lz.size(); ++i) cum[i + 1] = cum[i] + lz[i].len + 64;
    std::vector<size_t> cutp(xworld + 1, lz.size());
    cutp[0] = 0;
    for (uint32_t r = 1; r < xworld; ++r)
        cutp[r] = std::lower_bound(cum.begin(), cum.end(), cum.back() / xworld * r) - cum.begin();
    for (uint32_t r = 1; r <= xworld; ++r) if (cutp[r] < cutp[r - 1]) cutp[r] = cutp[r - 1];
    cutp[xworld] = lz.size();
    const size_t lo = cutp[xrank], hi = cutp[xrank + 1];
    std::vector<uint8_t> my_d; std::vector<uint64_t> my_o;
    if (!lz_encode_local(lz.data() + lo, hi - lo, my_d, my_o)) return false;
    std::vector<uint8_t> blob((hi - lo + 1) * 8 + my_o.back());
    memcpy(blob.data(), my_o.data(), (hi - lo + 1) * 8);
    if (my_o.back()) memcpy(blob.data() + (hi - lo + 1) * 8, my_d.data(), my_o.back());
    std::vector<std::vector<uint8_t>> all;
    if (!exchange(blob, all)) return false;
    doffs.assign(lz.size() + 1, 0); deltas.clear();
    for (uint32_t r = 0; r < xworld; ++r) {
        const size_t n = cutp[r + 1] - cutp[r];
        if (all[r].size() < (n + 1) * 8) return fail("exchange: truncated delta block");
        const uint64_t* o = (const uint64_t*)all[r].data();
        if (all[r].size() != (n + 1) * 8 + o[n]) return fail("exchange: delta block has the wrong size");
        const uint64_t base = deltas.size();
        for (size_t i = 0; i < n; ++i) doffs[cutp[r] + i + 1] = base + o[i + 1];
        deltas.insert(deltas.end(), all[r].begin() + (n + 1) * 8, all[r].end());
    }
    return true;
}

// The residual coder runs behind the pipeline (agcgpu_zstd_submit / agcgpu_zstd_collect), as the reference's compression threads do.
// Not with the callback exchange of agcgpu_set_exchange (its all-gather is a host call per batch): that path keeps the drain.
bool CAGCCompressor::async_coder() const
{
    static const bool drain_only = getenv("AGCGPU_ZSTD_DRAIN") != nullptr;       // diagnostics: code parts only at the drains
    return !drain_only && (xworld <= 1 || agcgpu_comm_world() > 1);
}

bool CAGCCompressor::submit_pending(bool with_extra)
{
    std::vector<ZTask*> tasks;
    for (size_t i = jobs_submitted; i < jobs.size(); ++i) for (auto& t : jobs[i].tasks) tasks.push_back(&t);
    jobs_submitted = jobs.size();
    if (with_extra) { for (auto* t : extra_tasks) tasks.push_back(t); extra_tasks.clear(); }
    if (tasks.empty()) return true;
    PhaseTimer pt("residual coder submit");
    std::vector<const uint8_t*> ptrs(tasks.size());
    std::vector<uint64_t> sizes(tasks.size());
    std::vector<int32_t> levels(tasks.size());
    for (size_t i = 0; i < tasks.size(); ++i) { ptrs[i] = tasks[i]->raw.data(); sizes[i] = tasks[i]->raw.size(); levels[i] = tasks[i]->level; }
    if (!gpu_ok(agcgpu_zstd_submit_parts(ctx, ptrs.data(), sizes.data(), levels.data(), (uint32_t)tasks.size()), "zstd_submit")) return false;
    inflight.insert(inflight.end(), tasks.begin(), tasks.end());
    return true;
}

bool CAGCCompressor::collect_inflight()
{
    if (inflight.empty()) return true;
    PhaseTimer pt("residual coder collect");
    std::vector<ZTask*> tasks;
    tasks.swap(inflight);
    std::vector<uint64_t> offs(tasks.size() + 1, 0);
    for (size_t i = 0; i < tasks.size(); ++i) offs[i + 1] = offs[i] + tasks[i]->raw.size();
    const uint64_t cap = offs.back() + offs.back() / 128 + 1024 * (tasks.size() + 1);
    std::vector<uint8_t> dst(cap);
    std::vector<uint64_t> doffs(tasks.size() + 1, 0);
    if (!gpu_ok(agcgpu_zstd_collect(ctx, (uint32_t)tasks.size(), dst.data(), cap, doffs.data()), "zstd_collect")) return false;
    for (size_t i = 0; i < tasks.size(); ++i) tasks[i]->packed.assign(dst.begin() + doffs[i], dst.begin() + doffs[i + 1]);
    if (verify) {                                            // decode-and-compare: the frames must give back exactly what went in
        PhaseTimer pv("residual coder self check");
        std::vector<uint8_t> back(offs.back() + 1);
        std::vector<uint64_t> boffs(tasks.size() + 1, 0);
        if (!gpu_ok(agcgpu_zstd_decompress_batch(ctx, dst.data(), doffs.data(), (uint32_t)tasks.size(), back.data(), offs.back(), boffs.data()),
                    "zstd_decompress_batch (self check)")) return false;
        for (size_t i = 0; i < tasks.size(); ++i)
            if (boffs[i + 1] - boffs[i] != tasks[i]->raw.size() || (!tasks[i]->raw.empty() && memcmp(back.data() + boffs[i], tasks[i]->raw.data(), tasks[i]->raw.size()) != 0))
                return fail("self check: frame " + std::to_string(i) + " of a residual-coder batch does not decode to its input");
    }
    return true;
}

bool CAGCCompressor::compress_tasks_local(std::vector<ZTask*>& tasks)
{
    PhaseTimer pt("residual coder batch");
    if (tasks.empty()) return true;
    std::vector<uint64_t> offs(tasks.size() + 1, 0);
    std::vector<int32_t> levels(tasks.size());
    for (size_t i = 0; i < tasks.size(); ++i) { offs[i + 1] = offs[i] + tasks[i]->raw.size(); levels[i] = tasks[i]->level; }
    std::vector<uint8_t> src(offs.back() + 1);
    for (size_t i = 0; i < tasks.size(); ++i) if (!tasks[i]->raw.empty()) memcpy(src.data() + offs[i], tasks[i]->raw.data(), tasks[i]->raw.size());
    uint64_t cap = offs.back() + offs.back() / 128 + 1024 * (tasks.size() + 1);
    std::vector<uint8_t> dst(cap);
    std::vector<uint64_t> doffs(tasks.size() + 1, 0);
    // with an NCCL communicator the frames are dealt out over the ranks and all-gathered between device buffers (every rank passes
    // the same batch); without one this is the plain call
    if (!gpu_ok(agcgpu_zstd_compress_batch_sharded(ctx, src.data(), offs.data(), levels.data(), (uint32_t)tasks.size(), dst.data(), cap, doffs.data()),
                "zstd_compress_batch")) return false;
    for (size_t i = 0; i < tasks.size(); ++i) tasks[i]->packed.assign(dst.begin() + doffs[i], dst.begin() + doffs[i + 1]);
    if (verify) {                                            // decode-and-compare: the frames must give back exactly what went in
        PhaseTimer pv("residual coder self check");
        std::vector<uint8_t> back(offs.back() + 1);
        std::vector<uint64_t> boffs(tasks.size() + 1, 0);
        if (!gpu_ok(agcgpu_zstd_decompress_batch(ctx, dst.data(), doffs.data(), (uint32_t)tasks.size(), back.data(), offs.back(), boffs.data()),
                    "zstd_decompress_batch (self check)")) return false;
        for (size_t i = 0; i < tasks.size(); ++i)
            if (boffs[i + 1] - boffs[i] != tasks[i]->raw.size() || (!tasks[i]->raw.empty() && memcmp(back.data() + boffs[i], tasks[i]->raw.data(), tasks[i]->raw.size()) != 0))
                return fail("self check: frame " + std::to_string(i) + " of a residual-coder batch does not decode to its input");
    }
    return true;
}

static void dump_u64(FILE* f, uint64_t v) { fwrite(&v, 8, 1, f); }
static void dump_bytes(FILE* f, const void* p, size_t n) { dump_u64(f, n); if (n) fwrite(p, 1, n, f); }

// write all pending parts in the reference's flush order: (registration epoch, stream id, call order)
// Parts are independent residual-coder inputs, so they are queued and coded in as few device batches as possible
// (every frame of a batch runs concurrently); the queue is drained when forced or when it holds flush_threshold bytes.
// Registration epochs only grow, so one sorted drain writes the parts in the same order as many small ones would.
bool CAGCCompressor::flush_jobs(bool force)
{
    if (!oversize_error.empty()) return fail(oversize_error);
    if (jobs.empty() && extra_tasks.empty()) return true;
    if (!force && !dump_f && !discard_parts) {
        // not a drain: hand what was queued since the last call to the device (asynchronous: it is coded while the next samples
        // are processed) and go on; the host copies stay until the drain
        // (only once a worthwhile batch has gathered: a submit costs device allocations and a stream of its own, and small parts
        // gain nothing from an early start -- AGCGPU_ZSTD_ASYNC_MIN overrides the 32 MiB)
        static const uint64_t async_min = getenv("AGCGPU_ZSTD_ASYNC_MIN") ? strtoull(getenv("AGCGPU_ZSTD_ASYNC_MIN"), nullptr, 10) : (32ull << 20);
        if (async_coder() && pending_job_bytes - submitted_job_bytes >= async_min) {
            if (!submit_pending(false)) return false;
            submitted_job_bytes = pending_job_bytes;
        }
        if (pending_job_bytes < flush_threshold) return true;
    }
    pending_job_bytes = 0; submitted_job_bytes = 0;
    if (!dump_f && !discard_parts && async_coder() && !submit_pending(true)) return false;       // before the sort moves the jobs around
    std::stable_sort(jobs.begin(), jobs.end(), [](const PartJob& a, const PartJob& b) {
        if (a.epoch != b.epoch) return a.epoch < b.epoch;
        if (a.stream_id != b.stream_id) return a.stream_id < b.stream_id;
        return a.seq < b.seq; });
    if (discard_parts) { jobs.clear(); return true; }
    if (dump_f) {
        for (auto& j : jobs) {
            fwrite("PART", 1, 4, dump_f);
            dump_u64(dump_f, (uint64_t)j.stream_id); dump_u64(dump_f, j.epoch); dump_u64(dump_f, (uint64_t)j.kind); dump_u64(dump_f, j.raw_size);
            dump_u64(dump_f, j.tasks.size());
            for (auto& t : j.tasks) { dump_u64(dump_f, (uint64_t)t.level); dump_bytes(dump_f, t.raw.data(), t.raw.size()); }
            dump_bytes(dump_f, j.fallback_raw.data(), j.fallback_raw.size());
        }
        jobs.clear();
        return true;
    }
    if (async_coder()) {
        // the parts queued since the last submit join the batches already in flight; one collect returns every frame
        if (!submit_pending(true)) return false;
        if (!collect_inflight()) return false;
    } else {
        std::vector<ZTask*> tasks;
        for (auto& j : jobs) for (auto& t : j.tasks) tasks.push_back(&t);
        for (auto* t : extra_tasks) tasks.push_back(t);
        extra_tasks.clear();
        if (!compress_tasks(tasks)) return false;
    }
    jobs_submitted = 0;
    for (auto& j : jobs) {
        if (j.kind == 0 || j.kind == 1) {                       // add_to_archive / add_to_archive_tuples (segment.h:172-215)
            auto& pk = j.tasks[0].packed;
            if ((uint32_t)pk.size() + 1u < (uint32_t)j.fallback_size) {
                pk.push_back((uint8_t)j.kind);
                out_archive.AddPart(j.stream_id, pk, j.fallback_size);
            } else {
                if (j.fallback_raw.size() != j.fallback_size) return fail("internal: raw fallback of a large reference was not fetched");
                out_archive.AddPart(j.stream_id, j.fallback_raw, 0);
            }
        } else if (j.kind == 4) out_archive.AddPart(j.stream_id, j.fallback_raw, j.raw_size);
        else if (j.kind == 2) out_archive.AddPart(j.stream_id, j.tasks[0].packed, j.raw_size);
        else {                                                  // store_batch_contig_details (collection_v3.cpp:225-257)
            std::vector<uint8_t> v;
            for (auto& t : j.tasks) { CCollection_V3::append(v, (uint32_t)t.raw.size()); CCollection_V3::append(v, (uint32_t)t.packed.size()); }
            for (auto& t : j.tasks) v.insert(v.end(), t.packed.begin(), t.packed.end());
            out_archive.AddPart(j.stream_id, v, 0);
        }
    }
    jobs.clear();
    return true;
}

bool CAGCCompressor::AddSampleFiles(std::vector<std::pair<std::string, std::string>> files, uint32_t)
{
    if (!working) return false;
    if (files.empty()) return true;
    if (!concatenated_genomes) return add_sample_files_arena(files);
    std::vector<std::vector<uint8_t>> raws;
    std::vector<BatchContig> owners;
    uint64_t raw_in_batch = 0;
    // -c: every contig is its own sample (named after the contig) and the synchronisation tokens come every
    // pack_cardinality contigs, across files (agc_compressor.cpp:2148-2156, 2180-2199)
    uint32_t cnt_contigs_in_sample = concatenated_genomes ? processed_samples % pack_cardinality : 0, unit = 0;
    for (auto& sf : files) {
        collection.reset_prev_sample_name();
        CGenomeIO gio;
        if (!gio.Open(sf.second)) { std::cerr << "Cannot open file: " << sf.second << std::endl; continue; }
        std::string id; std::vector<uint8_t> contig;
        bool any_read = false, any_added = false;
        while (gio.ReadContigRaw(id, contig)) {
            any_read = true;
            if (concatenated_genomes) {
                if (!collection.register_sample_contig("", id)) { std::cerr << "Error: Pair sample_name:contig_name " << id << ":" << id << " is already in the archive!\n"; continue; }
                uint32_t sid = (uint32_t)collection.sample_desc.size() - 1;
                owners.push_back(BatchContig{ sid, (uint32_t)collection.sample_desc[sid].contigs.size() - 1, unit });
                raw_in_batch += contig.size();
                raws.emplace_back(std::move(contig));
                contig.clear();
                any_added = true;
                if (++cnt_contigs_in_sample >= pack_cardinality) {
                    cnt_contigs_in_sample = 0; ++unit;
                    if (raw_in_batch >= batch_bases || adaptive_compression) { if (!process_batch(raws, owners)) return false; raws.clear(); owners.clear(); raw_in_batch = 0; }
                }
            } else if (collection.register_sample_contig(sf.first, id)) {
                uint32_t sid = (uint32_t)collection.sample_desc.size() - 1;
                owners.push_back(BatchContig{ sid, (uint32_t)collection.sample_desc[sid].contigs.size() - 1, sid });
                raw_in_batch += contig.size();
                raws.emplace_back(std::move(contig));
                contig.clear();
                any_added = true;
            } else std::cerr << "Error: Pair sample_name:contig_name " << sf.first << ":" << id << " is already in the archive!\n";
        }
        if (!any_read) std::cerr << "Warning: Pair sample_name:file_path " << sf.first << ":" << sf.second << " contains no contigs and will not be included in the archive!\n";
        if (!any_added) std::cerr << "Warning: Pair sample_name:file_path " << sf.first << ":" << sf.second << " contains only contigs already present in the archive!\n";
        // -a: the splitter set may grow at every sample's synchronisation point (new_splitters stage, 1187-1229), so a
        // device batch is one sample
        if (!concatenated_genomes && (raw_in_batch >= batch_bases || (adaptive_compression && !raws.empty()))) {
            if (!process_batch(raws, owners)) return false;
            raws.clear(); owners.clear(); raw_in_batch = 0;
        }
    }
    // -c: one more token after the last file (2241-2249), sent even when the last unit is complete -- the registration it
    // triggers is then empty but still runs the processed_samples bookkeeping (1142-1156)
    const bool trailing_empty_unit = concatenated_genomes && (owners.empty() || owners.back().unit != unit);
    if (!raws.empty()) if (!process_batch(raws, owners)) return false;
    if (trailing_empty_unit) { account_registration(); if (!flush_jobs(false)) return false; }
    if (concatenated_genomes) processed_samples = (uint32_t)collection.get_no_samples();        // 2255-2256
    if (processed_samples % pack_cardinality != 0)                                     // agc_compressor.cpp:2258-2259
        store_contig_batch((processed_samples / pack_cardinality) * pack_cardinality, processed_samples, epoch);
    ++epoch;
    return flush_jobs(false);
}

// ---------------------------------------------------------------------------------------------------------------------
// Ingest without -c: the raw contigs of a device batch gather back to back in ONE page-locked buffer (agcgpu_host_alloc) that
// agcgpu_scan_contigs uploads by DMA.  A plain FASTA file is read() straight into that buffer and cut into records in place with
// the rules of CGenomeIO::ReadContigRaw (the bodies slide down over the header lines); gzipped files go through ReadContigRaw.
// ---------------------------------------------------------------------------------------------------------------------
bool CAGCCompressor::arena_reserve(uint64_t need)
{
    if (need <= arena_cap) return true;
    uint64_t cap = 0;
    uint8_t* np = (uint8_t*)agcgpu_host_alloc(std::max<uint64_t>(need, 2 * arena_cap), &cap);
    if (!np) return fail("agcgpu_host_alloc(" + std::to_string(need) + ") failed");
    if (arena_used) memcpy(np, arena, arena_used);
    if (arena) agcgpu_host_free(arena, arena_cap);
    arena = np; arena_cap = cap;
    return true;
}

bool CAGCCompressor::add_sample_files_arena(std::vector<std::pair<std::string, std::string>>& files)
{
    std::vector<BatchContig> owners;
    std::vector<uint64_t> offs(1, 0);
    arena_used = 0;
    auto flush_batch = [&]() -> bool {
        if (owners.empty()) return true;
        const bool ok = process_batch_raw(arena, false, offs, owners);
        owners.clear(); offs.assign(1, 0); arena_used = 0;
        return ok;
    };
    for (auto& sf : files) {
        collection.reset_prev_sample_name();
        bool any_read = false, any_added = false, opened = false;
        auto add_contig = [&](const std::string& id) -> bool {          // the body already lies at arena[offs.back() .. arena_used)
            if (collection.register_sample_contig(sf.first, id)) {
                const uint32_t sid = (uint32_t)collection.sample_desc.size() - 1;
                owners.push_back(BatchContig{ sid, (uint32_t)collection.sample_desc[sid].contigs.size() - 1, sid });
                offs.push_back(arena_used);
                any_added = true;
                return true;
            }
            std::cerr << "Error: Pair sample_name:contig_name " << sf.first << ":" << id << " is already in the archive!\n";
            arena_used = offs.back();                                    // drop the body
            return false;
        };
        int fd = ::open(sf.second.c_str(), O_RDONLY);
        struct stat st;
        uint8_t magic[2] = { 0, 0 };
        const bool plain = fd >= 0 && fstat(fd, &st) == 0 && S_ISREG(st.st_mode) && ::pread(fd, magic, 2, 0) >= 0 && !(magic[0] == 0x1f && magic[1] == 0x8b);
        if (plain) {
            opened = true;
            const uint64_t fsize = (uint64_t)st.st_size, start = arena_used;
            if (!arena_reserve(start + fsize + 64)) { ::close(fd); return false; }
            uint64_t got = 0;
            while (got < fsize) { const ssize_t r = ::read(fd, arena + start + got, (size_t)std::min<uint64_t>(fsize - got, 1ull << 30)); if (r <= 0) break; got += (uint64_t)r; }
            ::close(fd);
            // records in place: rd = read cursor, arena_used = write cursor (never ahead of rd)
            const uint8_t* end = arena + start + got;
            const uint8_t* rd = arena + start;
            std::string id;
            while (true) {
                id.clear();
                bool eof = false;
                while (true) {                                           // header line (ReadContigRaw: up to '\n' or '\r')
                    if (rd >= end) { eof = true; break; }
                    const uint8_t c = *rd++;
                    if (c == '\n' || c == '\r') break;
                    id.push_back((char)c);
                }
                if (eof) break;
                if (!id.empty()) id.erase(id.begin());
                const uint8_t* q = (const uint8_t*)memchr(rd, '>', (size_t)(end - rd));
                const uint8_t* body_end = q ? q : end;
                const uint64_t blen = (uint64_t)(body_end - rd);
                if (id.empty() || blen == 0) break;
                any_read = true;
                memmove(arena + arena_used, rd, blen);
                arena_used += blen;
                add_contig(id);
                rd = body_end;
            }
            if (arena_used < start) arena_used = start;
        } else {
            if (fd >= 0) ::close(fd);
            CGenomeIO gio;
            if (gio.Open(sf.second)) {
                opened = true;
                std::string id; std::vector<uint8_t> contig;
                while (gio.ReadContigRaw(id, contig)) {
                    any_read = true;
                    if (!arena_reserve(arena_used + contig.size() + 64)) return false;
                    memcpy(arena + arena_used, contig.data(), contig.size());
                    arena_used += contig.size();
                    add_contig(id);
                }
            }
        }
        if (!opened) { std::cerr << "Cannot open file: " << sf.second << std::endl; continue; }
        if (!any_read) std::cerr << "Warning: Pair sample_name:file_path " << sf.first << ":" << sf.second << " contains no contigs and will not be included in the archive!\n";
        if (!any_added) std::cerr << "Warning: Pair sample_name:file_path " << sf.first << ":" << sf.second << " contains only contigs already present in the archive!\n";
        // -a: the splitter set may grow at every sample's synchronisation point (new_splitters stage, 1187-1229), so a
        // device batch is one sample
        if (arena_used >= batch_bases || (adaptive_compression && !owners.empty())) if (!flush_batch()) return false;
    }
    if (!flush_batch()) return false;
    if (processed_samples % pack_cardinality != 0)                                     // agc_compressor.cpp:2258-2259
        store_contig_batch((processed_samples / pack_cardinality) * pack_cardinality, processed_samples, epoch);
    ++epoch;
    return flush_jobs(false);
}

// ---------------------------------------------------------------------------------------------------------------------
// One device batch = whole samples.  scan -> (per sample, in order) add_segment + registration -> LZ batch -> bookkeeping
// ---------------------------------------------------------------------------------------------------------------------
bool CAGCCompressor::process_batch(std::vector<std::vector<uint8_t>>& raws, std::vector<BatchContig>& owners)
{
    const uint32_t nc = (uint32_t)raws.size();
    std::vector<uint64_t> offs(nc + 1, 0);
    for (uint32_t i = 0; i < nc; ++i) offs[i + 1] = offs[i] + raws[i].size();
    std::vector<uint8_t> cat(offs[nc] + 1);
    for (uint32_t i = 0; i < nc; ++i) { memcpy(cat.data() + offs[i], raws[i].data(), raws[i].size()); std::vector<uint8_t>().swap(raws[i]); }
    return process_batch_raw(cat.data(), false, offs, owners);
}

// -f mode.  find_splitters_in_contig (agc_compressor.cpp:762-825) remembers, for every splitter it finds, the k-mers that pass
// the fallback filter (and are not their own reverse complement) seen since the previous splitter: v_fallbacks entries
// (previous splitter, this splitter, k-mer, orientation).  The device reports where the splitters of its last
// determine / find_new call were found and the filtered k-mers of the contigs; the intervals are put together here.
bool CAGCCompressor::collect_fallbacks(const std::vector<uint32_t>& contigs)
{
    if (contigs.empty()) return true;
    uint64_t cap = 1024, n = 0;
    std::vector<uint32_t> sc; std::vector<uint64_t> sp, sk; std::vector<uint8_t> sl;
    for (int attempt = 0; attempt < 2; ++attempt) {
        sc.resize(cap); sp.resize(cap); sk.resize(cap); sl.resize(cap);
        int rc = agcgpu_last_splitter_positions(ctx, sc.data(), sp.data(), sk.data(), sl.data(), cap, &n);
        if (rc == AGCGPU_EOVERFLOW && attempt == 0) { cap = n + 16; continue; }
        if (!gpu_ok(rc, "last_splitter_positions")) return false;
        break;
    }
    std::vector<agcgpu_seg_req> rq;
    for (auto c : contigs) { agcgpu_seg_req q; memset(&q, 0, sizeof q); q.contig = c; q.start = 0; q.len = 0xffffffffu; rq.push_back(q); }   // whole contigs
    std::vector<uint64_t> offs(rq.size() + 1, 0);
    std::vector<agcgpu_fkmer> fk(1 << 16);
    int rc = agcgpu_filtered_kmers(ctx, rq.data(), (uint32_t)rq.size(), fallback_thr, fk.data(), fk.size(), offs.data());
    if (rc == AGCGPU_EOVERFLOW) { fk.resize(offs.back() + 16); rc = agcgpu_filtered_kmers(ctx, rq.data(), (uint32_t)rq.size(), fallback_thr, fk.data(), fk.size(), offs.data()); }
    if (!gpu_ok(rc, "filtered_kmers")) return false;
    for (size_t ci = 0; ci < contigs.size(); ++ci) {
        const agcgpu_fkmer* f = fk.data() + offs[ci]; const size_t nf = offs[ci + 1] - offs[ci];
        size_t fi = 0;
        uint64_t prev_splitter = EMPTY, lo = 0;                 // k-mers ending before `lo` were cut off by the reset after a splitter
        bool have_last = false; uint64_t last_kmer = 0;
        for (uint64_t i = 0; i < n; ++i) {
            if (sc[i] != contigs[ci]) continue;
            if (sl[i]) { have_last = true; last_kmer = sk[i]; continue; }
            for (; fi < nf && f[fi].pos <= sp[i]; ++fi)
                if (f[fi].pos >= lo && !f[fi].is_symmetric) pending_fallbacks.push_back({ prev_splitter, sk[i], f[fi].kmer, (uint64_t)f[fi].is_dir_oriented });
            prev_splitter = sk[i]; lo = sp[i] + kmer_length;
        }
        if (have_last)
            for (; fi < nf; ++fi)
                if (f[fi].pos >= lo && !f[fi].is_symmetric) pending_fallbacks.push_back({ prev_splitter, last_kmer, f[fi].kmer, (uint64_t)f[fi].is_dir_oriented });
    }
    return true;
}

// the registration token's add_fallback_mapping pass (agc_compressor.cpp:1121-1128, 413-427)
void CAGCCompressor::apply_pending_fallbacks()
{
    for (auto& x : pending_fallbacks) {
        auto& v = map_fallback_minimizers[x[2]];
        auto to_add = x[3] ? std::make_pair(x[0], x[1]) : std::make_pair(x[1], x[0]);
        if (std::count(v.begin(), v.end(), to_add) == 0) v.push_back(to_add);
    }
    pending_fallbacks.clear();
}

// find_cand_segment_using_fallback_minimizers (agc_compressor.cpp:1812-1963).  The segment is the resident range
// (bc, start, len); rc_view: the caller's "segment" is its reverse complement (call site 1352).
bool CAGCCompressor::find_cand_segment_using_fallback_minimizers(uint32_t bc, uint64_t start, uint32_t len, bool rc_view, uint64_t max_val,
                                                                 std::pair<uint64_t, uint64_t>& pk, bool& store_rc)
{
    const auto pk_empty = std::make_pair(EMPTY, EMPTY);
    pk = pk_empty; store_rc = false;
    const size_t max_num_to_estimate = 10;
    const bool short_segments = segment_size <= 10000;
    agcgpu_seg_req q; memset(&q, 0, sizeof q); q.contig = bc; q.start = start; q.len = len;
    uint64_t offs[2] = { 0, 0 };
    std::vector<agcgpu_fkmer> fk((size_t)((double)len * ((double)fallback_thr / 18446744073709551616.0) * 2.0) + 256);
    int rc = agcgpu_filtered_kmers(ctx, &q, 1, fallback_thr, fk.data(), fk.size(), offs);
    if (rc == AGCGPU_EOVERFLOW) { fk.resize(offs[1] + 16); rc = agcgpu_filtered_kmers(ctx, &q, 1, fallback_thr, fk.data(), fk.size(), offs); }
    if (!gpu_ok(rc, "filtered_kmers")) return false;
    std::map<std::pair<uint64_t, uint64_t>, std::vector<uint64_t>> cand_seg_counts;
    for (uint64_t i = 0; i < offs[1]; ++i) {
        auto p = map_fallback_minimizers.find(fk[i].kmer);
        if (p == map_fallback_minimizers.end()) continue;
        const bool dir_oriented = rc_view ? (fk[i].is_symmetric ? true : !fk[i].is_dir_oriented) : (fk[i].is_dir_oriented != 0);
        for (auto y : p->second)
            if (y.first != EMPTY && y.second != EMPTY) {
                if (!dir_oriented) std::swap(y.first, y.second);
                cand_seg_counts[y].push_back(fk[i].kmer);
            }
    }
    std::vector<std::pair<uint64_t, std::pair<uint64_t, uint64_t>>> pruned;
    for (auto& x : cand_seg_counts) {
        std::sort(x.second.begin(), x.second.end());
        size_t x_size = std::unique(x.second.begin(), x.second.end()) - x.second.begin();
        if (x_size >= max_val) pruned.emplace_back((uint64_t)x_size, x.first);
    }
    if (pruned.empty()) return true;
    std::sort(pruned.begin(), pruned.end(), std::greater<std::pair<uint64_t, std::pair<uint64_t, uint64_t>>>());
    if (pruned.size() > max_num_to_estimate) pruned.resize(max_num_to_estimate);
    while (pruned.back().first * 2 < pruned.front().first) pruned.pop_back();
    auto best_pair = pk_empty;
    uint64_t best_es = len;
    for (auto& x : pruned) {
        const bool is_seg_rc = x.second.first > x.second.second;
        auto p = map_segments.find(is_seg_rc ? std::make_pair(x.second.second, x.second.first) : x.second);
        uint64_t es = 0;
        if (p != map_segments.end()) {                      // can fail if the mappings are to a segment of the same sample
            if (short_segments) { best_pair = x.second; best_es = 0; break; }
            if (v_segments[p->second].lazy) { es = 0; }         // not unpacked yet: CSegment::estimate returns 0
            else {
            agcgpu_seg_req e; memset(&e, 0, sizeof e);
            e.contig = bc; e.start = start; e.len = len; e.is_rc = is_seg_rc != rc_view; e.group_id = (uint32_t)p->second; e.bound = (uint32_t)best_es;
            uint32_t est = 0;
            if (!gpu_ok(agcgpu_lz_estimate_batch(ctx, &e, 1, &est), "lz_estimate")) return false;
            es = est;
            }
        }
        if (es && es < best_es) { best_es = es; best_pair = x.second; }
    }
    if (adaptive_compression) {                             // better left as a new reference (1944-1957)
        if (short_segments) { if ((double)best_es >= len * 0.9) return true; }
        else if ((double)best_es >= len * 0.2) return true;
    }
    if (best_pair.first <= best_pair.second) { pk = best_pair; store_rc = false; }
    else { pk = std::make_pair(best_pair.second, best_pair.first); store_rc = true; }
    return true;
}

// what the registration token does once the segments are stored (agc_compressor.cpp:1136-1179)
void CAGCCompressor::account_registration()
{
    apply_pending_fallbacks();
    if (!concatenated_genomes) ++processed_samples;
    else {
        processed_samples = processed_samples / pack_cardinality * pack_cardinality + pack_cardinality;
        const uint32_t max_ps = (uint32_t)collection.get_no_samples();
        if (max_ps < processed_samples) processed_samples = max_ps;
    }
    if (processed_samples % pack_cardinality == 0) store_contig_batch(processed_samples - pack_cardinality, processed_samples, epoch);
    ++epoch;
}

bool CAGCCompressor::AddSamplesFromMemory(const std::vector<std::string>& sample_names, const std::vector<uint32_t>& sample_of_contig,
                                          const std::vector<std::string>& contig_ids, const uint8_t* raw, const uint64_t* offsets, bool raw_is_device)
{
    if (!working) return false;
    if (concatenated_genomes) return fail("agc-b200: -c is only implemented for AddSampleFiles");
    std::vector<BatchContig> owners;
    uint32_t prev = ~0u;
    for (size_t i = 0; i < contig_ids.size(); ++i) {
        if (sample_of_contig[i] != prev) { collection.reset_prev_sample_name(); prev = sample_of_contig[i]; }
        if (!collection.register_sample_contig(sample_names[sample_of_contig[i]], contig_ids[i]))
            return fail("Error: Pair sample_name:contig_name " + sample_names[sample_of_contig[i]] + ":" + contig_ids[i] + " is already in the archive!");
        uint32_t sid = (uint32_t)collection.sample_desc.size() - 1;
        owners.push_back(BatchContig{ sid, (uint32_t)collection.sample_desc[sid].contigs.size() - 1, sid });
    }
    std::vector<uint64_t> offs(offsets, offsets + contig_ids.size() + 1);
    if (!adaptive_compression) { if (!process_batch_raw(raw, raw_is_device, offs, owners)) return false; }
    else {                                                           // one device batch per sample (see AddSampleFiles)
        for (size_t a = 0; a < owners.size();) {
            size_t b = a;
            while (b < owners.size() && owners[b].sample_id == owners[a].sample_id) ++b;
            std::vector<BatchContig> ow(owners.begin() + a, owners.begin() + b);
            std::vector<uint64_t> of(offs.begin() + a, offs.begin() + b + 1);
            const uint8_t* base = raw;
            if (!raw_is_device) { base = raw + of[0]; const uint64_t o0 = of[0]; for (auto& x : of) x -= o0; }   // device pointers keep their alignment
            if (!process_batch_raw(base, raw_is_device, of, ow)) return false;
            a = b;
        }
    }
    if (processed_samples % pack_cardinality != 0)
        store_contig_batch((processed_samples / pack_cardinality) * pack_cardinality, processed_samples, epoch);
    ++epoch;
    return flush_jobs(false);
}

bool CAGCCompressor::process_batch_raw(const uint8_t* cat, bool is_device, const std::vector<uint64_t>& offs, std::vector<BatchContig>& owners)
{
    const uint32_t nc = (uint32_t)owners.size();
    std::vector<uint64_t> clen(nc + 1);
    uint64_t cap_cuts = offs[nc] / std::max<uint32_t>(segment_size / 4, 16) + 4ull * nc + 64, n_cuts = 0;
    std::vector<agcgpu_cut> cuts(cap_cuts);
    auto do_scan = [&]() {
        return is_device ? agcgpu_scan_contigs_dev(ctx, cat, offs[nc], offs.data(), nc, clen.data(), cuts.data(), cap_cuts, &n_cuts)
                         : agcgpu_scan_contigs(ctx, cat, offs.data(), nc, clen.data(), cuts.data(), cap_cuts, &n_cuts); };
    int rc; { PhaseTimer pt("scan_contigs"); rc = do_scan(); }
    if (rc == AGCGPU_EOVERFLOW && n_cuts > cap_cuts) { cap_cuts = n_cuts + 16; cuts.resize(cap_cuts); rc = do_scan(); }
    if (!gpu_ok(rc, "scan_contigs")) return false;
    cuts.resize(n_cuts);
    for (uint32_t i = 0; i < nc; ++i) total_bases += clen[i];

    // -a mode.  compress_contig (agc_compressor.cpp:2038-2044) sets aside every contig in which the scan met no splitter
    // (a contig with a hit always yields >= 2 cuts) and, if it is at least segment_size long, looks for new splitters in it
    // (find_new_splitters, 2054-2082).  At the sample's new_splitters token (1187-1229) they join the splitter set and
    // the contigs set aside are compressed again (hard_contigs stage); the other contigs keep their first segmentation.
    if (adaptive_compression) {
        std::vector<uint32_t> n_cuts_of(nc, 0);
        for (auto& c : cuts) ++n_cuts_of[c.contig];
        std::vector<uint32_t> hard, searched;
        uint64_t cap_new = 16;
        for (uint32_t c = 0; c < nc; ++c) if (n_cuts_of[c] <= 1) {
            hard.push_back(c);
            if (clen[c] >= segment_size) { searched.push_back(c); cap_new += clen[c] / std::max<uint32_t>(segment_size, 1) + 2; }
        }
        bool grown = false;
        if (!searched.empty()) {
            PhaseTimer pt("find_new_splitters");
            std::vector<uint64_t> fresh_spl(cap_new); uint64_t n_new = 0;
            if (!gpu_ok(agcgpu_find_new_splitters(ctx, searched.data(), (uint32_t)searched.size(), fresh_spl.data(), cap_new, &n_new), "find_new_splitters")) return false;
            fresh_spl.resize(n_new);
            if (fallback_thr && !collect_fallbacks(searched)) return false;
            std::vector<uint64_t> merged; merged.reserve(splitters.size() + n_new);
            std::set_union(splitters.begin(), splitters.end(), fresh_spl.begin(), fresh_spl.end(), std::back_inserter(merged));
            if (merged.size() != splitters.size()) {
                grown = true; splitters.swap(merged);
                if (!gpu_ok(agcgpu_set_splitters(ctx, splitters.data(), splitters.size()), "set_splitters")) return false;
                if (verbosity > 1 && is_app_mode) std::cerr << "No. of splitters: " << splitters.size() << std::endl;
            }
        }
        if (grown) {
            uint64_t cap2 = cap_cuts, n2 = 0;
            std::vector<agcgpu_cut> again(cap2);
            int rc2 = agcgpu_rescan_contigs(ctx, again.data(), cap2, &n2);
            if (rc2 == AGCGPU_EOVERFLOW && n2 > cap2) { cap2 = n2 + 16; again.resize(cap2); rc2 = agcgpu_rescan_contigs(ctx, again.data(), cap2, &n2); }
            if (!gpu_ok(rc2, "rescan_contigs")) return false;
            again.resize(n2);
            std::vector<uint8_t> is_hard(nc, 0);
            for (auto c : hard) is_hard[c] = 1;
            std::vector<agcgpu_cut> mixed; mixed.reserve(cuts.size() + again.size());
            size_t a = 0, b = 0;
            for (uint32_t c = 0; c < nc; ++c) {
                for (; a < cuts.size() && cuts[a].contig == c; ++a) if (!is_hard[c]) mixed.push_back(cuts[a]);
                for (; b < again.size() && again[b].contig == c; ++b) if (is_hard[c]) mixed.push_back(again[b]);
            }
            cuts.swap(mixed); n_cuts = cuts.size();
        }
    }

    // cut ranges per contig
    std::vector<uint64_t> cut_first(nc + 1, 0);
    { uint64_t p = 0; for (uint32_t c = 0; c < nc; ++c) { cut_first[c] = p; while (p < n_cuts && cuts[p].contig == c) ++p; } cut_first[nc] = p; }

    // device-side hash-assign of every cut against the current map; redone after each registration that changes the map
    std::vector<agcgpu_assign> assign(n_cuts);
    auto run_assign = [&](uint64_t from_cut) -> bool {
        if (from_cut >= n_cuts) return true;
        return gpu_ok(agcgpu_assign_cuts(ctx, cuts.data() + from_cut, n_cuts - from_cut, assign.data() + from_cut), "assign_cuts");
    };
    { PhaseTimer pt("assign_cuts"); if (!run_assign(0)) return false; }
    PhaseTimer pt_rest("add_segment..packs");

    struct RegGroup { uint32_t group; std::vector<Item> items; };
    struct SampleReg { uint32_t sample_id; std::vector<RegGroup> groups; };
    std::vector<SampleReg> regs;

    auto canon = [](uint64_t d, uint64_t r) { return d < r ? d : r; };
    auto seg_req = [&](uint32_t bc, uint64_t start, uint32_t len, bool is_rc, uint32_t group, uint32_t bound) {
        agcgpu_seg_req q; q.contig = bc; q.is_rc = is_rc; q.start = start; q.len = len; q.group_id = group; q.bound = bound; q.reserved = 0; return q; };

    // find_cand_segment_with_one_splitter (agc_compressor.cpp:1630-1808).  The CSegment::estimate calls of ALL one-sided
    // segments of the batch are issued as one agcgpu_lz_estimate_batch (re-issued after a registration changed the map).
    struct Cand { uint64_t a, b; bool rc; uint32_t group; };
    auto candidates_of = [&](uint64_t kd, std::vector<Cand>& cands) {
        cands.clear();
        auto p = map_segments_terminators.find(kd);
        if (p == map_segments_terminators.end()) return;
        for (auto ck : p->second) {
            Cand c;
            if (ck < kd) { c.a = ck; c.b = kd; c.rc = true; } else { c.a = kd; c.b = ck; c.rc = false; }
            c.group = (uint32_t)map_segments[std::make_pair(c.a, c.b)];
            cands.push_back(c);
        }
    };
    std::vector<std::vector<uint32_t>> est_cache(n_cuts);
    std::vector<uint8_t> est_valid(n_cuts, 0);
    auto prefetch_estimates = [&](uint64_t from_cut) -> bool {
        std::vector<agcgpu_seg_req> rq; std::vector<uint64_t> owner; std::vector<Cand> cands;
        for (uint64_t x = from_cut; x < n_cuts; ++x) {
            const agcgpu_cut& c = cuts[x];
            est_valid[x] = 0; est_cache[x].clear();
            if (c.has_front == c.has_back) continue;
            const bool front = c.has_front != 0;
            const uint64_t kd = front ? canon(c.front_dir, c.front_rc) : canon(c.back_dir, c.back_rc);
            const bool dir_is_rc = !front;
            const uint32_t len = (uint32_t)c.len;
            const uint32_t bound = len < 16 ? len : len - 16u;
            candidates_of(kd, cands);
            for (auto& cd : cands) { rq.push_back(seg_req(c.contig, c.start, len, cd.rc ? !dir_is_rc : dir_is_rc, cd.group, bound)); owner.push_back(x); }
            est_valid[x] = 1;
        }
        if (rq.empty()) return true;
        std::vector<uint32_t> est(rq.size(), 0);
        // groups reloaded by Append and not unpacked yet estimate to 0 (CSegment::estimate returns before it unpacks, segment.cpp:84-86)
        std::vector<agcgpu_seg_req> live; std::vector<size_t> live_idx;
        for (size_t i = 0; i < rq.size(); ++i) if (!v_segments[rq[i].group_id].lazy) { live.push_back(rq[i]); live_idx.push_back(i); }
        if (!live.empty()) {
            std::vector<uint32_t> le(live.size());
            if (!gpu_ok(agcgpu_lz_estimate_batch(ctx, live.data(), (uint32_t)live.size(), le.data()), "lz_estimate")) return false;
            for (size_t i = 0; i < live.size(); ++i) est[live_idx[i]] = le[i];
        }
        for (size_t i = 0; i < rq.size(); ++i) est_cache[owner[i]].push_back(est[i]);
        return true;
    };
    { AccTimer at("prefetch_estimates(0)"); if (!prefetch_estimates(0)) return false; }
    auto one_splitter = [&](uint64_t x, uint64_t kdir, uint64_t krc, uint32_t len,
                            std::pair<uint64_t, uint64_t>& best_pk, bool& is_best_rc) -> bool {
        const uint64_t kd = canon(kdir, krc);
        const bool dir_oriented = kdir <= krc;
        best_pk = std::make_pair(EMPTY, EMPTY); is_best_rc = false;
        uint64_t best_est = len < 16 ? len : len - 16u;
        std::vector<Cand> cands;
        candidates_of(kd, cands);
        if (!cands.empty()) {
            if (!est_valid[x] || est_cache[x].size() != cands.size()) return fail("internal: estimate cache out of date");
            const std::vector<uint32_t>& est = est_cache[x];
            for (size_t i = 0; i < cands.size(); ++i) if ((uint64_t)est[i] < best_est) best_est = est[i];
            for (size_t i = 0; i < cands.size(); ++i) {
                auto cpk = std::make_pair(cands[i].a, cands[i].b);
                if (est[i] < best_est || (est[i] == best_est && cpk < best_pk) || (est[i] == best_est && cpk == best_pk && !cands[i].rc)) {
                    best_est = est[i]; best_pk = cpk; is_best_rc = cands[i].rc;
                }
            }
        }
        if (best_pk == std::make_pair(EMPTY, EMPTY)) {
            if (dir_oriented) best_pk = std::make_pair(kd, EMPTY);
            else { best_pk = std::make_pair(EMPTY, kd); is_best_rc = true; }
        }
        return true;
    };

    // The decisions of one registration unit are independent of each other (the maps only change at the registration), so all of
    // them go to the device as ONE agcgpu_lz_cost_split_batch call before the unit's add_segment pass; the device returns the
    // split position of each (prefix / suffix sums and the argmin run there), 8 bytes per decision.
    struct SplitPlan { uint64_t mid; uint32_t g1, g2, flags; };
    auto split_plan = [&](uint64_t k1d, uint64_t k2d, bool dir_is_rc, SplitPlan& sp) -> bool {      // false: no shared splitter
        auto pf = map_segments_terminators.find(k1d), pb = map_segments_terminators.find(k2d);
        if (pf == map_segments_terminators.end() || pb == map_segments_terminators.end()) return false;
        std::vector<uint64_t> shared;
        std::set_intersection(pf->second.begin(), pf->second.end(), pb->second.begin(), pb->second.end(), std::back_inserter(shared));
        shared.erase(std::remove(shared.begin(), shared.end(), EMPTY), shared.end());
        if (shared.empty()) return false;
        sp.mid = shared.front();
        sp.g1 = (uint32_t)map_segments[std::minmax(k1d, sp.mid)]; sp.g2 = (uint32_t)map_segments[std::minmax(sp.mid, k2d)];
        sp.flags = 0;
        if (k1d < sp.mid) sp.flags |= (dir_is_rc ? AGCGPU_SPLIT_RC1 : 0u) | AGCGPU_SPLIT_PREFIX1;
        else sp.flags |= (dir_is_rc ? 0u : AGCGPU_SPLIT_RC1) | AGCGPU_SPLIT_REV1;
        if (sp.mid < k2d) sp.flags |= (dir_is_rc ? AGCGPU_SPLIT_RC2 : 0u);
        else sp.flags |= (dir_is_rc ? 0u : AGCGPU_SPLIT_RC2) | AGCGPU_SPLIT_PREFIX2 | AGCGPU_SPLIT_REV2;
        return true;
    };
    std::unordered_map<uint64_t, std::pair<uint64_t, uint32_t>> split_cache;        // cut -> (middle splitter, best position before snapping)
    auto prefetch_splits = [&](uint32_t c_from, uint32_t c_to) -> bool {
        split_cache.clear();
        if (concatenated_genomes) return true;
        AccTimer at("missing_middle (batched)");
        std::vector<agcgpu_split_req> rq; std::vector<uint64_t> owner; std::vector<uint64_t> mids;
        for (uint64_t x = cut_first[c_from]; x < cut_first[c_to]; ++x) {
            const agcgpu_cut& c = cuts[x];
            if (!c.has_front || !c.has_back || assign[x].group_id >= 0) continue;
            const uint64_t fc = canon(c.front_dir, c.front_rc), bcn = canon(c.back_dir, c.back_rc);
            if (fc == bcn || fc == EMPTY || bcn == EMPTY) continue;
            if (!map_segments_terminators.count(fc) || !map_segments_terminators.count(bcn)) continue;
            SplitPlan sp;
            if (!split_plan(std::min(fc, bcn), std::max(fc, bcn), fc > bcn, sp)) continue;
            if (v_segments[sp.g1].lazy || v_segments[sp.g2].lazy) continue;          // append: decided on the host (no cost vector exists)
            agcgpu_split_req q; q.contig = c.contig; q.len = (uint32_t)c.len; q.start = c.start; q.group1 = sp.g1; q.group2 = sp.g2;
            q.flags = sp.flags; q.reserved = 0;
            rq.push_back(q); owner.push_back(x); mids.push_back(sp.mid);
        }
        if (rq.empty()) return true;
        std::vector<uint32_t> bpos(rq.size()), bsum(rq.size());
        if (!gpu_ok(agcgpu_lz_cost_split_batch(ctx, rq.data(), (uint32_t)rq.size(), bpos.data(), bsum.data()), "lz_cost_split_batch")) return false;
        for (size_t i = 0; i < rq.size(); ++i) split_cache[owner[i]] = std::make_pair(mids[i], bpos[i]);
        return true;
    };

    // find_cand_segment_with_missing_middle_splitter (agc_compressor.cpp:1502-1627); dir_is_rc tells which orientation
    // of the resident segment plays "segment_dir"
    auto missing_middle = [&](uint64_t x, uint64_t k1d, uint64_t k2d, uint32_t bc, uint64_t start, uint32_t len, bool dir_is_rc,
                              uint64_t& middle, uint32_t& best_pos) -> bool {
        middle = EMPTY; best_pos = 0;
        auto snap = [&](uint32_t bp, uint32_t n_costs) {          // 1621-1624
            if (bp < kmer_length + 1u) bp = 0;
            if ((size_t)bp + kmer_length + 1u > n_costs) bp = n_costs;
            return bp; };
        auto pc = split_cache.find(x);
        if (pc != split_cache.end()) { middle = pc->second.first; best_pos = snap(pc->second.second, len); return true; }
        SplitPlan sp;
        if (!split_plan(k1d, k2d, dir_is_rc, sp)) return true;
        // a group reloaded by Append and not unpacked yet gives no cost vector (get_coding_cost returns on ref_size == 0, segment.cpp:101-103)
        const bool lazy1 = v_segments[sp.g1].lazy, lazy2 = v_segments[sp.g2].lazy;
        if (lazy1 != lazy2) return true;                          // 1600-1603: vectors of different sizes, no split
        uint32_t bp = 0, n_costs = 0;
        if (!lazy1) {
            AccTimer at("missing_middle (single)");
            agcgpu_split_req q; q.contig = bc; q.len = len; q.start = start; q.group1 = sp.g1; q.group2 = sp.g2; q.flags = sp.flags; q.reserved = 0;
            uint32_t bs = 0;
            if (!gpu_ok(agcgpu_lz_cost_split_batch(ctx, &q, 1, &bp, &bs), "lz_cost_split_batch")) return false;
            n_costs = len;
        }
        middle = sp.mid; best_pos = snap(bp, n_costs);
        return true;
    };

    uint32_t ci = 0;
    while (ci < nc) {
        const uint32_t unit = owners[ci].unit;           // contigs registered together: one sample, or (-c) pack_cardinality contigs
        uint32_t cj = ci;
        while (cj < nc && owners[cj].unit == unit) ++cj;
        std::vector<Item> known, fresh;
        if (!prefetch_splits(ci, cj)) return false;
        // ---- add_segment for every cut of the sample (agc_compressor.cpp:1275-1499)
        for (uint32_t bc = ci; bc < cj; ++bc) {
            uint32_t part_no = 0;
            const uint32_t sid = owners[bc].sample_id;
            const std::string* cname = &collection.sample_desc[sid].contigs[owners[bc].contig_idx].name;
            for (uint64_t x = cut_first[bc]; x < cut_first[bc + 1]; ++x) {
                const agcgpu_cut& cut = cuts[x];
                const agcgpu_assign& as = assign[x];
                Item it; it.sample_id = sid; it.contig_idx = owners[bc].contig_idx; it.seg_part_no = part_no; it.batch_contig = bc;
                it.start = cut.start; it.len = (uint32_t)cut.len; it.contig_name = cname; it.is_rc = false; it.group = -1;
                std::pair<uint64_t, uint64_t> pk(EMPTY, EMPTY);
                bool store_rc = false, have_second = false;
                Item it2 = it;
                const uint64_t fc = canon(cut.front_dir, cut.front_rc), bcn = canon(cut.back_dir, cut.back_rc);
                const auto pk_empty = std::make_pair(EMPTY, EMPTY);
                if (!cut.has_front && !cut.has_back) {
                    pk = pk_empty;
                    if (fallback_thr && !find_cand_segment_using_fallback_minimizers(bc, it.start, it.len, false, 1, pk, store_rc)) return false;   // 1290-1298
                }
                else if (cut.has_front && cut.has_back) { pk = std::make_pair(as.key1, as.key2); store_rc = as.is_rc; }
                else if (cut.has_front) {
                    if (!one_splitter(x, cut.front_dir, cut.front_rc, it.len, pk, store_rc)) return false;
                    if (fallback_thr && (pk.first == EMPTY || pk.second == EMPTY)) {                  // 1322-1336
                        std::pair<uint64_t, uint64_t> pk_alt; bool rc_alt = false;
                        if (!find_cand_segment_using_fallback_minimizers(bc, it.start, it.len, false, 5, pk_alt, rc_alt)) return false;
                        if (pk_alt != pk_empty) { pk = pk_alt; store_rc = rc_alt; }
                    }
                } else {
                    bool store_dir = false;      // kmer = kmer_back with swap_dir_rc; "segment_dir" is the reverse complement
                    if (!one_splitter(x, cut.back_rc, cut.back_dir, it.len, pk, store_dir)) return false;
                    store_rc = !store_dir;
                    if (fallback_thr && (pk.first == EMPTY || pk.second == EMPTY)) {                  // 1347-1361
                        std::pair<uint64_t, uint64_t> pk_alt; bool dir_alt = false;
                        if (!find_cand_segment_using_fallback_minimizers(bc, it.start, it.len, true, 5, pk_alt, dir_alt)) return false;
                        if (pk_alt != pk_empty) { pk = pk_alt; store_rc = !dir_alt; }
                    }
                }
                // map_segments.find(pk): the device hash-assign already answered it for the two-splitter / no-splitter classes
                auto p = map_segments.end();
                bool found = false;
                if (as.klass == 0 || (as.klass == 3 && !fallback_thr)) { found = as.group_id >= 0; if (found) p = map_segments.find(pk); }
                else { p = map_segments.find(pk); found = p != map_segments.end(); }
                if (found && p == map_segments.end()) return fail("internal: device and host segment maps disagree");
                int32_t segment_id = -1, segment_id2 = -1;
                if (!concatenated_genomes && p == map_segments.end() && pk.first != EMPTY && pk.second != EMPTY &&
                    map_segments_terminators.count(pk.first) && map_segments_terminators.count(pk.second)) {
                    if (fc == bcn) { if (!(cut.front_dir <= cut.front_rc)) store_rc = true; }
                    else {
                        uint64_t k1d = fc, k2d = bcn; bool use_rc = false;
                        if (k1d > k2d) { std::swap(k1d, k2d); use_rc = true; }
                        uint64_t mid; uint32_t bp;
                        if (!missing_middle(x, k1d, k2d, bc, it.start, it.len, use_rc, mid, bp)) return false;
                        if (mid != EMPTY) {
                            uint32_t left = bp, right = it.len - bp;
                            if (left == 0) { store_rc = (mid < k2d) ? use_rc : !use_rc; pk = std::minmax(mid, k2d); }
                            else if (right == 0) { store_rc = (k1d < mid) ? use_rc : !use_rc; pk = std::minmax(k1d, mid); }
                            else {
                                if (use_rc) std::swap(left, right);
                                uint32_t seg2_start = left - kmer_length / 2;
                                it2.start = it.start + seg2_start; it2.len = it.len - seg2_start; it2.seg_part_no = part_no + 1;
                                it.len = seg2_start + kmer_length;
                                if (fc < mid) { store_rc = false; pk = std::make_pair(fc, mid); } else { store_rc = true; pk = std::make_pair(mid, fc); }
                                segment_id = map_segments.at(pk);
                                std::pair<uint64_t, uint64_t> pk2;
                                if (mid < bcn) { it2.is_rc = false; pk2 = std::make_pair(mid, bcn); } else { it2.is_rc = true; pk2 = std::make_pair(bcn, mid); }
                                segment_id2 = map_segments.at(pk2);
                                it2.k1 = pk2.first; it2.k2 = pk2.second; it2.group = segment_id2;
                                have_second = true;
                            }
                        }
                    }
                    p = map_segments.find(pk);
                }
                if (p == map_segments.end() && fallback_thr) {                                       // 1461-1477
                    std::pair<uint64_t, uint64_t> pk_fb; bool rc_fb = false;
                    if (!find_cand_segment_using_fallback_minimizers(bc, it.start, it.len, false, 2, pk_fb, rc_fb)) return false;
                    if (pk_fb != pk_empty) { pk = pk_fb; store_rc = rc_fb; p = map_segments.find(pk); }
                }
                it.is_rc = store_rc; it.k1 = pk.first; it.k2 = pk.second;
                if (p == map_segments.end()) { it.group = -1; fresh.push_back(it); }
                else { it.group = have_second ? segment_id : p->second; known.push_back(it); if (have_second) known.push_back(it2); }
                part_no += have_second ? 2 : 1;
            }
        }
        // ---- register_segments (agc_compressor.cpp:954-971): canonical order = (sample, contig name, part) (agc_compressor.h:112-119)
        auto item_less = [](const Item& a, const Item& b) {
            if (*a.contig_name != *b.contig_name) return *a.contig_name < *b.contig_name;
            return a.seg_part_no < b.seg_part_no; };
        std::stable_sort(fresh.begin(), fresh.end(), item_less);
        fresh.erase(std::unique(fresh.begin(), fresh.end(), [](const Item& a, const Item& b) {
            return *a.contig_name == *b.contig_name && a.seg_part_no == b.seg_part_no; }), fresh.end());    // std::set semantics
        SampleReg reg; reg.sample_id = unit;
        std::map<uint32_t, std::vector<Item>> by_group;
        for (auto& it : known) by_group[(uint32_t)it.group].push_back(it);
        for (auto& kv : by_group) std::sort(kv.second.begin(), kv.second.end(), item_less);            // sort_known
        std::map<std::pair<uint64_t, uint64_t>, uint32_t> m_kmers;                                      // process_new (agc_compressor.h:384-415)
        uint32_t group_id = no_segments;
        for (auto& it : fresh) { auto key = std::make_pair(it.k1, it.k2); if (!m_kmers.count(key)) m_kmers[key] = group_id++; }
        const uint32_t no_new = group_id - no_segments;
        for (auto& it : fresh) { it.group = (int32_t)m_kmers[std::make_pair(it.k1, it.k2)]; by_group[(uint32_t)it.group].push_back(it); }
        if (v_segments.size() < group_id) v_segments.resize(group_id);
        std::vector<uint64_t> ins_k1, ins_k2; std::vector<int32_t> ins_g;
        std::vector<agcgpu_seg_req> new_refs;
        for (uint32_t i = 0; i < no_new; ++i) {
            GroupState& g = v_segments[no_segments + i];
            g.stream_ref = out_archive.RegisterStream(ss_base(no_segments + i) + "r");              // RegisterStreams, 960-961
            g.stream_delta = out_archive.RegisterStream(ss_base(no_segments + i) + "d");
        }
        // distribute_segments(0, 0, 16) (agc_compressor.h:417-435): the first n - ceil(n/16) items go round-robin to groups 1..15
        {
            auto g0 = by_group.find(0);
            if (g0 != by_group.end()) {
                std::vector<Item> all = std::move(g0->second), keep;
                uint32_t n = (uint32_t)all.size(), dest = 0, taken = 0;
                for (uint32_t i = 0; i < n; ++i) {
                    if (dest != 0) { Item it = all[taken++]; it.group = (int32_t)dest; by_group[dest].push_back(it); }
                    if (++dest == NO_RAW_GROUPS) dest = 0;
                }
                keep.assign(all.begin() + taken, all.end());
                if (keep.empty()) by_group.erase(0); else by_group[0] = std::move(keep);
            }
        }
        // store_segments part 1 (agc_compressor.cpp:1001-1027): group creation, map_segments / terminators update
        bool map_changed = false;
        for (auto& kv : by_group) {
            GroupState& g = v_segments[kv.first];
            if (g.lazy) { g.lazy = false; map_changed = true; }     // store_segments' add() unpacks the group: estimates of later samples are real
            if (!g.exists) {
                g.exists = true;
                const Item& first = kv.second.front();
                auto key = std::make_pair(first.k1, first.k2);
                auto pm = map_segments.find(key);
                if (pm == map_segments.end()) map_segments[key] = (int32_t)kv.first; else if (pm->second > (int32_t)kv.first) pm->second = (int32_t)kv.first;
                if (first.k1 != EMPTY && first.k2 != EMPTY) {
                    auto& t1 = map_segments_terminators[first.k1]; t1.push_back(first.k2); std::sort(t1.begin(), t1.end());
                    if (first.k1 != first.k2) { auto& t2 = map_segments_terminators[first.k2]; t2.push_back(first.k1); std::sort(t2.begin(), t2.end()); }
                }
                ins_k1.push_back(first.k1); ins_k2.push_back(first.k2); ins_g.push_back((int32_t)kv.first);
                new_refs.push_back(seg_req(first.batch_contig, first.start, first.len, first.is_rc, kv.first, 0));
                g.ref_size = first.len + 1;
                map_changed = true;
            }
            reg.groups.push_back(RegGroup{ kv.first, std::move(kv.second) });
        }
        no_segments += no_new;
        if (map_changed) {
            AccTimer at("map_insert+put_ref+reassign");
            if (!ins_g.empty()) {
                if (!gpu_ok(agcgpu_map_insert(ctx, ins_k1.data(), ins_k2.data(), ins_g.data(), ins_g.size()), "map_insert")) return false;
                if (!gpu_ok(agcgpu_group_put_reference_batch(ctx, new_refs.data(), (uint32_t)new_refs.size()), "put_reference")) return false;
                if (!run_assign(cut_first[cj])) return false;
            }
            if (!prefetch_estimates(cut_first[cj])) return false;
        }
        regs.push_back(std::move(reg));
        apply_pending_fallbacks();                 // the token's add_fallback_mapping pass: visible to the next unit's add_segment calls
        ci = cj;
    }

    // ---- one LZ batch for the whole device batch: everything except each group's first sequence (its reference)
    std::vector<agcgpu_seg_req> lz;
    std::vector<uint32_t> ref_groups;
    {
        std::vector<uint8_t> seen_ref(v_segments.size(), 0);
        for (uint32_t g = 0; g < v_segments.size(); ++g) seen_ref[g] = v_segments[g].no_seqs > 0;
        for (auto& reg : regs) for (auto& rg : reg.groups) {
            if (rg.group < NO_RAW_GROUPS) continue;
            for (auto& it : rg.items) {
                if (!seen_ref[rg.group]) { seen_ref[rg.group] = 1; ref_groups.push_back(rg.group); continue; }
                lz.push_back(seg_req(it.batch_contig, it.start, it.len, it.is_rc, rg.group, 0));
            }
        }
    }
    std::vector<uint8_t> deltas; std::vector<uint64_t> doffs;
    { AccTimer at("lz_encode"); if (!lz_encode(lz, deltas, doffs)) return false; }
    if (verify && !lz.empty()) {                             // decode-and-compare: every delta must give back its segment
        PhaseTimer pv("LZ self check");
        std::vector<uint32_t> gids(lz.size());
        uint64_t cap = 0;
        for (size_t i = 0; i < lz.size(); ++i) { gids[i] = lz[i].group_id; cap += lz[i].len; }
        std::vector<uint8_t> back(cap + 1), sym;
        std::vector<uint64_t> boffs(lz.size() + 1, 0);
        if (!gpu_ok(agcgpu_lz_decode_batch(ctx, gids.data(), deltas.data(), doffs.data(), (uint32_t)lz.size(), back.data(), cap, boffs.data()),
                    "lz_decode_batch (self check)")) return false;
        for (size_t i = 0; i < lz.size(); ++i) {
            if (doffs[i + 1] == doffs[i]) continue;          // empty delta = "equal to the reference" (segment.cpp:61-64)
            sym.resize(lz[i].len ? lz[i].len : 1);
            if (!gpu_ok(agcgpu_get_segment(ctx, lz[i].contig, lz[i].start, lz[i].len, lz[i].is_rc, sym.data()), "get_segment")) return false;
            if (boffs[i + 1] - boffs[i] != lz[i].len || memcmp(back.data() + boffs[i], sym.data(), lz[i].len) != 0)
                return fail("self check: LZ delta " + std::to_string(i) + " (group " + std::to_string(lz[i].group_id) + ") does not decode to its segment");
        }
    }
    // reference payloads (tuples or raw symbols) of the groups created in this batch
    std::vector<uint8_t> refpay; std::vector<uint64_t> roffs(ref_groups.size() + 1, 0); std::vector<uint8_t> ruse(ref_groups.size() + 1, 0);
    if (!ref_groups.empty()) {
        uint64_t cap = 64; for (auto g : ref_groups) cap += v_segments[g].ref_size + 2;
        refpay.resize(cap);
        AccTimer at("pack_ref_batch");
        if (!gpu_ok(agcgpu_pack_ref_batch(ctx, ref_groups.data(), (uint32_t)ref_groups.size(), refpay.data(), cap, roffs.data(), ruse.data()), "pack_ref_batch")) return false;
    }

    // ---- store_segments part 2 (CSegment::add / add_raw, segment.cpp:14-80) in sample order + collection placement
    AccTimer at_store("store_segments part 2");
    size_t lz_i = 0, ref_i = 0;
    for (auto& reg : regs) {
        for (auto& rg : reg.groups) {
            GroupState& g = v_segments[rg.group];
            if (!rg.items.empty()) g.packed_pending = false;        // CSegment::add / add_raw unpack the group (segment.cpp:18-19, 37-38)
            for (auto& it : rg.items) {
                uint32_t in_group_id;
                if (rg.group < NO_RAW_GROUPS) {                                   // add_raw
                    if (g.pack.size() == pack_cardinality) store_pack(rg.group, g, epoch);
                    std::vector<uint8_t> sym(it.len ? it.len : 1);
                    if (!gpu_ok(agcgpu_get_segment(ctx, it.batch_contig, it.start, it.len, it.is_rc, sym.data()), "get_segment")) return false;
                    sym.resize(it.len);
                    ++g.no_seqs; g.pack.emplace_back(std::move(sym));
                    in_group_id = g.no_seqs - 1;
                } else if (g.no_seqs == 0) {                                      // first sequence = reference (segment.cpp:39-48)
                    if (ref_i >= ref_groups.size() || ref_groups[ref_i] != rg.group) return fail("internal: reference order mismatch");
                    PartJob j; j.epoch = epoch; j.stream_id = g.stream_ref; j.kind = ruse[ref_i] ? 1 : 0;
                    j.tasks.emplace_back(); j.tasks[0].level = ruse[ref_i] ? 13 : 19;
                    j.tasks[0].raw.assign(refpay.begin() + roffs[ref_i], refpay.begin() + roffs[ref_i + 1]);
                    j.raw_size = it.len;
                    // fallback ("packed+1 >= raw"): the raw symbols; only tiny references can hit it
                    // (tuples of >= 256 symbols are <= n/4+2 bytes: the zstd frame can never reach n-1 bytes, so no fetch)
                    j.fallback_size = it.len;
                    if (!ruse[ref_i]) j.fallback_raw = j.tasks[0].raw;
                    else if (it.len < 256) { j.fallback_raw.resize(it.len ? it.len : 1); if (!gpu_ok(agcgpu_get_segment(ctx, it.batch_contig, it.start, it.len, it.is_rc, j.fallback_raw.data()), "get_segment")) return false; j.fallback_raw.resize(it.len); }
                    add_job(std::move(j));
                    ++ref_i; ++g.no_seqs;
                    in_group_id = 0;
                } else {
                    if (g.pack.size() == pack_cardinality) store_pack(rg.group, g, epoch);
                    std::vector<uint8_t> delta(deltas.begin() + doffs[lz_i], deltas.begin() + doffs[lz_i + 1]);
                    ++lz_i;
                    if (delta.empty()) in_group_id = 0;                            // IMPROVED_LZ_ENCODING (segment.cpp:61-64)
                    else {
                        auto p = std::find(g.pack.begin(), g.pack.end(), delta);
                        if (p != g.pack.end()) in_group_id = g.no_seqs - (uint32_t)std::distance(p, g.pack.end());
                        else { g.pack.emplace_back(std::move(delta)); ++g.no_seqs; in_group_id = g.no_seqs - 1; }
                    }
                }
                segment_desc_t d; d.group_id = rg.group; d.in_group_id = in_group_id; d.is_rev_comp = it.is_rc; d.raw_length = it.len;
                collection.add_segment_placed(it.sample_id, it.contig_idx, it.seg_part_no, d);
            }
        }
        account_registration();
    }
    return flush_jobs(false);
}

bool CAGCCompressor::Close(uint32_t)
{
    PhaseTimer pt("close");
    g_acc.report();
    if (g_acc.on && ctx) { agcgpu_stats st; if (!agcgpu_get_stats(ctx, &st)) fprintf(stderr, "[agcgpu] LZ encode: %llu segments chunk-parallel, %llu of them redone by the sequential kernel; %u launches, %.1f ms of kernels\n",
        (unsigned long long)st.lz_chunk_segments, (unsigned long long)st.lz_sequential_segments, st.lz_encode_launches, st.lz_kernel_ms_total); }
    if (!working) return false;
    working = false;
    // close_compression (agc_compressor.cpp:2094-2114): CSegment::finish for all groups, flush, metadata
    for (uint32_t i = 0; i < no_segments; ++i) {
        GroupState& g = v_segments[i];
        if (g.packed_pending) {                                  // untouched since Append: store_compressed_delta_in_archive (segment.h:283-292)
            PartJob j; j.epoch = epoch; j.kind = 4; j.stream_id = g.stream_delta; j.raw_size = g.packed_meta; j.fallback_raw = std::move(g.packed_delta);
            add_job(std::move(j));
        } else if (!g.pack.empty()) store_pack(i, g, epoch);
    }
    ++epoch;
    auto a32 = [](std::vector<uint8_t>& v, uint32_t x) { for (int i = 0; i < 4; ++i) { v.push_back(x & 0xff); x >>= 8; } };
    auto a64 = [](std::vector<uint8_t>& v, uint64_t x) { for (int i = 0; i < 8; ++i) { v.push_back(x & 0xff); x >>= 8; } };
    auto astr = [](std::vector<uint8_t>& v, const std::string& s) { v.insert(v.end(), s.begin(), s.end()); v.push_back(0); };
    std::vector<uint8_t> v_params; a32(v_params, kmer_length); a32(v_params, min_match_len); a32(v_params, pack_cardinality); a32(v_params, segment_size);
    std::vector<uint8_t> v_spl; for (auto x : splitters) a64(v_spl, x);                     // already sorted (agc_compressor.cpp:220-226)
    std::vector<uint8_t> v_map;                                                            // std::map iterates sorted (232-252)
    for (auto& x : map_segments) { a64(v_map, x.first.first); a64(v_map, x.first.second); a32(v_map, (uint32_t)x.second); }
    // collection-samples (collection_v3.cpp:155-165), file_type_info (agc_compressor.cpp:286-300, fields 53-59)
    PartJob js; js.epoch = epoch; js.kind = 2; js.stream_id = collection_samples_id; js.tasks.emplace_back(); js.tasks[0].level = 19;
    collection.serialize_sample_names(js.tasks[0].raw); js.raw_size = js.tasks[0].raw.size();
    std::map<std::string, std::string> fti;
    if (appending) fti = file_type_info;                         // store_file_type_info writes what load_file_type_info read
    else {
    fti["producer"] = "agc"; fti["producer_version_major"] = "3"; fti["producer_version_minor"] = "2"; fti["producer_version_build"] = "20260326.1";
    fti["file_version_major"] = "3"; fti["file_version_minor"] = "0";
    fti["comment"] = "AGC (Assembled Genomes Compressor) v. 3.2.2 [build 20260326.1]";
    }
    std::vector<uint8_t> v_fti; for (auto& x : fti) { astr(v_fti, x.first); astr(v_fti, x.second); }
    if (discard_parts) { flush_jobs(true); out_archive.Close(); return true; }
    if (dump_f) {
        if (!flush_jobs(true)) return false;
        auto dump_imm = [&](const char* name, const std::vector<uint8_t>& d, uint64_t meta) {
            fwrite("IMMD", 1, 4, dump_f); dump_bytes(dump_f, name, strlen(name)); dump_u64(dump_f, meta); dump_bytes(dump_f, d.data(), d.size()); };
        dump_imm("params", v_params, 0); dump_imm("splitters", v_spl, splitters.size()); dump_imm("segment-splitters", v_map, map_segments.size());
        dump_imm("file_type_info", v_fti, fti.size());
        add_job(std::move(js)); flush_jobs(true);
        for (const char* nm : { "params", "splitters", "segment-splitters", "file_type_info" }) out_archive.RegisterStream(nm);
        for (size_t i = 0; i < out_archive.NoStreams(); ++i) { fwrite("STRM", 1, 4, dump_f); dump_bytes(dump_f, out_archive.StreamName(i).data(), out_archive.StreamName(i).size()); }
        fclose(dump_f); dump_f = nullptr;
        out_archive.Close();
        return true;
    }
    extra_tasks.push_back(&js.tasks[0]);                        // coded in the same device batch as the last packs
    if (!flush_jobs(true)) return false;
    out_archive.AddPart(out_archive.RegisterStream("params"), v_params, 0);
    out_archive.AddPart(out_archive.RegisterStream("splitters"), v_spl, splitters.size());
    out_archive.AddPart(out_archive.RegisterStream("segment-splitters"), v_map, map_segments.size());
    // collection-samples is *buffered* (AddPartBuffered) while file_type_info is written immediately, so it lands last
    out_archive.AddPart(out_archive.RegisterStream("file_type_info"), v_fti, fti.size());
    out_archive.AddPart(collection_samples_id, js.tasks[0].packed, js.raw_size);
    return out_archive.Close();
}

}  // namespace agc_b200
