// append.cpp -- CAGCCompressor::Append (src/core/agc_compressor.cpp:2330-2374): reload the state of an existing archive so that
// AddSampleFiles / Close continue it.  Mirrors load_file_type_info / load_metadata (src/common/agc_basic.cpp:52-100, 199-246),
// CCollection_V3::prepare_for_appending_copy / prepare_for_appending_load_last_batch (collection_v3.cpp:48-109) with the
// deserializers (332-345, 424-466, 498-536, 589-680), CAGCCompressor::appending_init (303-380) and
// CSegment::appending_init / unpack (src/common/segment.cpp:418-470, 500-577).  Every zstd frame that has to be opened (sample
// names, the last contig batch, the references and the last pack of every group) goes through agcgpu_zstd_decompress_batch in
// two device batches; the references are handed back to the device as LZ references (agcgpu_group_put_reference).
#include "compressor.h"
#include <stdexcept>
#include <algorithm>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>

namespace agc_b200 {

static const uint64_t EMPTY_K = ~0ull;

// ---- input side of src/common/archive.{h,cpp}: footer, streams, parts ----------------------------------------------------
struct InArchive {
    struct Stream { std::string name; std::vector<std::pair<uint64_t, uint64_t>> parts; size_t cursor = 0; };
    std::vector<uint8_t> b;
    std::vector<Stream> streams;
    std::unordered_map<std::string, int> ids;

    bool varint(uint64_t& p, uint64_t& v) const
    {
        if (p >= b.size()) return false;
        uint32_t n = b[p++];
        if (n > 8 || p + n > b.size()) return false;
        v = 0;
        for (uint32_t i = 0; i < n; ++i) v = (v << 8) | b[p++];
        return true;
    }
    bool Open(const std::string& fn)
    {
        FILE* f = fopen(fn.c_str(), "rb");
        if (!f) return false;
        fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
        if (sz < 8) { fclose(f); return false; }
        b.resize((size_t)sz);
        bool ok = fread(b.data(), 1, b.size(), f) == b.size();
        fclose(f);
        if (!ok) return false;
        uint64_t fs = 0;
        for (int i = 0; i < 8; ++i) fs |= (uint64_t)b[b.size() - 8 + i] << (8 * i);
        if (fs + 8 > b.size()) return false;
        uint64_t p = b.size() - 8 - fs, ns;
        if (!varint(p, ns)) return false;
        for (uint64_t s = 0; s < ns; ++s) {
            Stream st;
            while (p < b.size() && b[p]) st.name.push_back((char)b[p++]);
            ++p;
            uint64_t np, raw;
            if (!varint(p, np) || !varint(p, raw)) return false;
            for (uint64_t i = 0; i < np; ++i) { uint64_t off, size; if (!varint(p, off) || !varint(p, size)) return false; st.parts.emplace_back(off, size); }
            ids[st.name] = (int)streams.size();
            streams.push_back(std::move(st));
        }
        return true;
    }
    int GetStreamId(const std::string& name) const { auto p = ids.find(name); return p == ids.end() ? -1 : p->second; }
    size_t GetNoParts(int id) const { return id < 0 ? 0 : streams[id].parts.size(); }
    bool GetPart(int id, size_t idx, std::vector<uint8_t>& data, uint64_t& meta) const
    {
        if (id < 0 || idx >= streams[id].parts.size()) return false;
        uint64_t p = streams[id].parts[idx].first;
        if (!varint(p, meta)) return false;
        uint64_t size = streams[id].parts[idx].second;
        if (p + size > b.size()) return false;
        data.assign(b.begin() + p, b.begin() + p + size);
        return true;
    }
};

// collection.h:163-203 read()
static bool rd_u32(const uint8_t*& p, const uint8_t* e, uint32_t& num)
{
    const uint32_t thr_1 = 1u << 7, thr_2 = thr_1 + (1u << 14), thr_3 = thr_2 + (1u << 21), thr_4 = thr_3 + (1u << 28);
    if (p >= e) return false;
    if ((*p & 0x80u) == 0) { num = *p++; return true; }
    if ((*p & 0xC0u) == 0x80u) { if (p + 2 > e) return false; num = ((uint32_t)p[0] << 8) + p[1] + thr_1 - (0x80u << 8); p += 2; return true; }
    if ((*p & 0xE0u) == 0xC0u) { if (p + 3 > e) return false; num = ((uint32_t)p[0] << 16) + ((uint32_t)p[1] << 8) + p[2] + thr_2 - (0xC0u << 16); p += 3; return true; }
    if ((*p & 0xF0u) == 0xE0u) { if (p + 4 > e) return false; num = ((uint32_t)p[0] << 24) + ((uint32_t)p[1] << 16) + ((uint32_t)p[2] << 8) + p[3] + thr_3 - (0xE0u << 24); p += 4; return true; }
    if (p + 5 > e) return false;
    num = ((uint32_t)p[1] << 24) + ((uint32_t)p[2] << 16) + ((uint32_t)p[3] << 8) + p[4] + thr_4; p += 5;
    return true;
}
static bool rd_str(const uint8_t*& p, const uint8_t* e, std::string& s)
{
    const uint8_t* q = p;
    while (q < e && *q) ++q;
    if (q >= e) return false;
    s.assign((const char*)p, q - p);
    p = q + 1;
    return true;
}
static uint64_t zigzag_decode(uint64_t x_val, uint64_t x_prev)      // utils.h:125-135
{
    if (x_val >= 2 * x_prev) return x_val;
    if (x_val & 1) return (2 * x_prev - x_val) / 2;
    return (x_val + 2 * x_prev) / 2;
}

bool CCollection_V3::deserialize_sample_names(const std::vector<uint8_t>& v)
{
    const uint8_t* p = v.data(); const uint8_t* e = p + v.size();
    uint32_t n;
    if (!rd_u32(p, e, n)) return false;
    sample_desc.assign(n, sample_desc_t());
    for (uint32_t i = 0; i < n; ++i) { if (!rd_str(p, e, sample_desc[i].name)) return false; sample_ids[sample_desc[i].name] = i; }
    return true;
}

// decode_split (collection_v3.cpp:424-466)
static std::string decode_split(std::vector<std::string>& prev, std::vector<std::string>& curr)
{
    std::string dec, cmp;
    for (size_t i = 0; i < curr.size(); ++i) {
        if (i >= prev.size()) throw std::runtime_error("damaged archive: contig name refers to a missing previous name");
        if (curr[i].size() == 1 && (signed char)curr[i].front() == -127) { dec.append(prev[i]); curr[i] = prev[i]; }
        else {
            cmp.clear();
            size_t pp = 0;                                                   // position in prev[i] (a damaged archive may run past its end)
            for (signed char c : curr[i]) {
                if (c >= 0) { cmp.push_back(c); ++pp; }
                else {
                    const size_t cnt = (size_t)(-(int)c);
                    if (pp > prev[i].size() || cnt > prev[i].size() - pp) throw std::runtime_error("damaged archive: contig name copy beyond the previous name");
                    cmp.append(prev[i], pp, cnt); pp += cnt;
                }
            }
            dec.append(cmp);
            curr[i] = cmp;
        }
        dec.push_back(' ');
    }
    if (!dec.empty()) dec.pop_back();
    return dec;
}

bool CCollection_V3::deserialize_contig_names(const std::vector<uint8_t>& v, size_t i_sample, uint32_t& n_in_batch)
{
    const uint8_t* p = v.data(); const uint8_t* e = p + v.size();
    if (!rd_u32(p, e, n_in_batch) || i_sample + n_in_batch > sample_desc.size()) return false;
    for (uint32_t i = 0; i < n_in_batch; ++i) {
        uint32_t nc;
        if (!rd_u32(p, e, nc)) return false;
        auto& s = sample_desc[i_sample + i];
        s.contigs.assign(nc, contig_desc_t());
        std::vector<std::string> prev_split, curr_split;
        std::string enc;
        for (uint32_t j = 0; j < nc; ++j) {
            if (!rd_str(p, e, enc)) return false;
            curr_split = split_string(enc);
            if (curr_split.size() != prev_split.size()) s.contigs[j].name = enc;
            else s.contigs[j].name = decode_split(prev_split, curr_split);
            prev_split = std::move(curr_split);
        }
    }
    return true;
}

bool CCollection_V3::deserialize_contig_details(const std::vector<uint8_t> (&v)[5], size_t i_sample)
{
    const uint8_t* p = v[0].data(); const uint8_t* e = p + v[0].size();
    uint32_t ns;
    if (!rd_u32(p, e, ns) || i_sample + ns > sample_desc.size()) return false;
    size_t no_items = 0;
    for (uint32_t i = 0; i < ns; ++i) {
        uint32_t nc;
        if (!rd_u32(p, e, nc)) return false;
        auto& s = sample_desc[i_sample + i];
        s.contigs.resize(nc);
        for (uint32_t j = 0; j < nc; ++j) { uint32_t nseg; if (!rd_u32(p, e, nseg)) return false; s.contigs[j].segments.assign(nseg, segment_desc_t()); no_items += nseg; }
    }
    std::vector<uint32_t> det[5];
    for (int i = 1; i < 5; ++i) {
        det[i].resize(no_items);
        const uint8_t* q = v[i].data(); const uint8_t* qe = q + v[i].size();
        for (size_t j = 0; j < no_items; ++j) if (!rd_u32(q, qe, det[i][j])) return false;
    }
    v_in_group_ids.clear();
    auto get_igi = [&](uint32_t pos) -> int { return pos >= v_in_group_ids.size() ? -1 : v_in_group_ids[pos]; };
    auto set_igi = [&](uint32_t pos, int val) {
        if (pos >= v_in_group_ids.size()) v_in_group_ids.resize((size_t)((double)pos * 1.2) + 1, -1);
        v_in_group_ids[pos] = val;
    };
    const uint32_t pred_raw_length = segment_size + kmer_length;
    size_t it = 0;
    for (uint32_t i = 0; i < ns; ++i)
        for (auto& c : sample_desc[i_sample + i].contigs)
            for (auto& seg : c.segments) {
                const uint32_t g = det[1][it], e_igi = det[2][it];
                if (g > (1u << 28)) return false;                           // (group ids are checked against the streams once those are loaded)
                seg.group_id = g;
                const int prev = get_igi(g);
                uint32_t igi;
                if (prev == -1) igi = e_igi;
                else if (e_igi == 0) igi = 0;
                else if (e_igi == 1) igi = (uint32_t)(prev + 1);
                else igi = (uint32_t)zigzag_decode(e_igi - 1u, (uint64_t)(prev + 1));
                seg.in_group_id = igi;
                seg.raw_length = (uint32_t)zigzag_decode(det[3][it], pred_raw_length);
                seg.is_rev_comp = det[4][it] != 0;
                if ((int)igi > prev && igi > 0) set_igi(g, (int)igi);
                ++it;
            }
    return true;
}

// tuples2bytes (src/common/segment.h:94-170)
static void tuples2bytes(const std::vector<uint8_t>& t, std::vector<uint8_t>& out)
{
    out.clear();
    if (t.size() < 2) return;
    const uint8_t marker = t.back(), nb = marker >> 4, trailing = marker & 0xf;
    if (nb != 4 && nb != 3 && nb != 2) { out.assign(t.begin(), t.end() - 1); return; }
    const uint32_t mult = nb == 4 ? 4 : (nb == 3 ? 6 : 16);
    const size_t n = (t.size() - 2) * nb + trailing;
    out.resize(n);
    size_t i = 0, j = 0;
    for (; j + nb <= n; ++i, j += nb) { uint8_t c = t[i]; for (int k = nb - 1; k >= 0; --k) { out[j + k] = c % mult; c /= mult; } }
    uint8_t c = t[i];
    const uint32_t rem = (uint32_t)(n % nb);
    for (int k = (int)rem - 1; k >= 0; --k) { out[j + k] = c % mult; c /= mult; }
}

bool CAGCCompressor::Append(const std::string& in_archive_fn, const std::string& out_archive_fn, uint32_t _verbosity, bool /*prefetch_archive*/,
                            bool _concatenated_genomes, bool _adaptive_compression, uint32_t /*no_threads*/, double fallback_frac)
{
    if (working) return false;
    verbosity = _verbosity; concatenated_genomes = _concatenated_genomes; adaptive_compression = _adaptive_compression;
    fallback_thr = fallback_frac == 0.0 ? 0ull : (uint64_t)(((double)~0ull) * fallback_frac);
    map_fallback_minimizers.clear(); pending_fallbacks.clear();
    InArchive in;
    if (!in.Open(in_archive_fn)) return fail("Cannot open archive " + in_archive_fn);
    std::vector<uint8_t> d; uint64_t meta = 0;
    // ---- load_file_type_info (agc_basic.cpp:52-100)
    if (!in.GetPart(in.GetStreamId("file_type_info"), 0, d, meta)) return fail("archive has no file_type_info");
    {   const uint8_t* p = d.data(); const uint8_t* e = p + d.size(); std::string k, v;
        file_type_info.clear();
        for (uint64_t i = 0; i < meta; ++i) { if (!rd_str(p, e, k) || !rd_str(p, e, v)) return fail("bad file_type_info"); file_type_info[k] = v; }
        if (file_type_info["file_version_major"] != "3") return fail("only archives of file version 3.x can be extended"); }
    // ---- load_metadata (agc_basic.cpp:199-246)
    if (!in.GetPart(in.GetStreamId("params"), 0, d, meta) || d.size() < 16) return fail("Archive does not contain parameters section");
    auto u32 = [&](size_t o) { return (uint32_t)d[o] | ((uint32_t)d[o + 1] << 8) | ((uint32_t)d[o + 2] << 16) | ((uint32_t)d[o + 3] << 24); };
    kmer_length = u32(0); min_match_len = u32(4); pack_cardinality = u32(8); segment_size = u32(12);
    agcgpu_params prm; memset(&prm, 0, sizeof prm);
    prm.kmer_length = kmer_length; prm.min_match_len = min_match_len; prm.segment_size = segment_size;
    prm.pack_cardinality = pack_cardinality; prm.device = device;
    prm.flags = adaptive_compression ? AGCGPU_F_ADAPTIVE : 0;
    if (agcgpu_create(&prm, &ctx)) return fail(std::string("agcgpu_create: ") + agcgpu_last_error(nullptr));
    if (agcgpu_comm_world() > 1) xrank = (uint32_t)agcgpu_comm_rank();
    if (!out_archive.Open(xrank == 0 ? out_archive_fn : std::string("/dev/null"))) return fail("Cannot create archive " + out_archive_fn);
    collection.set_params(pack_cardinality, segment_size, kmer_length);

    // frames that have to be opened, decoded in one device batch
    struct Frame { std::vector<uint8_t> packed; uint64_t raw_size; std::vector<uint8_t> raw; };
    auto decode_all = [&](std::vector<Frame*>& fr) -> bool {
        std::vector<uint64_t> so(1, 0); std::vector<uint8_t> src; std::vector<Frame*> todo;
        for (auto* f : fr) {
            if (f->raw_size == 0) { f->raw = f->packed; continue; }
            src.insert(src.end(), f->packed.begin(), f->packed.end()); so.push_back(src.size()); todo.push_back(f);
        }
        if (todo.empty()) return true;
        // sizes come from the frame headers (a reference part's metadata is its symbol count, not the size of its tuples)
        std::vector<uint64_t> dof(todo.size() + 1, 0);
        std::vector<uint8_t> dst(1);
        int rc = agcgpu_zstd_decompress_batch(ctx, src.data(), so.data(), (uint32_t)todo.size(), dst.data(), 0, dof.data());
        if (rc == AGCGPU_EOVERFLOW) {
            uint64_t announced = 0;
            for (auto* f : todo) announced += f->raw_size + 1;
            if (dof.back() > announced + 4 * (uint64_t)todo.size())         // frame headers of a damaged archive may claim anything
                return fail("archive parts announce more decoded bytes than their metadata");
            dst.resize(dof.back() + 1);
            rc = agcgpu_zstd_decompress_batch(ctx, src.data(), so.data(), (uint32_t)todo.size(), dst.data(), dof.back(), dof.data());
        }
        if (!gpu_ok(rc, "zstd_decompress_batch")) return false;
        for (size_t i = 0; i < todo.size(); ++i) {
            if (dof[i + 1] - dof[i] > todo[i]->raw_size + 1) return fail("archive part decodes to more than its metadata announces");
            todo[i]->raw.assign(dst.begin() + dof[i], dst.begin() + dof[i + 1]);
        }
        return true;
    };

    // ---- CCollection_V3::prepare_for_appending_copy (collection_v3.cpp:48-77)
    const int in_samples = in.GetStreamId("collection-samples"), in_contigs = in.GetStreamId("collection-contigs"), in_details = in.GetStreamId("collection-details");
    if (in_samples < 0 || in_contigs < 0 || in_details < 0) return fail("archive has no v3 collection streams");
    collection_samples_id = out_archive.RegisterStream("collection-samples");
    collection_contig_id = out_archive.RegisterStream("collection-contigs");
    collection_details_id = out_archive.RegisterStream("collection-details");
    const size_t no_contig_batches = in.GetNoParts(in_contigs);
    if (no_contig_batches == 0 || in.GetNoParts(in_details) != no_contig_batches) return fail("archive has no contig batches");
    Frame f_samples, f_names, f_det[5];
    if (!in.GetPart(in_samples, in.GetNoParts(in_samples) - 1, f_samples.packed, f_samples.raw_size)) return fail("cannot read collection-samples");
    for (size_t i = 0; i + 1 < no_contig_batches; ++i) {
        in.GetPart(in_contigs, i, d, meta); out_archive.AddPart(collection_contig_id, d, meta);
        in.GetPart(in_details, i, d, meta); out_archive.AddPart(collection_details_id, d, meta);
    }
    // ---- prepare_for_appending_load_last_batch (80-109): frames of the last batch
    std::vector<uint8_t> last_names_part, last_details_part; uint64_t last_names_meta = 0, last_details_meta = 0;
    in.GetPart(in_contigs, no_contig_batches - 1, last_names_part, last_names_meta);
    in.GetPart(in_details, no_contig_batches - 1, last_details_part, last_details_meta);
    f_names.packed = last_names_part; f_names.raw_size = last_names_meta;
    {   const uint8_t* p = last_details_part.data(); const uint8_t* e = p + last_details_part.size();
        uint32_t raw_sz[5], pk_sz[5];
        for (int i = 0; i < 5; ++i) if (!rd_u32(p, e, raw_sz[i]) || !rd_u32(p, e, pk_sz[i])) return fail("bad collection-details part");
        for (int i = 0; i < 5; ++i) { if (p + pk_sz[i] > e) return fail("bad collection-details part"); f_det[i].packed.assign(p, p + pk_sz[i]); f_det[i].raw_size = raw_sz[i]; p += pk_sz[i]; } }
    // zstd frames with raw_size 0 cannot occur for the collection (always coded); an empty stream decodes to nothing
    {   std::vector<Frame*> fr{ &f_samples, &f_names, &f_det[0], &f_det[1], &f_det[2], &f_det[3], &f_det[4] };
        for (auto* f : fr) if (f->raw_size == 0) { f->raw.clear(); f->packed.clear(); }
        std::vector<Frame*> nz; for (auto* f : fr) if (!f->packed.empty()) nz.push_back(f);
        if (!decode_all(nz)) return false; }
    if (!collection.deserialize_sample_names(f_samples.raw)) return fail("cannot deserialize the sample names");
    uint32_t no_samples_in_last_batch = 0;
    const size_t i_sample = (no_contig_batches - 1) * (size_t)pack_cardinality;
    if (!collection.deserialize_contig_names(f_names.raw, i_sample, no_samples_in_last_batch)) return fail("cannot deserialize the contig names");
    {   std::vector<uint8_t> v5[5]; for (int i = 0; i < 5; ++i) v5[i] = f_det[i].raw;
        if (!collection.deserialize_contig_details(v5, i_sample)) return fail("cannot deserialize the contig details"); }
    if (no_samples_in_last_batch == pack_cardinality) {
        out_archive.AddPart(collection_contig_id, last_names_part, last_names_meta);
        out_archive.AddPart(collection_details_id, last_details_part, last_details_meta);
        collection.clear_batch((uint32_t)i_sample, (uint32_t)std::min<size_t>(collection.sample_desc.size(), i_sample + pack_cardinality));
    }

    // ---- -a: build_candidate_kmers_from_archive (agc_compressor.cpp:828-847) needs the reference sample (sample 0) again:
    // its contig batch tells which references / pack items it is made of
    CCollection_V3 ref_coll;
    std::unordered_map<uint32_t, std::vector<uint8_t>> ref_syms;            // symbols of the references sample 0 uses
    if (adaptive_compression) {
        ref_coll.set_params(pack_cardinality, segment_size, kmer_length);
        if (!ref_coll.deserialize_sample_names(f_samples.raw) || ref_coll.sample_desc.empty()) return fail("cannot deserialize the sample names");
        Frame b_names, b_det[5];
        std::vector<uint8_t> part; uint64_t pm = 0;
        in.GetPart(in_contigs, 0, b_names.packed, b_names.raw_size);
        in.GetPart(in_details, 0, part, pm);
        const uint8_t* p = part.data(); const uint8_t* e = p + part.size();
        uint32_t raw_sz[5], pk_sz[5];
        for (int i = 0; i < 5; ++i) if (!rd_u32(p, e, raw_sz[i]) || !rd_u32(p, e, pk_sz[i])) return fail("bad collection-details part");
        for (int i = 0; i < 5; ++i) { if (p + pk_sz[i] > e) return fail("bad collection-details part"); b_det[i].packed.assign(p, p + pk_sz[i]); b_det[i].raw_size = raw_sz[i]; p += pk_sz[i]; }
        std::vector<Frame*> fr{ &b_names, &b_det[0], &b_det[1], &b_det[2], &b_det[3], &b_det[4] }, nz;
        for (auto* f : fr) if (f->raw_size && !f->packed.empty()) nz.push_back(f);
        if (!decode_all(nz)) return false;
        uint32_t nb = 0;
        std::vector<uint8_t> v5[5]; for (int i = 0; i < 5; ++i) v5[i] = b_det[i].raw;
        if (!ref_coll.deserialize_contig_names(b_names.raw, 0, nb) || !ref_coll.deserialize_contig_details(v5, 0)) return fail("cannot deserialize the first contig batch");
        for (auto& c : ref_coll.sample_desc[0].contigs) for (auto& sg : c.segments) if (sg.group_id >= 16) ref_syms[sg.group_id];
    }

    // ---- appending_init (agc_compressor.cpp:303-380) + CSegment::appending_init (segment.cpp:418-470)
    v_segments.clear(); no_segments = 0;
    std::vector<std::unique_ptr<Frame>> ref_frames, pack_frames;
    std::vector<uint32_t> ref_group, pack_group;
    while (true) {
        const int rs = in.GetStreamId(ss_base(no_segments) + "r"), ds = in.GetStreamId(ss_base(no_segments) + "d");
        if (rs < 0 && ds < 0) break;
        v_segments.emplace_back();
        GroupState& g = v_segments.back();
        g.exists = true;
        if (rs >= 0) g.stream_ref = out_archive.RegisterStream(ss_base(no_segments) + "r");
        if (ds >= 0) g.stream_delta = out_archive.RegisterStream(ss_base(no_segments) + "d");
        if (rs >= 0) {
            auto f = std::make_unique<Frame>();
            if (!in.GetPart(rs, 0, f->packed, f->raw_size)) return fail("cannot read a reference part");
            out_archive.AddPart(g.stream_ref, f->packed, f->raw_size);
            g.no_seqs = 1; g.lazy = true;
            ref_group.push_back(no_segments); ref_frames.push_back(std::move(f));
        }
        if (ds >= 0) {
            const size_t np = in.GetNoParts(ds);
            for (size_t i = 0; i + 1 < np; ++i) { in.GetPart(ds, i, d, meta); out_archive.AddPart(g.stream_delta, d, meta); g.no_seqs += pack_cardinality; }
            if (np) {
                auto f = std::make_unique<Frame>();
                in.GetPart(ds, np - 1, f->packed, f->raw_size);
                g.packed_delta = f->packed; g.packed_meta = f->raw_size; g.packed_pending = true;
                pack_group.push_back(no_segments); pack_frames.push_back(std::move(f));
            }
        }
        ++no_segments;
    }
    // CSegment::unpack (500-577) for every group at once: references -> LZ references on the device, last packs -> v_lzp / v_raw
    {   std::vector<Frame*> fr;
        std::vector<uint8_t> markers(ref_frames.size(), 0);
        for (size_t i = 0; i < ref_frames.size(); ++i)
            if (ref_frames[i]->raw_size) {
                if (ref_frames[i]->packed.empty()) return fail("damaged archive: empty reference part");
                markers[i] = ref_frames[i]->packed.back(); ref_frames[i]->packed.pop_back();                                   // the marker byte follows the frame
            }
        for (auto& f : ref_frames) fr.push_back(f.get());
        for (auto& f : pack_frames) fr.push_back(f.get());
        if (!decode_all(fr)) return false;
        std::vector<uint8_t> sym;
        for (size_t i = 0; i < ref_frames.size(); ++i) {
            Frame& f = *ref_frames[i];
            if (f.raw_size && markers[i] == 1) tuples2bytes(f.raw, sym); else sym = f.raw;
            if (!gpu_ok(agcgpu_group_put_reference(ctx, ref_group[i], sym.empty() ? (const uint8_t*)"" : sym.data(), (uint32_t)sym.size()), "group_put_reference")) return false;
            v_segments[ref_group[i]].ref_size = (uint32_t)sym.size() + 1;
            auto rs = ref_syms.find(ref_group[i]);
            if (rs != ref_syms.end()) rs->second = sym;
        }
        for (size_t i = 0; i < pack_frames.size(); ++i) {
            GroupState& g = v_segments[pack_group[i]];
            const std::vector<uint8_t>& dl = pack_frames[i]->raw;
            if (pack_cardinality > 1) {
                size_t b_pos = 0;
                for (size_t j = 0; j < dl.size(); ++j) if (dl[j] == 0xff) { g.pack.emplace_back(dl.begin() + b_pos, dl.begin() + j); b_pos = j + 1; }
            } else if (!dl.empty()) g.pack.emplace_back(dl.begin(), dl.end() - 1);
            g.no_seqs += (uint32_t)g.pack.size();
        } }

    // ---- -a: the reference sample's contigs, put together as CAGCDecompressorLibrary::decompress_contig does
    // (src/common/agc_decompressor_lib.cpp:172-288: segments in order, reverse-complemented where flagged, k symbols of overlap
    // dropped), go through agcgpu_determine_splitters once more: it leaves the sorted k-mer list of the reference sample on the
    // device (count_kmers, agc_compressor.cpp:566-630); the splitter set it computes is replaced by the archive's below
    if (adaptive_compression) {
        auto& ctgs = ref_coll.sample_desc[0].contigs;
        // pack items that are needed: (group, pack index) -> decoded pack
        std::map<std::pair<uint32_t, uint32_t>, std::unique_ptr<Frame>> packs;
        for (auto& c : ctgs) for (auto& sg : c.segments) {
            if (sg.group_id >= 16 && sg.in_group_id == 0) continue;
            const uint32_t idx = sg.group_id < 16 ? sg.in_group_id : sg.in_group_id - 1;
            auto key = std::make_pair(sg.group_id, idx / pack_cardinality);
            if (packs.count(key)) continue;
            auto f = std::make_unique<Frame>();
            if (!in.GetPart(in.GetStreamId(ss_base(sg.group_id) + "d"), key.second, f->packed, f->raw_size)) return fail("the reference sample points at a pack that is not in the archive");
            packs[key] = std::move(f);
        }
        {   std::vector<Frame*> fr; for (auto& kv : packs) fr.push_back(kv.second.get());
            if (!decode_all(fr)) return false; }
        auto item_of = [&](uint32_t group, uint32_t idx, std::vector<uint8_t>& out) -> bool {
            const std::vector<uint8_t>& dl = packs[std::make_pair(group, idx / pack_cardinality)]->raw;
            uint32_t want = idx % pack_cardinality, cur = 0; size_t b_pos = 0;
            if (pack_cardinality == 1) { if (dl.empty()) return false; out.assign(dl.begin(), dl.end() - 1); return true; }
            for (size_t j = 0; j < dl.size(); ++j) if (dl[j] == 0xff) { if (cur == want) { out.assign(dl.begin() + b_pos, dl.begin() + j); return true; } ++cur; b_pos = j + 1; }
            return false;
        };
        // deltas of sample 0 (two segments of the reference genome with the same splitter pair): one device decode batch
        std::vector<uint32_t> dg; std::vector<uint8_t> dblob; std::vector<uint64_t> doff(1, 0); std::vector<uint8_t> item;
        for (auto& c : ctgs) for (auto& sg : c.segments) if (sg.group_id >= 16 && sg.in_group_id > 0) {
            if (!item_of(sg.group_id, sg.in_group_id - 1, item)) return fail("the reference sample points at a pack item that is not in the archive");
            dg.push_back(sg.group_id); dblob.insert(dblob.end(), item.begin(), item.end()); doff.push_back(dblob.size());
        }
        std::vector<uint8_t> dec(1); std::vector<uint64_t> deco(dg.size() + 1, 0);
        if (!dg.empty()) {
            dblob.push_back(0);
            int rc = agcgpu_lz_decode_batch(ctx, dg.data(), dblob.data(), doff.data(), (uint32_t)dg.size(), dec.data(), 0, deco.data());
            if (rc == AGCGPU_EOVERFLOW) { dec.resize(deco.back() + 1); rc = agcgpu_lz_decode_batch(ctx, dg.data(), dblob.data(), doff.data(), (uint32_t)dg.size(), dec.data(), deco.back(), deco.data()); }
            if (!gpu_ok(rc, "lz_decode_batch")) return false;
        }
        static const char alpha[] = "ACGTNRYSWKMBDHVU";
        std::vector<uint8_t> raw; std::vector<uint64_t> roffs(1, 0); std::vector<uint8_t> seg;
        size_t di = 0;
        for (auto& c : ctgs) {
            bool first = true;
            for (auto& sg : c.segments) {
                if (sg.group_id < 16) { if (!item_of(sg.group_id, sg.in_group_id, seg)) return fail("the reference sample points at a raw item that is not in the archive"); }
                else if (sg.in_group_id == 0) seg = ref_syms[sg.group_id];
                else {
                    if (doff[di + 1] == doff[di]) seg = ref_syms[sg.group_id];       // (never stored: an empty delta is in_group_id 0)
                    else seg.assign(dec.begin() + deco[di], dec.begin() + deco[di + 1]);
                    ++di;
                }
                if (sg.is_rev_comp) { std::reverse(seg.begin(), seg.end()); for (auto& x : seg) if (x < 4) x = 3 - x; }
                size_t from = first ? 0 : std::min<size_t>(kmer_length, seg.size());
                for (size_t j = from; j < seg.size(); ++j) raw.push_back(seg[j] < 16 ? (uint8_t)alpha[seg[j]] : (uint8_t)'N');
                first = false;
            }
            roffs.push_back(raw.size());
        }
        if (raw.empty()) raw.push_back(0);
        std::vector<uint64_t> tmp_spl(raw.size() / std::max<uint32_t>(segment_size, 1) + 2 * roffs.size() + 64); uint64_t n_tmp = 0;
        if (!gpu_ok(agcgpu_determine_splitters(ctx, raw.data(), roffs.data(), (uint32_t)roffs.size() - 1, tmp_spl.data(), tmp_spl.size(), &n_tmp), "determine_splitters (reference k-mers)")) return false;
    }

    // splitters and segment map (agc_compressor.cpp:332-378)
    if (!in.GetPart(in.GetStreamId("splitters"), 0, d, meta) || d.size() < meta * 8) return fail("archive has no splitters");
    splitters.resize(meta);
    for (uint64_t i = 0; i < meta; ++i) { uint64_t x = 0; for (int j = 0; j < 8; ++j) x |= (uint64_t)d[i * 8 + j] << (8 * j); splitters[i] = x; }
    std::sort(splitters.begin(), splitters.end());
    if (!gpu_ok(agcgpu_set_splitters(ctx, splitters.data(), splitters.size()), "set_splitters")) return false;
    if (!in.GetPart(in.GetStreamId("segment-splitters"), 0, d, meta) || d.size() < meta * 20) return fail("archive has no segment map");
    map_segments.clear(); map_segments_terminators.clear();
    map_segments[std::make_pair(EMPTY_K, EMPTY_K)] = 0;
    std::vector<uint64_t> k1, k2; std::vector<int32_t> gv;
    for (uint64_t i = 0; i < meta; ++i) {
        uint64_t x1 = 0, x2 = 0; uint32_t x3 = 0;
        for (int j = 0; j < 8; ++j) { x1 |= (uint64_t)d[i * 20 + j] << (8 * j); x2 |= (uint64_t)d[i * 20 + 8 + j] << (8 * j); }
        for (int j = 0; j < 4; ++j) x3 |= (uint32_t)d[i * 20 + 16 + j] << (8 * j);
        if (x3 >= no_segments) return fail("damaged archive: segment-splitters refers to group " + std::to_string(x3) + " of " + std::to_string(no_segments));
        map_segments[std::make_pair(x1, x2)] = (int32_t)x3;
        if (x1 != EMPTY_K && x2 != EMPTY_K) {
            map_segments_terminators[x1].push_back(x2);
            if (x1 != x2) map_segments_terminators[x2].push_back(x1);
        }
        if (!(x1 == EMPTY_K && x2 == EMPTY_K)) { k1.push_back(x1); k2.push_back(x2); gv.push_back((int32_t)x3); }     // the context is born with (~0,~0) -> 0
    }
    for (auto& t : map_segments_terminators) std::sort(t.second.begin(), t.second.end());
    if (!k1.empty() && !gpu_ok(agcgpu_map_insert(ctx, k1.data(), k2.data(), gv.data(), gv.size()), "map_insert")) return false;

    processed_samples = (uint32_t)collection.get_no_samples();
    collection.reset_prev_sample_name();
    appending = true; working = true; epoch = 0;
    return true;
}

}  // namespace agc_b200
