// agcgpu_mock.cpp -- TEST INFRASTRUCTURE ONLY (never built by agc_b200/csrc/Makefile, never shipped, never loaded by agc_b200/).
//
// The device-level entry points of include/agcgpu.h restated over the C oracle (oracle/agc_oracle.c) and the host build of
// the residual coder (agc_b200/csrc/zstd_enc.cuh, as tests/zstd_host builds it), so that the HOST side of the path --
// agc_b200/csrc/host/compressor.cpp: add_segment's rare branches, registration order, pack bookkeeping, CCollection_V3,
// CArchive -- can be run end to end in the CPU test-suite and its archives compared byte for byte with the reference
// binary's.  tests/test_host_pipeline.py links this file with the product's host objects into tests/mock/libagcgpu_mock.so.
// The product library has no such path: without an sm_100 device agcgpu_create fails there.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>
#include <numeric>
#include "../../include/agcgpu.h"

#define ZE_NS ze
#define ZE_WN_W 512
#include "../../agc_b200/csrc/zstd_enc.cuh"
#undef ZE_NS
#undef ZE_WN_W
#define ZE_NS zen
#define ZE_WN_W 32
#include "../../agc_b200/csrc/zstd_enc.cuh"
#undef ZE_NS
#undef ZE_WN_W

#include "../../agc_b200/csrc/zstd_dec.cuh"

extern "C" {
typedef struct olz olz_t;
typedef struct { uint64_t start, len, front_dir, front_rc, back_dir, back_rc; uint32_t has_front, has_back; } orc_cut_t;
uint64_t orc_preprocess(const uint8_t* raw, uint64_t len, uint8_t* out);
void orc_reverse_complement(const uint8_t* src, uint64_t n, uint8_t* dst);
uint64_t orc_scan_contig(const uint8_t* ctg, uint64_t n, uint32_t k, const uint64_t* spl, uint64_t n_spl, orc_cut_t* cuts);
uint64_t orc_enumerate_kmers(const uint8_t* ctg, uint64_t n, uint32_t k, uint64_t* out);
uint64_t orc_determine_splitters(const uint8_t* ctgs, const uint64_t* offs, uint32_t n_ctg, uint32_t k, uint64_t segment_size,
                                 uint64_t* out, uint64_t* singletons_out, uint64_t* n_singletons);
uint64_t orc_find_new_splitters_pos(const uint8_t* ctg, uint64_t n, uint32_t k, uint64_t segment_size, const uint64_t* ref_kmers,
                                    uint64_t n_ref, uint64_t* out, uint64_t* out_pos, uint8_t* out_last);
uint64_t orc_find_splitters_pos(const uint8_t* ctg, uint64_t n, uint32_t k, uint64_t segment_size, const uint64_t* cand, uint64_t n_cand,
                                uint64_t* out, uint64_t* out_pos, uint8_t* out_last);
uint64_t orc_filtered_kmers(const uint8_t* ctg, uint64_t n, uint32_t k, uint64_t thr, uint64_t* out_pos, uint64_t* out_kmer, uint8_t* out_flags);
olz_t* orc_lz_prepare(const uint8_t* ref, uint32_t m, uint32_t min_match_len);
void orc_lz_free(olz_t* z);
uint64_t orc_lz_ht_size(const olz_t* z);
void orc_lz_get_ht(const olz_t* z, uint32_t* out);
uint64_t orc_lz_encode(const olz_t* z, const uint8_t* text, uint32_t n, uint8_t** out);
void orc_free(void* p);
uint64_t orc_lz_estimate(const olz_t* z, const uint8_t* text, uint32_t n, uint32_t bound);
uint64_t orc_lz_cost_vector(const olz_t* z, const uint8_t* text, uint32_t n, int prefix_costs, uint32_t* v);
uint64_t orc_bytes2tuples(const uint8_t* b, uint64_t n, uint8_t* out);
uint64_t orc_lz_decode(const uint8_t* ref, uint32_t m, const uint8_t* enc, uint64_t en, uint32_t mml, uint8_t* out);
int orc_ref_use_tuples(const uint8_t* d, uint64_t n);
}

struct agcgpu_ctx {
    agcgpu_params prm;
    std::string err;
    agcgpu_stats stats;
    std::vector<uint64_t> splitters;                          // sorted
    std::vector<std::vector<uint8_t>> contigs;                // resident batch, preprocessed symbols
    std::map<std::pair<uint64_t, uint64_t>, int32_t> map;
    struct Group { std::vector<uint8_t> ref; olz_t* lz = nullptr; };
    std::map<uint32_t, Group> groups;
    std::vector<uint64_t> ref_kmers;                          // AGCGPU_F_ADAPTIVE: sorted k-mers of the reference sample
    struct SplFound { uint32_t contig; uint64_t pos, kmer; uint8_t is_last; };
    std::vector<SplFound> last_spl;
    struct ZBatch { std::vector<uint8_t> src; std::vector<uint64_t> offs; std::vector<int32_t> levels; };
    std::vector<ZBatch> zwaves;                               // agcgpu_zstd_submit: batches waiting for agcgpu_zstd_collect
};

static thread_local std::string g_create_err;
static std::map<std::string, std::pair<uint64_t, uint64_t>> g_calls;     // AGC_MOCK_COUNT=1: calls and items per entry point, printed at destroy
#define COUNT(name, items) do { auto& c_ = g_calls[name]; c_.first++; c_.second += (items); } while (0)

static int fail(agcgpu_ctx* c, int code, const char* fmt, ...)
{
    char b[512]; va_list ap; va_start(ap, fmt); vsnprintf(b, sizeof b, fmt, ap); va_end(ap);
    if (c) c->err = b; else g_create_err = b;
    return code;
}

static void preprocess_all(agcgpu_ctx* ctx, const uint8_t* raw, const uint64_t* offs, uint32_t n)
{
    ctx->contigs.assign(n, {});
    for (uint32_t c = 0; c < n; ++c) {
        uint64_t len = offs[c + 1] - offs[c];
        ctx->contigs[c].resize(len + 1);
        uint64_t m = orc_preprocess(raw + offs[c], len, ctx->contigs[c].data());
        ctx->contigs[c].resize(m);
    }
}

// symbols of a request, padded so that the oracle's key reads stay inside the buffer
static int fetch(agcgpu_ctx* ctx, uint32_t contig, uint64_t start, uint32_t len, uint32_t is_rc, std::vector<uint8_t>& out)
{
    if (contig >= ctx->contigs.size() || start + len > ctx->contigs[contig].size())
        return fail(ctx, AGCGPU_EINVAL, "segment outside the resident batch (contig %u, start %llu, len %u)", contig, (unsigned long long)start, len);
    out.assign(len + 64, 0);
    const uint8_t* p = ctx->contigs[contig].data() + start;
    if (is_rc) orc_reverse_complement(p, len, out.data()); else if (len) memcpy(out.data(), p, len);
    return 0;
}

static void scan_all(agcgpu_ctx* ctx, std::vector<agcgpu_cut>& cuts)
{
    for (uint32_t c = 0; c < ctx->contigs.size(); ++c) {
        auto& s = ctx->contigs[c];
        uint64_t n = orc_scan_contig(s.data(), s.size(), ctx->prm.kmer_length, ctx->splitters.data(), ctx->splitters.size(), nullptr);
        std::vector<orc_cut_t> oc(n + 1);
        orc_scan_contig(s.data(), s.size(), ctx->prm.kmer_length, ctx->splitters.data(), ctx->splitters.size(), oc.data());
        for (uint64_t i = 0; i < n; ++i) {
            agcgpu_cut q; memset(&q, 0, sizeof q);
            q.contig = c; q.has_front = oc[i].has_front; q.has_back = oc[i].has_back; q.start = oc[i].start; q.len = oc[i].len;
            if (q.has_front) { q.front_dir = oc[i].front_dir; q.front_rc = oc[i].front_rc; }
            if (q.has_back) { q.back_dir = oc[i].back_dir; q.back_rc = oc[i].back_rc; }
            cuts.push_back(q);
        }
    }
}

static int emit_cuts(agcgpu_ctx* ctx, std::vector<agcgpu_cut>& cuts, agcgpu_cut* out, uint64_t cap, uint64_t* out_n)
{
    if (out_n) *out_n = cuts.size();
    if (cuts.size() > cap) return fail(ctx, AGCGPU_EOVERFLOW, "scan: %zu cuts, caller buffer holds %llu", cuts.size(), (unsigned long long)cap);
    if (!cuts.empty()) memcpy(out, cuts.data(), cuts.size() * sizeof(agcgpu_cut));
    return 0;
}

extern "C" {
#pragma GCC visibility push(default)

int agcgpu_create(const agcgpu_params* p, agcgpu_ctx** out)
{
    if (!p || !out) return AGCGPU_EINVAL;
    if (p->kmer_length < 1 || p->kmer_length > 32) return fail(nullptr, AGCGPU_EINVAL, "kmer_length out of range");
    agcgpu_ctx* c = new agcgpu_ctx();
    c->prm = *p; memset(&c->stats, 0, sizeof c->stats);
    c->map[std::make_pair(~0ULL, ~0ULL)] = 0;                  // raw groups (agc_compressor.cpp:2307), as api.cu seeds it
    *out = c;
    return 0;
}
void agcgpu_destroy(agcgpu_ctx* ctx)
{
    if (!ctx) return;
    for (auto& g : ctx->groups) orc_lz_free(g.second.lz);
    if (getenv("AGC_MOCK_COUNT")) for (auto& c : g_calls) fprintf(stderr, "[mock] %-28s %8llu calls %12llu items\n", c.first.c_str(), (unsigned long long)c.second.first, (unsigned long long)c.second.second);
    delete ctx;
}
const char* agcgpu_last_error(const agcgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
int agcgpu_sync(agcgpu_ctx*) { return 0; }
void* agcgpu_stream(agcgpu_ctx*) { return nullptr; }
int agcgpu_get_stats(agcgpu_ctx* ctx, agcgpu_stats* out) { if (!ctx || !out) return AGCGPU_EINVAL; *out = ctx->stats; return 0; }

int agcgpu_set_splitters(agcgpu_ctx* ctx, const uint64_t* s, uint64_t n)
{
    if (!ctx || (n && !s)) return AGCGPU_EINVAL;
    ctx->splitters.assign(s, s + n);
    std::sort(ctx->splitters.begin(), ctx->splitters.end());
    return 0;
}

int agcgpu_determine_splitters(agcgpu_ctx* ctx, const uint8_t* raw, const uint64_t* offs, uint32_t n, uint64_t* out, uint64_t cap, uint64_t* out_n)
{
    if (!ctx || !offs || !out_n) return AGCGPU_EINVAL;
    preprocess_all(ctx, raw, offs, n);
    std::vector<uint8_t> cat; std::vector<uint64_t> co(n + 1, 0);
    for (uint32_t c = 0; c < n; ++c) { cat.insert(cat.end(), ctx->contigs[c].begin(), ctx->contigs[c].end()); co[c + 1] = cat.size(); }
    cat.push_back(0);
    std::vector<uint64_t> spl(cat.size() + 2 * n + 16);
    std::vector<uint64_t> singles(cat.size() + 2); uint64_t n_singles = 0;
    uint64_t ns = orc_determine_splitters(cat.data(), co.data(), n, ctx->prm.kmer_length, ctx->prm.segment_size, spl.data(), singles.data(), &n_singles);
    ctx->last_spl.clear();
    for (uint32_t c = 0; c < n; ++c) {
        uint64_t len = co[c + 1] - co[c];
        std::vector<uint64_t> o(len + 2), op(len + 2); std::vector<uint8_t> ol(len + 2);
        uint64_t no = orc_find_splitters_pos(cat.data() + co[c], len, ctx->prm.kmer_length, ctx->prm.segment_size, singles.data(), n_singles, o.data(), op.data(), ol.data());
        for (uint64_t i = 0; i < no; ++i) ctx->last_spl.push_back(agcgpu_ctx::SplFound{ c, op[i], o[i], ol[i] });
    }
    if (ctx->prm.flags & AGCGPU_F_ADAPTIVE) {
        ctx->ref_kmers.assign(cat.size() + 1, 0);
        uint64_t nk = 0;
        for (uint32_t c = 0; c < n; ++c) nk += orc_enumerate_kmers(cat.data() + co[c], co[c + 1] - co[c], ctx->prm.kmer_length, ctx->ref_kmers.data() + nk);
        ctx->ref_kmers.resize(nk);
        std::sort(ctx->ref_kmers.begin(), ctx->ref_kmers.end());
    }
    *out_n = ns;                                               // the contigs stay resident until the next scan (agcgpu.h)
    if (ns > cap) return fail(ctx, AGCGPU_EOVERFLOW, "determine_splitters: %llu splitters, buffer holds %llu", (unsigned long long)ns, (unsigned long long)cap);
    if (ns) memcpy(out, spl.data(), ns * 8);
    ctx->splitters.assign(spl.begin(), spl.begin() + ns);
    return 0;
}

int agcgpu_find_new_splitters(agcgpu_ctx* ctx, const uint32_t* contigs, uint32_t n, uint64_t* out, uint64_t cap, uint64_t* out_n)
{
    if (!ctx || !out_n || (n && !contigs)) return AGCGPU_EINVAL;
    if (!(ctx->prm.flags & AGCGPU_F_ADAPTIVE)) return fail(ctx, AGCGPU_EINVAL, "find_new_splitters needs AGCGPU_F_ADAPTIVE");
    std::vector<uint64_t> all;
    ctx->last_spl.clear();
    for (uint32_t i = 0; i < n; ++i) {
        if (contigs[i] >= ctx->contigs.size()) return fail(ctx, AGCGPU_EINVAL, "find_new_splitters: contig %u is not resident", contigs[i]);
        auto& s = ctx->contigs[contigs[i]];
        std::vector<uint64_t> o(s.size() + 2), op(s.size() + 2); std::vector<uint8_t> ol(s.size() + 2);
        std::vector<uint8_t> padded(s); padded.push_back(0);
        uint64_t no = orc_find_new_splitters_pos(padded.data(), s.size(), ctx->prm.kmer_length, ctx->prm.segment_size, ctx->ref_kmers.data(), ctx->ref_kmers.size(),
                                                 o.data(), op.data(), ol.data());
        all.insert(all.end(), o.begin(), o.begin() + no);
        for (uint64_t j = 0; j < no; ++j) ctx->last_spl.push_back(agcgpu_ctx::SplFound{ contigs[i], op[j], o[j], ol[j] });
    }
    std::sort(all.begin(), all.end());
    all.erase(std::unique(all.begin(), all.end()), all.end());
    *out_n = all.size();
    if (all.size() > cap) return fail(ctx, AGCGPU_EOVERFLOW, "find_new_splitters: %zu splitters, buffer holds %llu", all.size(), (unsigned long long)cap);
    if (!all.empty()) memcpy(out, all.data(), all.size() * 8);
    return 0;
}

int agcgpu_filtered_kmers(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint64_t thr, agcgpu_fkmer* out, uint64_t cap, uint64_t* out_offsets)
{
    COUNT("agcgpu_filtered_kmers", n);
    if (!ctx || !out_offsets || (n && !reqs) || (cap && !out)) return AGCGPU_EINVAL;
    out_offsets[0] = 0;
    bool overflow = false;
    for (uint32_t i = 0; i < n; ++i) {
        std::vector<uint8_t> s;
        if (reqs[i].contig >= ctx->contigs.size() || reqs[i].start > ctx->contigs[reqs[i].contig].size()) return fail(ctx, AGCGPU_EINVAL, "filtered_kmers: range outside the resident batch");
        const uint32_t len = (uint32_t)std::min<uint64_t>(reqs[i].len, ctx->contigs[reqs[i].contig].size() - reqs[i].start);   // clipped like get_part
        if (int r = fetch(ctx, reqs[i].contig, reqs[i].start, len, 0, s)) return r;
        std::vector<uint64_t> pos(len + 1), km(len + 1); std::vector<uint8_t> fl(len + 1);
        uint64_t c = orc_filtered_kmers(s.data(), len, ctx->prm.kmer_length, thr, pos.data(), km.data(), fl.data());
        out_offsets[i + 1] = out_offsets[i] + c;
        if (out_offsets[i + 1] > cap) { overflow = true; continue; }
        for (uint64_t j = 0; j < c; ++j) { agcgpu_fkmer f; f.pos = pos[j]; f.kmer = km[j]; f.is_dir_oriented = fl[j] & 1; f.is_symmetric = (fl[j] >> 1) & 1; out[out_offsets[i] + j] = f; }
    }
    if (overflow) return fail(ctx, AGCGPU_EOVERFLOW, "filtered_kmers: buffer too small");
    return 0;
}
int agcgpu_last_splitter_positions(agcgpu_ctx* ctx, uint32_t* out_contig, uint64_t* out_pos, uint64_t* out_kmer, uint8_t* out_is_last, uint64_t cap, uint64_t* out_n)
{
    if (!ctx || !out_n) return AGCGPU_EINVAL;
    auto v = ctx->last_spl;
    std::sort(v.begin(), v.end(), [](const agcgpu_ctx::SplFound& a, const agcgpu_ctx::SplFound& b) { return a.contig != b.contig ? a.contig < b.contig : a.pos < b.pos; });
    *out_n = v.size();
    if (v.size() > cap) return fail(ctx, AGCGPU_EOVERFLOW, "last_splitter_positions: buffer too small");
    for (size_t i = 0; i < v.size(); ++i) { out_contig[i] = v[i].contig; out_pos[i] = v[i].pos; out_kmer[i] = v[i].kmer; out_is_last[i] = v[i].is_last; }
    return 0;
}

int agcgpu_scan_contigs(agcgpu_ctx* ctx, const uint8_t* raw, const uint64_t* offs, uint32_t n, uint64_t* out_len, agcgpu_cut* out_cuts,
                        uint64_t cap, uint64_t* out_n)
{
    if (!ctx || !offs || (n && !raw && offs[n])) return AGCGPU_EINVAL;
    preprocess_all(ctx, raw, offs, n);
    if (out_len) for (uint32_t c = 0; c < n; ++c) out_len[c] = ctx->contigs[c].size();
    std::vector<agcgpu_cut> cuts;
    scan_all(ctx, cuts);
    return emit_cuts(ctx, cuts, out_cuts, cap, out_n);
}
int agcgpu_scan_contigs_dev(agcgpu_ctx* ctx, const void* raw_dev, uint64_t, const uint64_t* offs, uint32_t n, uint64_t* out_len,
                            agcgpu_cut* out_cuts, uint64_t cap, uint64_t* out_n)
{
    return agcgpu_scan_contigs(ctx, (const uint8_t*)raw_dev, offs, n, out_len, out_cuts, cap, out_n);   // "device" memory is host memory here
}
int agcgpu_rescan_contigs(agcgpu_ctx* ctx, agcgpu_cut* out_cuts, uint64_t cap, uint64_t* out_n)
{
    if (!ctx) return AGCGPU_EINVAL;
    std::vector<agcgpu_cut> cuts;
    scan_all(ctx, cuts);
    return emit_cuts(ctx, cuts, out_cuts, cap, out_n);
}

int agcgpu_get_segment(agcgpu_ctx* ctx, uint32_t contig, uint64_t start, uint32_t len, uint32_t is_rc, uint8_t* out)
{
    COUNT("agcgpu_get_segment", 1);
    if (!ctx || (len && !out)) return AGCGPU_EINVAL;
    std::vector<uint8_t> s;
    if (int r = fetch(ctx, contig, start, len, is_rc, s)) return r;
    if (len) memcpy(out, s.data(), len);
    return 0;
}

int agcgpu_map_insert(agcgpu_ctx* ctx, const uint64_t* k1, const uint64_t* k2, const int32_t* gid, uint64_t n)
{
    COUNT("agcgpu_map_insert", n);
    if (!ctx || (n && (!k1 || !k2 || !gid))) return AGCGPU_EINVAL;
    for (uint64_t i = 0; i < n; ++i) {
        if (gid[i] < 0) return fail(ctx, AGCGPU_EINVAL, "map_insert: negative group id");
        auto key = std::make_pair(k1[i], k2[i]);
        auto p = ctx->map.find(key);
        if (p == ctx->map.end()) ctx->map[key] = gid[i]; else if (p->second > gid[i]) p->second = gid[i];
    }
    return 0;
}

int agcgpu_assign_cuts(agcgpu_ctx* ctx, const agcgpu_cut* cuts, uint64_t n, agcgpu_assign* out)
{
    COUNT("agcgpu_assign_cuts", n);
    if (!ctx || (n && (!cuts || !out))) return AGCGPU_EINVAL;
    for (uint64_t i = 0; i < n; ++i) {                          // add_segment key construction (agc_compressor.cpp:1287-1313)
        const agcgpu_cut& c = cuts[i];
        agcgpu_assign a; memset(&a, 0, sizeof a);
        uint64_t fc = std::min(c.front_dir, c.front_rc), bc = std::min(c.back_dir, c.back_rc);
        a.group_id = -1;
        if (c.has_front && c.has_back) { a.klass = 0; if (fc < bc) { a.key1 = fc; a.key2 = bc; } else { a.key1 = bc; a.key2 = fc; a.is_rc = 1; } }
        else if (c.has_front) { a.klass = 1; a.key1 = fc; a.key2 = ~0ULL; }
        else if (c.has_back) { a.klass = 2; a.key1 = ~0ULL; a.key2 = bc; }
        else { a.klass = 3; a.key1 = a.key2 = ~0ULL; }
        if (a.klass == 0 || a.klass == 3) { auto p = ctx->map.find(std::make_pair(a.key1, a.key2)); if (p != ctx->map.end()) a.group_id = p->second; }
        out[i] = a;
    }
    return 0;
}

int agcgpu_group_put_reference(agcgpu_ctx* ctx, uint32_t group_id, const uint8_t* symbols, uint32_t len)
{
    if (!ctx || (len && !symbols)) return AGCGPU_EINVAL;
    agcgpu_ctx::Group& g = ctx->groups[group_id];
    if (g.lz) { orc_lz_free(g.lz); g.lz = nullptr; }
    g.ref.assign(symbols, symbols + len);
    std::vector<uint8_t> padded(g.ref); padded.resize(len + 64, 0);
    g.lz = orc_lz_prepare(padded.data(), len, ctx->prm.min_match_len);
    return 0;
}
int agcgpu_group_put_reference_batch(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n)
{
    COUNT("agcgpu_group_put_reference_batch", n);
    if (!ctx || (n && !reqs)) return AGCGPU_EINVAL;
    for (uint32_t i = 0; i < n; ++i) {
        std::vector<uint8_t> s;
        if (int r = fetch(ctx, reqs[i].contig, reqs[i].start, reqs[i].len, reqs[i].is_rc, s)) return r;
        if (int r = agcgpu_group_put_reference(ctx, reqs[i].group_id, s.data(), reqs[i].len)) return r;
    }
    return 0;
}
int agcgpu_group_get_index(agcgpu_ctx* ctx, uint32_t group_id, uint32_t* out_slots, uint64_t cap, uint64_t* out_ht_size)
{
    if (!ctx || !out_ht_size) return AGCGPU_EINVAL;
    auto p = ctx->groups.find(group_id);
    if (p == ctx->groups.end()) return fail(ctx, AGCGPU_EINVAL, "group %u has no reference", group_id);
    *out_ht_size = orc_lz_ht_size(p->second.lz);
    if (!out_slots) return 0;
    if (*out_ht_size > cap) return fail(ctx, AGCGPU_EOVERFLOW, "index larger than the caller's buffer");
    orc_lz_get_ht(p->second.lz, out_slots);
    return 0;
}

static int lz_prep(agcgpu_ctx* ctx, const agcgpu_seg_req& q, std::vector<uint8_t>& text, olz_t** z)
{
    auto p = ctx->groups.find(q.group_id);
    if (p == ctx->groups.end()) return fail(ctx, AGCGPU_EINVAL, "group %u has no reference", q.group_id);
    *z = p->second.lz;
    return fetch(ctx, q.contig, q.start, q.len, q.is_rc, text);
}

int agcgpu_lz_encode_batch(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_offsets)
{
    COUNT("agcgpu_lz_encode_batch", n);
    if (!ctx || !out_offsets || (n && (!reqs || !out))) return AGCGPU_EINVAL;
    uint64_t o = 0; out_offsets[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        std::vector<uint8_t> t; olz_t* z;
        if (int r = lz_prep(ctx, reqs[i], t, &z)) return r;
        uint8_t* e = nullptr;
        uint64_t en = orc_lz_encode(z, t.data(), reqs[i].len, &e);
        if (o + en > out_cap) { orc_free(e); return fail(ctx, AGCGPU_EOVERFLOW, "lz_encode: output buffer too small"); }
        if (en) memcpy(out + o, e, en);
        orc_free(e);
        o += en; out_offsets[i + 1] = o;
    }
    return 0;
}
int agcgpu_lz_estimate_batch(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint32_t* out)
{
    COUNT("agcgpu_lz_estimate_batch", n);
    if (!ctx || (n && (!reqs || !out))) return AGCGPU_EINVAL;
    for (uint32_t i = 0; i < n; ++i) {
        std::vector<uint8_t> t; olz_t* z;
        if (int r = lz_prep(ctx, reqs[i], t, &z)) return r;
        out[i] = (uint32_t)orc_lz_estimate(z, t.data(), reqs[i].len, reqs[i].bound);
    }
    return 0;
}
int agcgpu_lz_cost_vector(agcgpu_ctx* ctx, const agcgpu_seg_req* req, int prefix_costs, uint32_t* out)
{
    COUNT("agcgpu_lz_cost_vector", 1);
    if (!ctx || !req || (req->len && !out)) return AGCGPU_EINVAL;
    std::vector<uint8_t> t; olz_t* z;
    if (int r = lz_prep(ctx, *req, t, &z)) return r;
    std::vector<uint32_t> v(req->len + 64, 0);
    uint64_t vn = orc_lz_cost_vector(z, t.data(), req->len, prefix_costs, v.data());
    if (vn != req->len) return fail(ctx, AGCGPU_EINVAL, "cost vector has %llu entries for %u symbols", (unsigned long long)vn, req->len);
    if (req->len) memcpy(out, v.data(), (size_t)req->len * 4);
    return 0;
}

int agcgpu_comm_world(void) { return 1; }
int agcgpu_comm_rank(void) { return 0; }
int agcgpu_lz_encode_batch_sharded(agcgpu_ctx* ctx, const agcgpu_seg_req* reqs, uint32_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_offsets)
{ return agcgpu_lz_encode_batch(ctx, reqs, n, out, out_cap, out_offsets); }
int agcgpu_zstd_compress_batch_sharded(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* src_offsets, const int32_t* levels, uint32_t n, uint8_t* dst,
                                       uint64_t dst_cap, uint64_t* dst_offsets)
{ return agcgpu_zstd_compress_batch(ctx, src, src_offsets, levels, n, dst, dst_cap, dst_offsets); }
int agcgpu_lz_cost_split_batch(agcgpu_ctx* ctx, const agcgpu_split_req* reqs, uint32_t n, uint32_t* out_best_pos, uint32_t* out_best_sum)
{
    COUNT("agcgpu_lz_cost_split_batch", n);
    if (!ctx || (n && (!reqs || !out_best_pos || !out_best_sum))) return AGCGPU_EINVAL;
    for (uint32_t i = 0; i < n; ++i) {
        const agcgpu_split_req& q = reqs[i];
        std::vector<uint32_t> c[2];
        for (int h = 0; h < 2; ++h) {
            const uint32_t f = h ? q.flags >> 3 : q.flags;
            agcgpu_seg_req r; memset(&r, 0, sizeof r);
            r.contig = q.contig; r.start = q.start; r.len = q.len; r.is_rc = f & 1u; r.group_id = h ? q.group2 : q.group1;
            c[h].assign((size_t)q.len + 64, 0);
            if (q.len) { if (int rc = agcgpu_lz_cost_vector(ctx, &r, (f >> 1) & 1, c[h].data())) return rc; }
            c[h].resize(q.len);
            if (f & 4u) std::reverse(c[h].begin(), c[h].end());
        }
        std::partial_sum(c[0].begin(), c[0].end(), c[0].begin());
        std::partial_sum(c[1].rbegin(), c[1].rend(), c[1].rbegin());
        uint32_t best = ~0u, bp = 0;
        for (uint32_t k = 0; k < q.len; ++k) { uint32_t cs = c[0][k] + c[1][k]; if (cs < best) { best = cs; bp = k; } }
        out_best_pos[i] = bp; out_best_sum[i] = best;
    }
    return 0;
}

int agcgpu_lz_decode_batch(agcgpu_ctx* ctx, const uint32_t* group_ids, const uint8_t* deltas, const uint64_t* dofs, uint32_t n, uint8_t* out,
                           uint64_t out_cap, uint64_t* out_offsets)
{
    if (!ctx || !out_offsets || !dofs || (n && (!group_ids || !deltas)) || (out_cap && !out)) return AGCGPU_EINVAL;
    out_offsets[0] = 0;
    std::vector<std::vector<uint8_t>> dec(n);
    for (uint32_t i = 0; i < n; ++i) {
        auto p = ctx->groups.find(group_ids[i]);
        if (p == ctx->groups.end()) return fail(ctx, AGCGPU_EINVAL, "lz_decode: group %u has no reference", group_ids[i]);
        const uint64_t en = dofs[i + 1] - dofs[i];
        std::vector<uint8_t> enc(deltas + dofs[i], deltas + dofs[i + 1]); enc.push_back(0);        // the oracle peeks one byte past a number
        uint64_t bound = en + 64;                                                                   // literals: one symbol per byte
        for (uint64_t j = 0; j < en; ++j) {
            if (enc[j] == '.') bound += p->second.ref.size();                                        // a match copies at most the whole reference
            if (enc[j] == 30) { uint64_t v = 0, q = j + 1; while (q < en && enc[q] >= '0' && enc[q] <= '9') v = v * 10 + (enc[q++] - '0'); bound += v + 4; }
        }
        dec[i].resize(bound);
        uint64_t o = orc_lz_decode(p->second.ref.data(), (uint32_t)p->second.ref.size(), enc.data(), en, ctx->prm.min_match_len, dec[i].data());
        dec[i].resize(o);
        out_offsets[i + 1] = out_offsets[i] + o;
    }
    if (out_offsets[n] > out_cap) return fail(ctx, AGCGPU_EOVERFLOW, "lz_decode: output buffer too small");
    for (uint32_t i = 0; i < n; ++i) if (!dec[i].empty()) memcpy(out + out_offsets[i], dec[i].data(), dec[i].size());
    return 0;
}

int agcgpu_pack_ref_batch(agcgpu_ctx* ctx, const uint32_t* group_ids, uint32_t n, uint8_t* out, uint64_t out_cap, uint64_t* out_offsets,
                          uint8_t* out_use_tuples)
{
    COUNT("agcgpu_pack_ref_batch", n);
    if (!ctx || !out_offsets || (n && (!group_ids || !out || !out_use_tuples))) return AGCGPU_EINVAL;
    uint64_t o = 0; out_offsets[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        auto p = ctx->groups.find(group_ids[i]);
        if (p == ctx->groups.end()) return fail(ctx, AGCGPU_EINVAL, "group %u has no reference", group_ids[i]);
        const std::vector<uint8_t>& r = p->second.ref;
        std::vector<uint8_t> pay(r.size() + 8);
        uint64_t pn;
        out_use_tuples[i] = (uint8_t)orc_ref_use_tuples(r.data(), r.size());
        if (out_use_tuples[i]) pn = orc_bytes2tuples(r.data(), r.size(), pay.data());
        else { pn = r.size(); if (pn) memcpy(pay.data(), r.data(), pn); }
        if (o + pn > out_cap) return fail(ctx, AGCGPU_EOVERFLOW, "pack_ref: output buffer too small");
        if (pn) memcpy(out + o, pay.data(), pn);
        o += pn; out_offsets[i + 1] = o;
    }
    return 0;
}

void* agcgpu_host_alloc(uint64_t bytes, uint64_t* out_cap) { if (!out_cap) return nullptr; *out_cap = bytes + 64; return malloc(bytes + 64); }
void agcgpu_host_free(void* p, uint64_t) { free(p); }

// the asynchronous coder: submit parks the batch, collect codes every parked batch in submission order
int agcgpu_zstd_submit(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* so, const int32_t* levels, uint32_t n)
{
    COUNT("agcgpu_zstd_submit", n);
    if (!ctx || !so || (n && (!src || !levels))) return AGCGPU_EINVAL;
    if (!n) return 0;
    agcgpu_ctx::ZBatch b;
    b.src.assign(src + so[0], src + so[n]); b.offs.resize(n + 1); b.levels.assign(levels, levels + n);
    for (uint32_t i = 0; i <= n; ++i) b.offs[i] = so[i] - so[0];
    ctx->zwaves.push_back(std::move(b));
    return 0;
}
int agcgpu_zstd_submit_parts(agcgpu_ctx* ctx, const uint8_t* const* ptrs, const uint64_t* sizes, const int32_t* levels, uint32_t n)
{
    COUNT("agcgpu_zstd_submit_parts", n);
    if (!ctx || (n && (!ptrs || !sizes || !levels))) return AGCGPU_EINVAL;
    if (!n) return 0;
    agcgpu_ctx::ZBatch b;
    b.offs.assign(1, 0); b.levels.assign(levels, levels + n);
    for (uint32_t i = 0; i < n; ++i) { if (sizes[i]) b.src.insert(b.src.end(), ptrs[i], ptrs[i] + sizes[i]); b.offs.push_back(b.src.size()); }
    ctx->zwaves.push_back(std::move(b));
    return 0;
}
int agcgpu_zstd_collect(agcgpu_ctx* ctx, uint32_t n_expected, uint8_t* dst, uint64_t dst_cap, uint64_t* dof)
{
    COUNT("agcgpu_zstd_collect", n_expected);
    if (!ctx || !dof || (n_expected && !dst)) return AGCGPU_EINVAL;
    std::vector<agcgpu_ctx::ZBatch> waves; waves.swap(ctx->zwaves);
    uint64_t total = 0; for (auto& w : waves) total += w.levels.size();
    if (total != n_expected) return fail(ctx, AGCGPU_EINVAL, "zstd collect: %llu inputs were submitted, the caller expects %u", (unsigned long long)total, n_expected);
    uint64_t base = 0, o = 0; dof[0] = 0;
    for (auto& w : waves) {
        const uint32_t n = (uint32_t)w.levels.size();
        std::vector<uint64_t> fo(n + 1, 0);
        w.src.push_back(0);
        if (int r = agcgpu_zstd_compress_batch(ctx, w.src.data(), w.offs.data(), w.levels.data(), n, dst + o, dst_cap - o, fo.data())) return r;
        for (uint32_t i = 0; i < n; ++i) dof[base + i + 1] = o + fo[i + 1];
        o += fo[n]; base += n;
    }
    return 0;
}

int agcgpu_zstd_compress_batch(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* so, const int32_t* levels, uint32_t n, uint8_t* dst,
                               uint64_t dst_cap, uint64_t* dof)
{
    COUNT("agcgpu_zstd_compress_batch", n);
    if (!ctx || !so || !dof || (n && (!src || !levels || !dst))) return AGCGPU_EINVAL;
    uint64_t o = 0; dof[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint64_t len = so[i + 1] - so[i];
        ze::Params cp = ze::get_params(levels[i], len);
        if (!cp.supported) return fail(ctx, AGCGPU_EUNSUPPORTED, "zstd: input %u (level %d, %llu bytes) is outside the implemented envelope", i, levels[i], (unsigned long long)len);
        std::vector<uint8_t> in(len + 64, 0), fr(ze::compress_bound(len) + 64);
        if (len) memcpy(in.data(), src + so[i], len);
        int err = 0; uint64_t r;
        if (len > 32768) {                                      // same routing as kernels_zstd.cu: wide coder above 32 KB
            ze::WorkSizes z = ze::work_sizes(cp);
            std::vector<uint8_t> mem(z.total + 64, 0);
            r = ze::compress_frame(in.data(), len, levels[i], fr.data(), fr.size(), mem.data(), &err);
        } else {
            zen::Params cpn = zen::get_params(levels[i], len);
            zen::WorkSizes z = zen::work_sizes(cpn);
            std::vector<uint8_t> mem(z.total + 64, 0);
            r = zen::compress_frame(in.data(), len, levels[i], fr.data(), fr.size(), mem.data(), &err);
        }
        if (err) return fail(ctx, AGCGPU_EUNSUPPORTED, "zstd: coder error %d on input %u", err, i);
        if (o + r > dst_cap) return fail(ctx, AGCGPU_EOVERFLOW, "zstd: output buffer too small");
        memcpy(dst + o, fr.data(), r);
        o += r; dof[i + 1] = o;
        ctx->stats.zstd_input_mb += (float)len / 1e6f;
    }
    return 0;
}

int agcgpu_zstd_decompress_batch(agcgpu_ctx* ctx, const uint8_t* src, const uint64_t* so, uint32_t n, uint8_t* dst, uint64_t dst_cap, uint64_t* dof)
{
    if (!ctx || !so || !dof || (n && !src) || (dst_cap && !dst)) return AGCGPU_EINVAL;
    dof[0] = 0;
    for (uint32_t i = 0; i < n; ++i) {
        int64_t fcs = zd::frame_content_size(src + so[i], so[i + 1] - so[i]);
        if (fcs < 0) return fail(ctx, AGCGPU_EUNSUPPORTED, "zstd decode: frame %u has no content size in its header", i);
        dof[i + 1] = dof[i] + (uint64_t)fcs;
    }
    if (dof[n] > dst_cap) return fail(ctx, AGCGPU_EOVERFLOW, "zstd decode: output buffer too small");
    std::vector<zd::Work> w(1);
    for (uint32_t i = 0; i < n; ++i) {
        int64_t r = zd::decompress_frame(src + so[i], so[i + 1] - so[i], dst + dof[i], dof[i + 1] - dof[i], w[0]);
        if (r < 0 || (uint64_t)r != dof[i + 1] - dof[i]) return fail(ctx, AGCGPU_EINVAL, "zstd decode: frame %u is malformed or truncated (code %lld)", i, (long long)r);
    }
    return 0;
}

#pragma GCC visibility pop
}
