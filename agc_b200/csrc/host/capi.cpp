// capi.cpp -- C facade over agc_b200::CAGCCompressor (include/agcgpu.h, "CAGCCompressor facade")
#include "compressor.h"
#include <string>
#include <chrono>
#include <cstdio>
#include <cstdlib>

// AGCGPU_TRACE=1: wall time of each facade call on stderr (diagnostics only)
namespace { struct CallTimer {
    const char* name; std::chrono::steady_clock::time_point t0; bool on;
    explicit CallTimer(const char* n) : name(n), t0(std::chrono::steady_clock::now()), on(getenv("AGCGPU_TRACE") != nullptr) {}
    ~CallTimer() { if (on) fprintf(stderr, "[agcgpu] call  %-22s %8.1f ms\n", name, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count()); }
}; }

struct agcgpu_compressor { agc_b200::CAGCCompressor impl; std::string err; };
static thread_local std::string g_err;
static agcgpu_stats g_last_stats = {};
static struct { uint32_t rank = 0, world = 1; agcgpu_allgather_fn fn = nullptr; void* user = nullptr; } g_exchange;

// No C++ exception crosses the C ABI (bad_alloc on a huge collection, out_of_range on a damaged archive fed to append ...): the
// body of every entry point that runs compressor code is wrapped; the message lands in the last-error slot.
#define AGC_GUARD_BEGIN try {
#define AGC_GUARD_END(c_) } catch (const std::exception& e) { g_err = std::string("exception: ") + e.what(); if (c_) (c_)->impl.SetLastError(g_err); return AGCGPU_ECUDA; } \
                          catch (...) { g_err = "unknown exception"; if (c_) (c_)->impl.SetLastError(g_err); return AGCGPU_ECUDA; }

extern "C" {

int agcgpu_set_exchange(uint32_t rank, uint32_t world, agcgpu_allgather_fn allgather, void* user)
{
    if (allgather && world > 1 && rank >= world) return AGCGPU_EINVAL;
    g_exchange.rank = rank; g_exchange.world = world; g_exchange.fn = allgather; g_exchange.user = user;
    return 0;
}

int agcgpu_compressor_create(const char* out_file, uint32_t pack_cardinality, uint32_t kmer_length, const char* reference_file,
                             uint32_t segment_size, uint32_t min_match_len, int concatenated_genomes, int adaptive_compression,
                             uint32_t verbosity, uint32_t no_threads, double fallback_frac, int device,
                             const char* dump_parts_path, agcgpu_compressor** out)
{
    if (!out_file || !reference_file || !out) { g_err = "null argument"; return AGCGPU_EINVAL; }
    CallTimer ct("compressor_create");
    agcgpu_compressor* c = nullptr;
    AGC_GUARD_BEGIN
    c = new agcgpu_compressor();
    c->impl.SetAppMode(false);
    c->impl.SetDevice(device);
    c->impl.SetExchange(g_exchange.rank, g_exchange.world, g_exchange.fn, g_exchange.user);
    if (dump_parts_path && *dump_parts_path) c->impl.SetDumpParts(dump_parts_path);
    if (!c->impl.Create(out_file, pack_cardinality, kmer_length, reference_file, segment_size, min_match_len,
                        concatenated_genomes != 0, adaptive_compression != 0, verbosity, no_threads, fallback_frac)) {
        g_err = c->impl.LastError();
        delete c;
        return AGCGPU_EUNSUPPORTED;
    }
    *out = c;
    return 0;
    } catch (const std::exception& e) { g_err = std::string("exception: ") + e.what(); delete c; return AGCGPU_ECUDA; }
    catch (...) { g_err = "unknown exception"; delete c; return AGCGPU_ECUDA; }
}

int agcgpu_compressor_append(const char* in_archive, const char* out_file, uint32_t verbosity, int prefetch_archive, int concatenated_genomes,
                             int adaptive_compression, uint32_t no_threads, double fallback_frac, int device, agcgpu_compressor** out)
{
    if (!in_archive || !out_file || !out) { g_err = "null argument"; return AGCGPU_EINVAL; }
    CallTimer ct("compressor_append");
    agcgpu_compressor* c = nullptr;
    AGC_GUARD_BEGIN
    c = new agcgpu_compressor();
    c->impl.SetAppMode(false);
    c->impl.SetDevice(device);
    c->impl.SetExchange(g_exchange.rank, g_exchange.world, g_exchange.fn, g_exchange.user);
    if (!c->impl.Append(in_archive, out_file, verbosity, prefetch_archive != 0, concatenated_genomes != 0, adaptive_compression != 0, no_threads, fallback_frac)) {
        g_err = c->impl.LastError();
        delete c;
        return AGCGPU_EUNSUPPORTED;
    }
    *out = c;
    return 0;
    } catch (const std::exception& e) { g_err = std::string("exception: ") + e.what(); delete c; return AGCGPU_ECUDA; }
    catch (...) { g_err = "unknown exception"; delete c; return AGCGPU_ECUDA; }
}

int agcgpu_compressor_add_sample_files(agcgpu_compressor* c, const char* const* sample_names, const char* const* file_names,
                                       uint32_t n, uint32_t no_threads)
{
    if (!c || (n && (!sample_names || !file_names))) return AGCGPU_EINVAL;
    AGC_GUARD_BEGIN
    std::vector<std::pair<std::string, std::string>> v;
    for (uint32_t i = 0; i < n; ++i) v.emplace_back(sample_names[i], file_names[i]);
    return c->impl.AddSampleFiles(v, no_threads) ? 0 : AGCGPU_ECUDA;
    AGC_GUARD_END(c)
}

int agcgpu_compressor_add_samples_memory(agcgpu_compressor* c, const char* const* sample_names, uint32_t n_samples,
                                         const uint32_t* sample_of_contig, const char* const* contig_ids, uint32_t n_contigs,
                                         const void* raw, const uint64_t* offsets, int raw_is_device)
{
    if (!c || !sample_names || !sample_of_contig || !contig_ids || !raw || !offsets) return AGCGPU_EINVAL;
    CallTimer ct("add_samples_memory");
    AGC_GUARD_BEGIN
    std::vector<std::string> sn(sample_names, sample_names + n_samples), ci(contig_ids, contig_ids + n_contigs);
    std::vector<uint32_t> soc(sample_of_contig, sample_of_contig + n_contigs);
    for (auto s : soc) if (s >= n_samples) return AGCGPU_EINVAL;
    return c->impl.AddSamplesFromMemory(sn, soc, ci, (const uint8_t*)raw, offsets, raw_is_device != 0) ? 0 : AGCGPU_ECUDA;
    AGC_GUARD_END(c)
}

int agcgpu_compressor_set_discard_parts(agcgpu_compressor* c, int discard)
{
    if (!c) return AGCGPU_EINVAL;
    c->impl.SetDiscardParts(discard != 0);
    return 0;
}

int agcgpu_compressor_set_verify(agcgpu_compressor* c, int verify)
{
    if (!c) return AGCGPU_EINVAL;
    c->impl.SetVerify(verify != 0);
    return 0;
}

int agcgpu_compressor_add_cmd_line(agcgpu_compressor* c, const char* cmd_line)
{
    if (!c || !cmd_line) return AGCGPU_EINVAL;
    c->impl.AddCmdLine(cmd_line);
    return 0;
}

int agcgpu_compressor_close(agcgpu_compressor* c, uint32_t no_threads)
{
    if (!c) return AGCGPU_EINVAL;
    CallTimer ct("compressor_close");
    bool ok = false;
    try { ok = c->impl.Close(no_threads); }
    catch (const std::exception& e) { c->impl.SetLastError(std::string("exception: ") + e.what()); }
    catch (...) { c->impl.SetLastError("unknown exception"); }
    if (!ok) g_err = c->impl.LastError();
    if (c->impl.Ctx()) agcgpu_get_stats(c->impl.Ctx(), &g_last_stats);
    {   CallTimer cd("compressor delete");
        delete c; }
    return ok ? 0 : AGCGPU_ECUDA;
}

const char* agcgpu_compressor_last_error(const agcgpu_compressor* c) { return c ? c->impl.LastError().c_str() : g_err.c_str(); }
uint64_t agcgpu_compressor_total_bases(const agcgpu_compressor* c) { return c ? c->impl.TotalBases() : 0; }
agcgpu_ctx* agcgpu_compressor_ctx(agcgpu_compressor* c) { return c ? c->impl.Ctx() : nullptr; }
int agcgpu_compressor_last_stats(agcgpu_stats* out) { if (!out) return AGCGPU_EINVAL; *out = g_last_stats; return 0; }

}
