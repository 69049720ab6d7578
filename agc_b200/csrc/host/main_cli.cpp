// agc-b200: `create` sub-command with the reference CLI's flags (src/app/application.cpp:125-169; defaults from
// src/app/application.h:24-84).  Everything per-base runs on the GPU through libagcgpu.
#include "compressor.h"
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <fstream>
#include <iostream>
#include <exception>
#include <unordered_set>

static void remove_common_suffixes(std::string& s)        // application.cpp:606-630
{
    const char* suf[] = { ".fna", ".gz", ".fa", ".fasta" };
    while (true) {
        bool removed = false;
        for (auto x : suf) {
            size_t l = strlen(x);
            if (s.length() <= l) continue;
            if (s.compare(s.length() - l, l, x) == 0) { s.resize(s.length() - l); removed = true; break; }
        }
        if (!removed) break;
    }
}

static int run(int argc, char** argv);
int main(int argc, char** argv)
{
    // a damaged input archive (append) or an allocation failure ends with a message and exit code 1, never with an abort
    try { return run(argc, argv); }
    catch (const std::exception& e) { std::cerr << "agc-b200: " << e.what() << std::endl; return 1; }
    catch (...) { std::cerr << "agc-b200: unknown error" << std::endl; return 1; }
}
static int run(int argc, char** argv)
{
    const bool is_append = argc >= 2 && std::string(argv[1]) == "append";
    if (argc < 3 || (std::string(argv[1]) != "create" && !is_append)) {
        std::cerr << "usage: agc-b200 append [-a] [-c] [-f frac] [-t n] [-v n] [-i list] [--device n] [--verify] -o out.agc in.agc [samples...]\n"
                     "       agc-b200 create [-k 31] [-l 20] [-s 60000] [-b 50] [-t n] [-v n] [-i list] [-d] [-a] [-c] [-f frac] [--device n] [--verify] [--dump-parts file] -o out.agc ref.fa [samples...]\n";
        return 1;
    }
    uint32_t k = 31, l = 20, s = 60000, b = 50, t = 1, v = 0; bool a = false, c = false, verify = false; double f = 0.0; int dev = 0;
    std::string out, list, dump;
    std::vector<std::string> inputs;
    for (int i = 2; i < argc; ++i) {
        std::string x = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) { std::cerr << "missing value for " << x << "\n"; exit(1); } return argv[++i]; };
        if (x == "-k") k = (uint32_t)atoi(next()); else if (x == "-l") l = (uint32_t)atoi(next()); else if (x == "-s") s = (uint32_t)atoi(next());
        else if (x == "-b") b = (uint32_t)atoi(next()); else if (x == "-t") t = (uint32_t)atoi(next()); else if (x == "-v") v = (uint32_t)atoi(next());
        else if (x == "-o") out = next(); else if (x == "-i") list = next(); else if (x == "-a") a = true; else if (x == "-c") c = true;
        else if (x == "-d") {} else if (x == "-f") f = atof(next()); else if (x == "--device") dev = atoi(next()); else if (x == "--dump-parts") dump = next(); else if (x == "--verify") verify = true;
        else inputs.push_back(x);
    }
    // b_value<T>::assign clamps every numeric option to its range (src/app/application.h:23-47, 63-71)
    k = std::clamp(k, 17u, 32u); b = std::clamp(b, 1u, 1000000000u); s = std::clamp(s, 100u, 1000000u); l = std::clamp(l, 15u, 32u);
    v = std::clamp(v, 0u, 2u); f = std::clamp(f, 0.0, 0.05);
    if (!list.empty()) { std::ifstream in(list); std::string ln; while (std::getline(in, ln)) if (!ln.empty()) inputs.push_back(ln); }
    if (inputs.empty()) { std::cerr << "need at least the reference FASTA (create) / the input archive (append)\n"; return 1; }   // no -o: the archive goes to stdout
    if (is_append) {                                           // src/app/main.cpp:124-160: the first positional argument is the archive to extend
        const std::string in_archive = inputs.front();
        inputs.erase(inputs.begin());
        { std::vector<std::string> u; std::unordered_set<std::string> seen; for (auto& x : inputs) if (seen.insert(x).second) u.push_back(x); inputs.swap(u); }
        agc_b200::CAGCCompressor agc;
        agc.SetDevice(dev); agc.SetVerify(verify);
        if (!agc.Append(in_archive, out, v, true, c, a, t, f)) { std::cerr << "Cannot extend archive " << in_archive << ": " << agc.LastError() << std::endl; return 1; }
        std::vector<std::pair<std::string, std::string>> files;
        for (auto& fn : inputs) { std::string nm = std::filesystem::path(fn).stem().string(); remove_common_suffixes(nm); files.emplace_back(nm, fn); }
        bool r = agc.AddSampleFiles(files, t);
        r &= agc.Close(t);
        return r ? 0 : 1;
    }
    { std::vector<std::string> u; std::unordered_set<std::string> seen; for (auto& x : inputs) if (seen.insert(x).second) u.push_back(x); inputs.swap(u); }
    agc_b200::CAGCCompressor agc;
    agc.SetDevice(dev);
    if (!dump.empty()) agc.SetDumpParts(dump);
    agc.SetVerify(verify);
    if (!agc.Create(out, b, k, inputs.front(), s, l, c, a, v, t, f)) { std::cerr << "Cannot create archive " << out << std::endl; return 1; }
    std::vector<std::pair<std::string, std::string>> files;
    for (auto& fn : inputs) { std::string nm = std::filesystem::path(fn).stem().string(); remove_common_suffixes(nm); files.emplace_back(nm, fn); }
    bool r = agc.AddSampleFiles(files, t);
    r &= agc.Close(t);
    return r ? 0 : 1;
}
