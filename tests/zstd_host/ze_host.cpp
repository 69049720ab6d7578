// TEST INFRASTRUCTURE ONLY: host build of agc_b200/csrc/zstd_enc.cuh so the CPU test-suite can diff the residual coder
// against the reference's libzstd without a GPU.  The product (libagcgpu.so) only contains the device build.
// Both instantiations the device library ships are built: the wide coder (512-slot window) and the narrow one (32 slots).
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
#include <cstdlib>
#include <cstring>
extern "C" {
__attribute__((visibility("default"))) long ze_host_compress(const unsigned char* src, unsigned long n, int level, unsigned char* dst, unsigned long cap)
{
    ze::Params cp = ze::get_params(level, n);
    if (!cp.supported) return -1;
    ze::WorkSizes z = ze::work_sizes(cp);
    unsigned char* mem = (unsigned char*)calloc(z.total + 64, 1);
    int err = 0;
    unsigned long r = ze::compress_frame(src, n, level, dst, cap, mem, &err);
    free(mem);
    return err ? -(long)err - 1 : (long)r;
}
__attribute__((visibility("default"))) long ze_host_compress_narrow(const unsigned char* src, unsigned long n, int level, unsigned char* dst, unsigned long cap)
{
    zen::Params cp = zen::get_params(level, n);
    if (!cp.supported) return -1;
    zen::WorkSizes z = zen::work_sizes(cp);
    unsigned char* mem = (unsigned char*)calloc(z.total + 64, 1);
    int err = 0;
    unsigned long r = zen::compress_frame(src, n, level, dst, cap, mem, &err);
    free(mem);
    return err ? -(long)err - 1 : (long)r;
}
__attribute__((visibility("default"))) unsigned long ze_host_bound(unsigned long n) { return ze::compress_bound(n); }
}
