// TEST INFRASTRUCTURE ONLY: host build of agc_b200/csrc/zstd_dec.cuh (the frame decoder of the residual coder) so the CPU
// suite can check it against frames written by the reference's libzstd.  The product only contains the device build.
#include "../../agc_b200/csrc/zstd_dec.cuh"
#include <cstdlib>
extern "C" __attribute__((visibility("default"))) long zd_host_decompress(const unsigned char* src, unsigned long n, unsigned char* dst, unsigned long cap)
{
    zd::Work* w = (zd::Work*)malloc(sizeof(zd::Work));
    long r = (long)zd::decompress_frame(src, n, dst, cap, *w);
    free(w);
    return r;
}
