// kernels_zstd.cu -- residual coder (zstd 1.5.5 frame producer).  Placeholder until the device coder lands:
// fails loudly (AGCGPU_EUNSUPPORTED) -- there is deliberately no host fallback.
#include "internal.cuh"
extern "C" int agcgpu_zstd_compress_batch(agcgpu_ctx* ctx, const uint8_t*, const uint64_t*, const int32_t*, uint32_t,
                                          uint8_t*, uint64_t, uint64_t*)
{
    if (!ctx) return AGCGPU_EINVAL;
    return agc_fail(ctx, AGCGPU_EUNSUPPORTED, "agcgpu_zstd_compress_batch: device residual coder not built yet");
}
