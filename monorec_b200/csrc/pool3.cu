// nn.MaxPool2d(kernel_size=3, stride=2, padding=1) on NHWC tensors: the pooling of torchvision's ResNet stem
// (reference: model/monorec/monorec_model.py:118-129 runs `encoder.maxpool` between conv1 and layer1).  One thread = one
// output pixel x 16 bytes of channels; taps outside the image are skipped (PyTorch pads with -inf).
#include "mr_common.cuh"
#include <cstdint>
#include <cuda_fp16.h>

namespace {

__device__ __forceinline__ uint4 vmax(uint4 a, uint4 b, bool half) {
    if (half) {
        const __half2* x = reinterpret_cast<const __half2*>(&a);
        const __half2* y = reinterpret_cast<const __half2*>(&b);
        uint4 r;
        __half2* o = reinterpret_cast<__half2*>(&r);
#pragma unroll
        for (int i = 0; i < 4; ++i) o[i] = __hmax2(x[i], y[i]);
        return r;
    }
    return make_uint4(__float_as_uint(fmaxf(__uint_as_float(a.x), __uint_as_float(b.x))),
                      __float_as_uint(fmaxf(__uint_as_float(a.y), __uint_as_float(b.y))),
                      __float_as_uint(fmaxf(__uint_as_float(a.z), __uint_as_float(b.z))),
                      __float_as_uint(fmaxf(__uint_as_float(a.w), __uint_as_float(b.w))));
}

template <bool HALF>
__global__ void __launch_bounds__(256) maxpool3s2_nhwc_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, int H, int W,
                                                              int CV, size_t total) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    const int c = (int)(i % CV);
    size_t r = i / CV;
    const int x = (int)(r % Wo); r /= Wo;
    const int y = (int)(r % Ho);
    const size_t b = r / Ho;
    // the centre tap (2y, 2x) is always inside the image: start from it
    uint4 acc = __ldg(src + ((b * H + 2 * y) * W + 2 * x) * CV + c);
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
        const int iy = 2 * y + dy;
        if (iy < 0 || iy >= H) continue;
#pragma unroll
        for (int dx = -1; dx <= 1; ++dx) {
            const int ix = 2 * x + dx;
            if (ix < 0 || ix >= W || (dy == 0 && dx == 0)) continue;
            acc = vmax(acc, __ldg(src + ((b * H + iy) * W + ix) * CV + c), HALF);
        }
    }
    dst[i] = acc;
}

template <bool HALF>
int launch(const void* src, void* dst, int B, int H, int W, int C, void* stream, const char* name) {
    constexpr int kVec = HALF ? 8 : 4;
    MR_REQUIRE(src && dst && B >= 1 && H >= 1 && W >= 1 && C >= kVec && (C % kVec) == 0, "%s: need C %% %d == 0 (got B=%d H=%d W=%d C=%d)", name,
               kVec, B, H, W, C);
    const size_t total = (size_t)B * ((H + 1) / 2) * ((W + 1) / 2) * (C / kVec);
    maxpool3s2_nhwc_kernel<HALF><<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const uint4*>(src), static_cast<uint4*>(dst), H, W, C / kVec, total);
    MR_LAUNCH_CHECK("maxpool3s2_nhwc_kernel");
    return MR_OK;
}

}  // namespace

extern "C" int mr_maxpool3s2_nhwc(const float* src, float* dst, int B, int H, int W, int C, void* stream) {
    return launch<false>(src, dst, B, H, W, C, stream, "mr_maxpool3s2_nhwc");
}

extern "C" int mr_maxpool3s2_nhwc_f16(const void* src, void* dst, int B, int H, int W, int C, void* stream) {
    return launch<true>(src, dst, B, H, W, C, stream, "mr_maxpool3s2_nhwc_f16");
}
