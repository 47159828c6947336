// Host-side packing of convolution weights for the tensor-core path (include/monorec_b200.h: mr_pack_conv_weights ...).
// Pure host code: a C/C++ host that reads a MonoRec checkpoint itself (no Python) needs exactly this to drive
// mr_conv2d_nhwc_tc.  Reference layers being packed: nn.Conv2d / ConvTranspose2d weights of model/layers.py:289-400 as used by
// MaskModule / DepthModule (model/monorec/monorec_model.py:287-385, :476-557).
#include "mr_common.cuh"
#include <cuda_fp16.h>
#include <cstdint>
#include <cstring>

namespace {

constexpr int kKC = 32;   // fp32 channels per 128-byte swizzle row (conv_tc.cu)

// K chunk width: 32 fp32 channels; 64 half channels, or 32 when every source has <= 32 channels (64-byte swizzle rows)
int chunk_channels(int n_src, const int* src_c, int dtype) {
    if (dtype != MR_DT_F16) return kKC;
    for (int s = 0; s < n_src; ++s)
        if (src_c[s] > 32) return 64;
    return 32;
}

float round_tf32(float x) {   // round-to-nearest onto the TF32 grid (10 explicit mantissa bits); the tensor core truncates the rest
    uint32_t b;
    std::memcpy(&b, &x, 4);
    b = (b + 0x1000u) & ~0x1FFFu;
    std::memcpy(&x, &b, 4);
    return x;
}

}  // namespace

extern "C" long long mr_pack_conv_weights_bytes(int Cout, int n_src, const int* src_c, int kh, int kw, int dtype, int* n_pad,
                                                int* k_pad) {
    if (Cout < 1 || n_src < 1 || n_src > MR_CONV_MAX_SRC || src_c == nullptr || kh < 1 || kw < 1 ||
        (dtype != MR_DT_F32 && dtype != MR_DT_F16))
        return -1;
    const int kc = chunk_channels(n_src, src_c, dtype);
    int kp = 0;
    for (int s = 0; s < n_src; ++s) {
        if (src_c[s] < 1) return -1;
        kp += (src_c[s] + kc - 1) / kc * kc;
    }
    const int np = (Cout + 15) / 16 * 16;
    if (n_pad) *n_pad = np;
    if (k_pad) *k_pad = kp;
    return (long long)kh * kw * np * kp * (dtype == MR_DT_F16 ? 2 : 4);
}

extern "C" int mr_pack_conv_weights(const float* w, int Cout, int n_src, const int* src_c, int kh, int kw, int dtype, void* out) {
    int n_pad = 0, k_pad = 0;
    const long long bytes = mr_pack_conv_weights_bytes(Cout, n_src, src_c, kh, kw, dtype, &n_pad, &k_pad);
    MR_REQUIRE(bytes > 0 && w != nullptr && out != nullptr, "mr_pack_conv_weights: bad arguments");
    const int kc = chunk_channels(n_src, src_c, dtype);
    int Cin = 0;
    for (int s = 0; s < n_src; ++s) Cin += src_c[s];
    std::memset(out, 0, (size_t)bytes);
    float* of = static_cast<float*>(out);
    __half* oh = static_cast<__half*>(out);
    for (int t = 0; t < kh * kw; ++t)
        for (int n = 0; n < Cout; ++n) {
            int ci = 0, ko = 0;
            for (int s = 0; s < n_src; ++s) {
                for (int c = 0; c < src_c[s]; ++c) {
                    const float v = w[((size_t)n * Cin + ci + c) * kh * kw + t];   // (Cout, Cin, kh, kw), tap t = ky * kw + kx
                    const size_t o = ((size_t)t * n_pad + n) * k_pad + ko + c;
                    if (dtype == MR_DT_F16) oh[o] = __float2half_rn(v);
                    else of[o] = round_tf32(v);
                }
                ci += src_c[s];
                ko += (src_c[s] + kc - 1) / kc * kc;
            }
        }
    return MR_OK;
}

// ConvTranspose2d(k = 4, s = 2) + crop (Refine, model/layers.py:380-400) as four 2x2 correlations, one per output phase
// (py, px): out[2 oy + py, 2 ox + px]; phase 0 uses kernel rows (3, 1) with one pixel of top padding, phase 1 rows (2, 0).
extern "C" int mr_subpixel_convt_k4s2(const float* w, int Cin, int Cout, int py, int px, float* out, int* pad_t, int* pad_l) {
    MR_REQUIRE(w && out && Cin >= 1 && Cout >= 1 && (py | 1) == 1 && (px | 1) == 1, "mr_subpixel_convt_k4s2: bad arguments");
    const int taps[2][2] = {{3, 1}, {2, 0}};
    for (int n = 0; n < Cout; ++n)
        for (int c = 0; c < Cin; ++c)
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b)
                    out[(((size_t)n * Cin + c) * 2 + a) * 2 + b] = w[(((size_t)c * Cout + n) * 4 + taps[py][a]) * 4 + taps[px][b]];
    if (pad_t) *pad_t = 1 - py;
    if (pad_l) *pad_l = 1 - px;
    return MR_OK;
}

// Upsample(x2, nearest) + pad(0,1,0,1) + Conv2d(k = 2) (Upconv, model/layers.py:338-356) per output phase: an even phase sees
// both taps of that axis on the same input pixel (the weights add up, kernel extent 1), an odd phase sees pixels o and o + 1.
extern "C" int mr_subpixel_upconv2(const float* w, int Cout, int Cin, int py, int px, float* out, int* kh_out, int* kw_out) {
    MR_REQUIRE(w && out && Cin >= 1 && Cout >= 1 && (py | 1) == 1 && (px | 1) == 1, "mr_subpixel_upconv2: bad arguments");
    const int kh = py ? 2 : 1, kw = px ? 2 : 1;
    for (int n = 0; n < Cout; ++n)
        for (int c = 0; c < Cin; ++c) {
            const float* s = w + ((size_t)n * Cin + c) * 4;
            for (int a = 0; a < kh; ++a)
                for (int b = 0; b < kw; ++b) {
                    float v;
                    if (py && px) v = s[a * 2 + b];
                    else if (py) v = s[a * 2] + s[a * 2 + 1];
                    else if (px) v = s[b] + s[2 + b];
                    else v = (s[0] + s[2]) + (s[1] + s[3]);
                    out[((size_t)n * Cin + c) * kh * kw + a * kw + b] = v;
                }
        }
    if (kh_out) *kh_out = kh;
    if (kw_out) *kw_out = kw;
    return MR_OK;
}

extern "C" long long mr_conv_workspace_bytes(const mr_conv_desc* desc) {
    (void)desc;
    return 0;   // both convolution kernels stage everything in shared / tensor memory: no device workspace is needed
}
