// Fused plane-sweep cost volume for sm_100a (B200).
//
// Replaces CostVolumeModule.forward (reference: model/monorec/monorec_model.py:150-280) together with
// Backprojection / point_projection (model/layers.py:43-71), F.grid_sample x2, SSIM (layers.py:119-137), the
// conv3d patch cost (:246-248) and the view weighting / fusion (:257-269).  Closed form: SURVEY.md Appendix C.
//
// Work decomposition ("warp-march"):
//   CTA        = one keyframe tile of 60 x TH output pixels (64 x (TH+4) with the 2-px stencil halo), all D planes,
//                all F source frames.  grid = (ceil(W/60), ceil(H/TH), B).
//   warp       = one depth plane at a time (planes d = warp, warp+NW, ...).  The warp marches down the tile's rows:
//                stage 1 (lane = column):  homography of the row's 64 pixels, 12 bilinear taps straight from the
//                                          L1/L2-resident source frame, warped row -> per-warp smem row buffer;
//                stage 2 (lane = 2 columns): 3x3 box sums of X, X^2, XY per channel as horizontal sums in registers
//                                          and a rolling vertical sum, SSIM error, channel weighting, second 3x3
//                                          box (horizontal neighbours by shuffle, vertical rolling) -> sad[d][row][col].
//   CTA phase 2 (thread = pixel): min_d / sum_d exp(..) view weight, single-frame volume written to HBM once,
//                weights kept in smem; after the last frame the fused volume is formed from the L2-hot single-frame
//                volumes this thread wrote itself (no intermediate tensor, each output element written once).
// Keyframe-only terms (mu_y, sigma_y + C2) are hoisted into a smem table per tile; pixels whose reprojection leaves
// the source for any plane (valid_f = 0) are found by a projection-only pre-pass and whole row ranges / frames of the
// tile are skipped.
#include "mr_common.cuh"
#include <cstdint>
#include <type_traits>

namespace {

constexpr int kTileCols = 64;   // buffer columns per tile row (output columns + 2-px halo each side)
constexpr int kOutCols = 60;    // output columns per tile
constexpr int kRowStride = 68;  // floats per smem image row: column b lives at index b+1 (so [2l-1, 2l+2] is 8B aligned)
#ifndef MR_CV_THREADS
#define MR_CV_THREADS 512
#endif
constexpr int kThreads = MR_CV_THREADS;
#ifndef MR_CV_MINBLOCKS
#define MR_CV_MINBLOCKS 1      // resident CTAs per SM the register allocator must leave room for
#endif

constexpr int kWarps = kThreads / 32;
constexpr float kC1 = 0.01f * 0.01f;  // layers.py:116
constexpr float kC2 = 0.03f * 0.03f;  // layers.py:117

struct CvArgs {
    const float* key;                    // [B,3,H,W]
    const float* frames[MR_MAX_FRAMES];  // each [B,3,H,W]
    const float* proj;                   // [B,F,12]
    const float* depths;                 // [D]
    float* cv;                           // [B,D,H,W]
    float* sfcv;                         // [F,B,D,H,W]
    const float4* packed;                // [F,B,H,W] (r,g,b,0) copies of the source frames, or nullptr (planar gather)
    int B, F, D, H, W, TH, b0;
    float alpha, inv_dm1;
    float cw0, cw1, cw2;                 // channel weights / 9
};

struct SmemLayout {
    int ytile, cst, xbuf, pjs, zs, vmask, rowrng, total;  // byte offsets
};

__host__ __device__ inline SmemLayout make_layout(int D, int TH, int F) {
    SmemLayout L;
    int off = 0;
    L.ytile = off; off += 3 * (TH + 4) * kRowStride * 4;
    L.cst = off;   off += 3 * (TH + 2) * kTileCols * 8;
    L.xbuf = off;  off += kWarps * 3 * kRowStride * 4;
    L.pjs = off;   off += F * 12 * 4;
    L.zs = off;    off += ((D + 3) / 4) * 16;
    L.vmask = off; off += F * TH * kTileCols;
    L.rowrng = off; off += F * 2 * 4;
    L.total = off;
    return L;
}

// ---- packed fp32x2 helpers (FADD2 / FMUL2 / FFMA2 on sm_100a; a pair is either two columns or two samples) ----------
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }

// single MUFU.RCP (flush-to-zero variant: no denormal pre/post-scaling code; operands here are never denormal)
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

// L2 residency hints: the single-frame volume is written by the march and read back once by the per-pixel phase of the same
// CTA, so its lines are asked to stay (evict_last); the fused volume is written once and never read here (evict_first).
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void st_hint_f2(float* ptr, float2 v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;" ::"l"(ptr), "f"(v.x), "f"(v.y), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_hint_f1(float* ptr, float v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.f32 [%0], %1, %2;" ::"l"(ptr), "f"(v), "l"(pol) : "memory");
}


constexpr int kChunk = 32;                 // planes the per-pixel phase keeps in registers at once
constexpr float kMagic = 12582912.0f;      // 1.5 * 2^23: adding it rounds to the nearest integer in the low mantissa bits
constexpr int kMagicBits = 0x4B400000;

// SSIM numerator / denominator for a pair of columns (layers.py:123-134 through 3x3 box sums; mu_y, sigma_y + C2 hoisted)
__device__ __forceinline__ void ssim_nd(float2 s1, float2 sxx, float2 sxy, float2 mu_y, float2 sy2, float2& n, float2& d) {
    const float2 k9 = bc2(1.0f / 9.0f);
    float2 mu_x = mul2(s1, k9);
    float2 mxy = mul2(mu_x, mu_y);
    float2 mxx = mul2(mu_x, mu_x);
    float2 sig_xy = fma2(sxy, k9, neg2(mxy));
    float2 sig_x = fma2(sxx, k9, neg2(mxx));
    n = mul2(fma2(bc2(2.0f), mxy, bc2(kC1)), fma2(bc2(2.0f), sig_xy, bc2(kC2)));
    d = mul2(add2(mxx, fma2(mu_y, mu_y, bc2(kC1))), add2(sig_x, sy2));
}

// ---- stage 1 of the march: homography of one tile row (pair = columns lane, lane + 32) and its 24 bilinear taps -------
struct Stage1Ctx {
    float2 pzx, pzy, pzz;             // per-lane column part of the projection (already times the plane depth)
    float rax, rbx, ray, rby, raz, rbz;  // per-row part: c = pz(u) + ra * v + rb
    const float* img;                 // source frame of this batch element, [3][H][W]
    const float4* img4;               // the same frame as [H][W] (r,g,b,0) pixels (packed gather) or nullptr
    int W, H, planei, v0, lane;
    float sx_lo, sx_hi, sy_lo, sy_hi; // == grid clamp(-2, 2), monorec_model.py:208
};

__device__ __forceinline__ void setup_stage1(Stage1Ctx& c, const float* m, const float* img, float z, float2 fu2) {
    // projection c = M [u v 1]^T z + p split into a per-lane column part and a per-row part
    c.pzx = mul2(mul2(bc2(m[0]), fu2), bc2(z));
    c.pzy = mul2(mul2(bc2(m[4]), fu2), bc2(z));
    c.pzz = mul2(mul2(bc2(m[8]), fu2), bc2(z));
    c.rax = m[1] * z; c.rbx = fmaf(m[2], z, m[3]);
    c.ray = m[5] * z; c.rby = fmaf(m[6], z, m[7]);
    c.raz = m[9] * z; c.rbz = fmaf(m[10], z, m[11]);
    c.img = img;
}

template <bool PACKED>
__device__ __forceinline__ void warp_row(const Stage1Ctx& c, const int r, float* __restrict__ xrow) {
    const int W = c.W, H = c.H;
    const float fv = (float)(c.v0 + r);
    const float rcx = fmaf(c.rax, fv, c.rbx), rcy = fmaf(c.ray, fv, c.rby), rcz = fmaf(c.raz, fv, c.rbz);
    const float2 cx = add2(c.pzx, bc2(rcx)), cy = add2(c.pzy, bc2(rcy)), cz = add2(c.pzz, bc2(rcz));
    const float2 inv = make_float2(fast_rcp(cz.x), fast_rcp(cz.y));
    const float2 ux = mul2(cx, inv), uy = mul2(cy, inv);
    // floor by magic-number rounding: rn(s - 0.5) differs from floor(s) only for integral s, where the interpolated value
    // is the same (weight 1 on the tap both conventions share)
    const float2 tx = add2(ux, bc2(kMagic - 1.0f)), ty = add2(uy, bc2(kMagic - 1.0f));
    const int x0a = __float_as_int(tx.x) - kMagicBits, x0b = __float_as_int(tx.y) - kMagicBits;
    const int y0a = __float_as_int(ty.x) - kMagicBits, y0b = __float_as_int(ty.y) - kMagicBits;
    const bool inb = ((unsigned)x0a <= (unsigned)(W - 2)) && ((unsigned)x0b <= (unsigned)(W - 2)) &&
                     ((unsigned)y0a <= (unsigned)(H - 2)) && ((unsigned)y0b <= (unsigned)(H - 2));
    float2 w00, w01, w10, w11;
    int oa, ob, dxa, dxb, dya, dyb;
    if (__all_sync(0xffffffffu, inb)) {
        // fast path: all 4 taps of every lane are inside the image
        const float2 x0f = add2(tx, bc2(-kMagic)), y0f = add2(ty, bc2(-kMagic));
        const float2 wx1 = add2(add2(ux, bc2(-0.5f)), neg2(x0f)), wy1 = add2(add2(uy, bc2(-0.5f)), neg2(y0f));
        const float2 wx0 = add2(bc2(1.0f), neg2(wx1)), wy0 = add2(bc2(1.0f), neg2(wy1));
        w00 = mul2(wx0, wy0); w01 = mul2(wx1, wy0); w10 = mul2(wx0, wy1); w11 = mul2(wx1, wy1);
        oa = y0a * W + x0a; ob = y0b * W + x0b;
        dxa = dxb = 1; dya = dyb = W;
    } else {
        // border path: per-tap zero padding exactly like F.grid_sample(padding_mode="zeros"): clamp the tap address, zero
        // the weight of every tap that falls outside the image
        float wx0s[2], wx1s[2], wy0s[2], wy1s[2];
        int os[2], dxs[2], dys[2];
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            float sxk = (k ? ux.y : ux.x) - 0.5f, syk = (k ? uy.y : uy.x) - 0.5f;
            sxk = fminf(fmaxf(sxk, c.sx_lo), c.sx_hi);
            syk = fminf(fmaxf(syk, c.sy_lo), c.sy_hi);
            const float x0f = floorf(sxk), y0f = floorf(syk);
            float wx1 = sxk - x0f, wy1 = syk - y0f;
            float wx0 = (x0f + 1.0f) - sxk, wy0 = (y0f + 1.0f) - syk;
            const int x0 = (int)x0f, y0 = (int)y0f;
            if ((unsigned)x0 >= (unsigned)W) wx0 = 0.f;
            if ((unsigned)(x0 + 1) >= (unsigned)W) wx1 = 0.f;
            if ((unsigned)y0 >= (unsigned)H) wy0 = 0.f;
            if ((unsigned)(y0 + 1) >= (unsigned)H) wy1 = 0.f;
            const int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
            const int ya = min(max(y0, 0), H - 1), yb = min(max(y0 + 1, 0), H - 1);
            wx0s[k] = wx0; wx1s[k] = wx1; wy0s[k] = wy0; wy1s[k] = wy1;
            os[k] = ya * W + xa; dxs[k] = xb - xa; dys[k] = (yb - ya) * W;
        }
        const float2 wx0 = make_float2(wx0s[0], wx0s[1]), wx1 = make_float2(wx1s[0], wx1s[1]);
        const float2 wy0 = make_float2(wy0s[0], wy0s[1]), wy1 = make_float2(wy1s[0], wy1s[1]);
        w00 = mul2(wx0, wy0); w01 = mul2(wx1, wy0); w10 = mul2(wx0, wy1); w11 = mul2(wx1, wy1);
        oa = os[0]; ob = os[1]; dxa = dxs[0]; dxb = dxs[1]; dya = dys[0]; dyb = dys[1];
    }
    if (PACKED) {
        // one 16-byte load per tap: 8 requests per row step instead of 24 (stage 1 is bound by L1/LSU requests)
        const float4* qa = c.img4 + oa;
        const float4* qb = c.img4 + ob;
        const float4 a00 = __ldg(qa), a01 = __ldg(qa + dxa), a10 = __ldg(qa + dya), a11 = __ldg(qa + dya + dxa);
        const float4 b00 = __ldg(qb), b01 = __ldg(qb + dxb), b10 = __ldg(qb + dyb), b11 = __ldg(qb + dyb + dxb);
        float2 v0 = fma2(make_float2(a00.x, b00.x), w00, bc2(0.5f));   // + 0.5: monorec_model.py:231
        float2 v1 = fma2(make_float2(a00.y, b00.y), w00, bc2(0.5f));
        float2 v2 = fma2(make_float2(a00.z, b00.z), w00, bc2(0.5f));
        v0 = fma2(make_float2(a01.x, b01.x), w01, v0); v1 = fma2(make_float2(a01.y, b01.y), w01, v1); v2 = fma2(make_float2(a01.z, b01.z), w01, v2);
        v0 = fma2(make_float2(a10.x, b10.x), w10, v0); v1 = fma2(make_float2(a10.y, b10.y), w10, v1); v2 = fma2(make_float2(a10.z, b10.z), w10, v2);
        v0 = fma2(make_float2(a11.x, b11.x), w11, v0); v1 = fma2(make_float2(a11.y, b11.y), w11, v1); v2 = fma2(make_float2(a11.z, b11.z), w11, v2);
        xrow[c.lane + 1] = v0.x;                  xrow[c.lane + 33] = v0.y;
        xrow[kRowStride + c.lane + 1] = v1.x;     xrow[kRowStride + c.lane + 33] = v1.y;
        xrow[2 * kRowStride + c.lane + 1] = v2.x; xrow[2 * kRowStride + c.lane + 33] = v2.y;
    } else {
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float* pa0 = c.img + (oa + ch * c.planei);
            const float* pb0 = c.img + (ob + ch * c.planei);
            const float* pa1 = pa0 + dya;
            const float* pb1 = pb0 + dyb;
            const float2 i00 = make_float2(__ldg(pa0), __ldg(pb0));
            const float2 i01 = make_float2(__ldg(pa0 + dxa), __ldg(pb0 + dxb));
            const float2 i10 = make_float2(__ldg(pa1), __ldg(pb1));
            const float2 i11 = make_float2(__ldg(pa1 + dxa), __ldg(pb1 + dxb));
            float2 val = fma2(i00, w00, bc2(0.5f));   // + 0.5: monorec_model.py:231
            val = fma2(i01, w01, val);
            val = fma2(i10, w10, val);
            val = fma2(i11, w11, val);
            xrow[ch * kRowStride + c.lane + 1] = val.x;
            xrow[ch * kRowStride + c.lane + 33] = val.y;
        }
    }
}

// ---- stage 2 of the march: SSIM + patch cost of one row; lane owns buffer columns 2l, 2l+1 (a pair) ---------------------
struct Stage2Ctx {
    const float* ys_l;       // keyframe tile (+0.5), this lane's columns
    const float4* cs_l;      // hoisted (mu_y, sigma_y + C2) table, this lane's column pair
    int ych, cch;            // channel strides of the two tables
    float2 cw0, cw1, cw2;    // channel weights / 9
    float* out_d;            // single-frame volume plane, tile row 0, this lane's columns
    int W, rows_left;        // rows_left = H - v0
    bool st_pair, st0, st1;
    uint64_t pol_keep;
};

// rolling state: horizontal 3-sums of X, X^2, XY per channel for the last rows, indexed by (row step mod 3) so that no
// register moves are needed
struct Stage2State {
    float2 hs1[3][3], hsx[3][3], hsy[3][3], hE[3];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            hE[i] = bc2(0.f);
#pragma unroll
            for (int c = 0; c < 3; ++c) hs1[i][c] = hsx[i][c] = hsy[i][c] = bc2(0.f);
        }
    }
};

// STAGE selects how far the two cascaded 3x3 windows are filled: 0 = rows 0,1 of a unit (only the horizontal sums are
// recorded), 1 = rows 2,3 (SSIM error row available, patch window not yet), 2 = steady state (a cost row is stored).
// Making it a template parameter keeps the steady-state step one branch-free block, so the three channels' dependency
// chains are scheduled against each other.
template <int P, int STAGE>
__device__ __forceinline__ void ssim_row(Stage2State& st, const Stage2Ctx& c, const int r, const float* __restrict__ xs_l) {
    constexpr int P1 = (P + 1) % 3, P2 = (P + 2) % 3;
    const float* yrow = c.ys_l + (r + 2) * kRowStride;
    float2 h1[3], hx[3], hy[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float2 xl = *reinterpret_cast<const float2*>(xs_l + ch * kRowStride);      // cols 2l-1, 2l
        const float2 xr = *reinterpret_cast<const float2*>(xs_l + ch * kRowStride + 2);  // cols 2l+1, 2l+2
        const float2 yl = *reinterpret_cast<const float2*>(yrow + ch * c.ych);
        const float2 yr = *reinterpret_cast<const float2*>(yrow + ch * c.ych + 2);
        const float2 xxl = mul2(xl, xl), xxr = mul2(xr, xr), xyl = mul2(xl, yl), xyr = mul2(xr, yr);
        const float m1 = xl.y + xr.x, mx = xxl.y + xxr.x, my = xyl.y + xyr.x;
        h1[ch] = make_float2(xl.x + m1, m1 + xr.y);
        hx[ch] = make_float2(xxl.x + mx, mx + xxr.y);
        hy[ch] = make_float2(xyl.x + my, my + xyr.y);
    }
    if (STAGE >= 1) {
        float2 nn[3], dd[3];
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float4 k4 = c.cs_l[ch * c.cch + r * (kTileCols / 2)];
            ssim_nd(add2(add2(st.hs1[P1][ch], st.hs1[P2][ch]), h1[ch]), add2(add2(st.hsx[P1][ch], st.hsx[P2][ch]), hx[ch]),
                    add2(add2(st.hsy[P1][ch], st.hsy[P2][ch]), hy[ch]), make_float2(k4.x, k4.y), make_float2(k4.z, k4.w),
                    nn[ch], dd[ch]);
        }
        const float2 q0 = mul2(nn[0], make_float2(fast_rcp(dd[0].x), fast_rcp(dd[0].y)));
        const float2 q1 = mul2(nn[1], make_float2(fast_rcp(dd[1].x), fast_rcp(dd[1].y)));
        const float2 q2 = mul2(nn[2], make_float2(fast_rcp(dd[2].x), fast_rcp(dd[2].y)));
        // clamp((1 - q) / 2, 0, 1)   (layers.py:137)
        const float2 e0 = make_float2(__saturatef(fmaf(-0.5f, q0.x, 0.5f)), __saturatef(fmaf(-0.5f, q0.y, 0.5f)));
        const float2 e1 = make_float2(__saturatef(fmaf(-0.5f, q1.x, 0.5f)), __saturatef(fmaf(-0.5f, q1.y, 0.5f)));
        const float2 e2 = make_float2(__saturatef(fmaf(-0.5f, q2.x, 0.5f)), __saturatef(fmaf(-0.5f, q2.y, 0.5f)));
        const float2 E = fma2(c.cw2, e2, fma2(c.cw1, e1, mul2(c.cw0, e0)));
        const float eL = __shfl_up_sync(0xffffffffu, E.y, 1);
        const float eR = __shfl_down_sync(0xffffffffu, E.x, 1);
        const float mid = E.x + E.y;
        const float2 hEc = make_float2(eL + mid, mid + eR);
        if (STAGE >= 2) {
            // single-frame volume 1 - 2 sad (monorec_model.py:251) straight to HBM; the validity mask is applied by the
            // per-pixel phase (which zeroes invalid pixels) once all planes are known
            const float2 sad = add2(add2(st.hE[P1], st.hE[P2]), hEc);
            const float2 sv = fma2(bc2(-2.0f), sad, bc2(1.0f));
            float* o = c.out_d + (size_t)(r - 2) * c.W;
            if (r - 2 < c.rows_left) {
                if (c.st_pair) {
                    st_hint_f2(o, sv, c.pol_keep);
                } else {
                    if (c.st0) st_hint_f1(o, sv.x, c.pol_keep);
                    if (c.st1) st_hint_f1(o + 1, sv.y, c.pol_keep);
                }
            }
        }
        st.hE[P] = hEc;
    }
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) { st.hs1[P][ch] = h1[ch]; st.hsx[P][ch] = hx[ch]; st.hsy[P][ch] = hy[ch]; }
}

template <bool PACKED>
__global__ void __launch_bounds__(kThreads, MR_CV_MINBLOCKS) cost_volume_kernel(const CvArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout L = make_layout(a.D, a.TH, a.F);
    float* ytile = reinterpret_cast<float*>(smem + L.ytile);
    float* cst = reinterpret_cast<float*>(smem + L.cst);
    float* pjs = reinterpret_cast<float*>(smem + L.pjs);
    float* zs = reinterpret_cast<float*>(smem + L.zs);
    unsigned char* vmask = smem + L.vmask;
    int* rowrng = reinterpret_cast<int*>(smem + L.rowrng);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, D = a.D, TH = a.TH, F = a.F;
    const int b = blockIdx.z + a.b0;
    const int u0 = blockIdx.x * kOutCols - 2;  // image column of buffer column 0
    const int v0 = blockIdx.y * TH;            // image row of tile row 0
    const size_t plane = (size_t)H * W;
    const int planei = H * W;
    float* xbuf = reinterpret_cast<float*>(smem + L.xbuf) + warp * 3 * kRowStride;

    // ---- keyframe tile (+0.5, monorec_model.py:232) and hoisted SSIM terms -------------------------------------
    const float* key = a.key + (size_t)b * 3 * plane;
    for (int i = tid; i < 3 * (TH + 4) * 66; i += kThreads) {
        int idx = i % 66, t = i / 66, rr = t % (TH + 4), ch = t / (TH + 4);
        int u = u0 + idx - 1, v = v0 - 2 + rr;
        float val = 0.f;
        if (u >= 0 && u < W && v >= 0 && v < H) val = __ldg(key + ch * plane + (size_t)v * W + u) + 0.5f;
        ytile[(ch * (TH + 4) + rr) * kRowStride + idx] = val;
    }
    for (int i = tid; i < D; i += kThreads) zs[i] = __ldg(a.depths + i);
    if (lane < 3) { xbuf[lane * kRowStride] = 0.f; xbuf[lane * kRowStride + kTileCols + 1] = 0.f; }
    __syncthreads();
    // table entry for the column pair (2j, 2j+1) of e-row er, channel ch: (mu_y[2j], mu_y[2j+1], sy2[2j], sy2[2j+1])
    for (int i = tid; i < 3 * (TH + 2) * kTileCols; i += kThreads) {
        int bc = i % kTileCols, t = i / kTileCols, er = t % (TH + 2), ch = t / (TH + 2);
        const float* y = ytile + (ch * (TH + 4) + er) * kRowStride + bc;  // rows er..er+2, idx bc..bc+2
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
                float q = y[dy * kRowStride + dx];
                s1 += q;
                s2 = fmaf(q, q, s2);
            }
        float mu = s1 * (1.0f / 9.0f);
        float sy2 = fmaf(s2, 1.0f / 9.0f, -mu * mu) + kC2;
        float* dst = cst + ((ch * (TH + 2) + er) * (kTileCols / 2) + (bc >> 1)) * 4 + (bc & 1);
        dst[0] = mu;
        dst[2] = sy2;
    }

    const float fW = (float)W, fH = (float)H;
    const float sx_lo = -(fW + 1.f) * 0.5f, sx_hi = (3.f * fW - 1.f) * 0.5f;  // == grid clamp(-2, 2), monorec_model.py:208
    const float sy_lo = -(fH + 1.f) * 0.5f, sy_hi = (3.f * fH - 1.f) * 0.5f;
    const float2 fu2 = make_float2((float)(u0 + lane), (float)(u0 + lane + 32));
    const float2 cw0 = bc2(a.cw0), cw1 = bc2(a.cw1), cw2 = bc2(a.cw2);
    // per-lane smem bases for stage 2 (columns 2l-1 .. 2l+2 live at float index 2l .. 2l+3 of a row)
    const float* ys_l = ytile + 2 * lane;
    const float4* cs_l = reinterpret_cast<const float4*>(cst) + lane;
    const int ych = (TH + 4) * kRowStride;        // ytile channel stride (floats)
    const int cch = (TH + 2) * (kTileCols / 2);   // cst channel stride (float4)
    // stage 2 writes the single-frame volume for output columns u0 + 2l, u0 + 2l + 1 (lanes 1..30)
    const int ucol = u0 + 2 * lane;
    const bool st0 = (lane >= 1) && (lane <= 30) && (ucol < W);
    const bool st1 = (lane >= 1) && (lane <= 30) && (ucol + 1 < W);
    const bool st_pair = st0 && st1 && ((W & 1) == 0);
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();

    // ---- validity pre-pass for every frame: valid_f(v,u) = interior(v,u) & all_d [ sample strictly inside
    //      (1,W-2)x(1,H-2) ]  (monorec_model.py:212-219: bilinear sample of the interior mask != 0 for every plane) ----
    if (tid < 2 * F) rowrng[tid] = (tid & 1) ? -1 : TH;
    if (tid < 12 * F) pjs[tid] = __ldg(a.proj + (size_t)b * F * 12 + tid);
    __syncthreads();
    for (int q = tid; q < F * TH * kTileCols; q += kThreads) {
        const int f = q / (TH * kTileCols), p = q - f * (TH * kTileCols);
        const int r = p >> 6, bc = p & 63;
        const int u = u0 + bc, v = v0 + r;
        bool ok = (bc >= 2) && (bc < 2 + kOutCols) && (u >= 2) && (u < W - 2) && (v >= 2) && (v < H - 2);
        if (ok) {
            const float* m = pjs + 12 * f;
            const float fu = (float)u, fv = (float)v;
            const float ax = fmaf(m[0], fu, fmaf(m[1], fv, m[2]));
            const float ay = fmaf(m[4], fu, fmaf(m[5], fv, m[6]));
            const float az = fmaf(m[8], fu, fmaf(m[9], fv, m[10]));
            const float m03 = m[3], m13 = m[7], m23 = m[11];
            for (int d = 0; d < D; ++d) {
                const float z = zs[d];
                const float inv = fast_rcp(fmaf(az, z, m23));
                const float sx = fmaf(fmaf(ax, z, m03), inv, -0.5f);
                const float sy = fmaf(fmaf(ay, z, m13), inv, -0.5f);
                ok = ok && (sx > 1.0f) && (sx < fW - 2.0f) && (sy > 1.0f) && (sy < fH - 2.0f);
            }
        }
        vmask[q] = ok ? 1 : 0;
        if (ok) { atomicMin(&rowrng[2 * f], r); atomicMax(&rowrng[2 * f + 1], r); }
    }
    __syncthreads();

    // ---- march over the F*D (frame, plane) units; no CTA-wide barrier in here ------------------------------------
    Stage2Ctx c2;
    c2.ys_l = ys_l; c2.cs_l = cs_l; c2.ych = ych; c2.cch = cch;
    c2.cw0 = cw0; c2.cw1 = cw1; c2.cw2 = cw2;
    c2.W = W; c2.rows_left = H - v0; c2.st_pair = st_pair; c2.st0 = st0; c2.st1 = st1; c2.pol_keep = pol_keep;
    Stage1Ctx c1;
    c1.W = W; c1.H = H; c1.planei = planei; c1.v0 = v0; c1.lane = lane;
    c1.sx_lo = sx_lo; c1.sx_hi = sx_hi; c1.sy_lo = sy_lo; c1.sy_hi = sy_hi;
#ifndef MR_CV_SKIP
#define MR_CV_SKIP 0     // timing experiments only: 1 = no march, 2 = no per-pixel phase, 3 = march without stage 2, 4 = march without stage 1
#endif
    for (int unit = warp; unit < F * D && MR_CV_SKIP != 1; unit += kWarps) {
        const int f = unit / D, d = unit - f * D;
        const int rlo = rowrng[2 * f], rhi = rowrng[2 * f + 1];
        if (rhi < rlo) continue;  // no valid pixel of this tile for frame f: the per-pixel phase zero-fills
        setup_stage1(c1, pjs + 12 * f, a.frames[f] + (size_t)b * 3 * plane, zs[d], fu2);
        c1.img4 = PACKED ? a.packed + ((size_t)f * a.B + b) * plane : nullptr;
        Stage2State st;
        st.clear();
        c2.out_d = a.sfcv + (((size_t)f * a.B + b) * D + d) * plane + (size_t)v0 * W + ucol;
        const int nsteps = rhi - rlo + 5;
        auto step = [&](auto tag, auto stage, const int t) {
            if (MR_CV_SKIP != 4) warp_row<PACKED>(c1, rlo - 2 + t, xbuf);
            __syncwarp();
            if (MR_CV_SKIP != 3) ssim_row<decltype(tag)::value, decltype(stage)::value>(st, c2, rlo - 2 + t, xbuf + 2 * lane);
            __syncwarp();
        };
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        using I2 = std::integral_constant<int, 2>;
        // window fill (nsteps >= 5 always), then the branch-free steady state
        step(I0{}, I0{}, 0);
        step(I1{}, I0{}, 1);
        step(I2{}, I1{}, 2);
        step(I0{}, I1{}, 3);
        int t = 4;
        for (; t + 2 < nsteps; t += 3) {
            step(I1{}, I2{}, t);
            step(I2{}, I2{}, t + 1);
            step(I0{}, I2{}, t + 2);
        }
        if (t < nsteps) step(I1{}, I2{}, t);
        if (t + 1 < nsteps) step(I2{}, I2{}, t + 1);
    }
    __syncthreads();  // the marching warps' global stores are visible to the whole CTA from here on

    // ---- per-pixel phase: view weights (monorec_model.py:257-260), zeroing of invalid pixels (:251) and fusion
    //      cv = sum_f w_f (1 - 2 sad_f) / sum_f w_f, 0 where sum_f w_f == 0 (:262-269).  Each thread reads back the
    //      L2-hot single-frame values of its pixel once per frame. ------------------------------------------------------
    const float na4 = -0.25f * a.alpha;
    for (int p = tid; p < TH * kTileCols && MR_CV_SKIP != 2; p += kThreads) {
        const int r = p >> 6, bc = p & 63;
        const int u = u0 + bc, v = v0 + r;
        const bool own = (bc >= 2) && (bc < 2 + kOutCols) && (u < W) && (v < H);
        if (!own) continue;
        const size_t pix = (size_t)v * W + u;
        float* cv_out = a.cv + (size_t)b * D * plane + pix;
        for (int d0 = 0; d0 < D; d0 += kChunk) {   // one pass when D <= kChunk (every shipped config)
            float acc[kChunk];
#pragma unroll
            for (int j = 0; j < kChunk; ++j) acc[j] = 0.f;
            float wsum = 0.f;
            for (int f = 0; f < F; ++f) {
                float* sf = a.sfcv + (((size_t)f * a.B + b) * D) * plane + pix;
                if (vmask[f * TH * kTileCols + p] == 0) {
                    if (d0 == 0)
                        for (int d = 0; d < D; ++d) sf[(size_t)d * plane] = 0.f;
                    continue;
                }
                // sad = (1 - sv) / 2, so (sad - min sad)^2 = ((max sv - sv) / 2)^2
                float m = -2.0f, sum = 0.f;
                float vv[kChunk];
                if (D <= kChunk) {
#pragma unroll
                    for (int j = 0; j < kChunk; ++j) vv[j] = (j < D) ? __ldcg(sf + (size_t)j * plane) : -2.0f;
#pragma unroll
                    for (int j = 0; j < kChunk; ++j) m = fmaxf(m, vv[j]);
#pragma unroll
                    for (int j = 0; j < kChunk; ++j) {
                        const float df = m - vv[j];
                        if (j < D) sum += __expf(na4 * df * df);
                    }
                } else {
                    for (int d = 0; d < D; ++d) m = fmaxf(m, __ldcg(sf + (size_t)d * plane));
                    for (int d = 0; d < D; ++d) {
                        const float df = m - __ldcg(sf + (size_t)d * plane);
                        sum += __expf(na4 * df * df);
                    }
#pragma unroll
                    for (int j = 0; j < kChunk; ++j) vv[j] = (d0 + j < D) ? __ldcg(sf + (size_t)(d0 + j) * plane) : 0.f;
                }
                // weight = 1 - 1/(D-1) * (sum - 1): separate roundings as in the reference so that flat-cost pixels
                // (sum == D) give exactly 0 (monorec_model.py:258, :265-269)
                const float w = __fsub_rn(1.0f, __fmul_rn(a.inv_dm1, __fsub_rn(sum, 1.0f)));
                wsum += w;
#pragma unroll
                for (int j = 0; j < kChunk; ++j) acc[j] = fmaf(w, vv[j], acc[j]);
            }
            const float inv = (wsum == 0.f) ? 0.f : 1.0f / wsum;
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
                if (d0 + j < D) st_hint_f1(cv_out + (size_t)(d0 + j) * plane, (wsum == 0.f) ? 0.f : acc[j] * inv, pol_stream);
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// projection tables (fp64 on device, one thread per (b,f)); see include/monorec_b200.h
// ----------------------------------------------------------------------------------------------------------------
struct PtrPack {
    const float* p[MR_MAX_FRAMES];
};

__device__ bool invert4(const float* src, double* out) {
    double m[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            m[i][j] = (double)src[i * 4 + j];
            m[i][4 + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        double best = fabs(m[c][c]);
        for (int r = c + 1; r < 4; ++r)
            if (fabs(m[r][c]) > best) { best = fabs(m[r][c]); piv = r; }
        if (best == 0.0) return false;
        if (piv != c)
            for (int j = 0; j < 8; ++j) { double t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t; }
        double inv = 1.0 / m[c][c];
        for (int j = 0; j < 8; ++j) m[c][j] *= inv;
        for (int r = 0; r < 4; ++r)
            if (r != c) {
                double fct = m[r][c];
                for (int j = 0; j < 8; ++j) m[r][j] -= fct * m[c][j];
            }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[i * 4 + j] = m[i][4 + j];
    return true;
}

__global__ void projection_tables_kernel(const float* kf_pose, const float* kf_K, PtrPack poses, PtrPack intr,
                                         int B, int F, int H, int W, float* proj, float* depths, int D,
                                         float lo, float hi) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (depths != nullptr && idx < D) {
        // torch.linspace (fp32, symmetric fill) followed by 1/x  -- monorec_model.py:184
        float step = __fdiv_rn(hi - lo, (float)(D - 1));
        float x = (idx < D / 2) ? fmaf(step, (float)idx, lo) : fmaf(-step, (float)(D - 1 - idx), hi);
        depths[idx] = __frcp_rn(x);
    }
    if (idx >= B * F) return;
    int b = idx / F, f = idx % F;
    double kinv[16], pinv[16], T[16], P[12];
    bool ok = invert4(kf_K + b * 16, kinv);
    ok = invert4(poses.p[f] + b * 16, pinv) && ok;
    const float* kp = kf_pose + b * 16;
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += pinv[i * 4 + k] * (double)kp[k * 4 + j];
            T[i * 4 + j] = s;
        }
    const float* Kf = intr.p[f] + b * 16;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += (double)Kf[i * 4 + k] * T[k * 4 + j];
            P[i * 4 + j] = s;
        }
    double sc[3] = {(double)W / (double)(W - 1), (double)H / (double)(H - 1), 1.0};
    float* o = proj + (size_t)idx * 12;
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += P[i * 4 + k] * kinv[k * 4 + j];
            o[i * 4 + j] = ok ? (float)(s * sc[i]) : __int_as_float(0x7fc00000);
        }
        double t = P[i * 4 + 3] + (i == 2 ? 1e-7 : 0.0);
        o[i * 4 + 3] = ok ? (float)(t * sc[i]) : __int_as_float(0x7fc00000);
    }
}

// Source frames re-laid as (r,g,b,0) pixels so that a bilinear tap is one 16-byte load (see warp_row<true>).
__global__ void repack_frames_kernel(PtrPack frames, float4* __restrict__ packed, int B, int HW, int b0) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int b = b0 + blockIdx.y, f = blockIdx.z;
    const float* src = frames.p[f] + (size_t)b * 3 * HW + p;
    packed[((size_t)f * B + b) * HW + p] = make_float4(__ldg(src), __ldg(src + HW), __ldg(src + 2 * (size_t)HW), 0.f);
}

#ifndef MR_CV_TILE_ROWS
#define MR_CV_TILE_ROWS 16
#endif
int pick_tile_rows(int D, int F) {
    const int limit = 227 * 1024;
    for (int th = MR_CV_TILE_ROWS; th >= 2; th >>= 1)
        if (make_layout(D, th, F).total <= limit) return th;
    return 0;
}

}  // namespace

extern "C" int mr_projection_tables(const float* keyframe_pose, const float* keyframe_K, const float* const* poses,
                                    const float* const* intrinsics, int B, int F, int H, int W, float* proj,
                                    float* depths, int D, float inv_depth_lo, float inv_depth_hi, void* stream) {
    MR_REQUIRE(keyframe_pose && keyframe_K && poses && intrinsics && proj, "mr_projection_tables: null pointer");
    MR_REQUIRE(B >= 1 && F >= 1 && F <= MR_MAX_FRAMES, "mr_projection_tables: need B>=1, 1<=F<=%d (got B=%d F=%d)",
               MR_MAX_FRAMES, B, F);
    MR_REQUIRE(H >= 5 && W >= 5, "mr_projection_tables: image too small (%dx%d)", H, W);
    MR_REQUIRE(depths == nullptr || D >= 2, "mr_projection_tables: D must be >= 2 (got %d)", D);
    PtrPack pp{}, ip{};
    for (int f = 0; f < F; ++f) {
        MR_REQUIRE(poses[f] && intrinsics[f], "mr_projection_tables: null pose/intrinsics pointer for frame %d", f);
        pp.p[f] = poses[f];
        ip.p[f] = intrinsics[f];
    }
    int n = B * F > D ? B * F : D;
    projection_tables_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(
        keyframe_pose, keyframe_K, pp, ip, B, F, H, W, proj, depths, depths ? D : 0, inv_depth_lo, inv_depth_hi);
    MR_LAUNCH_CHECK("projection_tables_kernel");
    return MR_OK;
}

int mr::launch_cost_volume(const float* keyframe, const float* const* frames, const float* proj,
                           const float* depths, float* out_cv, float* out_sfcv, int B, int F, int D, int H, int W,
                           float alpha, const float* chan_w, int b_begin, int b_count, void* workspace,
                           long long workspace_bytes, cudaStream_t stream) {
    MR_REQUIRE(keyframe && frames && proj && depths && out_cv && out_sfcv, "mr_cost_volume_fwd: null pointer");
    MR_REQUIRE(b_begin >= 0 && b_count >= 1 && b_begin + b_count <= B, "mr_cost_volume_fwd: bad batch range");
    MR_REQUIRE(B >= 1 && B <= 65535, "mr_cost_volume_fwd: batch %d out of range", B);
    MR_REQUIRE(F >= 1 && F <= MR_MAX_FRAMES, "mr_cost_volume_fwd: 1 <= F <= %d required (got %d)", MR_MAX_FRAMES, F);
    MR_REQUIRE(D >= 2 && D <= 128, "mr_cost_volume_fwd: 2 <= D <= 128 required (got %d)", D);
    MR_REQUIRE(H >= 5 && W >= 5, "mr_cost_volume_fwd: image too small (%dx%d)", H, W);
    CvArgs a{};
    a.key = keyframe;
    for (int f = 0; f < F; ++f) {
        MR_REQUIRE(frames[f] != nullptr, "mr_cost_volume_fwd: null frame pointer %d", f);
        a.frames[f] = frames[f];
    }
    a.proj = proj; a.depths = depths; a.cv = out_cv; a.sfcv = out_sfcv;
    a.B = B; a.F = F; a.D = D; a.H = H; a.W = W; a.b0 = b_begin;
    a.TH = pick_tile_rows(D, F);
    MR_REQUIRE(a.TH > 0, "mr_cost_volume_fwd: no tile height fits shared memory for D=%d F=%d", D, F);
    a.alpha = alpha;
    a.inv_dm1 = (float)(1.0 / (double)(D - 1));
    const float def_w[3] = {5.f / 32.f, 16.f / 32.f, 11.f / 32.f};  // monorec_model.py:133
    const float* cw = chan_w ? chan_w : def_w;
    a.cw0 = cw[0] / 9.f; a.cw1 = cw[1] / 9.f; a.cw2 = cw[2] / 9.f;  // monorec_model.py:141 (weights / patch_size^2)
    const SmemLayout L = make_layout(D, a.TH, F);
    dim3 grid((W + kOutCols - 1) / kOutCols, (H + a.TH - 1) / a.TH, b_count);
    if (workspace != nullptr) {
        MR_REQUIRE(workspace_bytes >= mr_cost_volume_workspace_bytes(B, F, H, W),
                   "mr_cost_volume_fwd_ws: workspace too small (%lld < %lld bytes)", workspace_bytes,
                   mr_cost_volume_workspace_bytes(B, F, H, W));
        MR_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "mr_cost_volume_fwd_ws: workspace must be 16-byte aligned");
        a.packed = static_cast<const float4*>(workspace);
        PtrPack fp{};
        for (int f = 0; f < F; ++f) fp.p[f] = frames[f];
        dim3 rgrid((H * W + 255) / 256, b_count, F);
        repack_frames_kernel<<<rgrid, 256, 0, stream>>>(fp, static_cast<float4*>(workspace), B, H * W, b_begin);
        MR_LAUNCH_CHECK("repack_frames_kernel");
        MR_CUDA(cudaFuncSetAttribute(cost_volume_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        cost_volume_kernel<true><<<grid, kThreads, L.total, stream>>>(a);
    } else {
        MR_CUDA(cudaFuncSetAttribute(cost_volume_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
        cost_volume_kernel<false><<<grid, kThreads, L.total, stream>>>(a);
    }
    MR_LAUNCH_CHECK("cost_volume_kernel");
    return MR_OK;
}

extern "C" long long mr_cost_volume_workspace_bytes(int B, int F, int H, int W) {
    if (B < 1 || F < 1 || H < 1 || W < 1) return 0;
    return (long long)F * B * H * W * 16;
}

extern "C" int mr_cost_volume_fwd_ws(const float* keyframe, const float* const* frames, const float* proj,
                                     const float* depths, float* out_cv, float* out_sfcv, int B, int F, int D, int H, int W,
                                     float alpha, const float* chan_w, void* workspace, long long workspace_bytes,
                                     void* stream) {
    MR_REQUIRE(workspace != nullptr, "mr_cost_volume_fwd_ws: null workspace");
    return mr::launch_cost_volume(keyframe, frames, proj, depths, out_cv, out_sfcv, B, F, D, H, W, alpha, chan_w, 0, B,
                                  workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int mr_cost_volume_fwd(const float* keyframe, const float* const* frames, const float* proj,
                                  const float* depths, float* out_cv, float* out_sfcv, int B, int F, int D, int H,
                                  int W, float alpha, const float* chan_w, void* stream) {
    return mr::launch_cost_volume(keyframe, frames, proj, depths, out_cv, out_sfcv, B, F, D, H, W, alpha, chan_w, 0,
                                  B, nullptr, 0, (cudaStream_t)stream);
}
