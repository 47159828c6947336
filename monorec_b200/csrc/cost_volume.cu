// Fused plane-sweep cost volume for sm_100a (B200), second generation: TMA-staged source windows.
//
// Replaces CostVolumeModule.forward (reference: model/monorec/monorec_model.py:150-280) together with
// Backprojection / point_projection (model/layers.py:43-71), F.grid_sample x2, SSIM (layers.py:119-137), the
// conv3d patch cost (:246-248) and the view weighting / fusion (:257-269).  Closed form: SURVEY.md Appendix C.
//
// Work decomposition:
//   CTA   = one keyframe tile of 60 x TH output pixels (64 x (TH+4) with the 2-px stencil halo), all D planes, all F
//           source frames.  grid = (ceil(W/60), ceil(H/TH), B).
//   plan  = per tile and source frame the D planes are cut into groups of consecutive planes whose source footprints
//           (projective image of the tile rectangle: extremes at its 4 corners) share one window of kPitch x kWinRows
//           pixels.  A window is the [3][rows][kPitch] fp32 copy of that frame region, brought into shared memory by
//           TMA boxes {kPitch, 8, 1} straight from the NCHW frame (cp.async.bulk.tensor, mbarrier complete_tx, kBuf
//           buffers in flight); out-of-image box parts are zero-filled by the TMA unit, which is exactly
//           F.grid_sample(padding_mode="zeros") once integer tap coordinates are clamped to the 2-px zero ring.
//   unit  = (frame, plane).  Warps claim units from a shared counter (units of one window are consecutive), wait for
//           the window's mbarrier, and march down the tile rows:
//             stage 1 (lane = columns l, l+32): homography, floor by magic-number rounding, 12 bilinear taps per
//                      sample as conflict-free LDS from the window (immediate offsets), warped row -> smem row buffer;
//             stage 2 (lane = columns 2l, 2l+1): 3x3 box sums of X, X^2, XY (horizontal in registers, vertical rolling),
//                      SSIM in 81x-scaled form, channel weights, second 3x3 box -> 1 - 2 sad streamed to HBM.
//           Stage 1 of row t+1 and stage 2 of row t are issued together (double-buffered row buffer, one __syncwarp
//           per row).  The last warp to finish a window's units re-arms the buffer with the window after next.
//           Units whose footprint does not fit a window (strong zoom), groups of fewer than kMinGroup planes and
//           launches whose frames TMA cannot address (W % 4 != 0, unaligned base) gather from global memory instead.
//   CTA phase 2 (thread = pixel): view weight from max_d / sum_d exp(..), zeroing of invalid pixels, fusion over
//           frames from the L2-hot single-frame volumes.
// Keyframe-only terms (9 mu_y, 81 (sigma_y + C2)) are hoisted into a smem table per tile; pixels whose reprojection leaves
// the source for any plane (valid_f = 0) are found by a projection pre-pass over the two extreme planes (the samples of
// one pixel lie on a line, monotone in depth, so the extremes decide) and whole row ranges / frames of a tile are skipped.
#include "mr_common.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <cstdint>
#include <type_traits>

namespace {

constexpr int kTileCols = 64;   // buffer columns per tile row (output columns + 2-px halo each side)
constexpr int kOutCols = 60;    // output columns per tile
constexpr int kRowStride = 68;  // floats per smem image row: column b lives at index b+1 (so [2l-1, 2l+2] is 8B aligned)
#ifndef MR_CV_WARPS
#define MR_CV_WARPS 16
#endif
#ifndef MR_CV_MINBLOCKS
#define MR_CV_MINBLOCKS 1       // resident CTAs per SM the register allocator must leave room for
#endif
#ifndef MR_CV_NBUF
#define MR_CV_NBUF 2            // source windows in flight per CTA
#endif
#ifndef MR_CV_WIN_ROWS
#define MR_CV_WIN_ROWS 40       // rows per window (multiple of 8)
#endif
#ifndef MR_CV_MIN_GROUP
#define MR_CV_MIN_GROUP 3       // plane groups smaller than this gather from global memory
#endif
#ifndef MR_CV_TILE_ROWS
#define MR_CV_TILE_ROWS 16
#endif
#ifndef MR_CV_SKIP
#define MR_CV_SKIP 0            // timing experiments only: 1 = no march, 2 = no per-pixel phase, 3 = march without stage 2, 4 = without stage 1
#endif
constexpr int kWarps = MR_CV_WARPS;
constexpr int kThreads = kWarps * 32;
constexpr int kBuf = MR_CV_NBUF;
constexpr int kPitch = 128;                       // pixels per window row (512 bytes)
constexpr int kWinRows = MR_CV_WIN_ROWS;
constexpr int kChanStride = kWinRows * kPitch;    // floats between the channel planes of a window
constexpr int kWinFloats = 3 * kChanStride;
constexpr int kBoxRows = 8;                       // rows per TMA box
constexpr int kMinGroup = MR_CV_MIN_GROUP;
static_assert(kWinRows % kBoxRows == 0, "window rows must be a multiple of the TMA box height");
static_assert(MR_MAX_FRAMES <= kWarps, "the plan gives every source frame its own warp");

constexpr float kC1 = 0.01f * 0.01f;  // layers.py:116
constexpr float kC2 = 0.03f * 0.03f;  // layers.py:117
constexpr float kMagic = 12582912.0f;      // 1.5 * 2^23: adding it rounds to the nearest integer in the low mantissa bits
constexpr int kMagicBits = 0x4B400000;
constexpr int kChunk = 32;                 // planes the per-pixel phase keeps in registers at once

struct CvArgs {
    const float* key;                    // [B,3,H,W]
    const float* frames[MR_MAX_FRAMES];  // each [B,3,H,W]
    const float* proj;                   // [B,F,12]
    const float* depths;                 // [D]
    float* cv;                           // [B,D,H,W]
    float* sfcv;                         // [F,B,D,H,W]
    void* sf_nhwc;                       // optional [F,B,H,W,D] copy of sfcv for the conv engine (D <= 32, D % 8 == 0) or nullptr
    int sf_nhwc_half;                    // 1: that copy is IEEE half, 0: fp32
    int B, F, D, H, W, TH, b0;
    int use_tma;                         // 0: every unit gathers from global memory
    float alpha, inv_dm1;
    float cw0, cw1, cw2;                 // channel weights / 9
};

struct CvMaps {
    CUtensorMap m[MR_MAX_FRAMES];        // frame f as a (W, H, 3B) fp32 tensor, box {kPitch, kBoxRows, 1}, zero fill
};

struct GroupInfo {                       // one window (a run of consecutive planes of one frame)
    short wx0, wy0;                      // image coordinates of the window origin (may be negative: zero ring)
    short nrows;                         // rows actually loaded (multiple of kBoxRows)
    short f;
    short count;                         // planes in the group
    short seq;                           // running number of the window inside the tile (buffer = seq % kBuf)
    short pad0, pad1;
};

struct SmemLayout {
    int win, ytile, cst, xbuf, pjs, zs, vmask, rowrng, bbox, gid, uflag, ginfo, seq2g, nwin, bars, ctr, total;  // byte offsets
};

__host__ __device__ inline SmemLayout make_layout(int D, int TH, int F, int use_tma) {
    SmemLayout L;
    int off = 0;
    auto take = [&](int bytes, int align) { off = (off + align - 1) / align * align; int o = off; off += bytes; return o; };
    L.win = take(use_tma ? kBuf * kWinFloats * 4 : 0, 128);
    L.ytile = take(3 * (TH + 4) * kRowStride * 4, 16);
    L.cst = take(3 * (TH + 2) * kTileCols * 8, 16);
    L.xbuf = take(kWarps * 2 * 3 * kRowStride * 4, 16);
    L.pjs = take(F * 12 * 4, 16);
    L.zs = take(((D + 3) / 4) * 16, 16);
    L.vmask = take(F * TH * kTileCols, 16);
    L.rowrng = take(F * 2 * 4, 16);
    L.bbox = take(F * D * 8, 8);
    L.gid = take(F * D * 2, 4);
    L.uflag = take(F * D, 4);
    L.ginfo = take(F * D * (int)sizeof(GroupInfo), 16);
    L.seq2g = take(F * D * 2, 4);
    L.nwin = take(MR_MAX_FRAMES * 4, 4);
    L.bars = take(kBuf * 8, 8);
    L.ctr = take((2 + 2 * kBuf) * 4, 4);
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

// ---- mbarrier / TMA wrappers ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a window that never arrives (a bug, not a load condition) traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// ---- explicit shared-memory accesses (32-bit shared addresses, immediate offsets) ---------------------------------------
// volatile: never hoisted over the mbarrier wait / __syncwarp that orders them; ptxas still schedules them freely
template <int OFF>
__device__ __forceinline__ float lds32(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ float2 lds64(uint32_t a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2+%3];" : "=f"(v.x), "=f"(v.y) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(OFF));
    return v;
}
template <int OFF>
__device__ __forceinline__ void sts32(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(OFF), "f"(v) : "memory");
}

constexpr int kXbBytes = 3 * kRowStride * 4;      // one warped-row buffer (3 channels)
constexpr int kYRowBytes = 3 * kRowStride * 4;    // keyframe tile: [row][channel][kRowStride]
constexpr int kCRowBytes = 3 * (kTileCols / 2) * 16;  // hoisted table: [e-row][channel][32 column pairs] float4

// ---- stage 1 of the march: homography of one tile row (pair = columns lane, lane + 32) and its 24 bilinear taps -------
struct Stage1Ctx {
    float2 pzx, pzy, pzz;                // per-lane column part of the projection (already times the plane depth)
    float rax, rbx, ray, rby, raz, rbz;  // per-row part: c = pz(u) + ra * v + rb
    uint32_t kaddr;                      // window modes: shared address of the window minus the unit's constant (see march)
    const float* img;                    // global mode: source frame of this batch element, [3][H][W]
    int W, H, planei;
    float sx_lo, sx_hi, sy_lo, sy_hi;    // == grid clamp(-2, 2) + 0.5, monorec_model.py:208
};

__device__ __forceinline__ void setup_stage1(Stage1Ctx& c, const float* m, float z, float2 fu2) {
    // projection c = M [u v 1]^T z + p split into a per-lane column part and a per-row part
    c.pzx = mul2(mul2(bc2(m[0]), fu2), bc2(z));
    c.pzy = mul2(mul2(bc2(m[4]), fu2), bc2(z));
    c.pzz = mul2(mul2(bc2(m[8]), fu2), bc2(z));
    c.rax = m[1] * z; c.rbx = fmaf(m[2], z, m[3]);
    c.ray = m[5] * z; c.rby = fmaf(m[6], z, m[7]);
    c.raz = m[9] * z; c.rbz = fmaf(m[10], z, m[11]);
}

struct Taps {                            // the 24 taps and 4 weight pairs of one row step (two samples per lane)
    float a[3][4], b[3][4];              // [channel][nw, ne, sw, se] of the sample at column lane / lane + 32
    float2 w00, w01, w10, w11;
};

// MODE 0: taps from the window, every sample of the unit strictly inside the image (decided by the plan): no clamps
// MODE 1: taps from the window, coordinates clamped to the 2-px zero ring (== zero padding of F.grid_sample)
// MODE 2: taps from global memory with per-tap zero padding
template <int MODE>
__device__ __forceinline__ void warp_row_issue(const Stage1Ctx& c, const float fv, Taps& t) {
    const float rcx = fmaf(c.rax, fv, c.rbx), rcy = fmaf(c.ray, fv, c.rby), rcz = fmaf(c.raz, fv, c.rbz);
    const float2 cx = add2(c.pzx, bc2(rcx)), cy = add2(c.pzy, bc2(rcy)), cz = add2(c.pzz, bc2(rcz));
    const float2 inv = make_float2(fast_rcp(cz.x), fast_rcp(cz.y));
    float2 ux = mul2(cx, inv), uy = mul2(cy, inv);   // sample position + 0.5
    if (MODE <= 1) {
        if (MODE == 1) {   // == .clamp(-2, 2) of the normalised grid (monorec_model.py:208); also maps NaN to the low bound
            ux.x = fminf(fmaxf(ux.x, c.sx_lo), c.sx_hi); ux.y = fminf(fmaxf(ux.y, c.sx_lo), c.sx_hi);
            uy.x = fminf(fmaxf(uy.x, c.sy_lo), c.sy_hi); uy.y = fminf(fmaxf(uy.y, c.sy_lo), c.sy_hi);
        }
        // floor by magic-number rounding: rn(s - 0.5) differs from floor(s) only for integral s, where the interpolated value
        // is the same (weight 1 on the tap both conventions share)
        const float2 tx = add2(ux, bc2(kMagic - 1.0f)), ty = add2(uy, bc2(kMagic - 1.0f));
        const float2 x0f = add2(tx, bc2(-kMagic)), y0f = add2(ty, bc2(-kMagic));
        const float2 wx1 = add2(add2(ux, bc2(-0.5f)), neg2(x0f)), wy1 = add2(add2(uy, bc2(-0.5f)), neg2(y0f));
        const float2 wx0 = add2(bc2(1.0f), neg2(wx1)), wy0 = add2(bc2(1.0f), neg2(wy1));
        t.w00 = mul2(wx0, wy0); t.w01 = mul2(wx1, wy0); t.w10 = mul2(wx0, wy1); t.w11 = mul2(wx1, wy1);
        uint32_t aa, ab;
        if (MODE == 0) {
            // address = window + 4 ((y0 - wy0) kPitch + x0 - wx0) with y0 = bits(ty) - kMagicBits: every constant is in kaddr
            aa = ((((uint32_t)__float_as_int(ty.x) << 7) + (uint32_t)__float_as_int(tx.x)) << 2) + c.kaddr;
            ab = ((((uint32_t)__float_as_int(ty.y) << 7) + (uint32_t)__float_as_int(tx.y)) << 2) + c.kaddr;
        } else {
            // integer tap origin clamped to [-2, W] x [-2, H]: taps of the ring [-2,-1] / [W, W+1] are zero-filled by TMA
            const int xa = min(max(__float_as_int(tx.x) - kMagicBits, -2), c.W);
            const int xb = min(max(__float_as_int(tx.y) - kMagicBits, -2), c.W);
            const int ya = min(max(__float_as_int(ty.x) - kMagicBits, -2), c.H);
            const int yb = min(max(__float_as_int(ty.y) - kMagicBits, -2), c.H);
            aa = ((((uint32_t)ya << 7) + (uint32_t)xa) << 2) + c.kaddr;
            ab = ((((uint32_t)yb << 7) + (uint32_t)xb) << 2) + c.kaddr;
        }
        static_assert(kPitch == 128, "the tap address uses a shift by 7");
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            // (ch is a compile-time constant after unrolling: the offsets are immediates)
            if (ch == 0) {
                t.a[0][0] = lds32<0>(aa); t.a[0][1] = lds32<4>(aa); t.a[0][2] = lds32<kPitch * 4>(aa); t.a[0][3] = lds32<kPitch * 4 + 4>(aa);
                t.b[0][0] = lds32<0>(ab); t.b[0][1] = lds32<4>(ab); t.b[0][2] = lds32<kPitch * 4>(ab); t.b[0][3] = lds32<kPitch * 4 + 4>(ab);
            } else if (ch == 1) {
                constexpr int o = kChanStride * 4;
                t.a[1][0] = lds32<o>(aa); t.a[1][1] = lds32<o + 4>(aa); t.a[1][2] = lds32<o + kPitch * 4>(aa); t.a[1][3] = lds32<o + kPitch * 4 + 4>(aa);
                t.b[1][0] = lds32<o>(ab); t.b[1][1] = lds32<o + 4>(ab); t.b[1][2] = lds32<o + kPitch * 4>(ab); t.b[1][3] = lds32<o + kPitch * 4 + 4>(ab);
            } else {
                constexpr int o = 2 * kChanStride * 4;
                t.a[2][0] = lds32<o>(aa); t.a[2][1] = lds32<o + 4>(aa); t.a[2][2] = lds32<o + kPitch * 4>(aa); t.a[2][3] = lds32<o + kPitch * 4 + 4>(aa);
                t.b[2][0] = lds32<o>(ab); t.b[2][1] = lds32<o + 4>(ab); t.b[2][2] = lds32<o + kPitch * 4>(ab); t.b[2][3] = lds32<o + kPitch * 4 + 4>(ab);
            }
        }
    } else {
        const int W = c.W, H = c.H;
        const float2 tx = add2(ux, bc2(kMagic - 1.0f)), ty = add2(uy, bc2(kMagic - 1.0f));
        const int x0a = __float_as_int(tx.x) - kMagicBits, x0b = __float_as_int(tx.y) - kMagicBits;
        const int y0a = __float_as_int(ty.x) - kMagicBits, y0b = __float_as_int(ty.y) - kMagicBits;
        const bool inb = ((unsigned)x0a <= (unsigned)(W - 2)) && ((unsigned)x0b <= (unsigned)(W - 2)) &&
                         ((unsigned)y0a <= (unsigned)(H - 2)) && ((unsigned)y0b <= (unsigned)(H - 2));
        int oa, ob, dxa, dxb, dya, dyb;
        if (__all_sync(0xffffffffu, inb)) {
            const float2 x0f = add2(tx, bc2(-kMagic)), y0f = add2(ty, bc2(-kMagic));
            const float2 wx1 = add2(add2(ux, bc2(-0.5f)), neg2(x0f)), wy1 = add2(add2(uy, bc2(-0.5f)), neg2(y0f));
            const float2 wx0 = add2(bc2(1.0f), neg2(wx1)), wy0 = add2(bc2(1.0f), neg2(wy1));
            t.w00 = mul2(wx0, wy0); t.w01 = mul2(wx1, wy0); t.w10 = mul2(wx0, wy1); t.w11 = mul2(wx1, wy1);
            oa = y0a * W + x0a; ob = y0b * W + x0b;
            dxa = dxb = 1; dya = dyb = W;
        } else {
            // border path: per-tap zero padding exactly like F.grid_sample(padding_mode="zeros"): clamp the tap address, zero
            // the weight of every tap that falls outside the image
            float wx0s[2], wx1s[2], wy0s[2], wy1s[2];
            int os[2], dxs[2], dys[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                float sxk = (k ? ux.y : ux.x), syk = (k ? uy.y : uy.x);
                sxk = fminf(fmaxf(sxk, c.sx_lo), c.sx_hi) - 0.5f;
                syk = fminf(fmaxf(syk, c.sy_lo), c.sy_hi) - 0.5f;
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
            t.w00 = mul2(wx0, wy0); t.w01 = mul2(wx1, wy0); t.w10 = mul2(wx0, wy1); t.w11 = mul2(wx1, wy1);
            oa = os[0]; ob = os[1]; dxa = dxs[0]; dxb = dxs[1]; dya = dys[0]; dyb = dys[1];
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            const float* pa0 = c.img + (oa + ch * c.planei);
            const float* pb0 = c.img + (ob + ch * c.planei);
            t.a[ch][0] = __ldg(pa0); t.a[ch][1] = __ldg(pa0 + dxa); t.a[ch][2] = __ldg(pa0 + dya); t.a[ch][3] = __ldg(pa0 + dya + dxa);
            t.b[ch][0] = __ldg(pb0); t.b[ch][1] = __ldg(pb0 + dxb); t.b[ch][2] = __ldg(pb0 + dyb); t.b[ch][3] = __ldg(pb0 + dyb + dxb);
        }
    }
}

// interpolation (same order as grid_sample: nw, ne, sw, se; + 0.5: monorec_model.py:231) and the warped row -> row buffer
// xw = shared address of this lane's first column in the row buffer to fill
__device__ __forceinline__ void warp_row_finish(const Taps& t, const uint32_t xw) {
    float va[3], vb[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        float a = fmaf(t.a[ch][0], t.w00.x, 0.5f), b = fmaf(t.b[ch][0], t.w00.y, 0.5f);
        a = fmaf(t.a[ch][1], t.w01.x, a); b = fmaf(t.b[ch][1], t.w01.y, b);
        a = fmaf(t.a[ch][2], t.w10.x, a); b = fmaf(t.b[ch][2], t.w10.y, b);
        a = fmaf(t.a[ch][3], t.w11.x, a); b = fmaf(t.b[ch][3], t.w11.y, b);
        va[ch] = a; vb[ch] = b;
    }
    sts32<0>(xw, va[0]);                  sts32<128>(xw, vb[0]);
    sts32<kRowStride * 4>(xw, va[1]);     sts32<kRowStride * 4 + 128>(xw, vb[1]);
    sts32<2 * kRowStride * 4>(xw, va[2]); sts32<2 * kRowStride * 4 + 128>(xw, vb[2]);
}

// ---- stage 2 of the march: SSIM + patch cost of one row; lane owns buffer columns 2l, 2l+1 (a pair) ---------------------
struct Stage2Ctx {
    float2 cw0, cw1, cw2;    // channel weights / 9
    int pairflag;            // 1: both columns of this lane are output pixels and W is even (one 8-byte store)
    bool st0, st1;
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

// One row step: horizontal sums of the new warped row, SSIM error of the row above it (windows complete from the third
// row of a unit on; earlier rows produce finite throw-away values from the zeroed state: 9 sxx >= s^2 for partial
// windows too), patch cost of the row above that.  `store` is warp-uniform (false for the first four rows of a unit).
//   xr: shared address of the warped row (this lane's columns 2l-1..2l+2); yr: keyframe tile row (+0.5), same columns;
//   cr: hoisted (Y, Y, Sg, Sg) table row of the SSIM row; out: single-frame volume, output row of this step.
template <int P>
__device__ __forceinline__ void ssim_row(Stage2State& st, const Stage2Ctx& c, const uint32_t xr, const uint32_t yr,
                                         const uint32_t cr, float* out, const bool store) {
    constexpr int P1 = (P + 1) % 3, P2 = (P + 2) % 3;
    float2 xl[3], xrr[3], yl[3], yrr[3];
    float4 k4[3];
    xl[0] = lds64<0>(xr);                  xrr[0] = lds64<8>(xr);
    xl[1] = lds64<kRowStride * 4>(xr);     xrr[1] = lds64<kRowStride * 4 + 8>(xr);
    xl[2] = lds64<2 * kRowStride * 4>(xr); xrr[2] = lds64<2 * kRowStride * 4 + 8>(xr);
    yl[0] = lds64<0>(yr);                  yrr[0] = lds64<8>(yr);
    yl[1] = lds64<kRowStride * 4>(yr);     yrr[1] = lds64<kRowStride * 4 + 8>(yr);
    yl[2] = lds64<2 * kRowStride * 4>(yr); yrr[2] = lds64<2 * kRowStride * 4 + 8>(yr);
    k4[0] = lds128<0>(cr); k4[1] = lds128<512>(cr); k4[2] = lds128<1024>(cr);
    float2 h1[3], hx[3], hy[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const float2 xxl = mul2(xl[ch], xl[ch]), xxr = mul2(xrr[ch], xrr[ch]), xyl = mul2(xl[ch], yl[ch]), xyr = mul2(xrr[ch], yrr[ch]);
        const float m1 = xl[ch].y + xrr[ch].x, mx = xxl.y + xxr.x, my = xyl.y + xyr.x;
        h1[ch] = make_float2(xl[ch].x + m1, m1 + xrr[ch].y);
        hx[ch] = make_float2(xxl.x + mx, mx + xxr.y);
        hy[ch] = make_float2(xyl.x + my, my + xyr.y);
    }
    float2 e[3];
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        // SSIM with every factor scaled by 81 (layers.py:123-137 through 3x3 box sums s = sum x, sxx, sxy; Y = 9 mu_y,
        // Sg = 81 (sigma_y + C2) hoisted):  n/d = (2 s Y + 81 C1)(2 (9 sxy - s Y) + 81 C2) / ((s^2 + Y^2 + 81 C1)(9 sxx - s^2 + Sg))
        const float2 Y = make_float2(k4[ch].x, k4[ch].y), Sg = make_float2(k4[ch].z, k4[ch].w);
        const float2 s = add2(add2(st.hs1[P1][ch], st.hs1[P2][ch]), h1[ch]);
        const float2 sxx = add2(add2(st.hsx[P1][ch], st.hsx[P2][ch]), hx[ch]);
        const float2 sxy = add2(add2(st.hsy[P1][ch], st.hsy[P2][ch]), hy[ch]);
        const float2 p = mul2(s, Y), q = mul2(s, s);
        const float2 n1h = add2(neg2(p), bc2(-40.5f * kC1));                 // -(N1 / 2)
        const float2 n2 = fma2(bc2(2.0f), fma2(bc2(9.0f), sxy, neg2(p)), bc2(81.0f * kC2));
        const float2 d1 = add2(q, fma2(Y, Y, bc2(81.0f * kC1)));
        const float2 d2 = add2(fma2(bc2(9.0f), sxx, Sg), neg2(q));
        const float2 num = mul2(n1h, n2), den = mul2(d1, d2);
        // clamp((1 - n/d) / 2, 0, 1)   (layers.py:137)
        e[ch] = make_float2(__saturatef(fmaf(num.x, fast_rcp(den.x), 0.5f)), __saturatef(fmaf(num.y, fast_rcp(den.y), 0.5f)));
    }
    const float2 E = fma2(c.cw2, e[2], fma2(c.cw1, e[1], mul2(c.cw0, e[0])));
    const float eL = __shfl_up_sync(0xffffffffu, E.y, 1);
    const float eR = __shfl_down_sync(0xffffffffu, E.x, 1);
    const float mid = E.x + E.y;
    const float2 hEc = make_float2(eL + mid, mid + eR);
    // single-frame volume 1 - 2 sad (monorec_model.py:251) straight to HBM; the validity mask is applied by the
    // per-pixel phase (which zeroes invalid pixels) once all planes are known
    const float2 sad = add2(add2(st.hE[P1], st.hE[P2]), hEc);
    const float2 sv = fma2(bc2(-2.0f), sad, bc2(1.0f));
    if (store) {
        if (c.pairflag) {
            st_hint_f2(out, sv, c.pol_keep);
        } else {
            if (c.st0) st_hint_f1(out, sv.x, c.pol_keep);
            if (c.st1) st_hint_f1(out + 1, sv.y, c.pol_keep);
        }
    }
    st.hE[P] = hEc;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) { st.hs1[P][ch] = h1[ch]; st.hsx[P][ch] = hx[ch]; st.hsy[P][ch] = hy[ch]; }
}

#ifndef MR_CV_ORDER
#define MR_CV_ORDER 0     // 0: stage 1 of row t+1 completes, then stage 2 of row t (measured faster: 1.07 vs 1.12 ms); 1: taps stay in flight across stage 2
#endif

// The march of one unit over tile rows rlo-2 .. rhi+2 (nsteps = rhi - rlo + 5 >= 5 rows).  Stage 1 of the next row is
// issued with stage 2 of the current one; the three-step loop body is entered at the slot that makes the last step end
// a triple (the rolling state is symmetric under rotation of its slots).
//   xb: shared address of this warp's two row buffers; yr / cr: keyframe row rlo-2 / table row rlo-2 (lane columns);
//   out: single-frame volume at output row rlo - 4 (advanced every step, stored from the fifth step on); wstride = W
template <int MODE>
__device__ __forceinline__ void march_unit(const Stage1Ctx& c1, const Stage2Ctx& c2, const uint32_t xb, const int lane,
                                           const float fv0, const int nsteps, uint32_t yr, uint32_t cr, float* out,
                                           const int wstride) {
    Stage2State st;
    st.clear();
    float fv = fv0;
    uint32_t off = 0;                      // byte offset of the row buffer stage 2 reads next
    const uint32_t xw = xb + 4 * (lane + 1), xr = xb + 8 * lane;
    {
        Taps t;
        warp_row_issue<MODE>(c1, fv, t);
        warp_row_finish(t, xw);
    }
    __syncwarp();
    const int n = nsteps - 1;              // steps that also run stage 1 of the following row
    int t = -((3 - n % 3) % 3);
    int done = 0;                          // rows stage 2 has consumed
    auto both = [&](auto tag) {
        fv += 1.0f;
        Taps tp;
        if (MR_CV_SKIP != 4) warp_row_issue<MODE>(c1, fv, tp);
        if (MR_CV_ORDER == 0 && MR_CV_SKIP != 4) warp_row_finish(tp, xw + (kXbBytes - off));
        if (MR_CV_SKIP != 3) ssim_row<decltype(tag)::value>(st, c2, xr + off, yr, cr, out, done >= 4);
        if (MR_CV_ORDER != 0 && MR_CV_SKIP != 4) warp_row_finish(tp, xw + (kXbBytes - off));
        __syncwarp();
        off = kXbBytes - off;
        yr += kYRowBytes;
        cr += kCRowBytes;
        out += wstride;
        ++done;
    };
    using I0 = std::integral_constant<int, 0>;
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    for (; t < n; t += 3) {
        if (t >= 0) both(I0{});
        if (t + 1 >= 0) both(I1{});
        both(I2{});
    }
    if (MR_CV_SKIP != 3) ssim_row<0>(st, c2, xr + off, yr, cr, out, true);
    __syncwarp();
}

// ---- per-pixel phase: view weights (monorec_model.py:257-260), zeroing of invalid pixels (:251) and fusion
//      cv = sum_f w_f (1 - 2 sad_f) / sum_f w_f, 0 where sum_f w_f == 0 (:262-269).  T lanes share a pixel, each with a chunk of
//      kChunk planes in registers (T = 1 for D <= 32, 2 for D <= 64, 4 for D <= 128): every (L2-hot) single-frame value is
//      read back exactly once; max / sum over the planes are combined across the T lanes by shuffles.
struct PixelPhase {
    float* cv;
    float* sfcv;
    void* sf_nhwc;
    int sf_nhwc_half;
    const unsigned char* vmask;
    int B, F, D, H, W, TH, b, u0, v0;
    float inv_dm1, kq;
    uint64_t pol_stream;
};

template <int T>
__device__ __forceinline__ void pixel_phase(const PixelPhase& c, const int tid) {
    constexpr int kSlots = 32 / T;                       // pixels per warp iteration
    const int lane = tid & 31, warp = tid >> 5;
    const int sub = lane % kSlots, chunk = lane / kSlots;
    const int D = c.D, F = c.F, TH = c.TH;
    const int d_lo = chunk * kChunk;                     // this lane's planes [d_lo, d_lo + kChunk) of D
    const size_t plane = (size_t)c.H * c.W;
    const size_t pstride = plane * sizeof(float);        // bytes between the planes of a pixel
    const size_t fstride = (size_t)c.B * D * pstride;    // bytes between the frames
    const int vstride = TH * kTileCols;
    const int per_iter = kWarps * kSlots;
    for (int p0 = 0; p0 < TH * kTileCols; p0 += per_iter) {
        const int p = p0 + warp * kSlots + sub;
        const int r = p >> 6, bc = p & 63;
        const int u = c.u0 + bc, v = c.v0 + r;
        const bool own = (p < TH * kTileCols) && (bc >= 2) && (bc < 2 + kOutCols) && (u < c.W) && (v < c.H);
        if (T == 1 && !own) continue;                    // (with T > 1 every lane stays for the shuffles)
        const size_t pix = own ? (size_t)v * c.W + u : 0;
        // addresses advance by pointer increments (one 64-bit add per access; an index expression costs a wide multiply each)
        char* cv_out = reinterpret_cast<char*>(c.cv + ((size_t)c.b * D + d_lo) * plane + pix);
        char* sf = reinterpret_cast<char*>(c.sfcv + ((size_t)c.b * D + d_lo) * plane + pix);       // frame f: + f * fstride
        float acc[kChunk], vv[kChunk];
#pragma unroll
        for (int j = 0; j < kChunk; ++j) acc[j] = 0.f;
        float wsum = 0.f;
        for (int f = 0; f < F; ++f, sf += fstride) {
            const bool valid = own && (c.vmask[f * vstride + p] != 0);
            char* nh = nullptr;                          // this pixel's D channels of frame f in the NHWC copy (T == 1 only)
            if (T == 1 && c.sf_nhwc != nullptr)
                nh = static_cast<char*>(c.sf_nhwc) + (((size_t)f * c.B + c.b) * plane + pix) * D * (c.sf_nhwc_half ? 2 : 4);
            char* q = sf;
            if (own && !valid) {                         // invalid pixel of frame f: the whole plane stack is 0
#pragma unroll 4
                for (int j = 0; j < kChunk; ++j, q += pstride)
                    if (d_lo + j < D) *reinterpret_cast<float*>(q) = 0.f;
                if (nh != nullptr)
                    for (int o = 0; o < D * (c.sf_nhwc_half ? 2 : 4); o += 16) *reinterpret_cast<uint4*>(nh + o) = make_uint4(0, 0, 0, 0);
            }
            if (T == 1 && !valid) continue;
            if (valid) {
                if (D == T * kChunk) {                   // 32 / 64 / 128 planes: no per-plane predicates
#pragma unroll
                    for (int j = 0; j < kChunk; ++j, q += pstride) vv[j] = __ldcg(reinterpret_cast<const float*>(q));
                } else {
#pragma unroll
                    for (int j = 0; j < kChunk; ++j, q += pstride) vv[j] = (d_lo + j < D) ? __ldcg(reinterpret_cast<const float*>(q)) : -2.0f;
                }
            } else {
#pragma unroll
                for (int j = 0; j < kChunk; ++j) vv[j] = -2.0f;
            }
            float m4[4] = {-2.0f, -2.0f, -2.0f, -2.0f};
#pragma unroll
            for (int j = 0; j < kChunk; ++j) m4[j & 3] = fmaxf(m4[j & 3], vv[j]);
            float m = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
#pragma unroll
            for (int s = kSlots; s < 32; s <<= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, s));
            const float km = c.kq * m;
            float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < kChunk; ++j) {
                const float t = fmaf(-c.kq, vv[j], km);
                float e;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-t * t));
                if (D == T * kChunk || d_lo + j < D) s4[j & 3] += e;
            }
            float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
#pragma unroll
            for (int s = kSlots; s < 32; s <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, s);
            // weight = 1 - 1/(D-1) * (sum - 1): separate roundings as in the reference so that flat-cost pixels
            // (sum == D) give exactly 0 (monorec_model.py:258, :265-269)
            const float w = valid ? __fsub_rn(1.0f, __fmul_rn(c.inv_dm1, __fsub_rn(sum, 1.0f))) : 0.f;
            wsum += w;
#pragma unroll
            for (int j = 0; j < kChunk; ++j) acc[j] = fmaf(w, vv[j], acc[j]);
            if (nh != nullptr) {                         // the MaskModule's input layout, written while the values are in registers
                if (c.sf_nhwc_half) {
#pragma unroll
                    for (int j = 0; j < kChunk; j += 8) {
                        if (j >= D) break;
                        uint4 pk;
                        __half2* h = reinterpret_cast<__half2*>(&pk);
#pragma unroll
                        for (int k = 0; k < 4; ++k) h[k] = __floats2half2_rn(vv[j + 2 * k], vv[j + 2 * k + 1]);
                        *reinterpret_cast<uint4*>(nh + 2 * j) = pk;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < kChunk; j += 4) {
                        if (j >= D) break;
                        *reinterpret_cast<float4*>(nh + 4 * j) = make_float4(vv[j], vv[j + 1], vv[j + 2], vv[j + 3]);
                    }
                }
            }
        }
        if (!own) continue;
        const float inv = (wsum == 0.f) ? 0.f : 1.0f / wsum;
        char* q = cv_out;
#pragma unroll
        for (int j = 0; j < kChunk; ++j, q += pstride)
            if (D == T * kChunk || d_lo + j < D) st_hint_f1(reinterpret_cast<float*>(q), acc[j] * inv, c.pol_stream);
    }
}

__global__ void __launch_bounds__(kThreads, MR_CV_MINBLOCKS)
cost_volume_kernel(const CvArgs a, const __grid_constant__ CvMaps maps) {
    extern __shared__ __align__(128) unsigned char smem[];
    const SmemLayout L = make_layout(a.D, a.TH, a.F, a.use_tma);
    float* win = reinterpret_cast<float*>(smem + L.win);
    float* ytile = reinterpret_cast<float*>(smem + L.ytile);
    float* cst = reinterpret_cast<float*>(smem + L.cst);
    float* pjs = reinterpret_cast<float*>(smem + L.pjs);
    float* zs = reinterpret_cast<float*>(smem + L.zs);
    unsigned char* vmask = smem + L.vmask;
    int* rowrng = reinterpret_cast<int*>(smem + L.rowrng);
    short4* bbox = reinterpret_cast<short4*>(smem + L.bbox);
    unsigned short* gid = reinterpret_cast<unsigned short*>(smem + L.gid);
    unsigned char* uflag = smem + L.uflag;
    GroupInfo* ginfo = reinterpret_cast<GroupInfo*>(smem + L.ginfo);
    unsigned short* seq2g = reinterpret_cast<unsigned short*>(smem + L.seq2g);
    int* nwin = reinterpret_cast<int*>(smem + L.nwin);
    // [0] next unit, [1] number of windows, [2 + buf] finished units of the window in buffer buf, [2 + kBuf + buf] number of
    // the last window whose TMA has been issued into buffer buf (-1: none)
    int* ctr = reinterpret_cast<int*>(smem + L.ctr);
    const uint32_t bars = smem_u32(smem + L.bars);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, D = a.D, TH = a.TH, F = a.F;
    const int b = blockIdx.z + a.b0;
    const int u0 = blockIdx.x * kOutCols - 2;  // image column of buffer column 0
    const int v0 = blockIdx.y * TH;            // image row of tile row 0
    const size_t plane = (size_t)H * W;
    const int planei = H * W;
    float* xbuf = reinterpret_cast<float*>(smem + L.xbuf) + warp * 2 * 3 * kRowStride;

    // ---- keyframe tile (+0.5, monorec_model.py:232) and hoisted SSIM terms -------------------------------------
    const float* key = a.key + (size_t)b * 3 * plane;
    for (int line = warp; line < 3 * (TH + 4); line += kWarps) {      // line = (tile row + 2) * 3 + channel
        const int rr = line / 3, ch = line - 3 * rr;
        const int v = v0 - 2 + rr;
        const float* src = key + ch * plane + (size_t)v * W;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int idx = lane + 32 * k;
            if (idx >= 66) break;
            const int u = u0 + idx - 1;
            float val = 0.f;
            if (u >= 0 && u < W && v >= 0 && v < H) val = __ldg(src + u) + 0.5f;
            ytile[line * kRowStride + idx] = val;
        }
    }
    for (int i = tid; i < D; i += kThreads) zs[i] = __ldg(a.depths + i);
    if (lane < 6) {   // columns -1 and 64 of both row buffers stay zero
        const int rb = lane / 3, ch = lane % 3;
        xbuf[(rb * 3 + ch) * kRowStride] = 0.f;
        xbuf[(rb * 3 + ch) * kRowStride + kTileCols + 1] = 0.f;
    }
    if (tid < 2 * F) rowrng[tid] = (tid & 1) ? -1 : TH;
    if (tid < 12 * F) pjs[tid] = __ldg(a.proj + (size_t)b * F * 12 + tid);
    if (tid < 2 + 2 * kBuf) ctr[tid] = (tid < 2 + kBuf) ? 0 : -1;
    if (tid == 0 && a.use_tma) {
        for (int i = 0; i < kBuf; ++i) mbar_init(bars + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // table entry for the column pair (2j, 2j+1) of e-row er, channel ch: (Y[2j], Y[2j+1], Sg[2j], Sg[2j+1]) with
    // Y = 9 mu_y = sum y, Sg = 81 (sigma_y + C2) = 9 sum y^2 - Y^2 + 81 C2
    for (int line = warp; line < 3 * (TH + 2); line += kWarps) {      // line = e-row * 3 + channel
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const int bc = lane + 32 * k;
            const float* y = ytile + line * kRowStride + bc;  // rows er..er+2 of this channel, idx bc..bc+2
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int dy = 0; dy < 3; ++dy)
#pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    float q = y[dy * 3 * kRowStride + dx];
                    s1 += q;
                    s2 = fmaf(q, q, s2);
                }
            float* dst = cst + (line * (kTileCols / 2) + (bc >> 1)) * 4 + (bc & 1);
            dst[0] = s1;
            dst[2] = fmaf(9.0f, s2, -s1 * s1) + 81.0f * kC2;
        }
    }

    const float fW = (float)W, fH = (float)H;
    const float sx_lo = -fW * 0.5f, sx_hi = 1.5f * fW;  // == grid clamp(-2, 2) in sample + 0.5 units, monorec_model.py:208
    const float sy_lo = -fH * 0.5f, sy_hi = 1.5f * fH;

    // ---- validity pre-pass for every frame: valid_f(v,u) = interior(v,u) & all_d [ sample strictly inside
    //      (1,W-2)x(1,H-2) ]  (monorec_model.py:212-219: bilinear sample of the interior mask != 0 for every plane).
    //      The D samples of a pixel lie on one line and move monotonically with the depth while the denominator keeps
    //      its sign, so the farthest and the nearest plane decide. -----------------------------------------------------
    for (int q = tid; q < F * TH * kTileCols; q += kThreads) {
        const int fr = q >> 6, bc = q & 63;            // fr = f * TH + r
        const int f = fr / TH, r = fr - f * TH;
        const int u = u0 + bc, v = v0 + r;
        bool ok = (bc >= 2) && (bc < 2 + kOutCols) && (u >= 2) && (u < W - 2) && (v >= 2) && (v < H - 2);
        if (ok) {
            const float* m = pjs + 12 * f;
            const float fu = (float)u, fv = (float)v;
            const float ax = fmaf(m[0], fu, fmaf(m[1], fv, m[2]));
            const float ay = fmaf(m[4], fu, fmaf(m[5], fv, m[6]));
            const float az = fmaf(m[8], fu, fmaf(m[9], fv, m[10]));
            const float m03 = m[3], m13 = m[7], m23 = m[11];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const float z = zs[k ? D - 1 : 0];
                const float den = fmaf(az, z, m23);
                const float inv = fast_rcp(den);
                const float sx = fmaf(fmaf(ax, z, m03), inv, -0.5f);
                const float sy = fmaf(fmaf(ay, z, m13), inv, -0.5f);
                ok = ok && (den > 0.f) && (sx > 1.0f) && (sx < fW - 2.0f) && (sy > 1.0f) && (sy < fH - 2.0f);
            }
        }
        vmask[q] = ok ? 1 : 0;
        // (a warp covers half a tile row of one frame: one pair of shared-memory atomics per warp, not per pixel)
        if (__any_sync(0xffffffffu, ok) && lane == 0) { atomicMin(&rowrng[2 * f], r); atomicMax(&rowrng[2 * f + 1], r); }
    }
    __syncthreads();

    // ---- plan: source footprint of every unit (4 corners of the rows / columns its march touches), then windows ------
    const int nunits = F * D;
    for (int u = tid; u < nunits; u += kThreads) {
        const int f = u / D, d = u - f * D;
        const int rlo = rowrng[2 * f], rhi = rowrng[2 * f + 1];
        short4 bb = make_short4(0, 0, 0, 0);
        unsigned char fl = 0;          // bit 0: footprint usable for a window, bit 1: strictly inside the image
        if (rhi >= rlo && a.use_tma) {
            const float* m = pjs + 12 * f;
            const float z = zs[d];
            float xmin = 3.0e38f, xmax = -3.0e38f, ymin = 3.0e38f, ymax = -3.0e38f;
            bool good = true;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float fu = (float)(u0 + ((k & 1) ? kTileCols - 1 : 0));
                const float fv = (float)(v0 + ((k & 2) ? rhi + 2 : rlo - 2));
                const float cx = fmaf(fmaf(m[0], fu, fmaf(m[1], fv, m[2])), z, m[3]);
                const float cy = fmaf(fmaf(m[4], fu, fmaf(m[5], fv, m[6])), z, m[7]);
                const float cz = fmaf(fmaf(m[8], fu, fmaf(m[9], fv, m[10])), z, m[11]);
                good = good && (cz > 1e-6f);
                const float inv = 1.0f / cz;
                const float sx = fminf(fmaxf(cx * inv, sx_lo), sx_hi) - 0.5f, sy = fminf(fmaxf(cy * inv, sy_lo), sy_hi) - 0.5f;
                xmin = fminf(xmin, sx); xmax = fmaxf(xmax, sx);
                ymin = fminf(ymin, sy); ymax = fmaxf(ymax, sy);
            }
            if (good) {
                // one pixel of slack on every side for the rounding differences between this estimate and stage 1
                // (the window origin is rounded down to a multiple of 4 pixels: TMA wants the innermost box coordinate
                // 16-byte aligned -- measured: tools/experiments_r02/tma_probe.cu -- negative coordinates are fine)
                const int xl = (max((int)floorf(xmin) - 1, -2) >> 2) << 2, xh = min((int)floorf(xmax) + 2, W + 1);
                const int yl = max((int)floorf(ymin) - 1, -2), yh = min((int)floorf(ymax) + 2, H + 1);
                if (xh >= xl && yh >= yl && xh - xl < kPitch && yh - yl < kWinRows) {
                    fl = 1;
                    if (xmin >= 0.05f && xmax <= fW - 1.05f && ymin >= 0.05f && ymax <= fH - 1.05f) fl = 3;
                    bb = make_short4((short)xl, (short)xh, (short)yl, (short)yh);
                }
            }
        }
        bbox[u] = bb;
        uflag[u] = fl;
        gid[u] = 0xFFFF;
    }
    __syncthreads();
    if (lane == 0 && warp < F) {   // one frame per warp: the F serial scans run side by side instead of as divergent lanes
        // greedy runs of consecutive planes whose union still fits one window
        const int f = warp;
        int ng = 0;
        const int rlo = rowrng[2 * f], rhi = rowrng[2 * f + 1];
        if (rhi >= rlo && a.use_tma) {
            int d = 0;
            while (d < D) {
                if (!(uflag[f * D + d] & 1)) { ++d; continue; }
                short4 g = bbox[f * D + d];
                int e = d + 1;
                while (e < D && (uflag[f * D + e] & 1)) {
                    const short4 o = bbox[f * D + e];
                    const int xl = min(g.x, o.x), xh = max(g.y, o.y), yl = min(g.z, o.z), yh = max(g.w, o.w);
                    if (xh - xl >= kPitch || yh - yl >= kWinRows) break;
                    g = make_short4((short)xl, (short)xh, (short)yl, (short)yh);
                    ++e;
                }
                if (e - d >= kMinGroup) {
                    GroupInfo gi;
                    gi.wx0 = g.x; gi.wy0 = g.z;
                    gi.nrows = (short)(((g.w - g.z + 1) + kBoxRows - 1) / kBoxRows * kBoxRows);
                    gi.f = (short)f; gi.count = (short)(e - d); gi.seq = (short)ng; gi.pad0 = gi.pad1 = 0;
                    ginfo[f * D + ng] = gi;
                    for (int k = d; k < e; ++k) gid[f * D + k] = (unsigned short)(f * D + ng);
                    ++ng;
                }
                d = e;
            }
        }
        nwin[f] = ng;
    }
    __syncthreads();
    if (lane == 0 && warp < F) {
        const int f = warp;
        int base = 0;
        for (int k = 0; k < f; ++k) base += nwin[k];
        for (int g = 0; g < nwin[f]; ++g) {
            ginfo[f * D + g].seq = (short)(base + g);
            seq2g[base + g] = (unsigned short)(f * D + g);
        }
        if (f == F - 1) ctr[1] = base + nwin[f];
    }
    __syncthreads();

    // window `seq` -> buffer seq % kBuf: 3 channel planes of nrows rows, TMA boxes of kBoxRows rows
    auto issue_window = [&](int seq) {
        const GroupInfo gi = ginfo[seq2g[seq]];
        const int buf = seq % kBuf;
        const uint32_t bar = bars + 8 * buf;
        const uint32_t dst = smem_u32(win + (size_t)buf * kWinFloats);
        atomicExch(&ctr[2 + kBuf + buf], seq);   // (an atomic, like its readers: a flag, not a data race)
        mbar_expect_tx(bar, (uint32_t)(3 * gi.nrows * kPitch * 4));
        for (int ch = 0; ch < 3; ++ch)
            for (int r8 = 0; r8 < gi.nrows; r8 += kBoxRows)
                tma_load_3d(dst + (uint32_t)((ch * kWinRows + r8) * kPitch * 4), &maps.m[gi.f], bar, gi.wx0, gi.wy0 + r8, b * 3 + ch);
    };
    if (tid == 0 && a.use_tma) {
        const int nw = ctr[1];
        for (int s = 0; s < kBuf && s < nw; ++s) issue_window(s);
    }

    const float2 fu2 = make_float2((float)(u0 + lane), (float)(u0 + lane + 32));
    // stage 2 writes the single-frame volume for output columns u0 + 2l, u0 + 2l + 1 (lanes 1..30)
    const int ucol = u0 + 2 * lane;
    const bool st0 = (lane >= 1) && (lane <= 30) && (ucol < W);
    const bool st1 = (lane >= 1) && (lane <= 30) && (ucol + 1 < W);
    const uint64_t pol_keep = l2_policy_evict_last(), pol_stream = l2_policy_evict_first();

    // ---- march over the F*D (frame, plane) units; no CTA-wide barrier in here ------------------------------------
    Stage2Ctx c2;
    c2.cw0 = bc2(a.cw0); c2.cw1 = bc2(a.cw1); c2.cw2 = bc2(a.cw2);
    c2.st0 = st0; c2.st1 = st1; c2.pairflag = (st0 && st1 && ((W & 1) == 0)) ? 1 : 0; c2.pol_keep = pol_keep;
    Stage1Ctx c1;
    c1.W = W; c1.H = H; c1.planei = planei;
    c1.sx_lo = sx_lo; c1.sx_hi = sx_hi; c1.sy_lo = sy_lo; c1.sy_hi = sy_hi;
    // per-lane shared addresses for stage 2 (columns 2l-1 .. 2l+2 live at float index 2l .. 2l+3 of a row)
    const uint32_t xb_s = smem_u32(xbuf);
    const uint32_t ys_s = smem_u32(ytile) + 8 * lane;
    const uint32_t cs_s = smem_u32(cst) + 16 * lane;
    const uint32_t win_s = smem_u32(win);
    for (; MR_CV_SKIP != 1;) {
        int unit = 0;
        if (lane == 0) unit = atomicAdd(&ctr[0], 1);
        unit = __shfl_sync(0xffffffffu, unit, 0);
        if (unit >= nunits) break;
        const int f = unit / D, d = unit - f * D;
        const int rlo = rowrng[2 * f], rhi = rowrng[2 * f + 1];
        if (rhi < rlo) continue;  // no valid pixel of this tile for frame f: the per-pixel phase zero-fills
        setup_stage1(c1, pjs + 12 * f, zs[d], fu2);
        const int nsteps = rhi - rlo + 5;
        const float fv0 = (float)(v0 + rlo - 2);
        const uint32_t yr = ys_s + rlo * kYRowBytes;          // tile row rlo-2 is keyframe-tile row rlo
        // the SSIM row of step t is tile row rlo-3+t, whose table row is rlo-2+t (t = 0, 1 read throw-away rows, possibly
        // in front of the table: still inside this CTA's shared memory, see make_layout)
        const uint32_t cr = cs_s + (rlo - 2) * kCRowBytes;
        float* out = a.sfcv + (((size_t)f * a.B + b) * D + d) * plane + ((ptrdiff_t)(v0 + rlo - 4) * W + ucol);
        const unsigned g = gid[unit];
        if (g != 0xFFFFu) {
            const GroupInfo gi = ginfo[g];
            const int buf = gi.seq % kBuf;
            // A parity wait is only meaningful on the phase in progress or the one before: first make sure this window's
            // load has been issued (the buffer's previous window is then complete and consumed), then wait for its bytes.
            for (uint32_t spin = 0;; ++spin) {
                int armed = 0;
                if (lane == 0) armed = atomicAdd(&ctr[2 + kBuf + buf], 0);
                if (__shfl_sync(0xffffffffu, armed, 0) >= gi.seq) break;
                if (spin > (1u << 22)) __trap();
            }
            mbar_wait(bars + 8 * buf, (uint32_t)((gi.seq / kBuf) & 1));
            const uint32_t wb = win_s + (uint32_t)buf * (kWinFloats * 4);
            if (uflag[unit] & 2) {
                // tap address = wb + 4 ((bits(ty) - kMagicBits - wy0) kPitch + bits(tx) - kMagicBits - wx0), mod 2^32
                c1.kaddr = wb - 4u * ((uint32_t)(kMagicBits + gi.wy0) * kPitch + (uint32_t)(kMagicBits + gi.wx0));
                march_unit<0>(c1, c2, xb_s, lane, fv0, nsteps, yr, cr, out, W);
            } else {
                c1.kaddr = wb - 4u * ((uint32_t)(int)gi.wy0 * kPitch + (uint32_t)(int)gi.wx0);
                march_unit<1>(c1, c2, xb_s, lane, fv0, nsteps, yr, cr, out, W);
            }
            // hand the buffer on: the last unit of the window re-arms it with the window after next
            if (lane == 0) {
                __threadfence_block();
                const int fin = atomicAdd(&ctr[2 + buf], 1) + 1;
                if (fin == gi.count) {
                    ctr[2 + buf] = 0;
                    __threadfence_block();
                    if (gi.seq + kBuf < ctr[1]) {
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        issue_window(gi.seq + kBuf);
                    }
                }
            }
            __syncwarp();
        } else {
            c1.img = a.frames[f] + (size_t)b * 3 * plane;
            march_unit<2>(c1, c2, xb_s, lane, fv0, nsteps, yr, cr, out, W);
        }
    }
    __syncthreads();  // the marching warps' global stores are visible to the whole CTA from here on

    // ---- per-pixel phase: view weights (monorec_model.py:257-260), zeroing of invalid pixels (:251) and fusion
    //      cv = sum_f w_f (1 - 2 sad_f) / sum_f w_f, 0 where sum_f w_f == 0 (:262-269).  Each thread reads back the
    //      L2-hot single-frame values of its pixel once per frame. ------------------------------------------------------
    PixelPhase pp;
    pp.cv = a.cv; pp.sfcv = a.sfcv; pp.sf_nhwc = a.sf_nhwc; pp.sf_nhwc_half = a.sf_nhwc_half;
    pp.vmask = vmask; pp.B = a.B; pp.F = F; pp.D = D; pp.H = H; pp.W = W; pp.TH = TH; pp.b = b; pp.u0 = u0; pp.v0 = v0;
    pp.inv_dm1 = a.inv_dm1;
    // exp(-alpha (sad - min sad)^2) with sad = (1 - sv) / 2 is ex2(-(k (max sv - sv))^2), k = sqrt(alpha log2(e)) / 2
    pp.kq = 0.5f * sqrtf(a.alpha * 1.4426950408889634f);
    pp.pol_stream = pol_stream;
    if (MR_CV_SKIP != 2) {
        if (D <= kChunk) pixel_phase<1>(pp, tid);
        else if (D <= 2 * kChunk) pixel_phase<2>(pp, tid);
        else pixel_phase<4>(pp, tid);
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

int pick_tile_rows(int D, int F, int use_tma) {
    const int limit = 227 * 1024;
    for (int th = MR_CV_TILE_ROWS; th >= 2; th >>= 1)
        if (make_layout(D, th, F, use_tma).total <= limit) return th;
    return 0;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;   // benign race: every thread resolves the same pointer
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
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
                           float alpha, const float* chan_w, int b_begin, int b_count, int gather_only,
                           cudaStream_t stream, void* sf_nhwc, int sf_nhwc_dtype) {
    MR_REQUIRE(keyframe && frames && proj && depths && out_cv && out_sfcv, "mr_cost_volume_fwd: null pointer");
    MR_REQUIRE(b_begin >= 0 && b_count >= 1 && b_begin + b_count <= B, "mr_cost_volume_fwd: bad batch range");
    MR_REQUIRE(B >= 1 && B <= 21845, "mr_cost_volume_fwd: batch %d out of range", B);
    MR_REQUIRE(F >= 1 && F <= MR_MAX_FRAMES, "mr_cost_volume_fwd: 1 <= F <= %d required (got %d)", MR_MAX_FRAMES, F);
    MR_REQUIRE(D >= 2 && D <= 128, "mr_cost_volume_fwd: 2 <= D <= 128 required (got %d)", D);
    MR_REQUIRE(H >= 5 && W >= 5 && H <= 16384 && W <= 16384, "mr_cost_volume_fwd: image size %dx%d out of range", H, W);
    CvArgs a{};
    a.key = keyframe;
    for (int f = 0; f < F; ++f) {
        MR_REQUIRE(frames[f] != nullptr, "mr_cost_volume_fwd: null frame pointer %d", f);
        a.frames[f] = frames[f];
    }
    a.proj = proj; a.depths = depths; a.cv = out_cv; a.sfcv = out_sfcv;
    MR_REQUIRE(sf_nhwc == nullptr || (D <= kChunk && (D % 8) == 0 && (reinterpret_cast<uintptr_t>(sf_nhwc) & 15) == 0 &&
                                      (sf_nhwc_dtype == MR_DT_F32 || sf_nhwc_dtype == MR_DT_F16)),
               "mr_cost_volume_fwd_nhwc: the NHWC copy needs D <= %d, D %% 8 == 0, a 16-byte aligned buffer and an fp32 / half type", kChunk);
    a.sf_nhwc = sf_nhwc; a.sf_nhwc_half = (sf_nhwc_dtype == MR_DT_F16) ? 1 : 0;
    a.B = B; a.F = F; a.D = D; a.H = H; a.W = W; a.b0 = b_begin;
    // TMA addresses the frames as (W, H, 3B) tensors: the row pitch must be a multiple of 16 bytes and the base 16-byte
    // aligned; otherwise (ragged widths) every unit takes the global gather of the same kernel.
    CvMaps local{};       // by-value kernel parameter (__grid_constant__): one tensor map per source frame
    a.use_tma = 0;
    EncodeTiledFn encode = gather_only ? nullptr : get_encode_fn();
    if (encode != nullptr && (W % 4) == 0) {
        bool ok = true;
        for (int f = 0; f < F && ok; ++f) {
            if (reinterpret_cast<uintptr_t>(frames[f]) & 15) { ok = false; break; }
            const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)3 * B};
            const cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4};
            const cuuint32_t box[3] = {(cuuint32_t)kPitch, (cuuint32_t)kBoxRows, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            const CUresult r = encode(&local.m[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(frames[f]), gdim, gstr,
                                      box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            ok = (r == CUDA_SUCCESS);
        }
        if (ok) {
            for (int f = F; f < MR_MAX_FRAMES; ++f) local.m[f] = local.m[0];
            a.use_tma = 1;
        }
    }
    a.TH = pick_tile_rows(D, F, a.use_tma);
    MR_REQUIRE(a.TH > 0, "mr_cost_volume_fwd: no tile height fits shared memory for D=%d F=%d", D, F);
    a.alpha = alpha;
    a.inv_dm1 = (float)(1.0 / (double)(D - 1));
    const float def_w[3] = {5.f / 32.f, 16.f / 32.f, 11.f / 32.f};  // monorec_model.py:133
    const float* cw = chan_w ? chan_w : def_w;
    a.cw0 = cw[0] / 9.f; a.cw1 = cw[1] / 9.f; a.cw2 = cw[2] / 9.f;  // monorec_model.py:141 (weights / patch_size^2)
    const SmemLayout L = make_layout(D, a.TH, F, a.use_tma);
    dim3 grid((W + kOutCols - 1) / kOutCols, (H + a.TH - 1) / a.TH, b_count);
    static int smem_set = 0;   // the attribute is per function and per device context; setting it again is harmless
    if (smem_set < L.total) {
        MR_CUDA(cudaFuncSetAttribute(cost_volume_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        smem_set = 227 * 1024;
    }
    cost_volume_kernel<<<grid, kThreads, L.total, stream>>>(a, local);
    MR_LAUNCH_CHECK("cost_volume_kernel");
    return MR_OK;
}

extern "C" int mr_cost_volume_fwd(const float* keyframe, const float* const* frames, const float* proj,
                                  const float* depths, float* out_cv, float* out_sfcv, int B, int F, int D, int H,
                                  int W, float alpha, const float* chan_w, void* stream) {
    return mr::launch_cost_volume(keyframe, frames, proj, depths, out_cv, out_sfcv, B, F, D, H, W, alpha, chan_w, 0,
                                  B, 0, (cudaStream_t)stream, nullptr, 0);
}

extern "C" int mr_cost_volume_fwd_gather(const float* keyframe, const float* const* frames, const float* proj,
                                         const float* depths, float* out_cv, float* out_sfcv, int B, int F, int D, int H,
                                         int W, float alpha, const float* chan_w, void* stream) {
    return mr::launch_cost_volume(keyframe, frames, proj, depths, out_cv, out_sfcv, B, F, D, H, W, alpha, chan_w, 0,
                                  B, 1, (cudaStream_t)stream, nullptr, 0);
}

extern "C" int mr_cost_volume_fwd_nhwc(const float* keyframe, const float* const* frames, const float* proj, const float* depths,
                                       float* out_cv, float* out_sfcv, void* out_sfcv_nhwc, int nhwc_dtype, int B, int F, int D,
                                       int H, int W, float alpha, const float* chan_w, void* stream) {
    MR_REQUIRE(out_sfcv_nhwc != nullptr, "mr_cost_volume_fwd_nhwc: null NHWC buffer");
    return mr::launch_cost_volume(keyframe, frames, proj, depths, out_cv, out_sfcv, B, F, D, H, W, alpha, chan_w, 0, B, 0,
                                  (cudaStream_t)stream, out_sfcv_nhwc, nhwc_dtype);
}
