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

namespace {

constexpr int kTileCols = 64;   // buffer columns per tile row (output columns + 2-px halo each side)
constexpr int kOutCols = 60;    // output columns per tile
constexpr int kRowStride = 68;  // floats per smem image row: column b lives at index b+1 (so [2l-1, 2l+2] is 8B aligned)
constexpr int kThreads = 512;
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
    int B, F, D, H, W, TH, b0;
    float alpha, inv_dm1;
    float cw0, cw1, cw2;                 // channel weights / 9
};

struct SmemLayout {
    int sad, ytile, cst, xbuf, wts, zs, vmask, misc, total;  // byte offsets
};

__host__ __device__ inline SmemLayout make_layout(int D, int TH, int F) {
    SmemLayout L;
    int off = 0;
    L.sad = off;   off += D * TH * kTileCols * 4;
    L.ytile = off; off += 3 * (TH + 4) * kRowStride * 4;
    L.cst = off;   off += 3 * (TH + 2) * kTileCols * 8;
    L.xbuf = off;  off += kWarps * 3 * kRowStride * 4;
    L.wts = off;   off += F * TH * kTileCols * 4;
    L.zs = off;    off += ((D + 3) / 4) * 16;
    L.vmask = off; off += TH * kTileCols;
    L.misc = off;  off += 16;
    L.total = off;
    return L;
}

__device__ __forceinline__ float ssim_err(float s1, float sxx, float sxy, float mu_y, float sy2) {
    // layers.py:123-137 with the 3x3 means expressed through box sums; mu_y and sy2 = sigma_y + C2 are hoisted.
    const float k9 = 1.0f / 9.0f;
    float mu_x = s1 * k9;
    float mxy = mu_x * mu_y;
    float mxx = mu_x * mu_x;
    float sig_xy = fmaf(sxy, k9, -mxy);
    float sig_x = fmaf(sxx, k9, -mxx);
    float n = fmaf(2.0f, mxy, kC1) * fmaf(2.0f, sig_xy, kC2);
    float d = (mxx + fmaf(mu_y, mu_y, kC1)) * (sig_x + sy2);
    float q = __fdividef(n, d);
    return __saturatef(fmaf(-0.5f, q, 0.5f));
}

__global__ void __launch_bounds__(kThreads, 1) cost_volume_kernel(const CvArgs a) {
    extern __shared__ __align__(16) unsigned char smem[];
    const SmemLayout L = make_layout(a.D, a.TH, a.F);
    float* sad_s = reinterpret_cast<float*>(smem + L.sad);
    float* ytile = reinterpret_cast<float*>(smem + L.ytile);
    float* cst = reinterpret_cast<float*>(smem + L.cst);
    float* wts = reinterpret_cast<float*>(smem + L.wts);
    float* zs = reinterpret_cast<float*>(smem + L.zs);
    unsigned char* vmask = smem + L.vmask;
    int* misc = reinterpret_cast<int*>(smem + L.misc);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = a.H, W = a.W, D = a.D, TH = a.TH, F = a.F;
    const int b = blockIdx.z + a.b0;
    const int u0 = blockIdx.x * kOutCols - 2;  // image column of buffer column 0
    const int v0 = blockIdx.y * TH;            // image row of tile row 0
    const size_t plane = (size_t)H * W;
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
        reinterpret_cast<float2*>(cst)[(ch * (TH + 2) + er) * kTileCols + bc] = make_float2(mu, sy2);
    }

    const float fW = (float)W, fH = (float)H;
    const float sx_lo = -(fW + 1.f) * 0.5f, sx_hi = (3.f * fW - 1.f) * 0.5f;  // == grid clamp(-2, 2), monorec_model.py:208
    const float sy_lo = -(fH + 1.f) * 0.5f, sy_hi = (3.f * fH - 1.f) * 0.5f;

    for (int f = 0; f < F; ++f) {
        const float* pj = a.proj + ((size_t)b * F + f) * 12;
        const float m00 = __ldg(pj + 0), m01 = __ldg(pj + 1), m02 = __ldg(pj + 2), m03 = __ldg(pj + 3);
        const float m10 = __ldg(pj + 4), m11 = __ldg(pj + 5), m12 = __ldg(pj + 6), m13 = __ldg(pj + 7);
        const float m20 = __ldg(pj + 8), m21 = __ldg(pj + 9), m22 = __ldg(pj + 10), m23 = __ldg(pj + 11);
        const float* img = a.frames[f] + (size_t)b * 3 * plane;

        if (tid == 0) { misc[0] = TH; misc[1] = -1; }
        __syncthreads();  // also orders the previous frame's phase 2 before sad/vmask are overwritten

        // ---- validity pre-pass: valid_f(v,u) = interior(v,u) & all_d [ sample strictly inside (1,W-2)x(1,H-2) ] ----
        // (monorec_model.py:212-219: bilinear sample of the interior mask != 0 for every plane)
        for (int p = tid; p < TH * kTileCols; p += kThreads) {
            int r = p >> 6, bc = p & 63;
            int u = u0 + bc, v = v0 + r;
            bool ok = (bc >= 2) && (bc < 2 + kOutCols) && (u >= 2) && (u < W - 2) && (v >= 2) && (v < H - 2);
            if (ok) {
                float fu = (float)u, fv = (float)v;
                float ax = fmaf(m00, fu, fmaf(m01, fv, m02));
                float ay = fmaf(m10, fu, fmaf(m11, fv, m12));
                float az = fmaf(m20, fu, fmaf(m21, fv, m22));
                for (int d = 0; d < D; ++d) {
                    float z = zs[d];
                    float cz = fmaf(az, z, m23);
                    float inv = __fdividef(1.0f, cz);
                    float sx = fmaf(fmaf(ax, z, m03), inv, -0.5f);
                    float sy = fmaf(fmaf(ay, z, m13), inv, -0.5f);
                    ok = ok && (sx > 1.0f) && (sx < fW - 2.0f) && (sy > 1.0f) && (sy < fH - 2.0f);
                }
            }
            vmask[p] = ok ? 1 : 0;
            if (ok) { atomicMin(&misc[0], r); atomicMax(&misc[1], r); }
        }
        __syncthreads();
        const int rlo = misc[0], rhi = misc[1];

        // ---- march: one plane per warp at a time -------------------------------------------------------------------
        if (rhi >= rlo) {
            for (int d = warp; d < D; d += kWarps) {
                const float z = zs[d];
                float h1a[3][2], h1b[3][2], hxa[3][2], hxb[3][2], hya[3][2], hyb[3][2];
                float hEa[2], hEb[2];
#pragma unroll
                for (int c = 0; c < 3; ++c)
#pragma unroll
                    for (int j = 0; j < 2; ++j) h1a[c][j] = h1b[c][j] = hxa[c][j] = hxb[c][j] = hya[c][j] = hyb[c][j] = 0.f;
                hEa[0] = hEa[1] = hEb[0] = hEb[1] = 0.f;

                const int nsteps = rhi - rlo + 5;
                for (int t = 0; t < nsteps; ++t) {
                    const int r = rlo - 2 + t;  // tile-relative row of the warped row produced in this step
                    // ---------- stage 1: warp one row (lane = column, two rounds) ----------
                    {
                        const float fv = (float)(v0 + r);
                        const float rx = fmaf(m01, fv, m02), ry = fmaf(m11, fv, m12), rz = fmaf(m21, fv, m22);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int bc = lane + 32 * k;
                            const float fu = (float)(u0 + bc);
                            float cx = fmaf(fmaf(m00, fu, rx), z, m03);
                            float cy = fmaf(fmaf(m10, fu, ry), z, m13);
                            float cz = fmaf(fmaf(m20, fu, rz), z, m23);
                            float inv = __fdividef(1.0f, cz);
                            float sx = fminf(fmaxf(fmaf(cx, inv, -0.5f), sx_lo), sx_hi);
                            float sy = fminf(fmaxf(fmaf(cy, inv, -0.5f), sy_lo), sy_hi);
                            float x0f = floorf(sx), y0f = floorf(sy);
                            float wx1 = sx - x0f, wy1 = sy - y0f;
                            float wx0 = (x0f + 1.0f) - sx, wy0 = (y0f + 1.0f) - sy;
                            int x0 = (int)x0f, y0 = (int)y0f;
                            // zero padding: taps outside the image contribute 0 (F.grid_sample padding_mode="zeros")
                            if ((unsigned)x0 >= (unsigned)W) wx0 = 0.f;
                            if ((unsigned)(x0 + 1) >= (unsigned)W) wx1 = 0.f;
                            if ((unsigned)y0 >= (unsigned)H) wy0 = 0.f;
                            if ((unsigned)(y0 + 1) >= (unsigned)H) wy1 = 0.f;
                            int xa = min(max(x0, 0), W - 1), xb = min(max(x0 + 1, 0), W - 1);
                            int ya = min(max(y0, 0), H - 1), yb = min(max(y0 + 1, 0), H - 1);
                            float w00 = wx0 * wy0, w01 = wx1 * wy0, w10 = wx0 * wy1, w11 = wx1 * wy1;
                            const float* p0 = img + (size_t)ya * W;
                            const float* p1 = img + (size_t)yb * W;
#pragma unroll
                            for (int c = 0; c < 3; ++c) {
                                float i00 = __ldg(p0 + c * plane + xa), i01 = __ldg(p0 + c * plane + xb);
                                float i10 = __ldg(p1 + c * plane + xa), i11 = __ldg(p1 + c * plane + xb);
                                float val = i00 * w00;
                                val = fmaf(i01, w01, val);
                                val = fmaf(i10, w10, val);
                                val = fmaf(i11, w11, val);
                                xbuf[c * kRowStride + bc + 1] = val + 0.5f;
                            }
                        }
                    }
                    __syncwarp();
                    // ---------- stage 2: lane owns buffer columns 2l, 2l+1 ----------
                    float E[2] = {0.f, 0.f};
                    {
                        const float* yrow = ytile + (r + 2) * kRowStride + 2 * lane;
                        const float* crow = cst + ((size_t)r * kTileCols + 2 * lane) * 2;
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            float2 xl = *reinterpret_cast<const float2*>(xbuf + c * kRowStride + 2 * lane);
                            float2 xr = *reinterpret_cast<const float2*>(xbuf + c * kRowStride + 2 * lane + 2);
                            float2 yl = *reinterpret_cast<const float2*>(yrow + c * (TH + 4) * kRowStride);
                            float2 yr = *reinterpret_cast<const float2*>(yrow + c * (TH + 4) * kRowStride + 2);
                            // columns 2l-1, 2l, 2l+1, 2l+2
                            float mid1 = xl.y + xr.x;
                            float h1_0 = xl.x + mid1, h1_1 = mid1 + xr.y;
                            float xx0 = xl.x * xl.x, xx1 = xl.y * xl.y, xx2 = xr.x * xr.x, xx3 = xr.y * xr.y;
                            float midx = xx1 + xx2;
                            float hx_0 = xx0 + midx, hx_1 = midx + xx3;
                            float xy0 = xl.x * yl.x, xy1 = xl.y * yl.y, xy2 = xr.x * yr.x, xy3 = xr.y * yr.y;
                            float midy = xy1 + xy2;
                            float hy_0 = xy0 + midy, hy_1 = midy + xy3;
                            if (t >= 2) {
                                float4 k4 = *reinterpret_cast<const float4*>(crow + (size_t)c * (TH + 2) * kTileCols * 2);
                                float cw = (c == 0) ? a.cw0 : ((c == 1) ? a.cw1 : a.cw2);
                                float e0 = ssim_err(h1a[c][0] + h1b[c][0] + h1_0, hxa[c][0] + hxb[c][0] + hx_0,
                                                    hya[c][0] + hyb[c][0] + hy_0, k4.x, k4.y);
                                float e1 = ssim_err(h1a[c][1] + h1b[c][1] + h1_1, hxa[c][1] + hxb[c][1] + hx_1,
                                                    hya[c][1] + hyb[c][1] + hy_1, k4.z, k4.w);
                                E[0] = fmaf(cw, e0, E[0]);
                                E[1] = fmaf(cw, e1, E[1]);
                            }
                            h1a[c][0] = h1b[c][0]; h1b[c][0] = h1_0; h1a[c][1] = h1b[c][1]; h1b[c][1] = h1_1;
                            hxa[c][0] = hxb[c][0]; hxb[c][0] = hx_0; hxa[c][1] = hxb[c][1]; hxb[c][1] = hx_1;
                            hya[c][0] = hyb[c][0]; hyb[c][0] = hy_0; hya[c][1] = hyb[c][1]; hyb[c][1] = hy_1;
                        }
                    }
                    if (t >= 2) {
                        float eL = __shfl_up_sync(0xffffffffu, E[1], 1);
                        float eR = __shfl_down_sync(0xffffffffu, E[0], 1);
                        float mid = E[0] + E[1];
                        float hE0 = eL + mid, hE1 = mid + eR;
                        if (t >= 4) {
                            float2 s = make_float2(hEa[0] + hEb[0] + hE0, hEa[1] + hEb[1] + hE1);
                            *reinterpret_cast<float2*>(sad_s + ((size_t)d * TH + (r - 2)) * kTileCols + 2 * lane) = s;
                        }
                        hEa[0] = hEb[0]; hEb[0] = hE0; hEa[1] = hEb[1]; hEb[1] = hE1;
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();

        // ---- phase 2: per-pixel view weight and single-frame volume (monorec_model.py:250-260) -------------------
        float* sf_out = a.sfcv + ((size_t)f * a.B + b) * D * plane;
        for (int p = tid; p < TH * kTileCols; p += kThreads) {
            int r = p >> 6, bc = p & 63;
            int u = u0 + bc, v = v0 + r;
            bool own = (bc >= 2) && (bc < 2 + kOutCols) && (u < W) && (v < H);
            if (!own) continue;
            const bool valid = vmask[p] != 0;
            float w = 0.f;
            float* out = sf_out + (size_t)v * W + u;
            if (valid) {
                const float* s = sad_s + (size_t)r * kTileCols + bc;
                float m = s[0];
                for (int d = 1; d < D; ++d) m = fminf(m, s[(size_t)d * TH * kTileCols]);
                float sum = 0.f;
                for (int d = 0; d < D; ++d) {
                    float sv = s[(size_t)d * TH * kTileCols];
                    float df = sv - m;
                    sum += __expf(-a.alpha * df * df);
                    out[(size_t)d * plane] = fmaf(-2.0f, sv, 1.0f);
                }
                // weight = 1 - 1/(D-1) * (sum - 1): separate roundings as in the reference so that flat-cost pixels
                // (sum == D) give exactly 0 (monorec_model.py:258, :265-269)
                w = __fsub_rn(1.0f, __fmul_rn(a.inv_dm1, __fsub_rn(sum, 1.0f)));
            } else {
                for (int d = 0; d < D; ++d) out[(size_t)d * plane] = 0.f;
            }
            wts[f * TH * kTileCols + p] = w;
        }
    }
    __syncthreads();

    // ---- fusion (monorec_model.py:262-269): cv = sum_f w_f (1 - 2 sad_f) / sum_f w_f, 0 where sum_f w_f == 0 -----
    for (int p = tid; p < TH * kTileCols; p += kThreads) {
        int r = p >> 6, bc = p & 63;
        int u = u0 + bc, v = v0 + r;
        bool own = (bc >= 2) && (bc < 2 + kOutCols) && (u < W) && (v < H);
        if (!own) continue;
        float wf[MR_MAX_FRAMES];
        float wsum = 0.f;
#pragma unroll
        for (int f = 0; f < MR_MAX_FRAMES; ++f) {
            wf[f] = (f < F) ? wts[f * TH * kTileCols + p] : 0.f;
            wsum += wf[f];
        }
        float* out = a.cv + (size_t)b * D * plane + (size_t)v * W + u;
        if (wsum == 0.f) {
            for (int d = 0; d < D; ++d) out[(size_t)d * plane] = 0.f;
        } else {
            const float inv = 1.0f / wsum;
            const float* sf = a.sfcv + (size_t)b * D * plane + (size_t)v * W + u;
            const size_t fstride = (size_t)a.B * D * plane;
            for (int d = 0; d < D; ++d) {
                float num = 0.f;
#pragma unroll
                for (int f = 0; f < MR_MAX_FRAMES; ++f)
                    if (f < F && wf[f] != 0.f) num = fmaf(wf[f], __ldcg(sf + f * fstride + (size_t)d * plane), num);
                out[(size_t)d * plane] = num * inv;
            }
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

int pick_tile_rows(int D, int F) {
    const int limit = 227 * 1024;
    for (int th = 16; th >= 2; th >>= 1)
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
                           float alpha, const float* chan_w, int b_begin, int b_count, cudaStream_t stream) {
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
    MR_CUDA(cudaFuncSetAttribute(cost_volume_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L.total));
    dim3 grid((W + kOutCols - 1) / kOutCols, (H + a.TH - 1) / a.TH, b_count);
    cost_volume_kernel<<<grid, kThreads, L.total, stream>>>(a);
    MR_LAUNCH_CHECK("cost_volume_kernel");
    return MR_OK;
}

extern "C" int mr_cost_volume_fwd(const float* keyframe, const float* const* frames, const float* proj,
                                  const float* depths, float* out_cv, float* out_sfcv, int B, int F, int D, int H,
                                  int W, float alpha, const float* chan_w, void* stream) {
    return mr::launch_cost_volume(keyframe, frames, proj, depths, out_cv, out_sfcv, B, F, D, H, W, alpha, chan_w, 0,
                                  B, (cudaStream_t)stream);
}
