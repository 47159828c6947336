// Point-cloud side of SURVEY.md section 8f row 3 (include/monorec_b200.h: mr_pointcloud_keep_mask, mr_pointcloud_add).
//
// Reference: create_pointcloud.py:65-105 (moving-object mask dilated by a 33x33 box, votes over a sliding window of frames,
// depth *= mask) and utils/ply_utils.py:34-53 PLYSaver.add_depthmap (1/x, distance / roi / dropout filter, Backprojection
// (model/layers.py:43-58) with inv(K), pose transform, boolean-mask compaction, .cpu().tolist() per frame).  Here the
// vertices stay on the device in one growing buffer, in the reference's order (batch element, then pixel, row-major).
#include "mr_common.cuh"
#include <cstdint>

namespace {

constexpr int kBlk = 256;          // pixels per block of the compaction kernels

// keep[b,p] = 1 iff no pixel with cv_mask >= thresh lies in the (fill+1) x (fill+1) window centred on p
// (create_pointcloud.py:77-78: conv2d(mask, ones(fill+1), padding = fill // 2) < 1; zero padding outside the image)
__global__ void keep_mask_kernel(const float* __restrict__ cv_mask, float* __restrict__ keep, int H, int W, int rad, float thresh) {
    // separable box test: a block handles one image row segment; rows are scanned directly (the mask is small and L2-resident)
    const int b = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const float* m = cv_mask + (size_t)b * H * W;
    bool hit = false;
    const int y0 = max(y - rad, 0), y1 = min(y + rad, H - 1), x0 = max(x - rad, 0), x1 = min(x + rad, W - 1);
    for (int yy = y0; yy <= y1 && !hit; ++yy) {
        const float* row = m + (size_t)yy * W;
        for (int xx = x0; xx <= x1; ++xx)
            if (__ldg(row + xx) >= thresh) { hit = true; break; }
    }
    keep[((size_t)b * H + y) * W + x] = hit ? 0.f : 1.f;
}

struct PtrPackPC {
    const float* p[16];
};

struct PcArgs {
    const float* inv_depth;   // [B,1,H,W] data_dict["result"]
    const float* image;       // [B,3,H,W] keyframe in [-0.5, 0.5]
    const float* K;           // [B,4,4] keyframe intrinsics
    const float* pose;        // [B,4,4] keyframe pose (camera -> world)
    PtrPackPC keeps;          // n_masks keep masks [B,1,H,W] (sliding window), may be empty
    int n_masks, min_hits;
    const float* rnd;         // [B,1,H,W] uniform numbers for the dropout, or nullptr
    float dropout, min_d, max_d;
    int B, H, W, r0, r1, c0, c1, use_roi;
    int* counts;              // [B * blocks_per_image] kept vertices per block, then exclusive offsets (in place)
    float* out;               // [capacity][6]
    long long capacity;
    long long base;           // vertices already in the buffer
    long long* total;         // device: number of vertices in the buffer after this call
};

__device__ __forceinline__ bool keep_vertex(const PcArgs& a, int b, int i, float& depth) {
    const size_t o = (size_t)b * a.H * a.W + i;
    float inv = __ldg(a.inv_depth + o);
    if (a.n_masks > 0) {     // mask = sum(mask_buffer) > buffer_length - min_hits; depth *= mask  (create_pointcloud.py:93-95)
        float s = 0.f;
        for (int k = 0; k < a.n_masks; ++k) s += __ldg(a.keeps.p[k] + o);
        inv *= (s > (float)(a.n_masks - a.min_hits)) ? 1.f : 0.f;
    }
    depth = __fdiv_rn(1.0f, inv);                           // ply_utils.py:36 (1 / 0 = inf fails the range test below)
    bool ok = (a.min_d <= depth) && (depth <= a.max_d);     // :38
    if (a.use_roi) {                                        // :39-43
        const int y = i / a.W, x = i - y * a.W;
        ok = ok && y >= a.r0 && y < a.r1 && x >= a.c0 && x < a.c1;
    }
    if (a.rnd != nullptr && a.dropout > 0.f) ok = ok && (__ldg(a.rnd + o) > a.dropout);   // :44-45
    return ok;
}

__global__ void pc_count_kernel(const PcArgs a) {
    const int b = blockIdx.y, i = blockIdx.x * kBlk + threadIdx.x;
    float depth;
    const bool ok = (i < a.H * a.W) && keep_vertex(a, b, i, depth);
    const int n = __syncthreads_count(ok ? 1 : 0);
    if (threadIdx.x == 0) a.counts[b * gridDim.x + blockIdx.x] = n;
}

// exclusive scan of the per-block counts (a few thousand entries: one block, sequential chunks per thread + block scan)
__global__ void pc_scan_kernel(int* counts, int n, long long base, long long capacity, long long* total) {
    __shared__ long long part[1024];
    const int t = threadIdx.x, per = (n + blockDim.x - 1) / blockDim.x;
    const int lo = min(t * per, n), hi = min(lo + per, n);
    long long s = 0;
    for (int i = lo; i < hi; ++i) s += counts[i];
    part[t] = s;
    __syncthreads();
    if (t == 0) {
        long long run = 0;
        for (int k = 0; k < (int)blockDim.x; ++k) { const long long v = part[k]; part[k] = run; run += v; }
        *total = (base + run <= capacity) ? base + run : -(base + run);    // negative: the buffer is too small (nothing is written)
    }
    __syncthreads();
    long long run = part[t];
    for (int i = lo; i < hi; ++i) { const int v = counts[i]; counts[i] = (int)run; run += v; }
}

__device__ bool invert4d(const float* src, double* out) {
    double m[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { m[i][j] = (double)src[i * 4 + j]; m[i][4 + j] = (i == j) ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        double best = fabs(m[c][c]);
        for (int r = c + 1; r < 4; ++r)
            if (fabs(m[r][c]) > best) { best = fabs(m[r][c]); piv = r; }
        if (best == 0.0) return false;
        if (piv != c)
            for (int j = 0; j < 8; ++j) { double tmp = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = tmp; }
        const double inv = 1.0 / m[c][c];
        for (int j = 0; j < 8; ++j) m[c][j] *= inv;
        for (int r = 0; r < 4; ++r)
            if (r != c) { const double f = m[r][c]; for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j]; }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) out[i * 4 + j] = m[i][4 + j];
    return true;
}

__global__ void pc_write_kernel(const PcArgs a) {
    __shared__ float kinv[9], pose[12];
    __shared__ int wsum[kBlk / 32];
    const int b = blockIdx.y, i = blockIdx.x * kBlk + threadIdx.x;
    if (*a.total < 0) return;                    // capacity exceeded: reported through *total, nothing written
    if (threadIdx.x == 0) {
        double inv[16];
        const bool okk = invert4d(a.K + b * 16, inv);
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) kinv[r * 3 + c] = okk ? (float)inv[r * 4 + c] : __int_as_float(0x7fc00000);
        for (int k = 0; k < 12; ++k) pose[k] = a.pose[b * 16 + k];
    }
    float depth = 0.f;
    const bool ok = (i < a.H * a.W) && keep_vertex(a, b, i, depth);
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) wsum[warp] = __popc(bal);
    __syncthreads();
    if (!ok) return;
    int rank = __popc(bal & ((1u << lane) - 1));
    for (int w = 0; w < warp; ++w) rank += wsum[w];
    const long long at = a.base + a.counts[b * gridDim.x + blockIdx.x] + rank;
    const int y = i / a.W, x = i - y * a.W;
    const float fx = (float)x, fy = (float)y;
    // Backprojection (layers.py:56-58): inv(K)[:3,:3] . (x, y, 1) * depth; then pose . (X, 1)   (ply_utils.py:47-49)
    const float cx = (kinv[0] * fx + kinv[1] * fy + kinv[2]) * depth;
    const float cy = (kinv[3] * fx + kinv[4] * fy + kinv[5]) * depth;
    const float cz = (kinv[6] * fx + kinv[7] * fy + kinv[8]) * depth;
    float* o = a.out + at * 6;
    o[0] = pose[0] * cx + pose[1] * cy + pose[2] * cz + pose[3];
    o[1] = pose[4] * cx + pose[5] * cy + pose[6] * cz + pose[7];
    o[2] = pose[8] * cx + pose[9] * cy + pose[10] * cz + pose[11];
    const size_t hw = (size_t)a.H * a.W;
    const float* im = a.image + (size_t)b * 3 * hw + i;
    o[3] = (__ldg(im) + 0.5f) * 255.0f;           // ply_utils.py:37
    o[4] = (__ldg(im + hw) + 0.5f) * 255.0f;
    o[5] = (__ldg(im + 2 * hw) + 0.5f) * 255.0f;
}

}  // namespace

extern "C" int mr_pointcloud_keep_mask(const float* cv_mask, float* keep, int B, int H, int W, int mask_fill, float thresh,
                                       void* stream) {
    MR_REQUIRE(cv_mask && keep, "mr_pointcloud_keep_mask: null pointer");
    MR_REQUIRE(B >= 1 && B <= 65535 && H >= 1 && H <= 65535 && W >= 1 && mask_fill >= 0 && (mask_fill % 2) == 0,
               "mr_pointcloud_keep_mask: bad shape / mask_fill must be even (the reference's conv2d keeps the size only then)");
    keep_mask_kernel<<<dim3((W + 127) / 128, H, B), 128, 0, (cudaStream_t)stream>>>(cv_mask, keep, H, W, mask_fill / 2, thresh);
    MR_LAUNCH_CHECK("keep_mask_kernel");
    return MR_OK;
}

extern "C" long long mr_pointcloud_workspace(int B, int H, int W) {
    if (B < 1 || H < 1 || W < 1) return 0;
    return (long long)B * ((H * W + kBlk - 1) / kBlk) * (long long)sizeof(int);
}

extern "C" int mr_pointcloud_add(const float* inv_depth, const float* keyframe, const float* K, const float* pose,
                                 const float* const* keep_masks, int n_masks, int min_hits, int B, int H, int W, float min_d,
                                 float max_d, const int* roi, const float* dropout_rand, float dropout, float* vertices,
                                 long long capacity, long long n_before, long long* n_after, void* workspace,
                                 long long workspace_bytes, void* stream) {
    MR_REQUIRE(inv_depth && keyframe && K && pose && vertices && n_after && workspace, "mr_pointcloud_add: null pointer");
    MR_REQUIRE(B >= 1 && B <= 65535 && H >= 1 && W >= 1 && n_masks >= 0 && n_masks <= 16 && (n_masks == 0 || keep_masks != nullptr),
               "mr_pointcloud_add: bad shape or more than 16 masks");
    MR_REQUIRE(capacity >= 0 && n_before >= 0 && n_before <= capacity, "mr_pointcloud_add: bad buffer position");
    if (workspace_bytes < mr_pointcloud_workspace(B, H, W)) {
        mr::set_error("mr_pointcloud_add: workspace too small (%lld < %lld bytes)", workspace_bytes, mr_pointcloud_workspace(B, H, W));
        return MR_ENOMEM;
    }
    PcArgs a{};
    a.inv_depth = inv_depth; a.image = keyframe; a.K = K; a.pose = pose;
    for (int k = 0; k < n_masks; ++k) {
        MR_REQUIRE(keep_masks[k] != nullptr, "mr_pointcloud_add: null mask %d", k);
        a.keeps.p[k] = keep_masks[k];
    }
    a.n_masks = n_masks; a.min_hits = min_hits;
    a.rnd = dropout_rand; a.dropout = dropout; a.min_d = min_d; a.max_d = max_d;
    a.B = B; a.H = H; a.W = W;
    a.use_roi = roi != nullptr;
    if (roi) { a.r0 = roi[0]; a.r1 = roi[1]; a.c0 = roi[2]; a.c1 = roi[3]; }
    a.counts = static_cast<int*>(workspace);
    a.out = vertices; a.capacity = capacity; a.base = n_before; a.total = n_after;
    const int nb = (H * W + kBlk - 1) / kBlk;
    cudaStream_t st = (cudaStream_t)stream;
    pc_count_kernel<<<dim3(nb, B), kBlk, 0, st>>>(a);
    MR_LAUNCH_CHECK("pc_count_kernel");
    pc_scan_kernel<<<1, 1024, 0, st>>>(a.counts, nb * B, n_before, capacity, n_after);
    MR_LAUNCH_CHECK("pc_scan_kernel");
    pc_write_kernel<<<dim3(nb, B), kBlk, 0, st>>>(a);
    MR_LAUNCH_CHECK("pc_write_kernel");
    return MR_OK;
}
