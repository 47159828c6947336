// Probe: which fp32 no-swizzle TMA boxes load correctly (rank, inner box width, negative / out-of-range coordinates,
// tensor map passed directly or inside a struct array).  One configuration per process (a faulting one kills the context).
//   tma_probe rank bw rows x0 y0 mode
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
struct Maps { CUtensorMap m[8]; };
struct Args { float* out; int rank, bw, rows, x0, y0, c, f; };
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ void run(const CUtensorMap* map, const Args& a) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float* buf = reinterpret_cast<float*>(smem);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 65536);
    const uint32_t b = smem_u32(bar), d = smem_u32(buf);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(a.bw * a.rows * 4) : "memory");
        if (a.rank == 3)
            asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                         ::"r"(d), "l"(map), "r"(b), "r"(a.x0), "r"(a.y0), "r"(a.c) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(d), "l"(map), "r"(b), "r"(a.x0), "r"(a.y0) : "memory");
    }
    uint32_t done = 0;
    for (uint32_t spin = 0; !done && spin < (1u << 22); ++spin)
        asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(b) : "memory");
    for (int i = threadIdx.x; i < a.bw * a.rows; i += blockDim.x) a.out[i] = done ? buf[i] : -777.f;
}
__global__ void k_direct(const __grid_constant__ CUtensorMap map, const Args a) { run(&map, a); }
__global__ void k_struct(const Args a, const __grid_constant__ Maps maps) { run(&maps.m[a.f], a); }
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
    if (argc < 7) return 2;
    const int rank = atoi(argv[1]), bw = atoi(argv[2]), rows = atoi(argv[3]), x0 = atoi(argv[4]), y0 = atoi(argv[5]), mode = atoi(argv[6]);
    const int W = 332, H = 48, C = 6, c = 4;
    std::vector<float> h((size_t)C * H * W);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)(i % 9973) + 1.f;
    float *d, *out;
    cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 65536);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    Maps maps{};
    CUresult r;
    if (rank == 3) {
        cuuint64_t gd[3] = {W, H, C}; cuuint64_t gs[2] = {W * 4, (cuuint64_t)H * W * 4};
        cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)rows, 1}; cuuint32_t es[3] = {1, 1, 1};
        r = enc(&maps.m[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    } else {
        cuuint64_t gd[2] = {W, (cuuint64_t)H * C}; cuuint64_t gs[1] = {W * 4};
        cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)rows}; cuuint32_t es[2] = {1, 1};
        r = enc(&maps.m[1], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, gd, gs, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) { printf("rank %d bw %d rows %d x0 %d y0 %d mode %d: ENCODE FAILED %d\n", rank, bw, rows, x0, y0, mode, (int)r); return 0; }
    Args a{out, rank, bw, rows, x0, rank == 3 ? y0 : y0 + c * H, c, 1};
    cudaFuncSetAttribute(k_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    cudaFuncSetAttribute(k_struct, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    if (mode == 0) k_direct<<<1, 128, 70000>>>(maps.m[1], a); else k_struct<<<1, 128, 70000>>>(a, maps);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("rank %d bw %d rows %d x0 %d y0 %d mode %d: CUDA ERROR %s\n", rank, bw, rows, x0, y0, mode, cudaGetErrorString(e)); return 0; }
    std::vector<float> o((size_t)bw * rows);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int rr = 0; rr < rows; ++rr)
        for (int x = 0; x < bw; ++x) {
            const int gx = x0 + x, gy = y0 + rr;
            // rank 2 folds (channel, row) into one axis: rows below 0 / above H-1 of channel c are the neighbours' rows
            float exp = 0.f;
            if (rank == 3) { if (gx >= 0 && gx < W && gy >= 0 && gy < H) exp = h[((size_t)c * H + gy) * W + gx]; }
            else { const int gg = gy + c * H; if (gx >= 0 && gx < W && gg >= 0 && gg < H * C) exp = h[(size_t)gg * W + gx]; }
            if (o[(size_t)rr * bw + x] != exp) ++bad;
        }
    printf("rank %d bw %d rows %d x0 %d y0 %d mode %d: %s (%d mismatches, first %.1f)\n", rank, bw, rows, x0, y0, mode, bad ? "WRONG" : "ok", bad, o[0]);
    return 0;
}
