// Micro-benchmarks that decide the conv engine's (K2) design trade-offs on the actual part:
//   1. tcgen05.mma issue cost vs UMMA shape (M, N), number of independent accumulators and resident CTAs per SM
//      (is a small-N MMA chain bound by a fixed per-instruction cost?  do independent chains overlap?);
//   2. TMA box-load throughput / latency vs box shape (128 vs 64-byte rows, tap box vs halo box, OOB-filled halves),
//      boxes in flight per CTA and CTAs per SM (what is the L2 -> SM rate for the boxes the conv kernels use?).
// Build + run:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench_tc tools/ubench_tc.cu && /tmp/ubench_tc
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "W_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra W_DONE;\n\t"
        "bra W_LOOP;\n\t"
        "W_DONE:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t row_bytes) {
    // K-major swizzled tile: LBO unused (1), SBO = 8 rows, version 1, layout 2 (SWIZZLE_128B) / 4 (SWIZZLE_64B)
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((8 * row_bytes) >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)(row_bytes == 128 ? 2 : 4) << 61);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- 1. MMA issue cost ---------------------------------------------------------------------------------------------------
// One CTA = one issuing thread; `nacc` accumulators used round-robin (independent chains), `per_commit` MMAs between
// commits (the conv kernels commit once per K chunk = 2..4 MMAs).  Operands: zeroed swizzled tiles in shared memory.
__global__ void __launch_bounds__(128) mma_kernel(int M, int N, int nacc, int iters, int per_commit, uint32_t tmem_cols, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    for (uint32_t i = threadIdx.x * 16; i < (128 + 256) * 128; i += blockDim.x * 16)
        asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i), "r"(0));
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);   // f16 x f16 -> f32, K-major
        const uint64_t da = make_desc(base, 128), db = make_desc(base + 128 * 128, 128);
        uint32_t phase = 0;
        // warm-up
        umma_f16(tmem_d, da, db, idesc, 0);
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), phase); phase ^= 1;
        const long long t0 = clock64();
        int n = 0;
        for (int i = 0; i < iters; ++i)
            for (int a = 0; a < nacc; ++a) {
                umma_f16(tmem_d + (uint32_t)(a * N), da + (uint64_t)(2 * (n & 3)), db + (uint64_t)(2 * (n & 3)), idesc, 1);
                if (++n % per_commit == 0) umma_commit(smem_u32(&bar));   // arrive::one on a count-1 barrier: phases just flip
            }
        // drain: one more commit after everything, wait for the phase it completes
        // (earlier commits flipped the barrier an unknown number of times: re-init a fresh barrier instead)
        __shared__ __align__(8) uint64_t bar2;
        mbar_init(smem_u32(&bar2), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        umma_commit(smem_u32(&bar2));
        mbar_wait(smem_u32(&bar2), 0);
        const long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(tmem_cols));
    }
}

// ---- 2. TMA box loads ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
// One thread per CTA keeps `inflight` boxes outstanding (ring of buffers + mbarriers) and walks the tensor tile by tile,
// `taps` boxes per tile (shifted by one pixel each, like the conv kernel's filter taps).
__global__ void __launch_bounds__(32) tma_kernel(const __grid_constant__ CUtensorMap tm, int box_bytes, int inflight, int boxes,
                                                 int tiles_x, int tiles_y, int images, int step_x, int step_y, int taps,
                                                 long long* out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bars[16];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t stride = ((uint32_t)box_bytes + 1023u) & ~1023u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < inflight; ++s) mbar_init(smem_u32(&bars[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const int per_img = tiles_x * tiles_y, total_tiles = per_img * images;
        long long t0 = clock64();
        int issued = 0, done = 0;
        int tile = blockIdx.x, tap = 0;
        while (done < boxes) {
            while (issued < boxes && issued - done < inflight) {
                const int st = issued % inflight;
                const int b = (tile / per_img) % images, t = tile % per_img;
                const int ty = t / tiles_x, tx = t % tiles_x;
                mbar_expect_tx(smem_u32(&bars[st]), (uint32_t)box_bytes);
                tma_load_4d(base + st * stride, &tm, smem_u32(&bars[st]), 0, tx * step_x + (tap % 3) - 1, ty * step_y + (tap / 3) - 1, b);
                ++issued;
                if (++tap == taps) { tap = 0; tile += gridDim.x; if (tile >= total_tiles) tile -= total_tiles; }
            }
            const int st = done % inflight;
            mbar_wait(smem_u32(&bars[st]), (uint32_t)(done / inflight) & 1u);
            ++done;
        }
        out[blockIdx.x] = clock64() - t0;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    int sms = 0, khz = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    printf("SMs %d, clock %.0f MHz\n", sms, khz / 1e3);
    long long* out;
    CK(cudaMallocManaged(&out, sizeof(long long) * 4096));

    // ---------------- 1. MMA ----------------
    CK(cudaFuncSetAttribute(mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    printf("\n== tcgen05.mma kind::f16 K=16, cycles per MMA seen by one CTA (avg over CTAs) | per-SM MMA rate\n");
    printf("%4s %4s %5s %5s %7s | %10s %12s %10s\n", "M", "N", "nacc", "ctas", "commit", "clk/MMA", "clk/MMA/SM", "pipe util");
    const int Ms[] = {128, 64}, Ns[] = {16, 32, 48, 64, 96, 128, 256}, naccs[] = {1, 2, 4}, ctass[] = {1, 2, 4}, commits[] = {4, 1000000};
    for (int M : Ms)
        for (int N : Ns)
            for (int nacc : naccs)
                for (int ctas : ctass)
                    for (int pc : commits) {
                        if (nacc * N > 512 / ctas) continue;
                        uint32_t cols = 32;
                        while (cols < (uint32_t)(nacc * N)) cols <<= 1;
                        if (cols * ctas > 512) continue;
                        const int iters = 512;
                        const int grid = sms * ctas;
                        mma_kernel<<<grid, 128, 50 * 1024, 0>>>(M, N, nacc, iters, pc, cols, out);
                        CK(cudaDeviceSynchronize());
                        double avg = 0;
                        for (int i = 0; i < grid; ++i) avg += (double)out[i];
                        avg /= grid;
                        const double per = avg / (iters * nacc);
                        // dense f16 peak: 8192 FMA/clk/SM (4 tensor cores x 2048); one MMA = M*N*16 FMA
                        const double util = (double)M * N * 16 / 8192.0 / (per / ctas);
                        printf("%4d %4d %5d %5d %7s | %10.1f %12.1f %9.1f%%\n", M, N, nacc, ctas, pc > 1000 ? "end" : "4", per, per / ctas, 100 * util);
                    }

    // ---------------- 2. TMA ----------------
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &qres));
    EncodeTiledFn encode = reinterpret_cast<EncodeTiledFn>(fnp);
    const int B = 32, H = 256, W = 512;
    __half* x;
    CK(cudaMalloc(&x, (size_t)B * H * W * 64 * sizeof(__half)));
    CK(cudaMemset(x, 0, (size_t)B * H * W * 64 * sizeof(__half)));
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    struct Case { const char* name; int C, boxc, bw, bh, step_x, step_y, taps; CUtensorMapSwizzle sw; };
    const Case cases[] = {
        {"tap box 64ch/128B rows (C=64), 128 rows, 9 taps", 64, 64, 16, 8, 16, 8, 9, CU_TENSOR_MAP_SWIZZLE_128B},
        {"tap box 64ch box over C=32 (half OOB), 9 taps", 32, 64, 16, 8, 16, 8, 9, CU_TENSOR_MAP_SWIZZLE_128B},
        {"tap box 32ch/64B rows (C=32), 128 rows, 9 taps", 32, 32, 16, 8, 16, 8, 9, CU_TENSOR_MAP_SWIZZLE_64B},
        {"halo box 64ch box over C=32, 18x16 rows, 1/tile", 32, 64, 16, 18, 8, 16, 1, CU_TENSOR_MAP_SWIZZLE_128B},
        {"halo box 32ch/64B rows, 18x16 rows, 1/tile", 32, 32, 16, 18, 8, 16, 1, CU_TENSOR_MAP_SWIZZLE_64B},
        {"halo box 64ch/128B rows (C=64), 18x16, 1/tile", 64, 64, 16, 18, 8, 16, 1, CU_TENSOR_MAP_SWIZZLE_128B},
        {"halo box 32ch/64B rows, 10x24 rows, 1/tile", 32, 32, 24, 10, 16, 8, 1, CU_TENSOR_MAP_SWIZZLE_64B},
    };
    printf("\n== TMA 4-D box loads (half NHWC [32,256,512,C]); bytes = smem bytes filled\n");
    printf("%-52s %5s %8s | %9s %12s %12s\n", "box", "ctas", "inflight", "us/box", "GB/s total", "B/clk/SM");
    for (const Case& c : cases) {
        CUtensorMap tm;
        const cuuint64_t gdim[4] = {(cuuint64_t)c.C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
        const cuuint64_t gstr[3] = {(cuuint64_t)c.C * 2, (cuuint64_t)W * c.C * 2, (cuuint64_t)H * W * c.C * 2};
        const cuuint32_t box[4] = {(cuuint32_t)c.boxc, (cuuint32_t)c.bw, (cuuint32_t)c.bh, 1};
        const cuuint32_t estr[4] = {1, 1, 1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, x, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-52s encode failed (%d)\n", c.name, (int)r); continue; }
        const int box_bytes = c.boxc * 2 * c.bw * c.bh;
        const int tiles_x = W / c.step_x, tiles_y = H / c.step_y;
        for (int ctas : {1, 2, 4})
            for (int inflight : {1, 2, 4, 8}) {
                const size_t stride = ((size_t)box_bytes + 1023) & ~(size_t)1023;
                const size_t smem = stride * inflight + 1024;
                if (smem * ctas > 200 * 1024) continue;
                const int boxes = 2048;
                const int grid = sms * ctas;
                tma_kernel<<<grid, 32, smem, 0>>>(tm, box_bytes, inflight, boxes, tiles_x, tiles_y, B, c.step_x, c.step_y, c.taps, out);
                CK(cudaDeviceSynchronize());
                double mx = 0;
                for (int i = 0; i < grid; ++i) mx = out[i] > mx ? (double)out[i] : mx;
                const double secs = mx / (khz * 1e3);
                const double total_bytes = (double)boxes * box_bytes * grid;
                printf("%-52s %5d %8d | %9.3f %12.1f %12.1f\n", c.name, ctas, inflight, secs / boxes * 1e6, total_bytes / secs / 1e9,
                       (double)boxes * box_bytes * ctas / mx);
            }
    }
    CK(cudaFree(x));
    CK(cudaFree(out));
    return 0;
}
