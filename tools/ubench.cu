// Micro-benchmarks that decide K1's design trade-offs on the actual part: issue rates of FFMA vs packed FFMA2,
// FADD2/FMUL2, MUFU, SHFL, LDS.32/64/128 and L1-hit LDG per SM.   nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void k(float* out, const float* in, int n) {
    __shared__ float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = in[i % n];
    __syncthreads();
    float a0 = in[threadIdx.x % n], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float b = 1.0001f, c = 0.0001f;
    float2 p0 = make_float2(a0, a1), p1 = make_float2(a2, a3), p2 = make_float2(a4, a5), p3 = make_float2(a6, a7);
    float2 pb = make_float2(b, b), pc = make_float2(c, c);
    int idx = threadIdx.x;
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0) {  // 8 scalar FFMA
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fmaf(a4, b, c); a5 = fmaf(a5, b, c); a6 = fmaf(a6, b, c); a7 = fmaf(a7, b, c);
        } else if (MODE == 1) {  // 4 FFMA2 (= 8 fma)
            p0 = __ffma2_rn(p0, pb, pc); p1 = __ffma2_rn(p1, pb, pc); p2 = __ffma2_rn(p2, pb, pc); p3 = __ffma2_rn(p3, pb, pc);
        } else if (MODE == 2) {  // 8 MUFU.RCP
            a0 = __fdividef(1.f, a0); a1 = __fdividef(1.f, a1); a2 = __fdividef(1.f, a2); a3 = __fdividef(1.f, a3);
            a4 = __fdividef(1.f, a4); a5 = __fdividef(1.f, a5); a6 = __fdividef(1.f, a6); a7 = __fdividef(1.f, a7);
        } else if (MODE == 3) {  // 8 SHFL
            a0 = __shfl_up_sync(~0u, a0, 1); a1 = __shfl_up_sync(~0u, a1, 1); a2 = __shfl_up_sync(~0u, a2, 1); a3 = __shfl_up_sync(~0u, a3, 1);
            a4 = __shfl_up_sync(~0u, a4, 1); a5 = __shfl_up_sync(~0u, a5, 1); a6 = __shfl_up_sync(~0u, a6, 1); a7 = __shfl_up_sync(~0u, a7, 1);
        } else if (MODE == 4) {  // 8 LDS.32 conflict-free, dependent address
            a0 += sm[(idx) & 4095]; a1 += sm[(idx + 32) & 4095]; a2 += sm[(idx + 64) & 4095]; a3 += sm[(idx + 96) & 4095];
            a4 += sm[(idx + 128) & 4095]; a5 += sm[(idx + 160) & 4095]; a6 += sm[(idx + 192) & 4095]; a7 += sm[(idx + 224) & 4095];
            idx += 256;
        } else if (MODE == 5) {  // 4 LDS.128
            float4 v0 = *reinterpret_cast<float4*>(&sm[(4 * idx) & 4095]);
            float4 v1 = *reinterpret_cast<float4*>(&sm[(4 * idx + 512) & 4095]);
            float4 v2 = *reinterpret_cast<float4*>(&sm[(4 * idx + 1024) & 4095]);
            float4 v3 = *reinterpret_cast<float4*>(&sm[(4 * idx + 1536) & 4095]);
            a0 += v0.x + v0.y + v0.z + v0.w; a1 += v1.x + v1.y + v1.z + v1.w; a2 += v2.x + v2.y + v2.z + v2.w; a3 += v3.x + v3.y + v3.z + v3.w;
            idx += 17;
        } else if (MODE == 6) {  // 8 LDG.32 L1-hit, coalesced
            a0 += __ldg(in + ((idx) & 1023)); a1 += __ldg(in + ((idx + 32) & 1023)); a2 += __ldg(in + ((idx + 64) & 1023)); a3 += __ldg(in + ((idx + 96) & 1023));
            a4 += __ldg(in + ((idx + 128) & 1023)); a5 += __ldg(in + ((idx + 160) & 1023)); a6 += __ldg(in + ((idx + 192) & 1023)); a7 += __ldg(in + ((idx + 224) & 1023));
            idx += 256;
        } else if (MODE == 7) {  // 4 FADD2 + 4 FMUL2
            p0 = __fadd2_rn(p0, pc); p1 = __fmul2_rn(p1, pb); p2 = __fadd2_rn(p2, pc); p3 = __fmul2_rn(p3, pb);
            p0 = __fmul2_rn(p0, pb); p1 = __fadd2_rn(p1, pc); p2 = __fmul2_rn(p2, pb); p3 = __fadd2_rn(p3, pc);
        } else if (MODE == 8) {  // 8 FMNMX (alu pipe)
            a0 = fminf(a0, b); a1 = fmaxf(a1, c); a2 = fminf(a2, b); a3 = fmaxf(a3, c);
            a4 = fminf(a4, a0); a5 = fmaxf(a5, a1); a6 = fminf(a6, a2); a7 = fmaxf(a7, a3);
        } else if (MODE == 9) {  // 4 FFMA + 4 FMNMX mix (do the pipes overlap?)
            a0 = fmaf(a0, b, c); a1 = fmaf(a1, b, c); a2 = fmaf(a2, b, c); a3 = fmaf(a3, b, c);
            a4 = fminf(a4, b); a5 = fmaxf(a5, c); a6 = fminf(a6, b); a7 = fmaxf(a7, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + p0.x + p0.y + p1.x + p1.y + p2.x + p2.y + p3.x + p3.y;
}

template <int MODE>
void run(const char* name, int ops_per_iter, float* out, const float* in) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    dim3 grid(sms * 2), block(512);
    k<MODE><<<grid, block>>>(out, in, 1024);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) k<MODE><<<grid, block>>>(out, in, 1024);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    double lane_ops = (double)grid.x * block.x * ITERS * ops_per_iter;
    double per_sm_per_ns = lane_ops / (ms * 1e6) / sms;
    printf("%-28s %8.3f ms  %7.1f lane-instr/ns/SM  (= %.1f per clk at max clock %.0f MHz)\n", name, ms, per_sm_per_ns,
           per_sm_per_ns / (clk * 1e-6), clk * 1e-3);
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 4096 * 4); cudaMalloc(&out, 1 << 22);
    float h[4096]; for (int i = 0; i < 4096; ++i) h[i] = 1.0f + i * 1e-4f;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("FFMA x8", 8, out, in);
    run<1>("FFMA2 x4 (instr)", 4, out, in);
    run<7>("FADD2/FMUL2 x8 (instr)", 8, out, in);
    run<2>("MUFU.RCP x8", 8, out, in);
    run<3>("SHFL x8", 8, out, in);
    run<4>("LDS.32 x8", 8, out, in);
    run<5>("LDS.128 x4 (instr)", 4, out, in);
    run<6>("LDG.32 L1-hit x8", 8, out, in);
    run<8>("FMNMX x8", 8, out, in);
    run<9>("FFMA x4 + FMNMX x4", 8, out, in);
    return 0;
}
