// pipes.cu -- micro-benchmarks of the per-SM issue rates this round's kernel decisions leaned on (not part of the library;
// B200 results: profiles/r01_ubench_pipes.txt):
//   * FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate per SMSP
//   * IADD3/LOP3 (integer pipe) issue rate per SMSP
//   * LDS.32 / LDS.128 with a warp-uniform address (broadcast) and with conflict-free per-lane addresses
//   * LDCU.128-fed FFMA2 (weights from a __grid_constant__ parameter at compile-time offsets), 2 FFMA2 per load
// build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o pipes tools/ubench/pipes.cu
// run  :  ./pipes        (prints warp-instructions per cycle per SM for 1..16 resident warps per SMSP)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096

struct Weights { float w[256][4]; };

template <int MODE>
__global__ void k(const __grid_constant__ Weights W, float *out, long long *cycles, int dummy)
{
    __shared__ __align__(16) float sm[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    float2 a0 = make_float2(1.f, 2.f), a1 = make_float2(3.f, 4.f), a2 = make_float2(5.f, 6.f), a3 = make_float2(7.f, 8.f);
    const float2 b = make_float2(1.0001f, 0.9999f);
    int x0 = threadIdx.x, x1 = dummy, x2 = 3, x3 = 5;
    const int lane = threadIdx.x & 31;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {            // 8 independent scalar FFMA
            a0.x = fmaf(a0.x, b.x, b.y); a0.y = fmaf(a0.y, b.x, b.y); a1.x = fmaf(a1.x, b.x, b.y); a1.y = fmaf(a1.y, b.x, b.y);
            a2.x = fmaf(a2.x, b.x, b.y); a2.y = fmaf(a2.y, b.x, b.y); a3.x = fmaf(a3.x, b.x, b.y); a3.y = fmaf(a3.y, b.x, b.y);
        } else if (MODE == 1) {     // 8 independent packed FFMA2 (two per accumulator pair and iteration)
            a0 = __ffma2_rn(a0, b, b); a1 = __ffma2_rn(a1, b, b); a2 = __ffma2_rn(a2, b, b); a3 = __ffma2_rn(a3, b, b);
            a0 = __ffma2_rn(a0, b, b); a1 = __ffma2_rn(a1, b, b); a2 = __ffma2_rn(a2, b, b); a3 = __ffma2_rn(a3, b, b);
        } else if (MODE == 2) {     // 8 integer-pipe ops
            x0 = (x0 + x1) ^ x2; x1 = (x1 + x2) ^ x3; x2 = (x2 + x3) ^ x0; x3 = (x3 + x0) ^ x1;
        } else if (MODE == 3) {     // 4 LDS.128, warp-uniform address (broadcast)
            const float4 *p = reinterpret_cast<const float4 *>(sm) + ((it + x1) & 255);
            const float4 v0 = p[0], v1 = p[1], v2 = p[2], v3 = p[3];
            a0.x += v0.x; a1.x += v1.y; a2.x += v2.z; a3.x += v3.w;
        } else if (MODE == 4) {     // 4 LDS.128, consecutive 16-byte chunks per lane (conflict free)
            const float4 *p = reinterpret_cast<const float4 *>(sm) + ((it + x1) & 127) + lane;
            const float4 v0 = p[0], v1 = p[32], v2 = p[64], v3 = p[96];
            a0.x += v0.x; a1.x += v1.y; a2.x += v2.z; a3.x += v3.w;
        } else if (MODE == 5) {     // 4 LDS.32 conflict free
            const float *p = sm + ((it + x1) & 1023) + lane;
            a0.x += p[0]; a1.x += p[32]; a2.x += p[64]; a3.x += p[96];
        } else {                    // LDCU.128-fed FFMA2: 16 uniform loads, 32 FFMA2 per iteration
#pragma unroll
            for (int q = 0; q < 16; q++) {
                const float4 w = *reinterpret_cast<const float4 *>(&W.w[q][0]);
                a0 = __ffma2_rn(make_float2(a2.x, a2.x), make_float2(w.x, w.y), a0);
                a1 = __ffma2_rn(make_float2(a3.x, a3.x), make_float2(w.z, w.w), a1);
            }
        }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0.x + a0.y + a1.x + a1.y + a2.x + a2.y + a3.x + a3.y + (float)(x0 + x1 + x2 + x3);
}

template <int MODE>
static void run(const char *name, int instr_per_iter)
{
    float *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(float)); cudaMalloc(&cyc, 8);
    Weights W = {};
    for (int warps = 4; warps <= 32; warps *= 2) {                      // warps per SM (4 SMSPs)
        k<MODE><<<148, warps * 32>>>(W, out, cyc, 0);
        k<MODE><<<148, warps * 32>>>(W, out, cyc, 0);
        long long c = 0;
        cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-28s warps/SM %2d : %.3f warp-instr / cycle / SM\n", name, warps, (double)warps * ITERS * instr_per_iter / (double)c);
    }
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<0>("FFMA (scalar)", 8);
    run<1>("FFMA2 (packed)", 8);
    run<2>("IADD3+LOP3 (8 int ops)", 8);
    run<3>("LDS.128 broadcast", 4);
    run<4>("LDS.128 conflict-free", 4);
    run<5>("LDS.32 conflict-free", 4);
    run<6>("LDCU.128 + 2 FFMA2", 48);
    return 0;
}
