// phase2.cu -- micro-benchmark of the alternatives for conv_b's phase 2 (o[16] += sum_k A[k] * W[k][16], k = 0..39 per pass,
// six passes with different weight blocks) at the kernel's real shape: CTAs of 160 threads, 3 or 4 resident per SM.
//   A  current form      : per-node thread, weights through uniform 128-bit loads (LDCU.128), 2 packed FFMA2 per load
//   B  scalar constants  : per-node thread, scalar FFMA with the weight as a constant-bank operand (no load instruction)
//   C  5-node tile       : A exchanged through shared memory in rounds of 8 k-values; warp q < 4 owns output quad q, every
//                          lane owns the five nodes {lane, 32+lane, ..., 128+lane}: one LDCU.128 feeds 10 FFMA2
//   D  smem broadcast    : per-node thread, weights through LDS.128 with a warp-uniform address
// build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o phase2 tools/ubench/phase2.cu
// prints cycles per (CTA, pass) per SM and the implied time of conv_b's phase 2 at config 2 (17920 CTAs x 6 passes).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define THREADS 160
#define NPASS 6
struct Weights { float w[NPASS][40][16]; };      // 15 KB, like dagr_l1b_params_t.w

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

template <int PASS>
__device__ __forceinline__ void pass_A(const Weights &W, const float (&A)[40], float2 (&o)[8])
{
#pragma unroll
    for (int k = 0; k < 40; k++)
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            const float4 w = *reinterpret_cast<const float4 *>(&W.w[PASS][k][4 * c4]);
            o[2 * c4] = ffma2(make_float2(A[k], A[k]), make_float2(w.x, w.y), o[2 * c4]);
            o[2 * c4 + 1] = ffma2(make_float2(A[k], A[k]), make_float2(w.z, w.w), o[2 * c4 + 1]);
        }
}

template <int PASS>
__device__ __forceinline__ void pass_B(const Weights &W, const float (&A)[40], float (&o)[16])
{
#pragma unroll
    for (int k = 0; k < 40; k++)
#pragma unroll
        for (int c = 0; c < 16; c++) o[c] = fmaf(A[k], W.w[PASS][k][c], o[c]);
}

template <int PASS>
__device__ __forceinline__ void pass_D(const float4 *sw, const float (&A)[40], float2 (&o)[8])
{
#pragma unroll
    for (int k = 0; k < 40; k++)
#pragma unroll
        for (int c4 = 0; c4 < 4; c4++) {
            const float4 w = sw[(PASS * 40 + k) * 4 + c4];
            o[2 * c4] = ffma2(make_float2(A[k], A[k]), make_float2(w.x, w.y), o[2 * c4]);
            o[2 * c4 + 1] = ffma2(make_float2(A[k], A[k]), make_float2(w.z, w.w), o[2 * c4 + 1]);
        }
}

// C: rounds of 8 k-values through double-buffered shared memory; warp q (< 4) multiplies quad q for 5 nodes per lane
template <int PASS>
__device__ __forceinline__ void pass_C(const Weights &W, const float (&A)[40], float2 (&acc)[10], float *sA, int &rnd)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < 5; j++) {
        float *buf = sA + (rnd & 1) * 8 * THREADS;
        rnd++;
#pragma unroll
        for (int c = 0; c < 8; c++) buf[c * THREADS + threadIdx.x] = A[8 * j + c];
        __syncthreads();
        if (wid < 4) {
#pragma unroll
            for (int c = 0; c < 8; c++) {
                float a[5];
#pragma unroll
                for (int i = 0; i < 5; i++) a[i] = buf[c * THREADS + 32 * i + lane];
#define QUAD(Q)                                                                                              \
    {                                                                                                        \
        const float4 w = *reinterpret_cast<const float4 *>(&W.w[PASS][8 * j + c][4 * (Q)]);                  \
        _Pragma("unroll") for (int i = 0; i < 5; i++) {                                                      \
            acc[2 * i] = ffma2(make_float2(a[i], a[i]), make_float2(w.x, w.y), acc[2 * i]);                  \
            acc[2 * i + 1] = ffma2(make_float2(a[i], a[i]), make_float2(w.z, w.w), acc[2 * i + 1]);          \
        }                                                                                                    \
    }
                if (wid == 0) QUAD(0) else if (wid == 1) QUAD(1) else if (wid == 2) QUAD(2) else QUAD(3)
#undef QUAD
            }
        }
    }
}

template <int MODE>
__global__ void __launch_bounds__(THREADS, 4) k(const __grid_constant__ Weights W, const float *wsrc, float *out, long long *cycles, int iters)
{
    extern __shared__ __align__(16) float smem[];
    float A[40];
#pragma unroll
    for (int k2 = 0; k2 < 40; k2++) A[k2] = 1.0f + 1e-3f * (float)((threadIdx.x + k2) & 7);
    if (MODE == 3)
        for (int i = threadIdx.x; i < NPASS * 40 * 16; i += blockDim.x) smem[i] = wsrc[i];
    __syncthreads();
    float2 o[8];
    float os[16];
    float2 acc[10];
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; i++) os[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 10; i++) acc[i] = make_float2(0.f, 0.f);
    int rnd = 0;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) { pass_A<0>(W, A, o); pass_A<1>(W, A, o); pass_A<2>(W, A, o); pass_A<3>(W, A, o); pass_A<4>(W, A, o); pass_A<5>(W, A, o); }
        if (MODE == 1) { pass_B<0>(W, A, os); pass_B<1>(W, A, os); pass_B<2>(W, A, os); pass_B<3>(W, A, os); pass_B<4>(W, A, os); pass_B<5>(W, A, os); }
        if (MODE == 2) { pass_C<0>(W, A, acc, smem, rnd); pass_C<1>(W, A, acc, smem, rnd); pass_C<2>(W, A, acc, smem, rnd);
                         pass_C<3>(W, A, acc, smem, rnd); pass_C<4>(W, A, acc, smem, rnd); pass_C<5>(W, A, acc, smem, rnd); }
        if (MODE == 3) { const float4 *sw = reinterpret_cast<const float4 *>(smem);
                         pass_D<0>(sw, A, o); pass_D<1>(sw, A, o); pass_D<2>(sw, A, o); pass_D<3>(sw, A, o); pass_D<4>(sw, A, o); pass_D<5>(sw, A, o); }
        // loop-carried dependence so that nothing is hoisted: the next iteration's A depends on this one's result
        float s = 0.f;
        if (MODE == 1) { for (int i = 0; i < 16; i++) s += os[i]; }
        else if (MODE == 2) { for (int i = 0; i < 10; i++) s += acc[i].x + acc[i].y; }
        else { for (int i = 0; i < 8; i++) s += o[i].x + o[i].y; }
        A[it % 40] = A[it % 40] * 0.999f + 1e-9f * s;
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    float s = 0.f;
    for (int i = 0; i < 8; i++) s += o[i].x + o[i].y;
    for (int i = 0; i < 16; i++) s += os[i];
    for (int i = 0; i < 10; i++) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + A[3];
}

template <int MODE>
static void run(const char *name, int ctas_per_sm, size_t smem_bytes)
{
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    const int grid = nsm * ctas_per_sm, iters = 200;
    float *out, *wsrc; long long *cyc;
    cudaMalloc(&out, (size_t)grid * THREADS * sizeof(float)); cudaMalloc(&cyc, (size_t)grid * 8); cudaMalloc(&wsrc, sizeof(Weights));
    Weights *W = new Weights;
    for (int p = 0; p < NPASS; p++) for (int k2 = 0; k2 < 40; k2++) for (int c = 0; c < 16; c++) W->w[p][k2][c] = 1e-3f * (float)((p + k2 + c) % 11);
    cudaMemcpy(wsrc, W, sizeof(Weights), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    for (int rep = 0; rep < 2; rep++) k<MODE><<<grid, THREADS, smem_bytes>>>(*W, wsrc, out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long *h = new long long[grid];
    cudaMemcpy(h, cyc, (size_t)grid * 8, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (int i = 0; i < grid; i++) mx = h[i] > mx ? h[i] : mx;
    const double per_cta_pass = mx / (iters * NPASS) / ctas_per_sm;             // SM cycles per (CTA, pass) with ctas_per_sm resident
    const double ms = per_cta_pass * 17920.0 * 6.0 / 148.0 / 1.965e6;
    printf("%-28s %d CTAs/SM : %8.0f SM-cycles per (CTA, pass)  -> phase 2 of conv_b at config 2 ~ %.3f ms   (%s)\n", name, ctas_per_sm,
           per_cta_pass, ms, cudaGetErrorString(e));
    cudaFree(out); cudaFree(cyc); cudaFree(wsrc); delete W; delete[] h;
}

int main()
{
    for (int c = 3; c <= 4; c++) {
        run<0>("A LDCU.128 + 2 FFMA2", c, 45 * 1024);
        run<1>("B scalar FFMA, const operand", c, 45 * 1024);
        run<2>("C 5-node tile via smem", c, 55 * 1024);
        run<3>("D LDS.128 broadcast weights", c, 45 * 1024);
    }
    return 0;
}
