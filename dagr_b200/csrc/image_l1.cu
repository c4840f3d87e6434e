// image_l1.cu -- event-level image fusion (use_image): bilinear sampling of ResNet feature maps at the
// events (net.py:15-17,193-221), conv_block1.conv_block1 on 1+16+2 input channels, and the per-voxel max
// of the 64-channel samples that are concatenated before pool1 (net.py:128-131).  sm_100a.
#include "common.cuh"

// grid_sample(align_corners=True) of one (x, y, batch) position: weights and corner offsets, mirroring
// _sample_features (net.py:207-221): normalise with the event resolution, unnormalise with the map size.
struct Bilin {
    int x0, y0, z0;
    float tx, ty, tz;
};
__device__ __forceinline__ Bilin bilin_setup(float posx, float posy, int b, float width, float height, int Bi, int h, int w)
{
    const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fmul_rn(posx, width)), width - 1.f), 1.f);
    const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fmul_rn(posy, height)), height - 1.f), 1.f);
    const float bs = (float)(Bi > 1 ? Bi : 2);
    const float gz = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, (float)b), bs - 1.f), 1.f);
    const float ix = __fmul_rn(__fmul_rn(__fadd_rn(gx, 1.f), 0.5f), (float)(w - 1));
    const float iy = __fmul_rn(__fmul_rn(__fadd_rn(gy, 1.f), 0.5f), (float)(h - 1));
    const float iz = __fmul_rn(__fmul_rn(__fadd_rn(gz, 1.f), 0.5f), (float)(Bi - 1));
    Bilin r;
    const float x0f = floorf(ix), y0f = floorf(iy), z0f = floorf(iz);
    r.x0 = (int)x0f; r.y0 = (int)y0f; r.z0 = (int)z0f;
    r.tx = ix - x0f; r.ty = iy - y0f; r.tz = iz - z0f;
    return r;
}
__device__ __forceinline__ float bilin_sample(const float *__restrict__ img, int Bi, int C, int h, int w, int c, const Bilin &q)
{
    float acc = 0.f;
#pragma unroll
    for (int dz = 0; dz < 2; dz++) {
        const int z = q.z0 + dz;
        const float wz = dz ? q.tz : 1.f - q.tz;
        if (z < 0 || z >= Bi) continue;
#pragma unroll
        for (int dy = 0; dy < 2; dy++) {
            const int y = q.y0 + dy;
            const float wy = dy ? q.ty : 1.f - q.ty;
            if (y < 0 || y >= h) continue;
#pragma unroll
            for (int dx = 0; dx < 2; dx++) {
                const int x = q.x0 + dx;
                const float wx = dx ? q.tx : 1.f - q.tx;
                if (x < 0 || x >= w) continue;
                acc += __ldg(img + (((int64_t)z * C + c) * h + y) * w + x) * (wx * wy * wz);
            }
        }
    }
    return acc;
}

// x0[chunk][p][8]: [pol, f0..f15, x/W, y/H, 0 x5]
__global__ void k_l1_x0_image(const dagr_geom_t g, int64_t N, const uint32_t *__restrict__ xyb, const float *__restrict__ feat_s,
                              const float *__restrict__ img0, int h, int w, float *__restrict__ x0)
{
    // 4 threads per node: thread q computes image channels 4q..4q+3
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t p = t >> 2;
    const int q = (int)(t & 3);
    if (p >= N) return;
    const uint32_t wd = xyb[p];
    const int x = wd & 0xfff, y = (wd >> 12) & 0xfff, b = wd >> 24;
    const float px = g.posx0[x], py = g.posy0[y];
    const Bilin bl = bilin_setup(px, py, b, (float)g.W, (float)g.H, g.B, h, w);
    float f[4];
#pragma unroll
    for (int k = 0; k < 4; k++) f[k] = bilin_sample(img0, g.B, 16, h, w, 4 * q + k, bl);
    // channel index in the 24-wide padded row: 0 = polarity, 1..16 = image, 17 = x, 18 = y
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int ch = 1 + 4 * q + k;
        x0[((int64_t)(ch >> 3) * N + p) * 8 + (ch & 7)] = f[k];
    }
    if (q == 0) {
        x0[p * 8] = feat_s[p];
        x0[((int64_t)2 * N + p) * 8 + 1] = px;
        x0[((int64_t)2 * N + p) * 8 + 2] = py;
#pragma unroll
        for (int k = 3; k < 8; k++) x0[((int64_t)2 * N + p) * 8 + k] = 0.f;
    }
}

extern "C" int dagr_l1_x0_image(const dagr_geom_t *g, int64_t N, const uint32_t *xyb, const float *feat_s, const float *img0,
                                int h, int w, float *x0, void *stream)
{
    if (N <= 0) return DAGR_OK;
    k_l1_x0_image<<<dagr_div_up(4 * N, 256), 256, 0, (cudaStream_t)stream>>>(*g, N, xyb, feat_s, img0, h, w, x0);
    DAGR_CHECK_LAUNCH();
    return DAGR_OK;
}

// ------------------------------------------------------------------------------------------------
// conv_block1.conv_block1 with 19 (padded 24) input channels: thread per node, 3 channel chunks x 3 slot
// groups = 9 passes over the ELL row (global gathers: this variant is bounded by the ResNet trunk anyway).
// ------------------------------------------------------------------------------------------------
#define CI_THREADS 128
#define CI_G 5

__global__ void __launch_bounds__(CI_THREADS)
k_l1_conv_a_image(const dagr_geom_t g, int64_t N, const float *__restrict__ x0, const int32_t *__restrict__ nbr,
                  const uint16_t *__restrict__ off, const float *__restrict__ tab, const dagr_l1img_params_t *__restrict__ P,
                  float *__restrict__ xa, float *__restrict__ skipv)
{
    extern __shared__ __align__(16) float s_tab[];                      // [3][ncell][8]
    for (int i = threadIdx.x; i < g.ncell * DAGR_TABW; i += blockDim.x) {
        const int c = i / DAGR_TABW, u = i % DAGR_TABW;
        if (u < DAGR_KU) s_tab[((size_t)(u / CI_G) * g.ncell + c) * 8 + (u % CI_G)] = tab[i];
    }
    __syncthreads();
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= N) return;
    const int n = nbr[(int64_t)(DAGR_ELL - 1) * N + p];
    float o[16], sk[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { o[k] = 0.f; sk[k] = 0.f; }
#pragma unroll 1
    for (int ch = 0; ch < 3; ch++) {
        {   // root + skip on this chunk of x_i
            const float4 *src = reinterpret_cast<const float4 *>(x0 + ((int64_t)ch * N + p) * 8);
            const float4 t0 = src[0], t1 = src[1];
            const float v[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
            for (int k = 0; k < 8; k++)
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    o[c] = fmaf(v[k], P->root[8 * ch + k][c], o[c]);
                    sk[c] = fmaf(v[k], P->skip[8 * ch + k][c], sk[c]);
                }
        }
#pragma unroll 1
        for (int grp = 0; grp < 3; grp++) {
            float A[CI_G][8];
#pragma unroll
            for (int u = 0; u < CI_G; u++)
#pragma unroll
                for (int k = 0; k < 8; k++) A[u][k] = 0.f;
            const float *tabg = s_tab + (size_t)grp * g.ncell * 8;
            for (int q = -1; q < n; q++) {
                const int row = q < 0 ? (int)p : nbr[(int64_t)q * N + p];
                const int c = q < 0 ? 0 : (int)off[(int64_t)q * N + p];
                const float4 *src = reinterpret_cast<const float4 *>(x0 + ((int64_t)ch * N + row) * 8);
                const float4 t0 = __ldg(src), t1 = __ldg(src + 1);
                const float e[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
                const float4 w0 = *reinterpret_cast<const float4 *>(tabg + c * 8);
                const float t[CI_G] = {w0.x, w0.y, w0.z, w0.w, tabg[c * 8 + 4]};
#pragma unroll
                for (int u = 0; u < CI_G; u++)
#pragma unroll
                    for (int k = 0; k < 8; k++) A[u][k] = fmaf(t[u], e[k], A[u][k]);
            }
#pragma unroll
            for (int u = 0; u < CI_G; u++)
#pragma unroll
                for (int k = 0; k < 8; k++)
#pragma unroll
                    for (int c = 0; c < 16; c++) o[c] = fmaf(A[u][k], P->w[grp * CI_G + u][8 * ch + k][c], o[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 16; c++) {
        const float r = fmaf(o[c], P->scale[c], P->shift[c]);
        o[c] = P->relu ? fmaxf(r, 0.f) : r;
        sk[c] = fmaf(sk[c], P->sscale[c], P->sshift[c]);
    }
    const int sw = XA_SWZ(p);
    float4 *dst = reinterpret_cast<float4 *>(xa + p * 8);
    dst[sw] = make_float4(o[0], o[1], o[2], o[3]);
    dst[sw ^ 1] = make_float4(o[4], o[5], o[6], o[7]);
    dst = reinterpret_cast<float4 *>(xa + (N + p) * 8);
    dst[sw] = make_float4(o[8], o[9], o[10], o[11]);
    dst[sw ^ 1] = make_float4(o[12], o[13], o[14], o[15]);
    float4 *sd = reinterpret_cast<float4 *>(skipv + p * 16);
    sd[0] = make_float4(sk[0], sk[1], sk[2], sk[3]);
    sd[1] = make_float4(sk[4], sk[5], sk[6], sk[7]);
    sd[2] = make_float4(sk[8], sk[9], sk[10], sk[11]);
    sd[3] = make_float4(sk[12], sk[13], sk[14], sk[15]);
}

extern "C" int dagr_l1_conv_a_image(const dagr_geom_t *g, int64_t N, const float *x0, const int32_t *nbr, const uint16_t *off,
                                    const float *tab, const dagr_l1img_params_t *p_dev, float *xa, float *skipv, void *stream)
{
    DAGR_CHECK_ARG(g && p_dev, "null argument");
    if (N <= 0) return DAGR_OK;
    const size_t smem = (size_t)3 * g->ncell * 8 * sizeof(float);
    DAGR_CUDA(cudaFuncSetAttribute(k_l1_conv_a_image, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_l1_conv_a_image<<<dagr_div_up(N, CI_THREADS), CI_THREADS, smem, (cudaStream_t)stream>>>(*g, N, x0, nbr, off, tab, p_dev, xa, skipv);
    DAGR_CHECK_LAUNCH();
    return DAGR_OK;
}

// ------------------------------------------------------------------------------------------------
// per-voxel max of image features sampled at the voxel's events: one CTA per voxel, lanes over channels
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_voxel_sample_max(const dagr_geom_t g, const int32_t *__restrict__ start, const uint32_t *__restrict__ xyb,
                   const float *__restrict__ img, int C, int h, int w, float *__restrict__ xg, int ldx, int c0)
{
    __shared__ float s_m[4][128];
    const int cell = blockIdx.x;
    const int p0 = start[(int64_t)cell * g.CP], p1 = start[(int64_t)(cell + 1) * g.CP];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int cb = 0; cb < C; cb += 128) {
        float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int p = p0 + wid; p < p1; p += 4) {
            const uint32_t wd = xyb[p];
            const int x = wd & 0xfff, y = (wd >> 12) & 0xfff, b = wd >> 24;
            const Bilin bl = bilin_setup(g.posx0[x], g.posy0[y], b, (float)g.W, (float)g.H, g.B, h, w);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c = cb + lane + 32 * k;
                if (c < C) m[k] = fmaxf(m[k], bilin_sample(img, g.B, C, h, w, c, bl));
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) s_m[wid][lane + 32 * k] = m[k];
        __syncthreads();
        const int c = cb + threadIdx.x;
        if (c < C) {
            float v = fmaxf(fmaxf(s_m[0][threadIdx.x], s_m[1][threadIdx.x]), fmaxf(s_m[2][threadIdx.x], s_m[3][threadIdx.x]));
            xg[(int64_t)cell * ldx + c0 + c] = (p1 > p0) ? v : 0.f;
        }
        __syncthreads();
    }
}

extern "C" int dagr_voxel_sample_max(const dagr_geom_t *g, int64_t N, const int32_t *start, const uint32_t *xyb, const float *img,
                                     int C, int h, int w, float *xg, int ldx, int c0, void *stream)
{
    (void)N;
    const int cells = g->B * g->ny1 * g->nx1;
    k_voxel_sample_max<<<cells, 128, 0, (cudaStream_t)stream>>>(*g, start, xyb, img, C, h, w, xg, ldx, c0);
    DAGR_CHECK_LAUNCH();
    return DAGR_OK;
}
