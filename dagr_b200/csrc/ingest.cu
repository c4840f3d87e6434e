// ingest.cu -- event ingest on the device: the steps between the raw DSEC stream and the graph
// builder, which the reference runs on the CPU (numba / numpy) per sample:
//   * 2x event down-sampler            scripts/downsample_events.py:91-124
//   * window slice, crop, relative t   src/dagr/data/dsec_data.py:141-147,177-179
//   * int16/int32 casts, fp32 normalise, denormalise   data/utils.py:6-20, utils/buffers.py:33-44, ev_tgn.py:11-16
// producing directly the int32 (batch, pos) + polarity arrays that dagr_graph_sort consumes.
#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// down-sampler.  The reference walks the events of a chunk in time order and keeps one signed fp32 accumulator per
// output pixel: += p/(fx*fy); when |acc| >= 1 the event passes and acc -= p.  Events of different output pixels never
// interact, so: bin the events by output pixel (histogram -> scan -> scatter -> in-bin rank by arrival index = a stable
// counting sort), then one thread per pixel replays ITS events in time order with the reference's arithmetic.
// ------------------------------------------------------------------------------------------------
__global__ void k_ds_hist(const uint16_t *__restrict__ x, const uint16_t *__restrict__ y, int64_t N, int fx, int fy, int ow,
                          int oh, int32_t *__restrict__ cell, int32_t *__restrict__ count)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int xl = min((int)x[i] / fx, ow - 1), yl = min((int)y[i] / fy, oh - 1);   // the reference would index out of bounds
    const int c = yl * ow + xl;
    cell[i] = c;
    atomicAdd(&count[c], 1);
}

__global__ void k_ds_scatter(const int32_t *__restrict__ cell, int64_t N, const int32_t *__restrict__ start,
                             int32_t *__restrict__ count, int32_t *__restrict__ tmp)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int c = cell[i];
    const int slot = atomicSub(&count[c], 1) - 1;                        // count returns to zero
    tmp[start[c] + slot] = (int)i;
}

__global__ void k_ds_rank(const int32_t *__restrict__ cell, const int32_t *__restrict__ tmp, int64_t N,
                          const int32_t *__restrict__ start, int32_t *__restrict__ sorted)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= N) return;
    const int i = tmp[q], c = cell[i];
    const int s = start[c], e = start[c + 1];
    int r = 0;
    for (int k = s; k < e; k++) r += tmp[k] < i;
    sorted[s + r] = i;
}

__global__ void k_ds_walk(const int8_t *__restrict__ p, const int32_t *__restrict__ start, const int32_t *__restrict__ sorted,
                          int cells, double denom, float *__restrict__ change_map, uint8_t *__restrict__ mask)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const int s = start[c], e = start[c + 1];
    if (s == e) return;
    float acc = change_map[c];
    for (int k = s; k < e; k++) {
        const int i = sorted[k];
        const float pi = (float)p[i];
        // numba: float32 array element += float64 expression  ->  sum in float64, rounded once to float32
        acc = (float)((double)acc + (double)pi * 1.0 / denom);
        const bool pass = fabsf(acc) >= 1.f;
        if (pass) acc = __fsub_rn(acc, pi);
        mask[i] = pass ? 1 : 0;
    }
    change_map[c] = acc;
}

extern "C" int dagr_downsample_events(const uint16_t *x, const uint16_t *y, const int8_t *p, int64_t N, int fx, int fy,
                                      int out_w, int out_h, float *change_map, int32_t *cell, int32_t *tmp, int32_t *sorted,
                                      int32_t *count, int32_t *start, int32_t *blocksums, uint8_t *mask, void *stream)
{
    DAGR_CHECK_ARG(fx >= 1 && fy >= 1 && out_w >= 1 && out_h >= 1, "bad down-sampling geometry");
    DAGR_CHECK_ARG(N < (1ll << 31), "N must fit int32");
    if (N <= 0) return DAGR_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int cells = out_w * out_h;
    k_ds_hist<<<dagr_div_up(N, 256), 256, 0, st>>>(x, y, N, fx, fy, out_w, out_h, cell, count);
    scan_exclusive(count, start, (int64_t)cells, blocksums, st);          // start[cells] = N
    k_ds_scatter<<<dagr_div_up(N, 256), 256, 0, st>>>(cell, N, start, count, tmp);
    k_ds_rank<<<dagr_div_up(N, 256), 256, 0, st>>>(cell, tmp, N, start, sorted);
    k_ds_walk<<<dagr_div_up(cells, 128), 128, 0, st>>>(p, start, sorted, cells, (double)(fx * fy), change_map, mask);
    DAGR_CHECK_LAUNCH();
    return DAGR_OK;
}

// ------------------------------------------------------------------------------------------------
// stable compaction of the events a mask keeps, with the output coordinates of the down-sampler
// (`(x / fx).astype("uint16")`, downsample_events.py:103-104; fx = fy = 1 leaves them unchanged)
// ------------------------------------------------------------------------------------------------
__global__ void k_mask_to_int(const uint8_t *__restrict__ mask, int64_t N, int32_t *__restrict__ flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) flag[i] = mask[i] ? 1 : 0;
}

__global__ void k_compact(const int32_t *__restrict__ flag, const int32_t *__restrict__ pos, int64_t N,
                          const uint16_t *__restrict__ x, const uint16_t *__restrict__ y, const int64_t *__restrict__ t,
                          const int8_t *__restrict__ p, int fx, int fy, uint16_t *__restrict__ xo, uint16_t *__restrict__ yo,
                          int64_t *__restrict__ to, int8_t *__restrict__ po)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || !flag[i]) return;
    const int j = pos[i];
    xo[j] = (uint16_t)(x[i] / fx); yo[j] = (uint16_t)(y[i] / fy); to[j] = t[i]; po[j] = p[i];
}

extern "C" int dagr_compact_events(const uint8_t *mask, int64_t N, const uint16_t *x, const uint16_t *y, const int64_t *t,
                                   const int8_t *p, int fx, int fy, int32_t *flag, int32_t *pos, int32_t *blocksums,
                                   uint16_t *xo, uint16_t *yo, int64_t *to, int8_t *po, int32_t *n_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    if (N <= 0) { DAGR_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), st)); return DAGR_OK; }
    k_mask_to_int<<<dagr_div_up(N, 256), 256, 0, st>>>(mask, N, flag);
    scan_exclusive(flag, pos, N, blocksums, st);
    k_compact<<<dagr_div_up(N, 256), 256, 0, st>>>(flag, pos, N, x, y, t, p, fx, fy, xo, yo, to, po);
    DAGR_CUDA(cudaMemcpyAsync(n_out, pos + N, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));   // total kept
    DAGR_CHECK_LAUNCH();
    return DAGR_OK;
}

// ------------------------------------------------------------------------------------------------
// window slice + crop + relative time + polarity + normalise/denormalise, fused: raw events of ONE sample ->
// batch i32[M], pos i32[M,3], polarity f32[M]  (the arrays DAGR.forward derives from a formatted Batch)
// ------------------------------------------------------------------------------------------------
__global__ void k_ing_flag(const uint16_t *__restrict__ y, const int64_t *__restrict__ t, int64_t N, int H, long long t_cut,
                           int32_t *__restrict__ flag, unsigned long long *__restrict__ tlast)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool keep = false;
    long long ti = 0;
    if (i < N) { ti = t[i]; keep = (int)y[i] < H && ti < t_cut; flag[i] = keep ? 1 : 0; }
    // t[-1] of the kept events (dsec_data.py:145): the stream is time-sorted, so the last kept event has the largest t.
    // (raw timestamps are non-negative microseconds)
    long long m = keep ? ti : -1;
    for (int d = 16; d > 0; d >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, d));
    if ((threadIdx.x & 31) == 0 && m >= 0) atomicMax(tlast, (unsigned long long)m);
}

__global__ void k_ing_emit(const int32_t *__restrict__ flag, const int32_t *__restrict__ pos, int64_t N,
                           const uint16_t *__restrict__ x, const uint16_t *__restrict__ y, const int64_t *__restrict__ t,
                           const int8_t *__restrict__ p, int p_is_01, float W, float H, int T,
                           const unsigned long long *__restrict__ tlast, int b, int32_t *__restrict__ batch_o,
                           int32_t *__restrict__ pos_o, float *__restrict__ feat_o)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N || !flag[i]) return;
    const int j = pos[i];
    // data/utils.py:12-13: xy -> int16, t -> int32;  dsec_data.py:145: t = time_window + t - t[-1]
    const int xi = (int)(int16_t)x[i], yi = (int)(int16_t)y[i];
    const int tr = (int)((long long)T + t[i] - (long long)*tlast);
    // buffers.py:43: int / int -> fp32 true division;  ev_tgn.py:15-16: (pos * [W,H,T] + 1e-3).int()
    const float Tf = (float)T;
    const float px = __fdiv_rn((float)xi, W), py = __fdiv_rn((float)yi, H), pt = __fdiv_rn((float)tr, Tf);
    pos_o[3 * (int64_t)j] = (int)__fadd_rn(__fmul_rn(W, px), 1e-3f);
    pos_o[3 * (int64_t)j + 1] = (int)__fadd_rn(__fmul_rn(H, py), 1e-3f);
    pos_o[3 * (int64_t)j + 2] = (int)__fadd_rn(__fmul_rn(Tf, pt), 1e-3f);
    batch_o[j] = b;
    const int pv = (int)p[i];
    feat_o[j] = (float)(p_is_01 ? (int)(int8_t)(2 * pv - 1) : pv);        // dsec_data.py:146
}

extern "C" int dagr_ingest_events(const uint16_t *x, const uint16_t *y, const int64_t *t, const int8_t *p, int64_t N,
                                  int p_is_01, int W, int H, int T, int64_t t_cut, int sample, int32_t *flag, int32_t *pos,
                                  int32_t *blocksums, unsigned long long *tlast, int32_t *batch_out, int32_t *pos_out,
                                  float *feat_out, int32_t *n_out, void *stream)
{
    cudaStream_t st = (cudaStream_t)stream;
    DAGR_CHECK_ARG(N < (1ll << 31), "N must fit int32");
    DAGR_CUDA(cudaMemsetAsync(tlast, 0, sizeof(unsigned long long), st));
    if (N <= 0) { DAGR_CUDA(cudaMemsetAsync(n_out, 0, sizeof(int32_t), st)); return DAGR_OK; }
    k_ing_flag<<<dagr_div_up(N, 256), 256, 0, st>>>(y, t, N, H, (long long)t_cut, flag, tlast);
    scan_exclusive(flag, pos, N, blocksums, st);
    k_ing_emit<<<dagr_div_up(N, 256), 256, 0, st>>>(flag, pos, N, x, y, t, p, p_is_01, (float)W, (float)H, T, tlast, sample,
                                                    batch_out, pos_out, feat_out);
    DAGR_CUDA(cudaMemcpyAsync(n_out, pos + N, sizeof(int32_t), cudaMemcpyDeviceToDevice, st));   // total kept
    DAGR_CHECK_LAUNCH();
    return DAGR_OK;
}
