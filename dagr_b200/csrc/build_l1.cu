// build_l1.cu -- fused event-level build: spiral probe on a shared-memory hashed grid + conv_a.
//
// One CTA per pool1 voxel.  Because events are stored in cell-major order, the voxel's own events are
// one contiguous range and the events of its 3x3 voxel neighbourhood are THREE contiguous runs (one per
// voxel row), so the (t, arrival idx, polarity) records the probe needs are staged in shared memory
// with fully coalesced loads.  A per-tile-pixel bin table (shared memory) maps each of the
// (CW+2r) x (CH+2r) pixels the voxel's events can reach to its FIFO column (newest <= Q entries,
// ev_graph.cu:201-211).  The spiral probe (ev_graph.cu:49-78) then runs entirely on chip; accepted
// neighbours are written to the column-major ELL and folded into conv_block1.conv_block1 on the fly
// (their features are (polarity, x/W, y/H): no gather at all).
// Voxels whose neighbourhood does not fit the staging buffer fall back to probing global memory.
#include "common.cuh"

// Two instances of the per-voxel routine (template parameters CAP = staged neighbourhood records, THREADS = CTA size):
//   regular : one CTA per pool1 voxel, 160 threads (>= events of a voxel at the nominal density, Poisson mean 134: one
//             pass), 2048 staged records, 4 CTAs per SM -- or, while no dense voxels are expected (defer = 0), the lean
//             launch with 1536 records (uniform 300k events/sample need ~1200-1450) and 5 CTAs per SM: the kernel is
//             latency bound, 25 instead of 20 warps per SM make it 7 % faster (1.57 -> 1.46 ms at config 2);
//   dense   : voxels whose 3x3 neighbourhood holds more records than that (moving edges in real / clustered streams: 58 %
//             of the events of the clustered benchmark stream) are pushed on a device work list by the regular kernel and
//             processed by a persistent second kernel (one 512-thread CTA per SM, 12288 staged records, dynamic pop) --
//             the global-memory probe remains only as the fallback behind that.
#define BL_THREADS 160
#define BL_CAP 2048
#define BL_CAP_LEAN 1536           // the count-only launch (defer = 0): 44 KB per CTA -> five CTAs per SM instead of four
#define BL_THREADS_BIG 512
#define BL_CAP_BIG 12288
#define BL_NB 8                  // time buckets of width delta_t kept per tile pixel

struct BLTile {
    int X0, Y0, TW, TH;          // tile origin (pixel) and extent
    int run_start[3], run_off[3], run_len[3];
    int smin, smax;              // slice (t / delta_t) range of the voxel's own events
    int unsorted;                // some pixel's records are not time-sorted -> no time bucketing
    int maxidx;                  // newest arrival index among the events this launch must process (-1: none)
};

__host__ __device__ __forceinline__ size_t bl_acc_offset(const dagr_geom_t &g, int cap, int threads)
{
    const size_t TW = g.CW + 2 * g.r, TH = g.CH + 2 * g.r, TP = TW * TH;
    const size_t o = (size_t)cap * 12 + TP * 4 + (TW + TH) * 4 + TP * BL_NB * 2 + (size_t)g.ncell * 4 + (size_t)threads * 2 + (TW + TH);
    return (o + 15) / 16 * 16;
}
static size_t bl_smem_bytes(const dagr_geom_t *g, int cap, int threads)
{
    const size_t TW = g->CW + 2 * g->r, TH = g->CH + 2 * g->r;
    return bl_acc_offset(*g, cap, threads) + (size_t)(DAGR_ELL - 1) * threads * 4 + (size_t)BL_NB * (TW + TH) * 4 + 16;
}

#define BL_R1 96                 // spiral cells walked one-thread-per-event before unsaturated events are handed
                                 // to the warp-cooperative continuation (saturated events need ~85 cells)

// Phase 1: warp-converged probe.  All lanes walk the spiral in lock step (same cell index c), four cells per
// iteration so the dependent shared-memory loads pipeline; the per-bin record loop runs to the warp-wide
// maximum with predication and the walk ends when every lane has its K-1 neighbours.  Lanes of a warp are
// events of similar age (threads are assigned in arrival order) and each tile pixel exposes only the
// sub-range of its FIFO column that can lie within delta_t of the event (time buckets), so the common
// iteration touches no record at all.
template <bool STAGED, int THREADS>
__device__ __forceinline__ void bl_probe(const dagr_geom_t &g, int64_t N, int p, bool active, const int2 me, int eb, int tidx0,
                                         int c_end, const uint32_t *s_pbin, const uint16_t *s_rng, const short *s_sp2,
                                         const int2 *s_ti, const int2 *__restrict__ ti, uint32_t *s_acc,
                                         int32_t *__restrict__ nbr, uint16_t *__restrict__ off, int &n_out)
{
    const int kmax = g.K - 1;
    int n = active ? 0 : kmax;
    for (int c0 = 0; c0 < c_end; c0 += 4) {
        if (__all_sync(0xffffffffu, n >= kmax)) break;
        uint32_t rg[4]; int pix[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int c = min(c0 + u, g.ncell - 1);
            pix[u] = tidx0 + s_sp2[c];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) rg[u] = s_rng[pix[u] * BL_NB + eb];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int c = c0 + u;
            const int lo = rg[u] & 0xff, hi = rg[u] >> 8;
            const int cnt = (c < c_end && n < kmax) ? hi - lo : 0;
            if (!__any_sync(0xffffffffu, cnt > 0)) continue;
            const int base = (int)(s_pbin[pix[u]] >> 8);
            const int vmax = __reduce_max_sync(0xffffffffu, cnt);
            for (int k = 0; k < vmax; k++) {
                if (k < cnt && n < kmax) {
                    // FIFO order: newest first.  Visible records of the pixel are [base, base+vis) in arrival order;
                    // the bucket range [lo,hi) counts from the OLDEST visible record.
                    const int j = base + hi - 1 - k;
                    const int2 o = STAGED ? s_ti[j] : __ldg(ti + j);
                    if (o.y < me.y && me.x - o.x <= g.dt_us) {              // ev_graph.cu:64-69
                        // accepted (record, cell) pairs wait in shared memory for phase B (staged mode)
                        if (STAGED) s_acc[n * THREADS + threadIdx.x] = ((uint32_t)j << 10) | (uint32_t)c;
                        else { nbr[(int64_t)n * N + p] = j; off[(int64_t)n * N + p] = (uint16_t)c; }
                        n++;
                    }
                }
            }
        }
    }
    n_out = active ? n : 0;
}


#ifndef BL_SHARED_ADDR
#define BL_SHARED_ADDR 1
#endif
// explicit 32-bit shared-memory accesses for the ring walk: with generic pointers ptxas re-materialises the shared window base
// (S2R SR_CgaCtaId + MOV + LEA) inside the pop loop instead of keeping it in a register: 1.457 -> 1.419 ms (same-box A/B).
// The same treatment of phase B's seven table bases (the conv_a accumulation loop) costs registers the 45 accumulators do not
// leave: stack 16 -> 40 bytes, 1.418 -> 1.555 ms -- that loop keeps its generic pointers.
__device__ __forceinline__ uint32_t bl_sa(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t bl_lds32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t bl_lds16(uint32_t a) { uint32_t v; asm volatile("{ .reg .u16 t; ld.shared.u16 t, [%1]; cvt.u32.u16 %0, t; }" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ int bl_lds16s(uint32_t a) { int v; asm volatile("{ .reg .s16 t; ld.shared.s16 t, [%1]; cvt.s32.s16 %0, t; }" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ int2 bl_lds64(uint32_t a) { int2 v; asm volatile("ld.shared.v2.s32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory"); return v; }
#ifndef BL_FAST_TABLES
#define BL_FAST_TABLES 1            // exact reciprocal division + 32-bit shared addressing in the occupancy scan: 1.420 -> 1.411 ms (A/B)
#endif
// t / d for 0 <= t < 2^24 through a float reciprocal and one exact fix-up step (the generic 32-bit division is ~20 instructions);
// anything else takes the plain division, so the result is always the C quotient
__device__ __forceinline__ int bl_div(int t, int d, float inv)
{
#if BL_FAST_TABLES
    if ((unsigned)t >= (1u << 24)) return t / d;
    int q = (int)((float)t * inv);
    const int r = t - q * d;
    if (r < 0) q--; else if (r >= d) q++;
    return q;
#else
    (void)inv;
    return t / d;
#endif
}
__device__ __forceinline__ void bl_sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }

// Ring walk with occupancy bitmasks.  For every time bucket e the CTA keeps, per tile row, a 32-bit mask of the
// pixels whose bucket range is non-empty (and the transposed per-column masks).  Ring d of the spiral consists of
// four straight segments (spiral.h: right column upwards, top row leftwards, left column downwards, bottom row
// rightwards, 2d cells each), so the candidate cells of a segment are one bit-field extract (+ a bit reversal
// for the two descending legs).  Each lane then visits ONLY its non-empty cells, in spiral order, by clearing
// the lowest set bit: empty pixels (~2/3 of the window) cost nothing and lanes are no longer held to the
// warp-wide maximum of per-cell work.  Unsaturated events simply run out of rings (no second phase needed).
template <bool STAGED, int THREADS>
__device__ __forceinline__ void bl_probe_rings(const dagr_geom_t &g, int64_t N, int p, bool active, const int2 me, int eb, int tx0, int ty0,
                                               int TW, int TH, const uint32_t *s_pbin, const uint16_t *s_rng, const uint32_t *s_occ_r,
                                               const uint32_t *s_occ_c, const short *s_sp2, const int2 *s_ti, const int2 *__restrict__ ti, uint32_t *s_acc,
                                               int32_t *__restrict__ nbr, uint16_t *__restrict__ off, int &n_out)
{
    const int kmax = g.K - 1;
    int n = active ? 0 : kmax;
    const uint32_t *occr = s_occ_r + eb * TH, *occc = s_occ_c + eb * TW;
    constexpr bool SA = STAGED && BL_SHARED_ADDR;
    const uint32_t a_rng = bl_sa(s_rng) + 2u * (uint32_t)eb, a_pbin = bl_sa(s_pbin), a_ti = bl_sa(s_ti), a_sp2 = bl_sa(s_sp2);
    const uint32_t a_acc = bl_sa(s_acc) + 4u * threadIdx.x;
    const int dt_us = g.dt_us;
    auto visit = [&](int pix, int c) {
        const uint32_t rg = SA ? bl_lds16(a_rng + (uint32_t)pix * (2u * BL_NB)) : (uint32_t)s_rng[pix * BL_NB + eb];
        const int lo = rg & 0xff, hi = rg >> 8;
        const int base = (int)((SA ? bl_lds32(a_pbin + 4u * (uint32_t)pix) : s_pbin[pix]) >> 8);
        for (int j = base + hi - 1; j >= base + lo && n < kmax; j--) {     // FIFO order: newest first
            const int2 o = SA ? bl_lds64(a_ti + 8u * (uint32_t)j) : (STAGED ? s_ti[j] : __ldg(ti + j));
            if (o.y < me.y && me.x - o.x <= dt_us) {                        // ev_graph.cu:64-69
                if (SA) bl_sts32(a_acc + (uint32_t)n * (4u * THREADS), ((uint32_t)j << 10) | (uint32_t)c);
                else if (STAGED) s_acc[n * THREADS + threadIdx.x] = ((uint32_t)j << 10) | (uint32_t)c;
                else { nbr[(int64_t)n * N + p] = j; off[(int64_t)n * N + p] = (uint16_t)c; }
                n++;
            }
        }
    };
    if (n < kmax && ((occr[ty0] >> tx0) & 1u)) visit(ty0 * TW + tx0, 0);   // spiral cell 0: own pixel
    const int tidx0 = ty0 * TW + tx0;
    for (int d = 1; d <= g.r; d++) {
        if (__all_sync(0xffffffffu, n >= kmax)) break;
        // one 8d-bit occupancy word per ring, bit i = spiral cell cbase + i (the four legs of spiral.h back to back), so a
        // lane runs ONE pop loop per ring instead of four: trip counts (popcount of a whole ring) vary much less between
        // the lanes of a warp than those of single legs
        const int l2 = 2 * d;
        const uint32_t fm = (l2 >= 32) ? 0xffffffffu : ((1u << l2) - 1u);
        const int cbase = (l2 - 1) * (l2 - 1);
        // the 8d mask bits live in two 32-bit words (d <= 8): bit i of `lo` = spiral cell cbase + i, of `hi` = cbase + 32 + i;
        // popping from a register pair costs a third of the 64-bit find-first-set / clear-lowest sequence
        uint32_t lo = 0, hi = 0;
        if (n < kmax) {
            const uint32_t leg0 = (occc[tx0 + d] >> (ty0 - d + 1)) & fm;                          // x = +d, y = -d+1 .. d
            const uint32_t leg1 = __brev((occr[ty0 + d] >> (tx0 - d)) & fm) >> (32 - l2);         // y = +d, x = d-1 .. -d
            const uint32_t leg2 = __brev((occc[tx0 - d] >> (ty0 - d)) & fm) >> (32 - l2);         // x = -d, y = d-1 .. -d
            const uint32_t leg3 = (occr[ty0 - d] >> (tx0 - d + 1)) & fm;                          // y = -d, x = -d+1 .. d
            const unsigned long long m = (unsigned long long)leg0 | ((unsigned long long)leg1 << l2) |
                                         ((unsigned long long)leg2 << (2 * l2)) | ((unsigned long long)leg3 << (3 * l2));
            lo = (uint32_t)m; hi = (uint32_t)(m >> 32);
        }
        while (lo | hi) {
            int i;
            if (lo) { i = __ffs((int)lo) - 1; lo &= lo - 1; }
            else    { i = 32 + __ffs((int)hi) - 1; hi &= hi - 1; }
            visit(tidx0 + (SA ? bl_lds16s(a_sp2 + 2u * (uint32_t)(cbase + i)) : (int)s_sp2[cbase + i]), cbase + i);
            if (n >= kmax) { lo = 0; hi = 0; }
        }
    }
    n_out = active ? n : 0;
}

// Phase 2: warp-cooperative continuation for one unsaturated event: the 32 lanes test 32 consecutive spiral
// cells at once; an exclusive scan over the lanes' accept counts restores the spiral order and the K cap.
template <bool STAGED, int THREADS>
__device__ __forceinline__ int bl_probe_coop(const dagr_geom_t &g, int64_t N, int p, const int2 me, int eb, int tidx0, int c_begin,
                                             int n, const uint32_t *s_pbin, const uint16_t *s_rng, const short *s_sp2,
                                             const int2 *s_ti, const int2 *__restrict__ ti, uint32_t *s_acc, int owner,
                                             int32_t *__restrict__ nbr, uint16_t *__restrict__ off)
{
    const int kmax = g.K - 1;
    const int lane = threadIdx.x & 31;
    for (int c0 = c_begin; c0 < g.ncell && n < kmax; c0 += 32) {
        const int c = c0 + lane;
        int cnt = 0, base = 0, hi = 0, acc = 0;
        if (c < g.ncell) {
            const int pix = tidx0 + s_sp2[c];
            const uint32_t rg = s_rng[pix * BL_NB + eb];
            hi = rg >> 8; cnt = hi - (int)(rg & 0xff);
            base = (int)(s_pbin[pix] >> 8);
            for (int k = 0; k < cnt; k++) {
                const int2 o = STAGED ? s_ti[base + hi - 1 - k] : __ldg(ti + base + hi - 1 - k);
                acc += (o.y < me.y && me.x - o.x <= g.dt_us) ? 1 : 0;
            }
        }
        int incl = acc;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, d); if (lane >= d) incl += v; }
        const int tot = __shfl_sync(0xffffffffu, incl, 31);
        if (acc > 0) {
            int slot = n + incl - acc;
            for (int k = 0; k < cnt && slot < kmax; k++) {
                const int j = base + hi - 1 - k;
                const int2 o = STAGED ? s_ti[j] : __ldg(ti + j);
                if (o.y < me.y && me.x - o.x <= g.dt_us) {
                    if (STAGED) s_acc[slot * THREADS + owner] = ((uint32_t)j << 10) | (uint32_t)c;
                    else { nbr[(int64_t)slot * N + p] = j; off[(int64_t)slot * N + p] = (uint16_t)c; }
                    slot++;
                }
            }
        }
        n = min(n + tot, kmax);
    }
    return n;
}

// work list: wl_hdr[0] = number of voxels beyond this instance's staging capacity (queued in wl_ids when `defer`, otherwise
// only counted and probed from global memory), wl_hdr[1] = pop cursor of the dense kernel
template <int CAP, int THREADS>
__device__ __forceinline__ void bl_voxel(const dagr_geom_t &g, int64_t N, const int32_t *__restrict__ start, const int2 *__restrict__ ti,
                                         const uint32_t *__restrict__ xyb, const float *__restrict__ feat_s, const float *__restrict__ tab,
                                         const dagr_l1a_params_t &P, const int do_conv, const int min_idx, const int32_t *__restrict__ flags,
                                         int32_t *__restrict__ nbr, uint16_t *__restrict__ off, uint32_t *__restrict__ cellmask,
                                         float *__restrict__ xa, const int cell, unsigned char *smem_raw, BLTile &T, uint32_t &s_mask,
                                         int32_t *__restrict__ wl_hdr, int32_t *__restrict__ wl_ids, const int defer)
{
    const int per = g.ny1 * g.nx1;
    const int b = cell / per, rem = cell % per, cy = rem / g.nx1, cx = rem % g.nx1;
    const int p0 = start[(int64_t)cell * g.CP], p1 = start[(int64_t)(cell + 1) * g.CP];
    if (p1 == p0) { if (threadIdx.x == 0 && min_idx <= 0) cellmask[cell] = 0; return; }      // block-uniform

    // ---- shared memory carve-up -------------------------------------------------------------------
    const int TWmax = g.CW + 2 * g.r, THmax = g.CH + 2 * g.r, TPmax = TWmax * THmax;
    int2 *s_ti = (int2 *)smem_raw;                                      // [CAP]
    float *s_feat = (float *)(s_ti + CAP);                              // [CAP]
    uint16_t *s_rng = (uint16_t *)(s_feat + CAP);                       // [TP][BL_NB]  lo | hi << 8   (16-byte aligned: CAP % 4 == 0)
    uint32_t *s_pbin = (uint32_t *)(s_rng + TPmax * BL_NB);             // [TP]  pos << 8 | visible count
    float *s_posx = (float *)(s_pbin + TPmax);                          // [TWmax]
    float *s_posy = s_posx + TWmax;                                     // [THmax]
    short *s_sp = (short *)(s_posy + THmax);                     // [ncell]  dx | dy << 8
    short *s_sp2 = s_sp + g.ncell;                                      // [ncell]  dy*TW + dx
    uint16_t *s_order = (uint16_t *)(s_sp2 + g.ncell);                  // [THREADS]
    uint32_t *s_acc = (uint32_t *)(smem_raw + bl_acc_offset(g, CAP, THREADS));   // [K-1][THREADS]  record << 10 | cell
    uint32_t *s_occ_r = s_acc + (DAGR_ELL - 1) * THREADS;               // [BL_NB][THmax] row occupancy bitmasks
    unsigned char *s_colv = (unsigned char *)(s_order + THREADS);       // [TWmax]
    unsigned char *s_rowv = s_colv + TWmax;                             // [THmax]

    const int dtw = max(g.dt_us, 1);
    const float dtw_inv = 1.0f / (float)dtw;
    if (threadIdx.x == 0) {
        const int X0 = g.vx0[cx], X1 = g.vx0[cx + 1], Y0 = g.vy0[cy], Y1 = g.vy0[cy + 1];
        T.X0 = X0 - g.r; T.Y0 = Y0 - g.r; T.TW = X1 - X0 + 2 * g.r; T.TH = Y1 - Y0 + 2 * g.r;
        const int clo = max(cx - 1, 0), chi = min(cx + 1, g.nx1 - 1);
        int o = 0;
        for (int rr = 0; rr < 3; rr++) {
            const int ry = cy - 1 + rr;
            if (ry < 0 || ry >= g.ny1) { T.run_start[rr] = 0; T.run_len[rr] = 0; T.run_off[rr] = o; continue; }
            const int64_t c0 = (int64_t)b * per + ry * g.nx1 + clo, c1 = (int64_t)b * per + ry * g.nx1 + chi + 1;
            const int s = start[c0 * g.CP], e = start[c1 * g.CP];
            T.run_start[rr] = s; T.run_len[rr] = e - s; T.run_off[rr] = o;
            o += e - s;
        }
        T.smin = 0x7fffffff; T.smax = -0x7fffffff; T.unsorted = 0; T.maxidx = -1;
        s_mask = 0;
    }
    __syncthreads();
    const int TW = T.TW, TH = T.TH, TP = TW * TH;
    uint32_t *s_occ_c = s_occ_r + BL_NB * TH;                           // [BL_NB][TW] column occupancy bitmasks
    const int total = T.run_off[2] + T.run_len[2];
    const bool staged = total <= CAP;                                   // block-uniform
    if (!staged && wl_hdr != nullptr) {
        // too many records for this instance's staging buffer: hand the voxel to the dense kernel (which runs next on
        // the stream) instead of probing global memory -- or, when the caller did not ask for that, just count it
        int slot = 0;
        if (threadIdx.x == 0) slot = atomicAdd(&wl_hdr[0], 1);
        if (defer) { if (threadIdx.x == 0) wl_ids[slot] = cell; return; }
    }
    const int bbase = b * per * g.CP;

    for (int i = threadIdx.x; i < g.ncell; i += blockDim.x) {
        const int dx = g.spiral[2 * i], dy = g.spiral[2 * i + 1];
        s_sp[i] = (short)((dx & 0xff) | (dy << 8));
        s_sp2[i] = (short)(dy * TW + dx);
    }
    // ---- tile tables: per-pixel FIFO bin, normalised positions, voxel direction ---------------------
    for (int i = threadIdx.x; i < TW; i += blockDim.x) {
        const int gx = T.X0 + i;
        const bool in = gx >= 0 && gx < g.W;
        s_posx[i] = in ? g.posx0[gx] : 0.f;
        s_colv[i] = (unsigned char)(in ? (g.xkey[gx] / g.CP - cx + 1) : 1);
    }
    for (int i = threadIdx.x; i < TH; i += blockDim.x) {
        const int gy = T.Y0 + i;
        const bool in = gy >= 0 && gy < g.H;
        s_posy[i] = in ? g.posy0[gy] : 0.f;
        s_rowv[i] = (unsigned char)(in ? (g.ykey[gy] / (g.nx1 * g.CP) - cy + 1) : 1);
    }
    for (int i = threadIdx.x; i < TP; i += blockDim.x) {
        const int ty = i / TW, tx = i % TW;
        const int gx = T.X0 + tx, gy = T.Y0 + ty;
        uint32_t v = 0;
        if (gx >= 0 && gx < g.W && gy >= 0 && gy < g.H) {
            const int ky = g.ykey[gy], kx = g.xkey[gx];
            const int k = bbase + ky + kx;
            const int s = start[k], e = start[k + 1];
            if (e > s) {
                const int lo = max(s, e - g.Q);                         // newest Q entries (ev_graph.cu:201-211)
                const int rr = ky / (g.nx1 * g.CP) - cy + 1;
                const int pos = staged ? (lo - T.run_start[rr] + T.run_off[rr]) : lo;
                v = ((uint32_t)pos << 8) | (uint32_t)(e - lo);
            }
        }
        s_pbin[i] = v;
    }
    // ---- stage the neighbourhood records (three coalesced runs) -------------------------------------
    if (staged) {
        for (int rr = 0; rr < 3; rr++) {
            const int s = T.run_start[rr], o = T.run_off[rr], len = T.run_len[rr];
            for (int i = threadIdx.x; i < len; i += blockDim.x) {
                s_ti[o + i] = ti[s + i];
                s_feat[o + i] = feat_s[s + i];
            }
        }
    }
    // slice range of the voxel's own events
    {
        int mn = 0x7fffffff, mx = -0x7fffffff, mi = -1;
        for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
            const int2 r = ti[p];
            if (r.y < min_idx) continue;                                // incremental mode: only new events are processed
            const int sl = bl_div(r.x, dtw, dtw_inv); mn = min(mn, sl); mx = max(mx, sl); mi = max(mi, r.y);
        }
        mn = __reduce_min_sync(0xffffffffu, mn); mx = __reduce_max_sync(0xffffffffu, mx); mi = __reduce_max_sync(0xffffffffu, mi);
        if ((threadIdx.x & 31) == 0) { atomicMin(&T.smin, mn); atomicMax(&T.smax, mx); atomicMax(&T.maxidx, mi); }
    }
    __syncthreads();
    if (T.maxidx < 0) return;                                           // block-uniform: no new event in this voxel
    // ---- per-pixel time-bucket ranges ---------------------------------------------------------------
    // bucket(t) = clamp(t/delta_t - (smin-1), 0, NB-1); an event in bucket e needs records of buckets {e-1, e}.
    const int sbase = T.smin - 1;
    bool bucketed = staged && (T.smax - sbase) < BL_NB && (flags == nullptr || flags[0] == 0);   // block-uniform
    const bool use_rings = TW <= 32 && TH <= 32 && g.r <= 8;            // block-uniform (a ring = 8d <= 64 mask bits)
    for (int pass = 0; pass < 2; pass++) {
        for (int i = threadIdx.x; i < TP; i += blockDim.x) {
            const uint32_t pb = s_pbin[i];
            const int vis = pb & 0xff, base = pb >> 8;
            unsigned char cum[BL_NB + 1];
#pragma unroll
            for (int q = 0; q <= BL_NB; q++) cum[q] = 0;
            if (bucketed) {
                int prev = 0;
                for (int k = 0; k < vis; k++) {
                    int bk = bl_div(s_ti[base + k].x, dtw, dtw_inv) - sbase;
                    bk = min(max(bk, 0), BL_NB - 1);
                    if (bk < prev) T.unsorted = 1;                      // benign race: any writer sets 1
                    prev = bk;
#pragma unroll
                    for (int q = 0; q <= BL_NB; q++) cum[q] += (bk < q) ? 1 : 0;
                }
            } else {
#pragma unroll
                for (int q = 1; q <= BL_NB; q++) cum[q] = (unsigned char)vis;
            }
            // the eight (lo | hi << 8) ranges of a pixel are one 16-byte store (eight 2-byte stores at a 16-byte lane stride
            // cost four wavefronts each)
            static_assert(BL_NB == 8, "one uint4 per pixel");
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < BL_NB; e += 2) {
                const uint32_t r0 = (uint32_t)cum[e > 0 ? e - 1 : 0] | ((uint32_t)cum[e + 1] << 8);
                const uint32_t r1 = (uint32_t)cum[e] | ((uint32_t)cum[e + 2] << 8);
                w[e >> 1] = r0 | (r1 << 16);
            }
            reinterpret_cast<uint4 *>(s_rng)[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
        __syncthreads();
        if (use_rings) {
            // occupancy bitmasks for the ring walk: one thread per (bucket, tile row) / (bucket, tile column) word scans its
            // pixels -- no atomics (one atomicOr per pixel and bucket serialised on the row word: 6x the ideal wavefronts)
            for (int idx = threadIdx.x; idx < BL_NB * (TH + TW); idx += blockDim.x) {
                uint32_t m = 0;
#if BL_FAST_TABLES
                if (idx < BL_NB * TH) {
                    const int e = idx / TH, row = idx % TH;
                    uint32_t a = bl_sa(s_rng) + 2u * (uint32_t)(row * TW * BL_NB + e);
                    for (int x = 0; x < TW; x++, a += 2u * BL_NB) {
                        const uint32_t rg = bl_lds16(a);
                        m |= ((rg >> 8) > (rg & 0xff)) ? (1u << x) : 0u;
                    }
                } else {
                    const int i2 = idx - BL_NB * TH, e = i2 / TW, col = i2 % TW;
                    uint32_t a = bl_sa(s_rng) + 2u * (uint32_t)(col * BL_NB + e);
                    for (int y = 0; y < TH; y++, a += 2u * BL_NB * (uint32_t)TW) {
                        const uint32_t rg = bl_lds16(a);
                        m |= ((rg >> 8) > (rg & 0xff)) ? (1u << y) : 0u;
                    }
                }
#else
                if (idx < BL_NB * TH) {
                    const int e = idx / TH, row = idx % TH;
                    for (int x = 0; x < TW; x++) {
                        const uint32_t rg = s_rng[(row * TW + x) * BL_NB + e];
                        m |= ((rg >> 8) > (rg & 0xff)) ? (1u << x) : 0u;
                    }
                } else {
                    const int i2 = idx - BL_NB * TH, e = i2 / TW, col = i2 % TW;
                    for (int y = 0; y < TH; y++) {
                        const uint32_t rg = s_rng[(y * TW + col) * BL_NB + e];
                        m |= ((rg >> 8) > (rg & 0xff)) ? (1u << y) : 0u;
                    }
                }
#endif
                s_occ_r[idx] = m;                                        // s_occ_c follows s_occ_r: [BL_NB][TH] then [BL_NB][TW]
            }
        }
        __syncthreads();
        if (!bucketed || !T.unsorted) break;                            // block-uniform
        bucketed = false;                                               // records not time-sorted: redo without buckets
    }
    // ---- thread <-> event assignment in arrival order (time-homogeneous warps) ------------------------
    const int nown = p1 - p0;
    const int own_off = staged ? (p0 - T.run_start[1] + T.run_off[1]) : 0;
    uint32_t mloc = 0;
    for (int pb0 = 0; pb0 < nown; pb0 += blockDim.x) {
        const int chunk = min((int)blockDim.x, nown - pb0);
        __syncthreads();
        // rank of each event of the chunk by arrival index (cell walk only: the ring walk does not need age-sorted warps)
        if (!use_rings && (int)threadIdx.x < chunk) {
            const int myidx = staged ? s_ti[own_off + pb0 + threadIdx.x].y : ti[p0 + pb0 + threadIdx.x].y;
            int rank = 0;
            for (int k = 0; k < chunk; k++) {
                const int oi = staged ? s_ti[own_off + pb0 + k].y : __ldg(&ti[p0 + pb0 + k].y);
                rank += (oi < myidx) ? 1 : 0;
            }
            s_order[rank] = (uint16_t)threadIdx.x;
        }
        __syncthreads();
        // arrival ranks are dealt round-robin to the warps: every warp gets the same share of the old
        // (unsaturated, slow) events of the voxel, so no warp is the straggler of the CTA
        const int nw = blockDim.x >> 5;
        // ring walk: warps of similar age leave the ring loop together -> contiguous arrival ranks per warp;
        // cell walk (fallback): ranks dealt round-robin so that no warp is the straggler
        const int rank = use_rings ? (int)threadIdx.x : (int)(threadIdx.x & 31) * nw + (int)(threadIdx.x >> 5);
        bool active = rank < chunk;
        const int p = p0 + pb0 + (active ? (use_rings ? rank : (int)s_order[rank]) : 0);
        int x = 0, y = 0;
        int2 me = make_int2(0, 0);
        if (active) {
            const uint32_t w = xyb[p];
            x = w & 0xfff; y = (w >> 12) & 0xfff;
            me = ti[p];
            active = me.y >= min_idx;
        }
        const int tx0 = active ? x - T.X0 : g.r, ty0 = active ? y - T.Y0 : g.r;
        int eb = 0;
        if (bucketed) eb = min(max(bl_div(me.x, dtw, dtw_inv) - sbase, 0), BL_NB - 1);
        int n;
        const int tidx0 = ty0 * TW + tx0;
        if (use_rings) {
            if (staged) bl_probe_rings<true, THREADS>(g, N, p, active, me, eb, tx0, ty0, TW, TH, s_pbin, s_rng, s_occ_r, s_occ_c, s_sp2, s_ti, ti, s_acc, nbr, off, n);
            else        bl_probe_rings<false, THREADS>(g, N, p, active, me, eb, tx0, ty0, TW, TH, s_pbin, s_rng, s_occ_r, s_occ_c, s_sp2, s_ti, ti, s_acc, nbr, off, n);
        } else {
            if (staged) bl_probe<true, THREADS>(g, N, p, active, me, eb, tidx0, BL_R1 < g.ncell ? BL_R1 : g.ncell, s_pbin, s_rng, s_sp2, s_ti, ti, s_acc, nbr, off, n);
            else        bl_probe<false, THREADS>(g, N, p, active, me, eb, tidx0, BL_R1 < g.ncell ? BL_R1 : g.ncell, s_pbin, s_rng, s_sp2, s_ti, ti, s_acc, nbr, off, n);
            // events still unsaturated after BL_R1 cells continue warp-cooperatively (32 cells per step), one at a time
            if (BL_R1 < g.ncell) {
                unsigned todo = __ballot_sync(0xffffffffu, active && n < g.K - 1);
                while (todo) {
                    const int src = __ffs(todo) - 1;
                    todo &= todo - 1;
                    const int ep = __shfl_sync(0xffffffffu, p, src);
                    const int2 eme = make_int2(__shfl_sync(0xffffffffu, me.x, src), __shfl_sync(0xffffffffu, me.y, src));
                    const int eeb = __shfl_sync(0xffffffffu, eb, src), etidx = __shfl_sync(0xffffffffu, tidx0, src);
                    const int en = __shfl_sync(0xffffffffu, n, src);
                    const int owner = (int)(threadIdx.x & ~31u) + src;
                    const int nn = staged ? bl_probe_coop<true, THREADS>(g, N, ep, eme, eeb, etidx, BL_R1, en, s_pbin, s_rng, s_sp2, s_ti, ti, s_acc, owner, nbr, off)
                                          : bl_probe_coop<false, THREADS>(g, N, ep, eme, eeb, etidx, BL_R1, en, s_pbin, s_rng, s_sp2, s_ti, ti, s_acc, owner, nbr, off);
                    if ((int)(threadIdx.x & 31) == src) n = nn;
                }
                __syncwarp();
            }
        }
        if (active) nbr[(int64_t)(DAGR_ELL - 1) * N + p] = n;
        // phase B: A_u = sum_e tab[c_e][u] * (polarity_src, x_src/W, y_src/H), converged over the ELL slots;
        // also translates the staged record index into the global sorted position and collects the voxel mask
        const float f0 = active ? feat_s[p] : 0.f, f1 = s_posx[tx0], f2 = s_posy[ty0];
        float A[DAGR_KU][3];
        {
#pragma unroll
            for (int u = 0; u < DAGR_KU; u++) { const float t = __ldg(tab + u); A[u][0] = t * f0; A[u][1] = t * f1; A[u][2] = t * f2; }
        }
        const int nmax = __reduce_max_sync(0xffffffffu, n);
        for (int q = 0; q < nmax; q++) {
            if (q < n) {
                int j, c;
                if (staged) { const uint32_t a = s_acc[q * THREADS + threadIdx.x]; j = (int)(a >> 10); c = (int)(a & 0x3ff); }
                else { j = nbr[(int64_t)q * N + p]; c = off[(int64_t)q * N + p]; }
                const int sp = s_sp[c];
                const int tx = tx0 + (int)(signed char)(sp & 0xff), ty = ty0 + (sp >> 8);
                const int rr = s_rowv[ty];
                float e0;
                if (staged) {
                    e0 = s_feat[j];
                    nbr[(int64_t)q * N + p] = j - T.run_off[rr] + T.run_start[rr];
                    off[(int64_t)q * N + p] = (uint16_t)c;
                } else e0 = __ldg(feat_s + j);
                const int dcx = (int)s_colv[tx] - 1, dcy = rr - 1;
                if (dcx | dcy) mloc |= 1u << ((dcy + 1) * 3 + (dcx + 1));
                const float e1 = s_posx[tx], e2 = s_posy[ty];
                const float4 *tr = reinterpret_cast<const float4 *>(tab + c * DAGR_TABW);
                const float4 t0 = __ldg(tr), t1 = __ldg(tr + 1), t2 = __ldg(tr + 2), t3 = __ldg(tr + 3);
                const float t[DAGR_KU] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w, t2.x, t2.y, t2.z, t2.w, t3.x, t3.y, t3.z};
#pragma unroll
                for (int u = 0; u < DAGR_KU; u++) {
                    A[u][0] = fmaf(t[u], e0, A[u][0]);
                    A[u][1] = fmaf(t[u], e1, A[u][1]);
                    A[u][2] = fmaf(t[u], e2, A[u][2]);
                }
            }
        }
        if (!active || !do_conv) continue;
        // conv_a phase 2: out = sum_u W_u^T A_u + W_root^T x_i, BN, act  (weights in the constant bank)
        float o[16];
#pragma unroll
        for (int k = 0; k < 16; k++) o[k] = 0.f;
#pragma unroll
        for (int u = 0; u < DAGR_KU; u++)
#pragma unroll
            for (int ci = 0; ci < 3; ci++)
#pragma unroll
                for (int k = 0; k < 16; k++) o[k] = fmaf(A[u][ci], P.w[u][ci][k], o[k]);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            float r = o[k];
            r = fmaf(f0, P.root[0][k], r);
            r = fmaf(f1, P.root[1][k], r);
            r = fmaf(f2, P.root[2][k], r);
            r = fmaf(r, P.scale[k], P.shift[k]);
            o[k] = P.relu ? fmaxf(r, 0.f) : r;
        }
        // xa is stored half-major [2][N][8] so that conv_b can stage one 32-byte channel half per pass
        const int sw = XA_SWZ(p);
        float4 *dst = reinterpret_cast<float4 *>(xa + (int64_t)p * 8);
        dst[sw] = make_float4(o[0], o[1], o[2], o[3]);
        dst[sw ^ 1] = make_float4(o[4], o[5], o[6], o[7]);
        dst = reinterpret_cast<float4 *>(xa + (N + (int64_t)p) * 8);
        dst[sw] = make_float4(o[8], o[9], o[10], o[11]);
        dst[sw ^ 1] = make_float4(o[12], o[13], o[14], o[15]);
    }
    mloc = __reduce_or_sync(0xffffffffu, mloc);
    if ((threadIdx.x & 31) == 0 && mloc) atomicOr(&s_mask, mloc);
    __syncthreads();
    if (threadIdx.x == 0) cellmask[cell] = (min_idx > 0 ? cellmask[cell] : 0u) | s_mask;
}



template <int CAP, int MIN_CTAS>
__global__ void __launch_bounds__(BL_THREADS, MIN_CTAS)
k_l1_build(const dagr_geom_t g, int64_t N, const int32_t *__restrict__ start, const int2 *__restrict__ ti,
           const uint32_t *__restrict__ xyb, const float *__restrict__ feat_s, const float *__restrict__ tab,
           const __grid_constant__ dagr_l1a_params_t P, const int do_conv, const int min_idx, const int32_t *__restrict__ flags,
           int32_t *__restrict__ nbr, uint16_t *__restrict__ off, uint32_t *__restrict__ cellmask, float *__restrict__ xa,
           int32_t *__restrict__ wl_hdr, int32_t *__restrict__ wl_ids, const int defer)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ BLTile T;
    __shared__ uint32_t s_mask;
    bl_voxel<CAP, BL_THREADS>(g, N, start, ti, xyb, feat_s, tab, P, do_conv, min_idx, flags, nbr, off, cellmask, xa,
                              (int)blockIdx.x, smem_raw, T, s_mask, wl_hdr, wl_ids, defer);
}

// dense voxels: persistent CTAs (one per SM) pop voxel ids from the work list the regular kernel filled
__global__ void __launch_bounds__(BL_THREADS_BIG, 1)
k_l1_build_dense(const dagr_geom_t g, int64_t N, const int32_t *__restrict__ start, const int2 *__restrict__ ti,
                 const uint32_t *__restrict__ xyb, const float *__restrict__ feat_s, const float *__restrict__ tab,
                 const __grid_constant__ dagr_l1a_params_t P, const int do_conv, const int min_idx, const int32_t *__restrict__ flags,
                 int32_t *__restrict__ nbr, uint16_t *__restrict__ off, uint32_t *__restrict__ cellmask, float *__restrict__ xa,
                 int32_t *__restrict__ wl_hdr, const int32_t *__restrict__ wl_ids)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ BLTile T;
    __shared__ uint32_t s_mask;
    __shared__ int s_next;
    const int count = wl_hdr[0];
    for (;;) {
        __syncthreads();                                                // everyone is done with the previous voxel
        if (threadIdx.x == 0) s_next = atomicAdd(&wl_hdr[1], 1);
        __syncthreads();
        const int i = s_next;
        if (i >= count) break;
        bl_voxel<BL_CAP_BIG, BL_THREADS_BIG>(g, N, start, ti, xyb, feat_s, tab, P, do_conv, min_idx, flags, nbr, off, cellmask, xa,
                                             wl_ids[i], smem_raw, T, s_mask, nullptr, nullptr, 0);
    }
}

extern "C" int dagr_l1_build(const dagr_geom_t *g, int64_t N, const int32_t *start, const int32_t *ti,
                             const uint32_t *xyb, const float *feat_s, const float *tab,
                             const dagr_l1a_params_t *p_host, const int32_t *flags, int min_idx, int32_t *nbr, uint16_t *off,
                             uint32_t *cellmask, float *xa, int32_t *wl_hdr, int32_t *wl_ids, int defer, void *stream)
{
    DAGR_CHECK_ARG(g, "null argument");
    static const dagr_l1a_params_t zero_params = {};
    const int do_conv = p_host != nullptr;
    if (!p_host) p_host = &zero_params;
    DAGR_CHECK_ARG(g->K >= 1 && g->K <= DAGR_ELL, "max_neighbors must be in [1,16]");
    DAGR_CHECK_ARG(g->r >= 0 && g->r <= 15 && g->Q <= 255, "radius must be <= 15 px and max_queue_size <= 255");
    DAGR_CHECK_ARG(N < (1ll << 24), "the staged probe packs positions in 24 bits (N < 16.7M per call)");
    const int cells = g->B * g->ny1 * g->nx1;
    const bool deferring = wl_hdr != nullptr && wl_ids != nullptr && defer;
    if (deferring) {
        const size_t smem = bl_smem_bytes(g, BL_CAP, BL_THREADS);
        auto kern = k_l1_build<BL_CAP, 4>;
        DAGR_CUDA(dagr_allow_smem(kern, smem, true));
        kern<<<cells, BL_THREADS, smem, (cudaStream_t)stream>>>(*g, N, start, (const int2 *)ti, xyb, feat_s, tab, *p_host, do_conv, min_idx,
                                                                flags, nbr, off, cellmask, xa, wl_hdr, wl_ids, 1);
    } else {
        const size_t smem = bl_smem_bytes(g, BL_CAP_LEAN, BL_THREADS);
        auto kern = k_l1_build<BL_CAP_LEAN, 5>;
        DAGR_CUDA(dagr_allow_smem(kern, smem, true));
        kern<<<cells, BL_THREADS, smem, (cudaStream_t)stream>>>(*g, N, start, (const int2 *)ti, xyb, feat_s, tab, *p_host, do_conv, min_idx,
                                                                flags, nbr, off, cellmask, xa, wl_hdr, wl_ids, 0);
    }
    DAGR_CHECK_LAUNCH();
    if (deferring) {
        const size_t smem_big = bl_smem_bytes(g, BL_CAP_BIG, BL_THREADS_BIG);
        static int n_sm = 0;
        if (n_sm == 0) {
            int dev = 0;
            DAGR_CUDA(cudaGetDevice(&dev));
            DAGR_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
        }
        DAGR_CUDA(dagr_allow_smem(k_l1_build_dense, smem_big));
        k_l1_build_dense<<<n_sm, BL_THREADS_BIG, smem_big, (cudaStream_t)stream>>>(*g, N, start, (const int2 *)ti, xyb, feat_s, tab,
                                                                                    *p_host, do_conv, min_idx, flags, nbr, off,
                                                                                    cellmask, xa, wl_hdr, wl_ids);
        DAGR_CHECK_LAUNCH();
    }
    return DAGR_OK;
}
