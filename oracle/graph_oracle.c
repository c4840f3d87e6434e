/*
 * oracle/graph_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Plain-C, single-threaded CPU restatement of the reference's causal radius-graph
 * builder.  The reference has no CPU implementation of this path (its wrapper moves
 * inputs to CUDA, src/dagr/graph/ev_graph.py:5-15), so this file restates the two
 * CUDA kernels + their Python driver op for op:
 *
 *   FIFO insert   : src/dagr/graph/ev_graph.cu:169-212 (batched), :130-166 (single, b=0)
 *                   driver src/dagr/graph/utils.py:6-18 (stable sort by pixel, unique, cumsum)
 *   edge search   : src/dagr/graph/ev_graph.cu:15-80, spiral order src/dagr/graph/spiral.h:1-16
 *                   driver src/dagr/graph/utils.py:20-23 (compaction edges[:, edges[1]>=0])
 *   state machine : src/dagr/graph/ev_graph.py:45-103 (AsyncGraph), :121-136 (delete_nodes)
 *
 * Parity status: on the GPU box this restatement is itself checked against the
 * reference's own kernels compiled from /root/reference into oracle/_ref/
 * (tests/test_graph_gpu.py), so the graph path is pinned by the real reference.
 *
 * Build: see oracle/Makefile  (gcc -O2 -shared -fPIC).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int B, Q, H, W;
    int32_t *queue;        /* [B][Q][H][W], -1 = empty             (ev_graph.py:50) */
    int32_t *ts;           /* all_timestamps, grows by appending   (ev_graph.py:75) */
    int64_t n_ts, cap_ts;
    int64_t max_index, min_index;
} graph_oracle_t;

graph_oracle_t *graph_oracle_create(int B, int Q, int H, int W)
{
    graph_oracle_t *g = (graph_oracle_t *)calloc(1, sizeof(*g));
    if (!g) return NULL;
    g->B = B; g->Q = Q; g->H = H; g->W = W;
    size_t n = (size_t)B * Q * H * W;
    g->queue = (int32_t *)malloc(n * sizeof(int32_t));
    if (!g->queue) { free(g); return NULL; }
    memset(g->queue, 0xff, n * sizeof(int32_t));
    return g;
}

void graph_oracle_destroy(graph_oracle_t *g)
{
    if (!g) return;
    free(g->queue); free(g->ts); free(g);
}

/* ev_graph.py:52-60 */
void graph_oracle_reset(graph_oracle_t *g)
{
    size_t n = (size_t)g->B * g->Q * g->H * g->W;
    memset(g->queue, 0xff, n * sizeof(int32_t));
    g->n_ts = 0; g->max_index = 0; g->min_index = 0;
}

/* ev_graph.py:121-136 (node part only; edge bookkeeping lives in the python wrapper) */
void graph_oracle_delete_nodes(graph_oracle_t *g, int64_t n_delete)
{
    if (n_delete > g->n_ts) n_delete = g->n_ts;
    memmove(g->ts, g->ts + n_delete, (size_t)(g->n_ts - n_delete) * sizeof(int32_t));
    g->n_ts -= n_delete;
    g->min_index += n_delete;
}

int64_t graph_oracle_num_nodes(const graph_oracle_t *g) { return g->n_ts; }
int64_t graph_oracle_min_index(const graph_oracle_t *g) { return g->min_index; }
int64_t graph_oracle_max_index(const graph_oracle_t *g) { return g->max_index; }
const int32_t *graph_oracle_queue(const graph_oracle_t *g) { return g->queue; }

/* one FIFO column update: shift by `counts`, newest first (ev_graph.cu:201-211) */
static void fifo_push(graph_oracle_t *g, int b, int y, int x, const int32_t *new_idx_ascending, int counts)
{
    const int Q = g->Q, H = g->H, W = g->W;
    for (int q = Q - 1; q >= 0; q--) {
        size_t index = (size_t)b * H * W * Q + (size_t)q * H * W + (size_t)y * W + x;
        if (q >= counts) {
            size_t shifted = (size_t)b * H * W * Q + (size_t)(q - counts) * H * W + (size_t)y * W + x;
            g->queue[index] = g->queue[shifted];
        } else {
            g->queue[index] = new_idx_ascending[counts - 1 - q];
        }
    }
}

typedef struct { int64_t key; int32_t idx; } keyed_t;

static int cmp_keyed(const void *a, const void *b)
{
    const keyed_t *x = (const keyed_t *)a, *y = (const keyed_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx ? 1 : 0);   /* stable: arrival order */
}

/*
 * AsyncGraph.forward (ev_graph.py:63-103).
 *   batch : int32[N]          pos : int32[N,3] = (x, y, t)
 *   edges_out : int64[2 * K * N] laid out as the compacted [2, E] result (row 0 then row 1,
 *               each of length E);  returns E (>= 0) or a negative error code.
 */
int64_t graph_oracle_forward(graph_oracle_t *g, const int32_t *batch, const int32_t *pos, int64_t N,
                             int K, int radius, int delta_t_us, int64_t *edges_src, int64_t *edges_dst)
{
    const int Q = g->Q, H = g->H, W = g->W;
    if (N == 0) return 0;                                            /* ev_graph.py:70-71 */

    /* all_timestamps = cat(all_timestamps, pos[:,2])                   ev_graph.py:75 */
    if (g->n_ts + N > g->cap_ts) {
        int64_t cap = (g->n_ts + N) * 2;
        int32_t *t = (int32_t *)realloc(g->ts, (size_t)cap * sizeof(int32_t));
        if (!t) return -1;
        g->ts = t; g->cap_ts = cap;
    }
    for (int64_t i = 0; i < N; i++) g->ts[g->n_ts + i] = pos[3 * i + 2];
    g->n_ts += N;

    /* indices = max_index + arange(N)                                   ev_graph.py:82-83 */
    const int64_t base = g->max_index;
    g->max_index += N;

    /* ---- insert ALL events before any search (ev_graph.py:85 precedes :90) ---- */
    if (N > 1) {                                                     /* graph/utils.py:7-14 */
        keyed_t *k = (keyed_t *)malloc((size_t)N * sizeof(keyed_t));
        int32_t *tmp = (int32_t *)malloc((size_t)N * sizeof(int32_t));
        if (!k || !tmp) { free(k); free(tmp); return -1; }
        for (int64_t i = 0; i < N; i++) {
            k[i].key = (int64_t)pos[3 * i] + (int64_t)W * pos[3 * i + 1] + (int64_t)W * H * batch[i];
            k[i].idx = (int32_t)(base + i);
        }
        qsort(k, (size_t)N, sizeof(keyed_t), cmp_keyed);             /* stable sort by pixel */
        int64_t s = 0;
        while (s < N) {
            int64_t e = s;
            while (e < N && k[e].key == k[s].key) e++;
            int counts = (int)(e - s);
            for (int c = 0; c < counts; c++) tmp[c] = k[s + c].idx;
            int64_t key = k[s].key;
            int x = (int)(key % W);                                   /* ev_graph.cu:196-198 */
            int y = (int)(((key - x) / W) % H);
            int b = (int)(key / ((int64_t)W * H));
            fifo_push(g, b, y, x, tmp, counts);
            s = e;
        }
        free(k); free(tmp);
    } else {
        /* single-event kernel ignores `batch`, b = 0                    ev_graph.cu:150-152 */
        int32_t idx = (int32_t)base;
        fifo_push(g, 0, pos[1], pos[0], &idx, 1);
    }

    /* ---- search (ev_graph.cu:15-80), output already compacted, dst ascending ---- */
    const int64_t min_index = g->min_index;
    int64_t E = 0;
    const int ncell = (2 * radius + 1) * (2 * radius + 1);
    for (int64_t i = 0; i < N; i++) {
        int b = batch[i], x = pos[3 * i], y = pos[3 * i + 1], ts_event = pos[3 * i + 2];
        int64_t own = base + i;
        int nn = 0;
        edges_src[E] = own - min_index; edges_dst[E] = own - min_index; E++; nn++;   /* :44-46 */
        /* SpiralOut state (spiral.h) */
        unsigned layer = 1, leg = 0; int sx = 0, sy = 0;
        for (int c = 0; c < ncell; c++) {
            if (nn >= K) break;                                                   /* :50 */
            for (int q = 0; q < Q; q++) {
                int xn = x + sx, yn = y + sy;
                if (!((xn >= 0) && (yn >= 0) && (xn < W) && (yn < H))) break;      /* :56 */
                size_t qi = (size_t)xn + (size_t)W * yn + (size_t)H * W * q + (size_t)H * W * Q * b;
                int32_t idx = g->queue[qi];
                if (idx < min_index) break;                                        /* :62 */
                if (own > idx) {                                                   /* :64 */
                    int32_t dt = ts_event - g->ts[idx - min_index];
                    if ((float)dt > (float)delta_t_us) continue;                   /* :66-69 */
                    edges_src[E] = idx - min_index; edges_dst[E] = own - min_index; E++; nn++;
                    if (nn >= K) break;                                            /* :74 */
                }
            }
            /* goNext (spiral.h:8-15) */
            switch (leg) {
            case 0: ++sx; if (sx == (int)layer) ++leg; break;
            case 1: ++sy; if (sy == (int)layer) ++leg; break;
            case 2: --sx; if (-sx == (int)layer) ++leg; break;
            case 3: --sy; if (-sy == (int)layer) { leg = 0; ++layer; } break;
            }
        }
    }
    return E;
}
