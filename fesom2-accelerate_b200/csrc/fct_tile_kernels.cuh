// Tile-staged fused kernels (the fast path of the device-resident step on manifold meshes).
//
// A CTA owns a *tile*: a run of consecutive entries of a node list (a compact patch of a
// space-filling-curve numbered mesh).  Everything a tile needs besides its own columns is described
// by contiguous, precomputed per-tile tables (plan): the rows to stage, one header per node and a
// fixed-width gather list per node in tile-local row indices.  The kernel
//   1. copies the tables to shared memory (connectivity staging, one coalesced pass) and, in the
//      same barrier interval,
//   2. stages the neighbour rows ONCE per tile in shared memory -- phase A: the a1 bounds
//      max/min(fct_LO, ttf) of the tile's own + halo nodes (the reference's a1 kernel fused in);
//      phase B: their fct_plus / fct_minus;
//   3. walks *work items* = (node, slot of VEC levels) pairs of ACTIVE levels only, so lanes are
//      not wasted on the part of a column below the sea floor; items of one node never straddle
//      a block iteration, which keeps the vertical stencil inside one barrier interval.
// Latency: every global load of an item (own columns, up to HB edge-flux rows) is issued before
// the shared-memory gather that precedes its first use, and each CTA prefetches into L2 the
// tables of the tile a later CTA will start with.
// On a triangulation the unique ring neighbours of a node are exactly the other ends of its edges,
// with the same depth (the elements holding both nodes are the edge's two elements), so ONE list
// in ascending edge order drives the a2/a3 bound gather, the b1h / c_h sums and the b3h limiter.
#pragma once
#include "fct_kernels.cuh"

namespace fct {

struct TileDev {
    const int *row_off;    // [ntiles+1] -> rows
    const int2 *rows;      // {node, active levels}: the tile's own nodes first, then its halo
    const int4 *hdr;       // [ntiles*TN] {node (-1: none), active levels, fillmin, self depth | entries << 16}
    const int *work_off;   // [ntiles*(TN+1)] first work item of each node; [TN] = items of the tile
    const int4 *ent;       // [ntiles*TN*TE] {edge, tile-local row of the other end, meta, other node}
    int TN, TE, max_rows, ntiles;
    int ahead;             // tiles between a CTA and the tile whose tables it prefetches into L2
};

constexpr int TILE_THREADS = 256;

__host__ __device__ inline size_t tile_smem_bytes(bool phaseA, int PW, int TN, int TE, int max_rows)
{
    size_t b = (size_t)2 * max_rows * PW * sizeof(double);              // staged rows (two arrays)
    if (phaseA) b += (size_t)2 * TN * (PW + 4) * sizeof(double);        // cluster bounds for the stencil
    b += (size_t)TN * TE * sizeof(int4) + (size_t)TN * sizeof(int4);    // gather lists, headers
    b += (size_t)(TN + 4) * sizeof(int);
    return (b + 15) & ~(size_t)15;
}

struct TileSmem {
    double *ra, *rb;     // staged rows
    double *tva, *tvb;   // phase A only
    int4 *ent, *hdr;
    int *wo;
};
__device__ __forceinline__ TileSmem tile_carve(unsigned char *raw, bool phaseA, int PW, const TileDev &T)
{
    TileSmem s;
    s.ra = reinterpret_cast<double *>(raw);
    s.rb = s.ra + (size_t)T.max_rows * PW;
    double *p = s.rb + (size_t)T.max_rows * PW;
    s.tva = s.tvb = nullptr;
    if (phaseA) {
        s.tva = p + 2;   // rows of PW + 4 doubles: [z-1] stays inside, [z0] stays 16-byte aligned
        s.tvb = s.tva + (size_t)T.TN * (PW + 4);
        p += (size_t)2 * T.TN * (PW + 4);
    }
    s.ent = reinterpret_cast<int4 *>(p);
    s.hdr = s.ent + (size_t)T.TN * T.TE;
    s.wo = reinterpret_cast<int *>(s.hdr + T.TN);
    return s;
}

__device__ __forceinline__ void prefetch_l2(const void *p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// copy the tile's tables to shared memory (visible after the next barrier) and pull the tables of
// tile + ahead into L2 for the CTA that will own it
__device__ __forceinline__ void tile_load_tables(const TileDev &T, const TileSmem &s, int tile)
{
    const int t = threadIdx.x;
    const int ne = T.TN * T.TE;
    const int4 *eg = T.ent + (size_t)tile * ne;
    for (int i = t; i < ne; i += TILE_THREADS) s.ent[i] = __ldg(eg + i);
    for (int i = t; i < T.TN; i += TILE_THREADS) s.hdr[i] = __ldg(T.hdr + (size_t)tile * T.TN + i);
    for (int i = t; i <= T.TN; i += TILE_THREADS) s.wo[i] = __ldg(T.work_off + (size_t)tile * (T.TN + 1) + i);
    const int fut = tile + T.ahead;
    if (T.ahead > 0 && fut < T.ntiles) {
        const char *fe = reinterpret_cast<const char *>(T.ent + (size_t)fut * ne);
        const int lines = (ne * (int)sizeof(int4) + 127) / 128;
        if (t < lines) prefetch_l2(fe + (size_t)t * 128);
        if (t == 32) prefetch_l2(T.hdr + (size_t)fut * T.TN);
        if (t == 33) prefetch_l2(T.work_off + (size_t)fut * (T.TN + 1));
        if (t == 34) prefetch_l2(T.rows + __ldg(T.row_off + fut));
        if (t == 35) prefetch_l2(reinterpret_cast<const char *>(T.rows + __ldg(T.row_off + fut)) + 128);
        if (t == 36) prefetch_l2(reinterpret_cast<const char *>(T.rows + __ldg(T.row_off + fut)) + 256);
    }
}

// tile-local node of work item i: largest ln with wo[ln] <= i
__device__ __forceinline__ int tile_find_node(const int *wo, int TN, int i)
{
    int lo = 0, hi = TN;   // invariant wo[lo] <= i < wo[hi] (wo[TN] > i for every live item)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (wo[mid] <= i) lo = mid;
        else hi = mid;
    }
    return lo;
}

template <int VEC>
__device__ __forceinline__ void lds_v(const double *p, double (&o)[VEC])
{
    if constexpr (VEC == 2) {
        const double2 t = *reinterpret_cast<const double2 *>(p);
        o[0] = t.x;
        o[1] = t.y;
    } else {
        o[0] = *p;
    }
}
template <int VEC>
__device__ __forceinline__ void sts_v(double *p, const double (&o)[VEC])
{
    if constexpr (VEC == 2) *reinterpret_cast<double2 *>(p) = make_double2(o[0], o[1]);
    else p[0] = o[0];
}

// Stage rows of two arrays (src_a, src_b) of the tile's own + halo nodes into smem.  PHASE_A
// converts (fct_LO, ttf) into the a1 bounds on the way (reference.cpp:315-316).
template <int VEC, bool PHASE_A>
__device__ __forceinline__ void tile_stage_rows(const Arrays &A, const TileDev &T, const TileSmem &s, int tile,
                                                const double *__restrict__ src_a, const double *__restrict__ src_b)
{
    const int PW = A.pitchL;
    const int t = threadIdx.x;
    const int CH = (A.nl - 1 + VEC - 1) / VEC;    // slots covering nl-1 levels
    const int rpp = TILE_THREADS / CH;             // rows per pass
    const int ty = t / CH, z0 = (t - ty * CH) * VEC;
    const size_t tb = blockIdx.y * A.ts_node + z0;
    const int r0 = __ldg(T.row_off + tile), U = __ldg(T.row_off + tile + 1) - r0;
    if (ty >= rpp) return;
#pragma unroll 6
    for (int u = ty; u < U; u += rpp) {
        const int2 rw = __ldg(T.rows + r0 + u);
        if (z0 < rw.y) {
            double a[VEC], b[VEC];
            const size_t r = tb + (size_t)rw.x * PW;
            ldv_ro<VEC>(src_a + r, a);
            ldv_ro<VEC>(src_b + r, b);
            if constexpr (PHASE_A) {
                double hi[VEC], lw[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    hi[v] = pick_max(a[v], b[v]);
                    lw[v] = pick_min(a[v], b[v]);
                }
                sts_v<VEC>(s.ra + (size_t)u * PW + z0, hi);
                sts_v<VEC>(s.rb + (size_t)u * PW + z0, lw);
            } else {
                sts_v<VEC>(s.ra + (size_t)u * PW + z0, a);
                sts_v<VEC>(s.rb + (size_t)u * PW + z0, b);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Phase A = a1 + a2 + a3 + b1 vertical + b1 horizontal + b2
// HB: edge-flux rows per item loaded ahead of the gather (0: load at use)
// ------------------------------------------------------------------------------------------------
template <int VEC, int HB, int MINB>
__global__ void __launch_bounds__(TILE_THREADS, MINB) k_phaseA_tile(Arrays A, TileDev T)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const int PW = A.pitchL;
    const TileSmem s = tile_carve(smraw, true, PW, T);
    const int t = threadIdx.x;
    tile_load_tables(T, s, blockIdx.x);
    tile_stage_rows<VEC, true>(A, T, s, blockIdx.x, A.lo, A.ttf);
    __syncthreads();

    const int nwork = s.wo[T.TN];
    const int W = PW + 4;
    for (int i0 = 0; i0 < nwork; i0 += TILE_THREADS) {
        const int i = i0 + t;
        int ln = 0, z0 = 0, nz = 0, n = 0, cnt = 0;
        bool act = false;
        size_t off = 0;
        double l[VEC], ai[VEC], f[VEC + 1];
        double hpre[HB > 0 ? HB : 1][VEC];
        if (i < nwork) {
            ln = tile_find_node(s.wo, T.TN, i);
            const int4 hd = s.hdr[ln];
            n = hd.x;
            nz = hd.y;
            z0 = (i - s.wo[ln]) * VEC;
            act = n >= 0 && z0 < nz;
            if (act) {
                const int4 *en = s.ent + ln * T.TE;
                // ---- issue every global load of this item now; first use is after the gather ----
                off = blockIdx.y * A.ts_node + (size_t)n * PW + z0;
                ldv_ro<VEC>(A.lo + off, l);
                ldv_ro<VEC>(A.area_inv + (size_t)n * A.pitchV + z0, ai);
                {
                    const double *vrow = A.adf_v + blockIdx.y * A.ts_nodev + (size_t)n * A.pitchV + z0;
                    if constexpr (VEC == 2) {
                        const double2 q = *reinterpret_cast<const double2 *>(vrow);
                        f[0] = q.x;
                        f[1] = q.y;
                        f[2] = (z0 + 2 <= nz) ? vrow[2] : 0.0;
                    } else {
#pragma unroll
                        for (int v = 0; v <= VEC; ++v) f[v] = (z0 + v <= nz) ? vrow[v] : 0.0;
                    }
                }
                if constexpr (HB > 0) {
                    const double *hb = A.adf_h_in + blockIdx.y * A.ts_edge + z0;
#pragma unroll
                    for (int k = 0; k < HB; ++k) {
                        if (k < T.TE) {
                            const int4 e = en[k];   // padding entries have depth 0
                            if (z0 < FCT_META_DEPTH(e.z)) ldv_ro<VEC>(hb + (size_t)e.x * A.pitchH, hpre[k]);
                        }
                    }
                }
                // ---- cluster bounds of this slot: own a1 bounds + the other end of every edge ----
                double hi[VEC], lw[VEC];
                const int sd = hd.w & 0xffff;
                cnt = hd.w >> 16;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    const bool fl = z0 + v >= hd.z;
                    hi[v] = fl ? -A.big : -CUDART_INF;
                    lw[v] = fl ? A.big : CUDART_INF;
                }
                if (z0 < sd) {
                    double x[VEC], y[VEC];
                    lds_v<VEC>(s.ra + (size_t)ln * PW + z0, x);
                    lds_v<VEC>(s.rb + (size_t)ln * PW + z0, y);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if (z0 + v < sd) {
                            hi[v] = pick_max(hi[v], x[v]);
                            lw[v] = pick_min(lw[v], y[v]);
                        }
                    }
                }
                for (int k = 0; k < cnt; ++k) {
                    const int4 e = en[k];
                    const int dg = FCT_META_DEPTH(e.z);
                    if (z0 < dg) {
                        double x[VEC], y[VEC];
                        lds_v<VEC>(s.ra + (size_t)e.y * PW + z0, x);
                        lds_v<VEC>(s.rb + (size_t)e.y * PW + z0, y);
#pragma unroll
                        for (int v = 0; v < VEC; ++v) {
                            if (z0 + v < dg) {
                                hi[v] = pick_max(hi[v], x[v]);
                                lw[v] = pick_min(lw[v], y[v]);
                            }
                        }
                    }
                }
                sts_v<VEC>(s.tva + (size_t)ln * W + z0, hi);
                sts_v<VEC>(s.tvb + (size_t)ln * W + z0, lw);
            }
        }
        __syncthreads();
        if (!act) continue;
        double bm[VEC], bn[VEC], p[VEC], m[VEC];
        {
            const double *tva = s.tva + (size_t)ln * W, *tvb = s.tvb + (size_t)ln * W;
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const int z = z0 + v;
                double x = tva[z], y = tvb[z];
                if (z > 0 && z < nz - 1) {
                    x = pick_max(pick_max(tva[z - 1], x), tva[z + 1]);
                    y = pick_min(pick_min(tvb[z - 1], y), tvb[z + 1]);
                }
                bm[v] = x - l[v];
                bn[v] = y - l[v];
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {   // b1 vertical, reference.cpp:397-398
            p[v] = pick_max(0., f[v]) + pick_max(0., -f[v + 1]);
            m[v] = pick_min(0., f[v]) + pick_min(0., -f[v + 1]);
        }
        {
            const double *hb = A.adf_h_in + blockIdx.y * A.ts_edge + z0;
            const int4 *en = s.ent + ln * T.TE;
            auto add_edge = [&](const int4 &e, const double(&h)[VEC]) {
                const int dg = FCT_META_DEPTH(e.z);
                const bool second = FCT_META_SECOND(e.z);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if (z0 + v < dg) {
                        const double q = second ? -h[v] : h[v];
                        p[v] += pick_max(0., q);
                        m[v] += pick_min(0., q);
                    }
                }
            };
            if constexpr (HB > 0) {
#pragma unroll
                for (int k = 0; k < HB; ++k)
                    if (k < cnt) add_edge(en[k], hpre[k]);
            }
            for (int k = HB; k < cnt; ++k) {
                const int4 e = en[k];
                if (z0 < FCT_META_DEPTH(e.z)) {
                    double h[VEC];
                    ldv_ro<VEC>(hb + (size_t)e.x * A.pitchH, h);
                    add_edge(e, h);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) b2_point(p[v], m[v], bm[v], bn[v], ai[v], A.dt, A.eps);
        const int c = min(VEC, nz - z0);
        stv<VEC>(A.ttf_max + off, bm, c);
        stv<VEC>(A.ttf_min + off, bn, c);
        stv<VEC>(A.plus + off, p, c);
        stv<VEC>(A.minus + off, m, c);
    }
}

// ------------------------------------------------------------------------------------------------
// Phase B = b3 vertical + b3 horizontal + c vertical + c horizontal
// ------------------------------------------------------------------------------------------------
template <int VEC, int HB, int MINB>
__global__ void __launch_bounds__(TILE_THREADS, MINB) k_phaseB_tile(Arrays A, TileDev T)
{
    extern __shared__ __align__(16) unsigned char smraw[];
    const int PW = A.pitchL;
    const TileSmem s = tile_carve(smraw, false, PW, T);
    const int t = threadIdx.x;
    tile_load_tables(T, s, blockIdx.x);
    tile_stage_rows<VEC, false>(A, T, s, blockIdx.x, A.plus, A.minus);
    __syncthreads();

    const int nwork = s.wo[T.TN];
    for (int i0 = 0; i0 < nwork; i0 += TILE_THREADS) {
        const int i = i0 + t;
        bool act = false;
        int nz = 0, z0 = 0;
        double fl[VEC + 1], dh[VEC], dv[VEC];
        size_t off = 0;
        double *vrow = nullptr;
        if (i < nwork) {
            const int ln = tile_find_node(s.wo, T.TN, i);
            const int4 hd = s.hdr[ln];
            const int n = hd.x;
            nz = hd.y;
            z0 = (i - s.wo[ln]) * VEC;
            act = n >= 0 && z0 < nz;
            if (act) {
                const int cnt = hd.w >> 16;
                const int4 *en = s.ent + ln * T.TE;
                off = blockIdx.y * A.ts_node + (size_t)n * PW + z0;
                const size_t offs = (size_t)n * PW + z0;
                vrow = A.adf_v + blockIdx.y * A.ts_nodev + (size_t)n * A.pitchV;
                const double *prow = s.ra + (size_t)ln * PW, *mrow = s.rb + (size_t)ln * PW;
                const double *hin = A.adf_h_in + blockIdx.y * A.ts_edge + z0;
                double *ho = A.adf_h_out + blockIdx.y * A.ts_edge + z0;
                double x[VEC], l[VEC], hn[VEC], hw[VEC], ar[VEC], pn[VEC], mn[VEC], fr[VEC + 1];
                double hpre[HB > 0 ? HB : 1][VEC];
                // ---- every global load of this item, issued together ----
                ldv<VEC>(A.del_v + off, dv);
                ldv<VEC>(A.del_h + off, dh);
                ldv_ro<VEC>(A.ttf + off, x);
                ldv_ro<VEC>(A.lo + off, l);
                ldv_ro<VEC>(A.hnode + offs, hn);
                ldv_ro<VEC>(A.hnode_new + offs, hw);
                ldv_ro<VEC>(A.area + (size_t)n * A.pitchV + z0, ar);
                if constexpr (VEC == 2) {
                    const double2 q = *reinterpret_cast<const double2 *>(vrow + z0);
                    fr[0] = q.x;
                    fr[1] = q.y;
                    fr[2] = (z0 + 2 <= nz) ? vrow[z0 + 2] : 0.0;
                } else {
#pragma unroll
                    for (int v = 0; v <= VEC; ++v) fr[v] = (z0 + v <= nz) ? vrow[z0 + v] : 0.0;
                }
                if constexpr (HB > 0) {
#pragma unroll
                    for (int k = 0; k < HB; ++k) {
                        if (k < T.TE) {
                            const int4 e = en[k];
                            if (z0 < FCT_META_DEPTH(e.z)) ldv_ro<VEC>(hin + (size_t)e.x * A.pitchH, hpre[k]);
                        }
                    }
                }
                // ---- b3 vertical + c vertical ----
#pragma unroll
                for (int v = 0; v <= VEC; ++v) {
                    const int z = z0 + v;
                    fl[v] = (z < nz) ? b3v_point(fr[v], z, prow, mrow) : fr[v];   // the bottom flux stays
                }
                lds_v<VEC>(prow + z0, pn);
                lds_v<VEC>(mrow + z0, mn);
#pragma unroll
                for (int v = 0; v < VEC; ++v) ar[v] = A.dt / ar[v];
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    dv[v] = dv[v] - x[v] * hn[v] + l[v] * hw[v] + (fl[v] - fl[v + 1]) * ar[v];
                // ---- b3 horizontal + c horizontal over the node's edges, ascending edge id ----
                auto do_edge = [&](const int4 &e, double(&h)[VEC]) {
                    const int dg = FCT_META_DEPTH(e.z);
                    double po[VEC], mo[VEC];
                    lds_v<VEC>(s.ra + (size_t)e.y * PW + z0, po);
                    lds_v<VEC>(s.rb + (size_t)e.y * PW + z0, mo);
                    const bool second = FCT_META_SECOND(e.z);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if (z0 + v < dg) {
                            h[v] = second ? b3h_point(h[v], po[v], mo[v], pn[v], mn[v])
                                          : b3h_point(h[v], pn[v], mn[v], po[v], mo[v]);
                            const double q = h[v] * ar[v];
                            dh[v] = second ? dh[v] - q : dh[v] + q;
                        }
                    }
                    if (FCT_META_WRITER(e.z)) stv<VEC>(ho + (size_t)e.x * A.pitchH, h, min(VEC, dg - z0));
                };
                if constexpr (HB > 0) {
#pragma unroll
                    for (int k = 0; k < HB; ++k) {
                        if (k < cnt) {
                            const int4 e = en[k];
                            if (z0 < FCT_META_DEPTH(e.z)) do_edge(e, hpre[k]);
                        }
                    }
                }
                for (int k = HB; k < cnt; ++k) {
                    const int4 e = en[k];
                    if (z0 < FCT_META_DEPTH(e.z)) {
                        double h[VEC];
                        ldv_ro<VEC>(hin + (size_t)e.x * A.pitchH, h);
                        do_edge(e, h);
                    }
                }
            }
        }
        __syncthreads();   // raw fct_adf_v reads of this iteration's columns precede their in-place update
        if (!act) continue;
        const int c = min(VEC, nz - z0);
        double fo[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) fo[v] = fl[v];
        stv<VEC>(A.adf_v_out + (vrow - A.adf_v) + z0, fo, c);
        stv<VEC>(A.del_v + off, dv, c);
        stv<VEC>(A.del_h + off, dh, c);
    }
}

}   // namespace fct
