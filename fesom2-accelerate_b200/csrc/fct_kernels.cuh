// fct_ale kernels for sm_100a (fp64, HBM-bound; no tensor cores -- nothing here is a contraction).
//
// Thread mapping shared by every kernel: blockDim = (LX, NY).  threadIdx.x is a *level slot* of VEC
// consecutive levels (VEC = 2 -> one aligned double2 per array per thread on the padded rows of the
// device-resident path, VEC = 1 -> dense odd-pitch rows of the legacy handle ABI), threadIdx.y is
// the node / element / edge inside the block.  LX covers one whole column, so a block owns NY
// complete consecutive columns: every global access is a contiguous run of a row (coalesced), the
// vertical 3-point stencils stay inside the block (shared memory), and consecutive columns of a
// space-filling-curve numbered mesh share their neighbour rows through L1/L2.
//
// Arithmetic follows the oracle's operation order exactly (compare-select max/min like std::max /
// std::min, no FMA contraction: the library is compiled with -fmad=false), so results are
// bit-identical to src/reference.cpp for a1..b2 and to the Fortran order for b3 / c.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

namespace fct {

struct Arrays {
    // per-tracer arrays: pointer to tracer 0, blockIdx.y selects the tracer
    const double *ttf, *lo;
    double *adf_v;
    double *adf_v_out;   // limited vertical fluxes of the fused kernels (== adf_v: in place)
    const double *adf_h_in;
    double *adf_h_out;
    double *ttf_max, *ttf_min, *plus, *minus, *del_v, *del_h;
    double2 *uv_rhs;
    size_t ts_node, ts_nodev, ts_edge, ts_uv;   // tracer strides (doubles; double2 for uv)
    // mesh-static arrays (shared by all tracers)
    const double *area, *area_inv, *hnode, *hnode_new;
    int flags;   // warp-item kernels (results unaffected): 1 = no per-item L2 prefetch of the c-vertical operands (timing
                 // experiment), 2 = area_inv probed one item ahead in phase A, 4 = set by the launcher when the issuers
                 // prefetch whole tiles (WT_OPT 128): the consumers skip their per-item probes
    int pitchL;   // row pitch of [node][nl-1] arrays
    int pitchV;   // row pitch of [node][nl] arrays (fct_adf_v, area, area_inv)
    int pitchH;   // row pitch of fct_adf_h
    int pitchU;   // row pitch of UV_rhs (double2 units)
    int nl;
    double dt, eps, big;
    // vlimit 2 / 3 and the iterative branch (docs/refactoring.md:113-148, :226-290)
    int vlimit;
    double *adf_v2, *adf_h2;   // rejected flux parts of b3 with iter_yn
};

struct MeshDev {
    const int *nlev_n, *nlev_e, *elem_nodes, *nie, *nie_num;
    int nie_dim;
    const int *edges, *edge_tri;
    // derived gather lists (plan)
    const int *nbr_off;
    const int2 *nbr;      // {node, depth}; entry 0 of every node is the node itself
    const int *fillmin;   // first level that sees the (-big, +big) fill of a ring element
    const int *edg_off;
    const int4 *edg;      // {edge, other node, meta, 0}; ascending edge id
};

// meta word of an edge entry
#define FCT_META_DEPTH(m) ((m) & 0xffff)
#define FCT_META_SECOND(m) (((m) >> 16) & 1)   // this node is edges[2g+1]
#define FCT_META_WRITER(m) (((m) >> 17) & 1)   // this node stores the limited flux of the edge

__device__ __forceinline__ double pick_max(double a, double b) { return (a < b) ? b : a; }
__device__ __forceinline__ double pick_min(double a, double b) { return (b < a) ? b : a; }

template <int VEC>
__device__ __forceinline__ void ldv(const double *p, double (&o)[VEC])
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
__device__ __forceinline__ void ldv_ro(const double *__restrict__ p, double (&o)[VEC])
{
    if constexpr (VEC == 2) {
        const double2 t = __ldg(reinterpret_cast<const double2 *>(p));
        o[0] = t.x;
        o[1] = t.y;
    } else {
        o[0] = __ldg(p);
    }
}
// store the first `cnt` (1..VEC) components
template <int VEC>
__device__ __forceinline__ void stv(double *p, const double (&o)[VEC], int cnt)
{
    if constexpr (VEC == 2) {
        if (cnt >= 2) *reinterpret_cast<double2 *>(p) = make_double2(o[0], o[1]);
        else if (cnt == 1) p[0] = o[0];
    } else {
        if (cnt >= 1) p[0] = o[0];
    }
}

struct Item {
    int idx;      // node / element / edge id (0-based), -1 when the thread has no item
    int z0;       // first level of this thread's slot
};
__device__ __forceinline__ Item my_item(const int *__restrict__ list, int first, int count, int vec)
{
    Item it;
    const int li = blockIdx.x * blockDim.y + threadIdx.y;
    it.idx = (li < count) ? (list ? __ldg(list + first + li) : first + li) : -1;
    it.z0 = threadIdx.x * vec;
    return it;
}

// ------------------------------------------------------------------------------------------------
// Stage kernels: one per reference kernel (kernels/fct_ale_*.cu), same inputs / outputs.
// ------------------------------------------------------------------------------------------------

// a1 -- reference.cpp:306-319
template <int VEC>
__global__ void k_a1(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int nz = __ldg(M.nlev_n + it.idx) - 1;
    if (it.z0 >= nz) return;
    const size_t off = blockIdx.y * A.ts_node + (size_t)it.idx * A.pitchL + it.z0;
    double l[VEC], t[VEC], hi[VEC], lw[VEC];
    ldv_ro<VEC>(A.lo + off, l);
    ldv_ro<VEC>(A.ttf + off, t);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        hi[v] = pick_max(l[v], t[v]);
        lw[v] = pick_min(l[v], t[v]);
    }
    const int cnt = min(VEC, nz - it.z0);
    stv<VEC>(A.ttf_max + off, hi, cnt);
    stv<VEC>(A.ttf_min + off, lw, cnt);
}

// a2 -- reference.cpp:321-351 (element-centric, materialises UV_rhs)
template <int VEC>
__global__ void k_a2(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int L = A.nl - 1;
    if (it.z0 >= L) return;
    const int e = it.idx;
    const int nlev = __ldg(M.nlev_e + e);
    const int nz = nlev - 1;
    double2 *uv = A.uv_rhs + blockIdx.y * A.ts_uv + (size_t)e * A.pitchU + it.z0;
    double hi[VEC], lw[VEC];
    if (it.z0 < nz) {
        const size_t tb = blockIdx.y * A.ts_node + it.z0;
        const size_t r0 = tb + (size_t)(__ldg(M.elem_nodes + 3 * e + 0) - 1) * A.pitchL;
        const size_t r1 = tb + (size_t)(__ldg(M.elem_nodes + 3 * e + 1) - 1) * A.pitchL;
        const size_t r2 = tb + (size_t)(__ldg(M.elem_nodes + 3 * e + 2) - 1) * A.pitchL;
        double a0[VEC], a1[VEC], a2[VEC], b0[VEC], b1[VEC], b2[VEC];
        ldv<VEC>(A.ttf_max + r0, a0);
        ldv<VEC>(A.ttf_max + r1, a1);
        ldv<VEC>(A.ttf_max + r2, a2);
        ldv<VEC>(A.ttf_min + r0, b0);
        ldv<VEC>(A.ttf_min + r1, b1);
        ldv<VEC>(A.ttf_min + r2, b2);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            hi[v] = pick_max(pick_max(a0[v], a1[v]), a2[v]);
            lw[v] = pick_min(pick_min(b0[v], b1[v]), b2[v]);
        }
    }
    const bool fill = nlev <= L;   // reference.cpp:341
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int z = it.z0 + v;
        if (z < nz) uv[v] = make_double2(hi[v], lw[v]);
        else if (z < L && fill) uv[v] = make_double2(-A.big, A.big);
    }
}

// cluster bounds of one node slot: max / min of UV_rhs over the ring elements (reference.cpp:362-378)
template <int VEC>
__device__ __forceinline__ void a3_ring_bounds(const Arrays &A, const MeshDev &M, int n, int z0, int nz,
                                               double (&hi)[VEC], double (&lw)[VEC])
{
    const int *ring = M.nie + (size_t)n * M.nie_dim;
    const int cnt = __ldg(M.nie_num + n);
    const double2 *uvb = A.uv_rhs + blockIdx.y * A.ts_uv + z0;
    {
        const double2 *uv = uvb + (size_t)(__ldg(ring) - 1) * A.pitchU;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (z0 + v < nz) {
                const double2 t = uv[v];
                hi[v] = t.x;
                lw[v] = t.y;
            }
        }
    }
    for (int k = 1; k < cnt; ++k) {
        const double2 *uv = uvb + (size_t)(__ldg(ring + k) - 1) * A.pitchU;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (z0 + v < nz) {
                const double2 t = uv[v];
                hi[v] = pick_max(hi[v], t.x);
                lw[v] = pick_min(lw[v], t.y);
            }
        }
    }
}

// a3 -- reference.cpp:353-392 (bounds only; b1 vertical is its own kernel like kernels/fct_ale_b1_vertical.cu)
// dynamic smem: NY * 2 * (LX*VEC + 2) doubles
template <int VEC>
__global__ void k_a3(Arrays A, MeshDev M, const int *list, int first, int count)
{
    extern __shared__ double sm[];
    const int W = blockDim.x * VEC + 2;
    double *tvmax = sm + (size_t)threadIdx.y * 2 * W + 1;
    double *tvmin = tvmax + W;
    const Item it = my_item(list, first, count, VEC);
    const int n = it.idx;
    const int nz = (n >= 0) ? __ldg(M.nlev_n + n) - 1 : 0;
    const bool act = it.z0 < nz;
    double hi[VEC], lw[VEC];
    if (act) {
        a3_ring_bounds<VEC>(A, M, n, it.z0, nz, hi, lw);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            if (it.z0 + v < nz) {
                tvmax[it.z0 + v] = hi[v];
                tvmin[it.z0 + v] = lw[v];
            }
        }
    }
    __syncthreads();
    if (!act) return;
    const size_t off = blockIdx.y * A.ts_node + (size_t)n * A.pitchL + it.z0;
    double l[VEC], omax[VEC], omin[VEC];
    ldv_ro<VEC>(A.lo + off, l);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int z = it.z0 + v;
        if (z < nz) {
            double bm = tvmax[z], bn = tvmin[z];
            if (z > 0 && z < nz - 1) {
                bm = pick_max(pick_max(tvmax[z - 1], bm), tvmax[z + 1]);
                bn = pick_min(pick_min(tvmin[z - 1], bn), tvmin[z + 1]);
            }
            omax[v] = bm - l[v];
            omin[v] = bn - l[v];
        }
    }
    const int cnt = min(VEC, nz - it.z0);
    stv<VEC>(A.ttf_max + off, omax, cnt);
    stv<VEC>(A.ttf_min + off, omin, cnt);
}

// a3 with vlimit == 2 (docs/refactoring.md:113-129) or 3 (md:131-148): the horizontal cluster bound
// of a level is widened (2) or narrowed (3) by the node's own a1 maxima of levels z-1..z+1 -- the
// listing takes maxval AND minval from fct_ttf_max (md:120-121, :139-140), restated as written.
// No executable form exists in the reference (reference.cpp:51-96 stubs, fct_ale_a3.py:152-155 pass).
// dynamic smem as k_a3 (one array used): the own a1 maxima, read before any level is overwritten
template <int VEC>
__global__ void k_a3_vlimit(Arrays A, MeshDev M, const int *list, int first, int count)
{
    extern __shared__ double sm[];
    const int W = blockDim.x * VEC + 2;
    double *amax = sm + (size_t)threadIdx.y * 2 * W + 1;
    const Item it = my_item(list, first, count, VEC);
    const int n = it.idx;
    const int nz = (n >= 0) ? __ldg(M.nlev_n + n) - 1 : 0;
    const bool act = it.z0 < nz;
    const size_t off = act ? blockIdx.y * A.ts_node + (size_t)n * A.pitchL + it.z0 : 0;
    double hi[VEC], lw[VEC];
    if (act) {
        a3_ring_bounds<VEC>(A, M, n, it.z0, nz, hi, lw);
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (it.z0 + v < nz) amax[it.z0 + v] = A.ttf_max[off + v];
    }
    __syncthreads();
    if (!act) return;
    double l[VEC], omax[VEC], omin[VEC];
    ldv_ro<VEC>(A.lo + off, l);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int z = it.z0 + v;
        if (z < nz) {
            double bm = hi[v], bn = lw[v];
            if (z > 0 && z < nz - 1) {
                const double vmax = pick_max(pick_max(amax[z - 1], amax[z]), amax[z + 1]);
                const double vmin = pick_min(pick_min(amax[z - 1], amax[z]), amax[z + 1]);
                if (A.vlimit == 2) {
                    bm = pick_max(bm, vmax);
                    bn = pick_min(bn, vmin);
                } else {
                    bm = pick_min(bm, vmax);
                    bn = pick_max(bn, vmin);
                }
            }
            omax[v] = bm - l[v];
            omin[v] = bn - l[v];
        }
    }
    const int cnt = min(VEC, nz - it.z0);
    stv<VEC>(A.ttf_max + off, omax, cnt);
    stv<VEC>(A.ttf_min + off, omin, cnt);
}

// vertical antidiffusive sums of one slot: plus/minus[v] for levels z0..z0+VEC-1 (reference.cpp:393-399)
template <int VEC>
__device__ __forceinline__ void b1v_slot(const double *vrow, int z0, int nz, double (&p)[VEC], double (&m)[VEC])
{
    double f[VEC + 1];
#pragma unroll
    for (int v = 0; v <= VEC; ++v) f[v] = (z0 + v <= nz) ? vrow[z0 + v] : 0.0;
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        p[v] = pick_max(0., f[v]) + pick_max(0., -f[v + 1]);
        m[v] = pick_min(0., f[v]) + pick_min(0., -f[v + 1]);
    }
}

// b1 vertical -- reference.cpp:393-399
template <int VEC>
__global__ void k_b1v(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int nz = __ldg(M.nlev_n + it.idx) - 1;
    if (it.z0 >= nz) return;
    const double *vrow = A.adf_v + blockIdx.y * A.ts_nodev + (size_t)it.idx * A.pitchV;
    double p[VEC], m[VEC];
    b1v_slot<VEC>(vrow, it.z0, nz, p, m);
    const size_t off = blockIdx.y * A.ts_node + (size_t)it.idx * A.pitchL + it.z0;
    const int cnt = min(VEC, nz - it.z0);
    stv<VEC>(A.plus + off, p, cnt);
    stv<VEC>(A.minus + off, m, cnt);
}

// gather of the horizontal antidiffusive sums of one node slot, ascending edge order
// (deterministic replacement of the 4 atomicAdd of kernels/fct_ale_b1_horizontal.cu:24-27;
//  same summation order as the sequential edge loop reference.cpp:406-425)
template <int VEC>
__device__ __forceinline__ void b1h_gather(const Arrays &A, const MeshDev &M, int n, int z0, int nz,
                                           double (&p)[VEC], double (&m)[VEC])
{
    const double *hb = A.adf_h_in + blockIdx.y * A.ts_edge + z0;
    const int b = __ldg(M.edg_off + n), e = __ldg(M.edg_off + n + 1);
    for (int k = b; k < e; ++k) {
        const int4 en = __ldg(M.edg + k);
        const int dg = FCT_META_DEPTH(en.z);
        if (z0 < dg) {
            double h[VEC];
            ldv_ro<VEC>(hb + (size_t)en.x * A.pitchH, h);
            const bool second = FCT_META_SECOND(en.z);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (z0 + v < dg) {
                    const double f = second ? -h[v] : h[v];
                    p[v] += pick_max(0., f);
                    m[v] += pick_min(0., f);
                }
            }
        }
    }
}

// b1 horizontal -- reference.cpp:406-425, node-centric
template <int VEC>
__global__ void k_b1h(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int nz = __ldg(M.nlev_n + it.idx) - 1;
    if (it.z0 >= nz) return;
    const size_t off = blockIdx.y * A.ts_node + (size_t)it.idx * A.pitchL + it.z0;
    double p[VEC], m[VEC];
    ldv<VEC>(A.plus + off, p);
    ldv<VEC>(A.minus + off, m);
    b1h_gather<VEC>(A, M, it.idx, it.z0, nz, p, m);
    const int cnt = min(VEC, nz - it.z0);
    stv<VEC>(A.plus + off, p, cnt);
    stv<VEC>(A.minus + off, m, cnt);
}

// Zalesak factors of one value pair -- reference.cpp:432-435
__device__ __forceinline__ void b2_point(double &p, double &m, double bmax, double bmin, double ai,
                                         double dt, double eps)
{
    double flux = p * dt * ai + eps;
    p = pick_min(1., bmax / flux);
    flux = m * dt * ai - eps;
    m = pick_min(1., bmin / flux);
}

// b2 -- reference.cpp:426-437
template <int VEC>
__global__ void k_b2(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int nz = __ldg(M.nlev_n + it.idx) - 1;
    if (it.z0 >= nz) return;
    const size_t off = blockIdx.y * A.ts_node + (size_t)it.idx * A.pitchL + it.z0;
    double p[VEC], m[VEC], bm[VEC], bn[VEC], ai[VEC];
    ldv<VEC>(A.plus + off, p);
    ldv<VEC>(A.minus + off, m);
    ldv_ro<VEC>(A.ttf_max + off, bm);
    ldv_ro<VEC>(A.ttf_min + off, bn);
    ldv_ro<VEC>(A.area_inv + (size_t)it.idx * A.pitchV + it.z0, ai);
#pragma unroll
    for (int v = 0; v < VEC; ++v) b2_point(p[v], m[v], bm[v], bn[v], ai[v], A.dt, A.eps);
    const int cnt = min(VEC, nz - it.z0);
    stv<VEC>(A.plus + off, p, cnt);
    stv<VEC>(A.minus + off, m, cnt);
}

// limited vertical flux at level z of a column (docs/refactoring.md:205-231): raw flux f,
// own-column factors p/m (rows of fct_plus / fct_minus)
__device__ __forceinline__ double b3v_factor(double f, int z, const double *p, const double *m)
{
    double ae = 1.;
    if (z == 0) {
        if (f >= 0.) ae = pick_min(ae, p[0]);
        else ae = pick_min(ae, m[0]);
    } else if (f >= 0.) {
        ae = pick_min(ae, m[z - 1]);
        ae = pick_min(ae, p[z]);
    } else {
        ae = pick_min(ae, p[z - 1]);
        ae = pick_min(ae, m[z]);
    }
    return ae;
}
__device__ __forceinline__ double b3v_point(double f, int z, const double *p, const double *m)
{
    return b3v_factor(f, z, p, m) * f;
}

// b3 vertical -- docs/refactoring.md:205-233 (in place; each level depends on its own raw flux only)
template <int VEC>
__global__ void k_b3v(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int nz = __ldg(M.nlev_n + it.idx) - 1;
    if (it.z0 >= nz) return;
    double *vrow = A.adf_v + blockIdx.y * A.ts_nodev + (size_t)it.idx * A.pitchV;
    const size_t off = blockIdx.y * A.ts_node + (size_t)it.idx * A.pitchL;
    const double *p = A.plus + off, *m = A.minus + off;
    double f[VEC];
    ldv<VEC>(vrow + it.z0, f);
#pragma unroll
    for (int v = 0; v < VEC; ++v)
        if (it.z0 + v < nz) f[v] = b3v_point(f[v], it.z0 + v, p, m);
    stv<VEC>(vrow + it.z0, f, min(VEC, nz - it.z0));
}

// limiter factor times flux of one edge level (docs/refactoring.md:246-261).
// p1/m1 belong to edges[2g], p2/m2 to edges[2g+1].
__device__ __forceinline__ double b3h_factor(double h, double p1, double m1, double p2, double m2)
{
    double ae = 1.;
    if (h >= 0.) {
        ae = pick_min(ae, p1);
        ae = pick_min(ae, m2);
    } else {
        ae = pick_min(ae, m1);
        ae = pick_min(ae, p2);
    }
    return ae;
}
__device__ __forceinline__ double b3h_point(double h, double p1, double m1, double p2, double m2)
{
    return b3h_factor(h, p1, m1, p2, m2) * h;
}

__device__ __forceinline__ int edge_depth_dev(const MeshDev &M, int g)
{
    const int el = __ldg(M.edge_tri + 2 * g) - 1, er = __ldg(M.edge_tri + 2 * g + 1) - 1;
    const int d1 = __ldg(M.nlev_e + el) - 1;
    const int d2 = (er >= 0) ? __ldg(M.nlev_e + er) - 1 : 0;
    return max(d1, d2);
}

// b3 horizontal -- docs/refactoring.md:238-263, edge-centric like kernels/fct_ale_b3_horizontal.cu
// reads adf_h_in, writes adf_h_out (the same buffer in the staged mode: in place)
template <int VEC>
__global__ void k_b3h(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int g = it.idx;
    const int dg = edge_depth_dev(M, g);
    if (it.z0 >= dg) return;
    const size_t tb = blockIdx.y * A.ts_node + it.z0;
    const size_t r1 = tb + (size_t)(__ldg(M.edges + 2 * g) - 1) * A.pitchL;
    const size_t r2 = tb + (size_t)(__ldg(M.edges + 2 * g + 1) - 1) * A.pitchL;
    const size_t ho = blockIdx.y * A.ts_edge + (size_t)g * A.pitchH + it.z0;
    double h[VEC], p1[VEC], m1[VEC], p2[VEC], m2[VEC];
    ldv<VEC>(A.adf_h_in + ho, h);
    ldv_ro<VEC>(A.plus + r1, p1);
    ldv_ro<VEC>(A.minus + r1, m1);
    ldv_ro<VEC>(A.plus + r2, p2);
    ldv_ro<VEC>(A.minus + r2, m2);
#pragma unroll
    for (int v = 0; v < VEC; ++v)
        if (it.z0 + v < dg) h[v] = b3h_point(h[v], p1[v], m1[v], p2[v], m2[v]);
    stv<VEC>(A.adf_h_out + ho, h, min(VEC, dg - it.z0));
}

// c vertical -- docs/refactoring.md:295-300, flux term grouped x*(dt/area) as kernels/fct_ale_c_vertical.cu:12
template <int VEC>
__global__ void k_cv(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int nz = __ldg(M.nlev_n + it.idx) - 1;
    if (it.z0 >= nz) return;
    const size_t off = blockIdx.y * A.ts_node + (size_t)it.idx * A.pitchL + it.z0;
    const size_t offs = (size_t)it.idx * A.pitchL + it.z0;
    const double *vrow = A.adf_v + blockIdx.y * A.ts_nodev + (size_t)it.idx * A.pitchV;
    double d[VEC], t[VEC], l[VEC], hn[VEC], hw[VEC], ar[VEC], f[VEC + 1];
    ldv<VEC>(A.del_v + off, d);
    ldv_ro<VEC>(A.ttf + off, t);
    ldv_ro<VEC>(A.lo + off, l);
    ldv_ro<VEC>(A.hnode + offs, hn);
    ldv_ro<VEC>(A.hnode_new + offs, hw);
    ldv_ro<VEC>(A.area + (size_t)it.idx * A.pitchV + it.z0, ar);
#pragma unroll
    for (int v = 0; v <= VEC; ++v) f[v] = (it.z0 + v <= nz) ? vrow[it.z0 + v] : 0.0;
#pragma unroll
    for (int v = 0; v < VEC; ++v)
        d[v] = d[v] - t[v] * hn[v] + l[v] * hw[v] + (f[v] - f[v + 1]) * (A.dt / ar[v]);
    stv<VEC>(A.del_v + off, d, min(VEC, nz - it.z0));
}

// c horizontal -- docs/refactoring.md:303-314, node-centric ascending-edge gather instead of the
// 2 atomicAdd of kernels/fct_ale_c_horizontal.cu:25-26.  Reads the limited fluxes from adf_h_out.
template <int VEC>
__global__ void k_ch(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int n = it.idx;
    const int nz = __ldg(M.nlev_n + n) - 1;
    if (it.z0 >= nz) return;
    const size_t off = blockIdx.y * A.ts_node + (size_t)n * A.pitchL + it.z0;
    double d[VEC], ar[VEC];
    ldv<VEC>(A.del_h + off, d);
    ldv_ro<VEC>(A.area + (size_t)n * A.pitchV + it.z0, ar);
#pragma unroll
    for (int v = 0; v < VEC; ++v) ar[v] = A.dt / ar[v];   // dt/area once per cell (fct_ale_c_horizontal.cu:25)
    const double *hb = A.adf_h_out + blockIdx.y * A.ts_edge + it.z0;
    const int b = __ldg(M.edg_off + n), e = __ldg(M.edg_off + n + 1);
    for (int k = b; k < e; ++k) {
        const int4 en = __ldg(M.edg + k);
        const int dg = FCT_META_DEPTH(en.z);
        if (it.z0 < dg) {
            double h[VEC];
            ldv<VEC>(hb + (size_t)en.x * A.pitchH, h);
            const bool second = FCT_META_SECOND(en.z);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (it.z0 + v < dg) {
                    const double x = h[v] * ar[v];
                    d[v] = second ? d[v] - x : d[v] + x;
                }
            }
        }
    }
    stv<VEC>(A.del_h + off, d, min(VEC, nz - it.z0));
}

// ------------------------------------------------------------------------------------------------
// The iterative branch (iter_yn, docs/refactoring.md:226-290): b3 also
// keeps the rejected part of every flux, the limited fluxes update the low-order solution, and the
// rejected parts become the antidiffusive fluxes of the next pass.  No executable form exists in
// the reference; operation order is the listing's.
// ------------------------------------------------------------------------------------------------

// b3 vertical with iter_yn -- md:205-233: adf_v2 = (1-ae)*flux below the surface level (md:228-230)
template <int VEC>
__global__ void k_b3v_iter(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int nz = __ldg(M.nlev_n + it.idx) - 1;
    if (it.z0 >= nz) return;
    const size_t vo = blockIdx.y * A.ts_nodev + (size_t)it.idx * A.pitchV;
    double *vrow = A.adf_v + vo, *v2row = A.adf_v2 + vo;
    const size_t off = blockIdx.y * A.ts_node + (size_t)it.idx * A.pitchL;
    const double *p = A.plus + off, *m = A.minus + off;
    double f[VEC];
    ldv<VEC>(vrow + it.z0, f);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int z = it.z0 + v;
        if (z < nz) {
            const double ae = b3v_factor(f[v], z, p, m);
            if (z > 0) v2row[z] = (1.0 - ae) * f[v];
            f[v] = ae * f[v];
        }
    }
    stv<VEC>(vrow + it.z0, f, min(VEC, nz - it.z0));
}

// b3 horizontal with iter_yn -- md:238-263
template <int VEC>
__global__ void k_b3h_iter(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int g = it.idx;
    const int dg = edge_depth_dev(M, g);
    if (it.z0 >= dg) return;
    const size_t tb = blockIdx.y * A.ts_node + it.z0;
    const size_t r1 = tb + (size_t)(__ldg(M.edges + 2 * g) - 1) * A.pitchL;
    const size_t r2 = tb + (size_t)(__ldg(M.edges + 2 * g + 1) - 1) * A.pitchL;
    const size_t ho = blockIdx.y * A.ts_edge + (size_t)g * A.pitchH + it.z0;
    double h[VEC], h2[VEC], p1[VEC], m1[VEC], p2[VEC], m2[VEC];
    ldv<VEC>(A.adf_h_in + ho, h);
    ldv_ro<VEC>(A.plus + r1, p1);
    ldv_ro<VEC>(A.minus + r1, m1);
    ldv_ro<VEC>(A.plus + r2, p2);
    ldv_ro<VEC>(A.minus + r2, m2);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        if (it.z0 + v < dg) {
            const double ae = b3h_factor(h[v], p1[v], m1[v], p2[v], m2[v]);
            h2[v] = (1.0 - ae) * h[v];
            h[v] = ae * h[v];
        }
    }
    const int cnt = min(VEC, dg - it.z0);
    stv<VEC>(A.adf_h_out + ho, h, cnt);
    stv<VEC>(A.adf_h2 + ho, h2, cnt);
}

// "c. Update the LO" -- md:265-287, node-centric: the vertical term first, then the node's edges in
// ascending edge id (the order in which the listing's two loops reach this node); x*dt/area/hnode_new
// is evaluated left to right as written.  Owned nodes only: halo rows of fct_LO are the caller's
// exchange before the next pass, as in the Fortran.
template <int VEC>
__global__ void k_lo_update(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int n = it.idx;
    const int nz = __ldg(M.nlev_n + n) - 1;
    if (it.z0 >= nz) return;
    const size_t off = blockIdx.y * A.ts_node + (size_t)n * A.pitchL + it.z0;
    const double *vrow = A.adf_v + blockIdx.y * A.ts_nodev + (size_t)n * A.pitchV;
    double *lo = const_cast<double *>(A.lo);
    double l[VEC], ar[VEC], hw[VEC], f[VEC + 1];
    ldv<VEC>(lo + off, l);
    ldv_ro<VEC>(A.area + (size_t)n * A.pitchV + it.z0, ar);
    ldv_ro<VEC>(A.hnode_new + (size_t)n * A.pitchL + it.z0, hw);
#pragma unroll
    for (int v = 0; v <= VEC; ++v) f[v] = (it.z0 + v <= nz) ? vrow[it.z0 + v] : 0.0;
#pragma unroll
    for (int v = 0; v < VEC; ++v) l[v] = l[v] + (f[v] - f[v + 1]) * A.dt / ar[v] / hw[v];
    const double *hb = A.adf_h_out + blockIdx.y * A.ts_edge + it.z0;
    const int b = __ldg(M.edg_off + n), e = __ldg(M.edg_off + n + 1);
    for (int k = b; k < e; ++k) {
        const int4 en = __ldg(M.edg + k);
        const int dg = FCT_META_DEPTH(en.z);
        if (it.z0 < dg) {
            double h[VEC];
            ldv<VEC>(hb + (size_t)en.x * A.pitchH, h);
            const bool second = FCT_META_SECOND(en.z);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                if (it.z0 + v < dg) {
                    const double x = h[v] * A.dt / ar[v] / hw[v];
                    l[v] = second ? l[v] - x : l[v] + x;
                }
            }
        }
    }
    stv<VEC>(lo + off, l, min(VEC, nz - it.z0));
}

// ------------------------------------------------------------------------------------------------
// Measured alternative only: the reference's edge-centric scatter with fp64 atomics (RED.ADD.F64 on
// sm_100a).  Summation order depends on the schedule, so results are NOT bit-reproducible (parity
// within 1e-12 relative); the product path never launches these, tools/stage_sweep.py times them
// next to the gathers.
// ------------------------------------------------------------------------------------------------

// b1 horizontal as kernels/fct_ale_b1_horizontal.cu:3-29 -- 4 atomics per edge level
template <int VEC>
__global__ void k_b1h_atomic(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int g = it.idx;
    const int dg = edge_depth_dev(M, g);
    if (it.z0 >= dg) return;
    const size_t tb = blockIdx.y * A.ts_node + it.z0;
    const size_t r1 = tb + (size_t)(__ldg(M.edges + 2 * g) - 1) * A.pitchL;
    const size_t r2 = tb + (size_t)(__ldg(M.edges + 2 * g + 1) - 1) * A.pitchL;
    double h[VEC];
    ldv_ro<VEC>(A.adf_h_in + blockIdx.y * A.ts_edge + (size_t)g * A.pitchH + it.z0, h);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        if (it.z0 + v < dg) {
            atomicAdd(A.plus + r1 + v, pick_max(0., h[v]));
            atomicAdd(A.minus + r1 + v, pick_min(0., h[v]));
            atomicAdd(A.plus + r2 + v, pick_max(0., -h[v]));
            atomicAdd(A.minus + r2 + v, pick_min(0., -h[v]));
        }
    }
}

// c horizontal as kernels/fct_ale_c_horizontal.cu:3-28 -- 2 atomics and 2 divisions per edge level
template <int VEC>
__global__ void k_ch_atomic(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    if (it.idx < 0) return;
    const int g = it.idx;
    const int dg = edge_depth_dev(M, g);
    if (it.z0 >= dg) return;
    const int n1 = __ldg(M.edges + 2 * g) - 1, n2 = __ldg(M.edges + 2 * g + 1) - 1;
    const size_t r1 = blockIdx.y * A.ts_node + (size_t)n1 * A.pitchL + it.z0;
    const size_t r2 = blockIdx.y * A.ts_node + (size_t)n2 * A.pitchL + it.z0;
    double h[VEC];
    ldv<VEC>(A.adf_h_out + blockIdx.y * A.ts_edge + (size_t)g * A.pitchH + it.z0, h);
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        if (it.z0 + v < dg) {
            const double a1 = __ldg(A.area + (size_t)n1 * A.pitchV + it.z0 + v);
            const double a2 = __ldg(A.area + (size_t)n2 * A.pitchV + it.z0 + v);
            atomicAdd(A.del_h + r1 + v, h[v] * (A.dt / a1));
            atomicAdd(A.del_h + r2 + v, -(h[v] * (A.dt / a2)));
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Fused phase kernels (docs/fct_ale_dependencies.dot: the only global dependency inside the chain
// is the neighbours' fct_plus / fct_minus between b2 and b3 horizontal).
// ------------------------------------------------------------------------------------------------

// Phase A = a1 + a2 + a3 + b1 vertical + b1 horizontal + b2, node-centric.
// UV_rhs and the a1 bounds are never materialised: the cluster bound of a node is taken directly
// over the unique nodes of its ring elements (plan list nbr, depth = deepest ring element holding
// that node) plus the (-big, +big) fill from level fillmin on -- identical to a2 followed by a3.
// dynamic smem: NY * 2 * (LX*VEC + 2) doubles
template <int VEC>
__global__ void k_phaseA(Arrays A, MeshDev M, const int *list, int first, int count)
{
    extern __shared__ double sm[];
    const int W = blockDim.x * VEC + 2;
    double *tvmax = sm + (size_t)threadIdx.y * 2 * W + 1;
    double *tvmin = tvmax + W;
    const Item it = my_item(list, first, count, VEC);
    const int n = it.idx;
    const int z0 = it.z0;
    const int nz = (n >= 0) ? __ldg(M.nlev_n + n) - 1 : 0;
    const bool act = z0 < nz;
    const size_t tb = blockIdx.y * A.ts_node + z0;
    double l[VEC];
    if (act) {
        double hi[VEC], lw[VEC];
        const int fm = __ldg(M.fillmin + n);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const bool f = z0 + v >= fm;
            hi[v] = f ? -A.big : -CUDART_INF;
            lw[v] = f ? A.big : CUDART_INF;
        }
        const int b = __ldg(M.nbr_off + n), e = __ldg(M.nbr_off + n + 1);
        for (int k = b; k < e; ++k) {
            const int2 nb = __ldg(M.nbr + k);
            double ll[VEC], tt[VEC];
            if (k == b || z0 < nb.y) {
                const size_t r = tb + (size_t)nb.x * A.pitchL;
                ldv_ro<VEC>(A.lo + r, ll);
                if (k == b) {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) l[v] = ll[v];
                }
                if (z0 < nb.y) {
                    ldv_ro<VEC>(A.ttf + r, tt);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        if (z0 + v < nb.y) {
                            hi[v] = pick_max(hi[v], pick_max(ll[v], tt[v]));
                            lw[v] = pick_min(lw[v], pick_min(ll[v], tt[v]));
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            tvmax[z0 + v] = hi[v];
            tvmin[z0 + v] = lw[v];
        }
    }
    __syncthreads();
    if (!act) return;
    double bm[VEC], bn[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const int z = z0 + v;
        double x = tvmax[z], y = tvmin[z];
        if (z > 0 && z < nz - 1) {
            x = pick_max(pick_max(tvmax[z - 1], x), tvmax[z + 1]);
            y = pick_min(pick_min(tvmin[z - 1], y), tvmin[z + 1]);
        }
        bm[v] = x - l[v];
        bn[v] = y - l[v];
    }
    double p[VEC], m[VEC], ai[VEC];
    b1v_slot<VEC>(A.adf_v + blockIdx.y * A.ts_nodev + (size_t)n * A.pitchV, z0, nz, p, m);
    ldv_ro<VEC>(A.area_inv + (size_t)n * A.pitchV + z0, ai);
    b1h_gather<VEC>(A, M, n, z0, nz, p, m);
#pragma unroll
    for (int v = 0; v < VEC; ++v) b2_point(p[v], m[v], bm[v], bn[v], ai[v], A.dt, A.eps);
    const size_t off = tb + (size_t)n * A.pitchL;
    const int cnt = min(VEC, nz - z0);
    stv<VEC>(A.ttf_max + off, bm, cnt);
    stv<VEC>(A.ttf_min + off, bn, cnt);
    stv<VEC>(A.plus + off, p, cnt);
    stv<VEC>(A.minus + off, m, cnt);
}

// Phase B = b3 vertical + b3 horizontal + c vertical + c horizontal, node-centric.
// Every node recomputes the limited flux of each of its edges (both end nodes get the identical
// value) and accumulates it in ascending edge order; the end node flagged WRITER stores it to
// adf_h_out.  fct_adf_v is limited in place: all raw reads of a column happen before the block
// barrier, all writes after it.
template <int VEC>
__global__ void k_phaseB(Arrays A, MeshDev M, const int *list, int first, int count)
{
    const Item it = my_item(list, first, count, VEC);
    const int n = it.idx;
    const int z0 = it.z0;
    const int nz = (n >= 0) ? __ldg(M.nlev_n + n) - 1 : 0;
    const bool act = z0 < nz;
    double fl[VEC + 1];   // limited vertical fluxes z0 .. z0+VEC
    double dh[VEC], dv[VEC];
    size_t off = 0;
    double *vrow = nullptr;
    if (act) {
        off = blockIdx.y * A.ts_node + (size_t)n * A.pitchL + z0;
        const size_t offs = (size_t)n * A.pitchL + z0;
        vrow = A.adf_v + blockIdx.y * A.ts_nodev + (size_t)n * A.pitchV;
        const double *prow = A.plus + (off - z0), *mrow = A.minus + (off - z0);
#pragma unroll
        for (int v = 0; v <= VEC; ++v) {
            const int z = z0 + v;
            const double f = (z <= nz) ? vrow[z] : 0.0;
            fl[v] = (z < nz) ? b3v_point(f, z, prow, mrow) : f;   // the bottom flux stays
        }
        double t[VEC], l[VEC], hn[VEC], hw[VEC], ar[VEC], pn[VEC], mn[VEC];
        ldv<VEC>(A.del_v + off, dv);
        ldv<VEC>(A.del_h + off, dh);
        ldv_ro<VEC>(A.ttf + off, t);
        ldv_ro<VEC>(A.lo + off, l);
        ldv_ro<VEC>(A.hnode + offs, hn);
        ldv_ro<VEC>(A.hnode_new + offs, hw);
        ldv_ro<VEC>(A.area + (size_t)n * A.pitchV + z0, ar);
        ldv_ro<VEC>(A.plus + off, pn);
        ldv_ro<VEC>(A.minus + off, mn);
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            ar[v] = A.dt / ar[v];
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            dv[v] = dv[v] - t[v] * hn[v] + l[v] * hw[v] + (fl[v] - fl[v + 1]) * ar[v];

        const double *hb = A.adf_h_in + blockIdx.y * A.ts_edge + z0;
        double *ho = A.adf_h_out + blockIdx.y * A.ts_edge + z0;
        const size_t tb = blockIdx.y * A.ts_node + z0;
        const int b = __ldg(M.edg_off + n), e = __ldg(M.edg_off + n + 1);
        for (int k = b; k < e; ++k) {
            const int4 en = __ldg(M.edg + k);
            const int dg = FCT_META_DEPTH(en.z);
            if (z0 < dg) {
                double h[VEC], po[VEC], mo[VEC];
                const size_t ro = tb + (size_t)en.y * A.pitchL;
                ldv_ro<VEC>(hb + (size_t)en.x * A.pitchH, h);
                ldv_ro<VEC>(A.plus + ro, po);
                ldv_ro<VEC>(A.minus + ro, mo);
                const bool second = FCT_META_SECOND(en.z);
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    if (z0 + v < dg) {
                        h[v] = second ? b3h_point(h[v], po[v], mo[v], pn[v], mn[v])
                                      : b3h_point(h[v], pn[v], mn[v], po[v], mo[v]);
                        const double x = h[v] * ar[v];
                        dh[v] = second ? dh[v] - x : dh[v] + x;
                    }
                }
                if (FCT_META_WRITER(en.z)) stv<VEC>(ho + (size_t)en.x * A.pitchH, h, min(VEC, dg - z0));
            }
        }
    }
    __syncthreads();   // every raw fct_adf_v read of the block's columns precedes the writes
    if (!act) return;
    const int cnt = min(VEC, nz - z0);
    double fo[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) fo[v] = fl[v];
    stv<VEC>(A.adf_v_out + blockIdx.y * A.ts_nodev + (size_t)n * A.pitchV + z0, fo, cnt);
    stv<VEC>(A.del_v + off, dv, cnt);
    stv<VEC>(A.del_h + off, dh, cnt);
}

}   // namespace fct
