// Warp-item fused kernels with TMA-staged tiles: the fast path of the device-resident step.
//
// The tile-staged kernels of fct_tile_kernels.cuh turned out to be bound by instruction issue
// (profiles/r1_v1_*: 123 M + 99 M warp instructions per CORE2 step, a quarter of them index
// arithmetic, DRAM traffic only 1.2x algorithmic).  Here every irregular decision is taken by the
// inspector (fct_plan.cu, build_warptiles) and the kernels only stream:
//
//   * A tile is a run of consecutive owned nodes.  Its plan data is ONE contiguous blob: header,
//     table of node rows to stage, table of edge-flux rows to stage, one header per node, the
//     nodes' edge entries with precomputed shared-memory BYTE offsets, and a warp-item schedule.
//   * Staging is done by the TMA unit: thread 0 bulk-copies the blob (cp.async.bulk, SASS UBLKCP)
//     against mbarrier 0; the lanes of warp 0 then issue one bulk copy per row -- exactly the
//     active levels, in 16-byte granules -- for the two gathered node arrays (phase A: fct_LO and
//     ttf, converted in place to the a1 bounds of reference.cpp:315-316; phase B: fct_plus and
//     fct_minus) and for the edge-flux rows of the tile, against mbarrier 1.  Every edge whose two
//     end nodes lie in the tile is fetched once instead of once per end node.
//   * A warp item is 32*NCH virtual lanes, each a (node, pair of ACTIVE levels) slot; slots of a
//     node are consecutive virtual lanes and never straddle an item, so the vertical 3-point
//     stencil of a3 is done with warp shuffles and warps never meet at a CTA barrier after the
//     staging.  Lane l holds virtual lanes l, l+32, ... (NCH independent chains for ILP; a column
//     may have up to 64*NCH levels).
//   * In the item loop every shared-memory address is "region base + precomputed offset + 8*z0",
//     level masks ride on the DSETP...AND predicates, and the +/- split of b1 horizontal is two
//     predicated DADDs (adding +0 is exact, and the sums never are -0).
//
// Arithmetic and its order are those of fct_kernels.cuh (bit-identical results).
#pragma once
#include <cstdint>

#include "fct_kernels.cuh"

namespace fct {

struct WarpTilesDev {
    const uint4 *blob;          // concatenated per-tile blobs
    const unsigned *blob_off;   // [ntiles+1] in 16-byte units
    int ntiles;
    int smem_bytes;             // dynamic shared memory of one CTA (max over tiles)
};

constexpr int WT_THREADS = 256;
constexpr int WT_WARPS = WT_THREADS / 32;
// blob header: 16 ints
//  [0] node rows  [1] edge rows  [2] nodes  [3] warp items
//  [4] byte offset of the edge-row table  [5] of the node headers  [6] of the entries  [7] of the schedule
//  [8] blob bytes  [9] bytes of ONE staged node-row region  [10] bytes of the edge-row region  [11] bytes the row copies deliver
// node-row table at byte 64: int2 {global element offset of the row, (16-byte units) smem offset | size << 16}
// node header int4: {node*pitch, nz | fillmin << 8 | self depth << 16, smem byte offset of the own row, first entry | entries << 16}
// entry int4: {smem byte offset of the edge row, smem byte offset of the other node's row,
//              depth | writer << 30 | second << 31, edge*pitch}
constexpr int WT_HDR_BYTES = 64;
constexpr unsigned WT_IDLE = 0xffffu;

// ---- PTX: mbarrier + 1-D bulk copy (TMA) ---------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WT_DONE;\n"
        "bra WT_WAIT;\n"
        "WT_DONE:\n"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}

__device__ __forceinline__ double flip_sign(double v, unsigned sgn)
{
    return __hiloint2double(__double2hiint(v) ^ (int)sgn, __double2loint(v));
}

// a / b exactly as IEEE division; a zero numerator over a normal denominator (a local extremum:
// bound == low-order value, very common) skips the division's slow path: (+-0) * b has the sign
// and value of (+-0) / b.
__device__ __forceinline__ double div_exact(double a, double b)
{
    const unsigned eb = (unsigned)__double2hiint(b) & 0x7ff00000u;
    if (a == 0. && eb != 0u && eb != 0x7ff00000u) return a * b;
    return a / b;
}


// ---- hot inner bodies in PTX: ptxas keeps the predicated form (one DSETP...AND + two FSEL per
// bound update, one DSETP + one predicated DADD per sum) where the C++ form was if-converted into
// branches and select chains (2.5x the instructions) ---------------------------------------------

// One edge of phase A for two levels z0, z0+1.  meta = depth | writer << 30 | second << 31.
//   bounds: hi = pick_max(hi, x), lw = pick_min(lw, y) for levels above the edge depth
//   b1 horizontal (reference.cpp:417-423): q = +-h;  p += max(0, q);  m += min(0, q)
//   (adding +0 is exact and p, m never are -0, so the sums are predicated DADDs)
__device__ __forceinline__ void wt_edge_a(int z0, int meta, const double2 &x, const double2 &y, const double2 &h,
                                          double &hi0, double &hi1, double &lw0, double &lw1, double &p0,
                                          double &p1, double &m0, double &m1)
{
    asm("{\n"
        ".reg .pred P0, P1, q;\n"
        ".reg .b32 dg, sg, a, b, z1;\n"
        ".reg .f64 t;\n"
        "and.b32 dg, %9, 0xffff;\n"
        "and.b32 sg, %9, 0x80000000;\n"
        "add.s32 z1, %8, 1;\n"
        "setp.lt.s32 P0, %8, dg;\n"
        "setp.lt.s32 P1, z1, dg;\n"
        "setp.lt.and.f64 q, %0, %10, P0;\n"
        "selp.f64 %0, %10, %0, q;\n"
        "setp.lt.and.f64 q, %1, %11, P1;\n"
        "selp.f64 %1, %11, %1, q;\n"
        "setp.lt.and.f64 q, %12, %2, P0;\n"
        "selp.f64 %2, %12, %2, q;\n"
        "setp.lt.and.f64 q, %13, %3, P1;\n"
        "selp.f64 %3, %13, %3, q;\n"
        "mov.b64 {a, b}, %14;\n"
        "xor.b32 b, b, sg;\n"
        "mov.b64 t, {a, b};\n"
        "setp.gt.and.f64 q, t, 0d0000000000000000, P0;\n"
        "@q add.rn.f64 %4, %4, t;\n"
        "setp.lt.and.f64 q, t, 0d0000000000000000, P0;\n"
        "@q add.rn.f64 %6, %6, t;\n"
        "mov.b64 {a, b}, %15;\n"
        "xor.b32 b, b, sg;\n"
        "mov.b64 t, {a, b};\n"
        "setp.gt.and.f64 q, t, 0d0000000000000000, P1;\n"
        "@q add.rn.f64 %5, %5, t;\n"
        "setp.lt.and.f64 q, t, 0d0000000000000000, P1;\n"
        "@q add.rn.f64 %7, %7, t;\n"
        "}"
        : "+d"(hi0), "+d"(hi1), "+d"(lw0), "+d"(lw1), "+d"(p0), "+d"(p1), "+d"(m0), "+d"(m1)
        : "r"(z0), "r"(meta), "d"(x.x), "d"(x.y), "d"(y.x), "d"(y.y), "d"(h.x), "d"(h.y));
}

// b1 vertical of two levels (reference.cpp:397-398): p = max(0, f[z]) + max(0, -f[z+1]),
// m = min(0, f[z]) + min(0, -f[z+1]) with compare-select max / min
__device__ __forceinline__ void wt_b1v(double f0, double f1, double f2, double &p0, double &p1, double &m0, double &m1)
{
    asm("{\n"
        ".reg .pred g0, l0, g1, l1, g2, l2;\n"
        "setp.gt.f64 g0, %4, 0d0000000000000000;\n"
        "setp.lt.f64 l0, %4, 0d0000000000000000;\n"
        "setp.gt.f64 g1, %5, 0d0000000000000000;\n"
        "setp.lt.f64 l1, %5, 0d0000000000000000;\n"
        "setp.gt.f64 g2, %6, 0d0000000000000000;\n"
        "setp.lt.f64 l2, %6, 0d0000000000000000;\n"
        "selp.f64 %0, %4, 0d0000000000000000, g0;\n"
        "selp.f64 %2, %4, 0d0000000000000000, l0;\n"
        "selp.f64 %1, %5, 0d0000000000000000, g1;\n"
        "selp.f64 %3, %5, 0d0000000000000000, l1;\n"
        "@l1 sub.rn.f64 %0, %0, %5;\n"
        "@g1 sub.rn.f64 %2, %2, %5;\n"
        "@l2 sub.rn.f64 %1, %1, %6;\n"
        "@g2 sub.rn.f64 %3, %3, %6;\n"
        "}"
        : "=&d"(p0), "=&d"(p1), "=&d"(m0), "=&d"(m1)
        : "d"(f0), "d"(f1), "d"(f2));
}

// One edge of phase B for two levels (docs/refactoring.md:246-261 + :303-314).  n1 = edges[2g],
// n2 = edges[2g+1]; h >= 0: ae = min(1, plus[n1], minus[n2]) else min(1, minus[n1], plus[n2]) in
// that order; the own node is n2 when `second`.  hl = ae*h (identical on both end nodes);
// dh +-= hl * (dt/area).
__device__ __forceinline__ void wt_edge_b(int z0, int meta, const double2 &po, const double2 &mo, const double2 &h,
                                          double pn0, double pn1, double mn0, double mn1, double ar0, double ar1,
                                          double &dh0, double &dh1, double &hl0, double &hl1)
{
    asm("{\n"
        ".reg .pred P0, P1, S, q, r;\n"
        ".reg .b32 dg, sg, a, b, z1;\n"
        ".reg .f64 own, oth, x1, x2, ae, t;\n"
        "and.b32 dg, %5, 0xffff;\n"
        "and.b32 sg, %5, 0x80000000;\n"
        "add.s32 z1, %4, 1;\n"
        "setp.lt.s32 P0, %4, dg;\n"
        "setp.lt.s32 P1, z1, dg;\n"
        "setp.lt.s32 S, %5, 0;\n"
        // level z0
        "setp.ge.xor.f64 q, %10, 0d0000000000000000, S;\n"
        "selp.f64 own, %12, %14, q;\n"
        "selp.f64 oth, %8, %6, q;\n"
        "selp.f64 x1, oth, own, S;\n"
        "selp.f64 x2, own, oth, S;\n"
        "setp.lt.f64 r, x1, 0d3FF0000000000000;\n"
        "selp.f64 ae, x1, 0d3FF0000000000000, r;\n"
        "setp.lt.f64 r, x2, ae;\n"
        "selp.f64 ae, x2, ae, r;\n"
        "mul.rn.f64 %2, ae, %10;\n"
        "mov.b64 {a, b}, %2;\n"
        "xor.b32 b, b, sg;\n"
        "mov.b64 t, {a, b};\n"
        "mul.rn.f64 t, t, %16;\n"
        "@P0 add.rn.f64 %0, %0, t;\n"
        // level z0+1
        "setp.ge.xor.f64 q, %11, 0d0000000000000000, S;\n"
        "selp.f64 own, %13, %15, q;\n"
        "selp.f64 oth, %9, %7, q;\n"
        "selp.f64 x1, oth, own, S;\n"
        "selp.f64 x2, own, oth, S;\n"
        "setp.lt.f64 r, x1, 0d3FF0000000000000;\n"
        "selp.f64 ae, x1, 0d3FF0000000000000, r;\n"
        "setp.lt.f64 r, x2, ae;\n"
        "selp.f64 ae, x2, ae, r;\n"
        "mul.rn.f64 %3, ae, %11;\n"
        "mov.b64 {a, b}, %3;\n"
        "xor.b32 b, b, sg;\n"
        "mov.b64 t, {a, b};\n"
        "mul.rn.f64 t, t, %17;\n"
        "@P1 add.rn.f64 %1, %1, t;\n"
        "}"
        : "+d"(dh0), "+d"(dh1), "=&d"(hl0), "=&d"(hl1)
        : "r"(z0), "r"(meta), "d"(po.x), "d"(po.y), "d"(mo.x), "d"(mo.y), "d"(h.x), "d"(h.y), "d"(pn0), "d"(pn1),
          "d"(mn0), "d"(mn1), "d"(ar0), "d"(ar1));
}

struct WtView {
    int n_rows, n_erows, n_nodes, n_witems;
    unsigned char *blob, *rowsA, *rowsB, *erows;
    const int4 *hdr, *ent;
    const unsigned short *sched;
    int rows_bytes;
};

// Load the tile's blob and stage its rows.  src_a / src_b: the two gathered node arrays of this
// tracer, src_e: the edge fluxes.  Returns after the rows have landed (all threads).
__device__ __forceinline__ WtView wt_stage(unsigned char *sm, const WarpTilesDev &T, const double *src_a,
                                           const double *src_b, const double *src_e)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned b0 = __ldg(T.blob_off + blockIdx.x), b1 = __ldg(T.blob_off + blockIdx.x + 1);
    const uint32_t bar0 = smem_u32(sm), bar1 = bar0 + 8;
    WtView v;
    v.blob = sm + 16;
    if (tid == 0) {
        mbar_init(bar0, 1);
        mbar_init(bar1, 1);
        fence_mbar_init();
        mbar_expect_tx(bar0, (b1 - b0) * 16u);
        bulk_g2s(smem_u32(v.blob), T.blob + b0, (b1 - b0) * 16u, bar0);
    }
    __syncthreads();
    mbar_wait(bar0, 0);
    const int4 h0 = reinterpret_cast<const int4 *>(v.blob)[0];
    const int4 h1 = reinterpret_cast<const int4 *>(v.blob)[1];
    const int4 h2 = reinterpret_cast<const int4 *>(v.blob)[2];
    v.n_rows = h0.x;
    v.n_erows = h0.y;
    v.n_nodes = h0.z;
    v.n_witems = h0.w;
    v.rows_bytes = h2.y;
    v.rowsA = v.blob + h2.x;
    v.rowsB = v.rowsA + h2.y;
    v.erows = v.rowsB + h2.y;
    v.hdr = reinterpret_cast<const int4 *>(v.blob + h1.y);
    v.ent = reinterpret_cast<const int4 *>(v.blob + h1.z);
    v.sched = reinterpret_cast<const unsigned short *>(v.blob + h1.w);
    if (warp == 0) {
        if (lane == 0) mbar_expect_tx(bar1, (uint32_t)h2.w);
        __syncwarp();
        const uint32_t sa = smem_u32(v.rowsA), sb = smem_u32(v.rowsB), se = smem_u32(v.erows);
        const int2 *rt = reinterpret_cast<const int2 *>(v.blob + WT_HDR_BYTES);
        for (int u = lane; u < v.n_rows; u += 32) {
            const int2 r = rt[u];
            const uint32_t so = ((uint32_t)r.y & 0xffffu) << 4, sz = ((uint32_t)r.y >> 16) << 4;
            if (sz) {
                bulk_g2s(sa + so, src_a + (uint32_t)r.x, sz, bar1);
                bulk_g2s(sb + so, src_b + (uint32_t)r.x, sz, bar1);
            }
        }
        const int2 *et = reinterpret_cast<const int2 *>(v.blob + h1.x);
        for (int u = lane; u < v.n_erows; u += 32) {
            const int2 r = et[u];
            const uint32_t so = ((uint32_t)r.y & 0xffffu) << 4, sz = ((uint32_t)r.y >> 16) << 4;
            if (sz) bulk_g2s(se + so, src_e + (uint32_t)r.x, sz, bar1);
        }
    }
    mbar_wait(bar1, 0);
    return v;
}

// previous / next virtual lane's value (virtual lane = chunk*32 + lane)
template <int NCH>
__device__ __forceinline__ void vl_neighbours(const double (&lo_end)[NCH], const double (&hi_end)[NCH], int lane,
                                              double (&prev)[NCH], double (&next)[NCH])
{
    // lo_end[c]: this virtual lane's value at its FIRST level, hi_end[c]: at its SECOND level
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        prev[c] = __shfl_up_sync(0xffffffffu, hi_end[c], 1);
        next[c] = __shfl_down_sync(0xffffffffu, lo_end[c], 1);
    }
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        if (c > 0) {
            const double w = __shfl_sync(0xffffffffu, hi_end[c - 1], 31);
            if (lane == 0) prev[c] = w;
        }
        if (c + 1 < NCH) {
            const double w = __shfl_sync(0xffffffffu, lo_end[c + 1], 0);
            if (lane == 31) next[c] = w;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Phase A = a1 + a2 + a3 + b1 vertical + b1 horizontal + b2
// ------------------------------------------------------------------------------------------------
template <int NCH, int MINB>
__global__ void __launch_bounds__(WT_THREADS, MINB) k_phaseA_warp(Arrays A, WarpTilesDev T)
{
    extern __shared__ __align__(128) unsigned char wt_sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t tn = blockIdx.y * A.ts_node;
    const WtView V = wt_stage(wt_sm, T, A.lo + tn, A.ttf + tn, A.adf_h_in + blockIdx.y * A.ts_edge);

    // a1 in place on the staged rows: (fct_LO, ttf) -> (max, min), reference.cpp:315-316
    {
        double2 *pa = reinterpret_cast<double2 *>(V.rowsA), *pb = reinterpret_cast<double2 *>(V.rowsB);
        const int n16 = V.rows_bytes >> 4;
        for (int g = tid; g < n16; g += WT_THREADS) {
            const double2 l = pa[g], t = pb[g];
            pa[g] = make_double2(pick_max(l.x, t.x), pick_max(l.y, t.y));
            pb[g] = make_double2(pick_min(l.x, t.x), pick_min(l.y, t.y));
        }
    }
    __syncthreads();

    const double *g_lo = A.lo + tn;
    const double *g_ai = A.area_inv;
    const double *g_v = A.adf_v + blockIdx.y * A.ts_nodev;
    for (int wi = warp; wi < V.n_witems; wi += WT_WARPS) {
        bool act[NCH];
        int z0[NCH], nz[NCH], cnt[NCH];
        unsigned grow[NCH];
        const int4 *en[NCH];
        const unsigned char *ra[NCH], *rb[NCH], *re[NCH];
        double hi[NCH][2], lw[NCH][2], p[NCH][2], m[NCH][2], l[NCH][2], ai[NCH][2];
        int kmax = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const unsigned d = V.sched[(wi * NCH + c) * 32 + lane];
            act[c] = d != WT_IDLE;
            const int4 hd = V.hdr[act[c] ? (d & 0xffu) : 0u];
            z0[c] = act[c] ? (int)(d >> 8) * 2 : 0;
            nz[c] = act[c] ? (hd.y & 0xff) : 0;
            cnt[c] = act[c] ? (int)((unsigned)hd.w >> 16) : 0;
            grow[c] = (unsigned)hd.x + (unsigned)z0[c];
            en[c] = V.ent + (hd.w & 0xffff);
            ra[c] = V.rowsA + z0[c] * 8;
            rb[c] = V.rowsB + z0[c] * 8;
            re[c] = V.erows + z0[c] * 8;
            kmax = max(kmax, cnt[c]);
            // ---- own column: global loads first, used after the gather ----
            double f0 = 0., f1 = 0., f2 = 0.;
            l[c][0] = l[c][1] = ai[c][0] = ai[c][1] = 0.;
            if (act[c]) {
                const double2 ll = __ldg(reinterpret_cast<const double2 *>(g_lo + grow[c]));
                const double2 aa = __ldg(reinterpret_cast<const double2 *>(g_ai + grow[c]));
                const double2 ff = __ldg(reinterpret_cast<const double2 *>(g_v + grow[c]));
                if (z0[c] + 2 <= nz[c]) f2 = __ldg(g_v + grow[c] + 2);
                l[c][0] = ll.x; l[c][1] = ll.y;
                ai[c][0] = aa.x; ai[c][1] = aa.y;
                f0 = ff.x; f1 = ff.y;
            }
            // cluster bounds start from the (-big, +big) fill of ring elements that already ended
            // (reference.cpp:341-349) and the node's own a1 bounds
            const int fm = (hd.y >> 8) & 0xff, sd = (hd.y >> 16) & 0xff;
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const bool fl = z0[c] + v >= fm;
                hi[c][v] = fl ? -A.big : -CUDART_INF;
                lw[c][v] = fl ? A.big : CUDART_INF;
            }
            {
                const double2 x = *reinterpret_cast<const double2 *>(ra[c] + hd.z);
                const double2 y = *reinterpret_cast<const double2 *>(rb[c] + hd.z);
                if (act[c] && z0[c] < sd) {
                    hi[c][0] = pick_max(hi[c][0], x.x);
                    lw[c][0] = pick_min(lw[c][0], y.x);
                }
                if (act[c] && z0[c] + 1 < sd) {
                    hi[c][1] = pick_max(hi[c][1], x.y);
                    lw[c][1] = pick_min(lw[c][1], y.y);
                }
            }
            wt_b1v(f0, f1, f2, p[c][0], p[c][1], m[c][0], m[c][1]);
        }
        // ---- the node's edges in ascending edge id: a2/a3 bounds + b1 horizontal ----
        for (int k = 0; k < kmax; ++k) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (NCH == 1 || k < cnt[c]) {
                    const int4 e = en[c][k];
                    const double2 x = *reinterpret_cast<const double2 *>(ra[c] + e.y);
                    const double2 y = *reinterpret_cast<const double2 *>(rb[c] + e.y);
                    const double2 h = *reinterpret_cast<const double2 *>(re[c] + e.x);
                    wt_edge_a(z0[c], e.z, x, y, h, hi[c][0], hi[c][1], lw[c][0], lw[c][1], p[c][0], p[c][1], m[c][0],
                              m[c][1]);
                }
            }
        }
        // ---- vertical 3-point stencil of a3 (reference.cpp:380-392) through warp shuffles ----
        double a0[NCH], a1[NCH], b0[NCH], b1[NCH], pmax[NCH], nmax[NCH], pmin[NCH], nmin[NCH];
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            a0[c] = hi[c][0]; a1[c] = hi[c][1];
            b0[c] = lw[c][0]; b1[c] = lw[c][1];
        }
        vl_neighbours<NCH>(a0, a1, lane, pmax, nmax);
        vl_neighbours<NCH>(b0, b1, lane, pmin, nmin);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            if (!act[c]) continue;
            double bm[2], bn[2];
            {
                double x = a0[c], y = b0[c];
                if (z0[c] > 0 && z0[c] < nz[c] - 1) {
                    x = pick_max(pick_max(pmax[c], x), a1[c]);
                    y = pick_min(pick_min(pmin[c], y), b1[c]);
                }
                bm[0] = x - l[c][0];
                bn[0] = y - l[c][0];
                x = a1[c];
                y = b1[c];
                if (z0[c] + 1 < nz[c] - 1) {
                    x = pick_max(pick_max(a0[c], x), nmax[c]);
                    y = pick_min(pick_min(b0[c], y), nmin[c]);
                }
                bm[1] = x - l[c][1];
                bn[1] = y - l[c][1];
            }
            // ---- b2, reference.cpp:432-435 ----
            double pf[2], mf[2];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                double flux = p[c][v] * A.dt * ai[c][v] + A.eps;
                pf[v] = pick_min(1., div_exact(bm[v], flux));
                flux = m[c][v] * A.dt * ai[c][v] - A.eps;
                mf[v] = pick_min(1., div_exact(bn[v], flux));
            }
            const size_t off = tn + grow[c];
            if (z0[c] + 1 < nz[c]) {
                *reinterpret_cast<double2 *>(A.ttf_max + off) = make_double2(bm[0], bm[1]);
                *reinterpret_cast<double2 *>(A.ttf_min + off) = make_double2(bn[0], bn[1]);
                *reinterpret_cast<double2 *>(A.plus + off) = make_double2(pf[0], pf[1]);
                *reinterpret_cast<double2 *>(A.minus + off) = make_double2(mf[0], mf[1]);
            } else {
                A.ttf_max[off] = bm[0];
                A.ttf_min[off] = bn[0];
                A.plus[off] = pf[0];
                A.minus[off] = mf[0];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Phase B = b3 vertical + b3 horizontal + c vertical + c horizontal
// ------------------------------------------------------------------------------------------------
template <int NCH, int MINB>
__global__ void __launch_bounds__(WT_THREADS, MINB) k_phaseB_warp(Arrays A, WarpTilesDev T)
{
    extern __shared__ __align__(128) unsigned char wt_sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t tn = blockIdx.y * A.ts_node;
    const WtView V = wt_stage(wt_sm, T, A.plus + tn, A.minus + tn, A.adf_h_in + blockIdx.y * A.ts_edge);
    (void)lane;

    const double *g_v = A.adf_v + blockIdx.y * A.ts_nodev;
    double *g_vout = A.adf_v_out + blockIdx.y * A.ts_nodev;
    double *g_ho = A.adf_h_out + blockIdx.y * A.ts_edge;
    for (int wi = warp; wi < V.n_witems; wi += WT_WARPS) {
        bool act[NCH];
        int z0[NCH], nz[NCH], cnt[NCH];
        unsigned grow[NCH];
        const int4 *en[NCH];
        const unsigned char *ra[NCH], *rb[NCH], *re[NCH];
        double dh[NCH][2], dv[NCH][2], fl[NCH][2], ar[NCH][2], pn[NCH][2], mn[NCH][2];
        int kmax = 0;
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            const unsigned d = V.sched[(wi * NCH + c) * 32 + lane];
            act[c] = d != WT_IDLE;
            const int4 hd = V.hdr[act[c] ? (d & 0xffu) : 0u];
            z0[c] = act[c] ? (int)(d >> 8) * 2 : 0;
            nz[c] = act[c] ? (hd.y & 0xff) : 0;
            cnt[c] = act[c] ? (int)((unsigned)hd.w >> 16) : 0;
            grow[c] = (unsigned)hd.x + (unsigned)z0[c];
            en[c] = V.ent + (hd.w & 0xffff);
            ra[c] = V.rowsA + z0[c] * 8;
            rb[c] = V.rowsB + z0[c] * 8;
            re[c] = V.erows + z0[c] * 8;
            kmax = max(kmax, cnt[c]);
            dh[c][0] = dh[c][1] = dv[c][0] = dv[c][1] = 0.;
            fl[c][0] = fl[c][1] = ar[c][0] = ar[c][1] = 0.;
            pn[c][0] = pn[c][1] = mn[c][0] = mn[c][1] = 0.;
            if (act[c]) {
                // ---- every global load of the own column ----
                const size_t off = tn + grow[c];
                const double2 q_dv = *reinterpret_cast<const double2 *>(A.del_v + off);
                const double2 q_dh = *reinterpret_cast<const double2 *>(A.del_h + off);
                const double2 q_t = __ldg(reinterpret_cast<const double2 *>(A.ttf + off));
                const double2 q_l = __ldg(reinterpret_cast<const double2 *>(A.lo + off));
                const double2 q_hn = __ldg(reinterpret_cast<const double2 *>(A.hnode + grow[c]));
                const double2 q_hw = __ldg(reinterpret_cast<const double2 *>(A.hnode_new + grow[c]));
                const double2 q_ar = __ldg(reinterpret_cast<const double2 *>(A.area + grow[c]));
                const double2 q_f = __ldg(reinterpret_cast<const double2 *>(g_v + grow[c]));
                const double f2 = (z0[c] + 2 <= nz[c]) ? __ldg(g_v + grow[c] + 2) : 0.;
                // own factors from the staged rows: levels z0-1 .. z0+2
                const unsigned char *pr = ra[c] + hd.z, *mr = rb[c] + hd.z;
                const double2 pp = *reinterpret_cast<const double2 *>(pr);
                const double2 mm = *reinterpret_cast<const double2 *>(mr);
                const double p_m1 = *reinterpret_cast<const double *>(pr - 8);
                const double m_m1 = *reinterpret_cast<const double *>(mr - 8);
                const double p_p2 = *reinterpret_cast<const double *>(pr + 16);
                const double m_p2 = *reinterpret_cast<const double *>(mr + 16);
                pn[c][0] = pp.x; pn[c][1] = pp.y;
                mn[c][0] = mm.x; mn[c][1] = mm.y;
                // ---- b3 vertical, docs/refactoring.md:205-231 (the bottom flux stays) ----
                const int z = z0[c];
                double l0, l1, l2;
                {
                    double ae = 1.;
                    if (z == 0) {
                        ae = pick_min(ae, (q_f.x >= 0.) ? pp.x : mm.x);
                    } else if (q_f.x >= 0.) {
                        ae = pick_min(ae, m_m1);
                        ae = pick_min(ae, pp.x);
                    } else {
                        ae = pick_min(ae, p_m1);
                        ae = pick_min(ae, mm.x);
                    }
                    l0 = ae * q_f.x;
                }
                if (z + 1 < nz[c]) {
                    double ae = 1.;
                    if (q_f.y >= 0.) {
                        ae = pick_min(ae, mm.x);
                        ae = pick_min(ae, pp.y);
                    } else {
                        ae = pick_min(ae, pp.x);
                        ae = pick_min(ae, mm.y);
                    }
                    l1 = ae * q_f.y;
                } else {
                    l1 = q_f.y;
                }
                if (z + 2 < nz[c]) {
                    double ae = 1.;
                    if (f2 >= 0.) {
                        ae = pick_min(ae, mm.y);
                        ae = pick_min(ae, p_p2);
                    } else {
                        ae = pick_min(ae, pp.y);
                        ae = pick_min(ae, m_p2);
                    }
                    l2 = ae * f2;
                } else {
                    l2 = f2;
                }
                fl[c][0] = l0;
                fl[c][1] = l1;
                // ---- c vertical, docs/refactoring.md:295-300 ----
                ar[c][0] = A.dt / q_ar.x;
                ar[c][1] = A.dt / q_ar.y;
                dv[c][0] = q_dv.x - q_t.x * q_hn.x + q_l.x * q_hw.x + (l0 - l1) * ar[c][0];
                dv[c][1] = q_dv.y - q_t.y * q_hn.y + q_l.y * q_hw.y + (l1 - l2) * ar[c][1];
                dh[c][0] = q_dh.x;
                dh[c][1] = q_dh.y;
            }
        }
        // ---- b3 horizontal + c horizontal over the node's edges, ascending edge id ----
        for (int k = 0; k < kmax; ++k) {
#pragma unroll
            for (int c = 0; c < NCH; ++c) {
                if (NCH == 1 || k < cnt[c]) {
                    const int4 e = en[c][k];
                    const double2 po = *reinterpret_cast<const double2 *>(ra[c] + e.y);
                    const double2 mo = *reinterpret_cast<const double2 *>(rb[c] + e.y);
                    const double2 h = *reinterpret_cast<const double2 *>(re[c] + e.x);
                    double hl0, hl1;
                    wt_edge_b(z0[c], e.z, po, mo, h, pn[c][0], pn[c][1], mn[c][0], mn[c][1], ar[c][0], ar[c][1],
                              dh[c][0], dh[c][1], hl0, hl1);
                    const int dg = e.z & 0xffff;
                    if ((e.z & 0x40000000) && z0[c] < dg) {
                        double *o = g_ho + (unsigned)e.w + z0[c];
                        if (z0[c] + 1 < dg) *reinterpret_cast<double2 *>(o) = make_double2(hl0, hl1);
                        else o[0] = hl0;
                    }
                }
            }
        }
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
            if (!act[c]) continue;
            const size_t off = tn + grow[c];
            if (z0[c] + 1 < nz[c]) {
                *reinterpret_cast<double2 *>(g_vout + grow[c]) = make_double2(fl[c][0], fl[c][1]);
                *reinterpret_cast<double2 *>(A.del_v + off) = make_double2(dv[c][0], dv[c][1]);
                *reinterpret_cast<double2 *>(A.del_h + off) = make_double2(dh[c][0], dh[c][1]);
            } else {
                g_vout[grow[c]] = fl[c][0];
                A.del_v[off] = dv[c][0];
                A.del_h[off] = dh[c][0];
            }
        }
    }
}

}   // namespace fct
