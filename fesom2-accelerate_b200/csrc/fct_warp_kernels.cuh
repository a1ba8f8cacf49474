// Persistent, warp-specialised fused kernels with TMA-staged tiles: the fast path of the
// device-resident step.
//
// The tile-staged kernels of fct_tile_kernels.cuh turned out to be bound by instruction issue
// (profiles/r1_v1_*: 123 M + 99 M warp instructions per CORE2 step, a quarter of them index
// arithmetic, DRAM traffic only 1.2x algorithmic), and a first CTA-per-tile TMA version spent 43 %
// of its warp time waiting for its own staging (profiles/r1_v2_*).  Here every irregular decision
// is taken by the inspector (fct_plan.cu, build_warptiles), the kernels only stream, and staging
// runs ahead of the arithmetic:
//
//   * A tile is a run of consecutive owned nodes.  Its plan data is ONE contiguous blob: header,
//     list of bulk copies, one header per node, the nodes' edge entries with precomputed
//     shared-memory BYTE offsets, and a warp-item schedule.  The same kernels serve the padded
//     layout (row = index * pitch) and the packed level storage (columns back to back, active
//     levels only): every global offset they use comes out of the blob.
//   * One persistent CTA per SM draws tiles from a device-wide counter (in curve order: the tiles
//     in flight are a compact patch, the tail is balanced) and runs them through an NSTAGE-deep
//     ring of shared-memory stages guarded by mbarriers.  Producer warps feed the ring with the TMA
//     unit (cp.async.bulk, SASS UBLKCP): warp 0 fetches the blob (prefetched into L2 while the
//     stage is still busy); NPW issuer warps share the tile's copy list: the columns of the two
//     gathered node arrays (phase A: fct_LO and ttf; phase B: fct_plus and fct_minus; own + halo
//     nodes) and the tile's edge-flux rows (an edge with both ends in the tile is fetched once, not
//     once per end), exactly the active levels in 16-byte granules.  The stage is laid out own
//     columns first, halo columns, edge rows by ascending id, so rows that are adjacent in global
//     memory are adjacent in shared memory and travel as ONE copy: about 40 copies per tile in the
//     packed storage, 240 in the padded layout, where a single issuer warp was the bottleneck
//     (profiles/r1_v3_*).  The a1 bounds of reference.cpp:315-316 are taken on the fly in phase A's
//     edge loop from the raw (fct_LO, ttf) rows (round 2; two converter warps that rewrite the landed
//     rows in place remain for the vlimit 2 / 3 variants).  The issuer warps also pull the own columns
//     of every array the consumers load straight from global memory into L2, tile by tile.  NWC
//     consumer warps draw warp items of the ready stage from a shared counter, so tiles k+1 and k+2
//     are in flight while tile k is computed, and every item loads the first-needed own-column
//     values of the warp's NEXT item before its own edge loop.
//   * A warp item is 32 lanes, each a (node, pair of ACTIVE levels) slot; consecutive lanes hold
//     consecutive slots of a node, so the vertical 3-point stencil of a3 is two warp shuffles.  A
//     column that does not fit the rest of an item is split, with one GHOST slot on either side of
//     the cut (it recomputes the neighbouring cluster bound and stores nothing), which keeps the
//     lanes > 90 % full at any depth.  Nodes of equal degree are scheduled next to each other, so
//     the lanes of an item walk edge lists of equal length.
//   * In the item loop every shared-memory address is "region base + precomputed offset + 8*z0",
//     level masks ride on the DSETP...AND predicates, the +/- split of b1 horizontal runs on the
//     FP64 pipe (adding +0 is exact, and the sums never are -0).
//
// Arithmetic and its order are those of fct_kernels.cuh (bit-identical results).
#pragma once
#include <cstdint>

#include "fct_kernels.cuh"
#include "fct_tile_kernels.cuh"   // prefetch_l2

namespace fct {

struct WarpTilesDev {
    const uint4 *blob;          // concatenated per-tile blobs
    const unsigned *blob_off;   // [ntiles+1] in 16-byte units
    int ntiles;
    int smem_bytes;             // shared memory of one stage (max over tiles)
    int diag;                   // timing experiments only (wrong results): 1 skip the edge-row copies, 2 skip all row copies
    int opt;                    // scheduling options (results unaffected): 1 producers wait suspended in hardware
                                // instead of polling, 2 consumers issue a tile's first loads before waiting for its rows,
                                // 4 copy lists travel ahead of their blobs (needs max_copies <= WT_PRE_MAX_COPIES),
                                // 8 fetcher and issuers probe a released stage every 40 ns (hand-over on the critical path),
                                // 16 / 32 / 64 other producer wait modes, 128 the issuers pull the own columns of every array
                                // the consumers load from global memory into L2, tile by tile (with lists ahead: two tile
                                // periods ahead, or one with 256)
    int max_copies;             // longest copy list of a tile
    long long *trace;           // profiling aid (knob WT_TRACE): SM-clock stamps of the pipeline events of CTA 0,
                                // WT_TRACE_SLOTS per tile iteration, at most WT_TRACE_ITERS iterations; null: off
};
constexpr int WT_TRACE_SLOTS = 10;
constexpr int WT_TRACE_ITERS = 4096;
// slots: 0 fetcher: stage seen empty, 1 fetcher: blob copy issued, 2 issuer: starts the row copies, 3 issuer: copies issued,
//        4 converter: rows landed, 5 converter: a1 done, 6 first consumer warp: tile ready, 7 first consumer warp: leaves
//        the tile, 8 last consumer warp: tile ready, 9 last consumer warp: leaves the tile
__device__ __forceinline__ void wt_trace(const WarpTilesDev &T, int it, int slot)
{
    if (T.trace != nullptr && blockIdx.x == 0 && it < WT_TRACE_ITERS && (threadIdx.x & 31) == 0)
        T.trace[it * WT_TRACE_SLOTS + slot] = clock64();
}

// warp roles: 0 blob fetcher, 1..NPW copy issuers, in phase A NCV a1 converters (0 in the default shape:
// a1 is folded into the consumers), then NWC consumers (registers are granted as if the CTA had a
// multiple of 4 warps: 16 warps -> 128 per thread, 20 -> 96, 24 -> 80)
// Head of a blob (header + copy list) copied ahead of the blob itself into a slot of its own, one
// per stage: the issuer warps pull the tile's rows into L2 while the stage is still being computed
// and start the bulk copies the moment it is released, next to the blob's copy instead of after it.
constexpr int WT_PRE_BYTES = 1024;
constexpr int WT_SMEM_CTRL = 256;     // mbarriers + per-stage counters
constexpr int WT_SMEM_HEAD = WT_SMEM_CTRL + 4 * WT_PRE_BYTES;   // in front of the stages
#ifndef WT_IDLE_NS
#define WT_IDLE_NS 400                // sleep of an idle producer between two probes of its barrier
#endif
constexpr int WT_CONVERTERS = 2;      // phase A: warps that turn the landed rows into the a1 bounds
constexpr int WT_SMEM_MAX = 227 * 1024;
// blob header: 16 ints
//  [0] bulk copies  [1] edge rows  [2] nodes  [3] warp items
//  [4] node rows  [5] byte offset of the node headers  [6] of the entries  [7] of the schedule
//  [8] blob bytes  [9] bytes of ONE staged node-row region  [10] bytes of the edge-row region  [11] bytes the copies deliver
// copy list at byte 64: int2 {global element offset of the first row,
//     (16-byte units) smem offset from the first row region | size << 14 | array << 28} with array 0 / 1
//     the two gathered node arrays (regions A, B) and 2 the edge fluxes (region E); regions are consecutive
// node header int4: {global element offset of the column, nz | fillmin << 8 | self depth << 16,
//                    smem byte offset of the own row, first entry | entries << 16}
// entry int4: {smem byte offset of the edge row, smem byte offset of the other node's row,
//              depth | writer << 30 | second << 31, global element offset of the edge row}
// schedule: one unsigned short per lane: node (8 bits) | level pair (7 bits) << 8 | ghost << 15; 0xffff idle
//  [12] L2 prefetch entries behind the copy list (array code 3, no shared-memory offset): runs of own columns
constexpr int WT_MAX_PREFETCH = 8;
constexpr int WT_HDR_BYTES = 64;
constexpr int WT_PRE_MAX_COPIES = (WT_PRE_BYTES - WT_HDR_BYTES) / 8;
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_proxy_async()
{
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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
// try_wait suspends the thread in hardware until the phase completes or the hint (ns) expires: a
// waiting warp issues an instruction every ~20 us instead of spinning, and wakes as soon as the
// barrier flips
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 20000;\n"
        "@p bra WT_DONE;\n"
        "bra WT_WAIT;\n"
        "WT_DONE:\n"
        "}" ::"r"(bar), "r"(parity)
        : "memory");
}
// producer-side wait: producers are idle most of the time; probe, then sleep between probes so that
// they leave the issue slots to the consumers (costs a fraction of a microsecond per hand-over)
// the same wait with a short suspend-time hint: one probe per HINT ns instead of one per ~26 ns (the opcode mix of
// round 1 shows the nanosleep loop below turning over every ~26 ns per producer warp: 15 % of all issued instructions)
template <int HINT>
__device__ __forceinline__ void mbar_wait_hint(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity), "n"(HINT)
            : "memory");
    } while (!done);
}
// mode (WarpTilesDev::opt): bit 0 suspended with a 20 us hint, bit 4 (16) a 1 us hint, bit 5 (32) a 250 ns hint,
// bit 6 (64) the polling loop with a 2 us sleep instead of 400 ns
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity, int mode = 0)
{
    if (mode & 1) {
        mbar_wait(bar, parity);
        return;
    }
    if (mode & 16) {
        mbar_wait_hint<1000>(bar, parity);
        return;
    }
    if (mode & 32) {
        mbar_wait_hint<250>(bar, parity);
        return;
    }
    const unsigned ns = (mode & 64) ? 2000u : (unsigned)WT_IDLE_NS;
    uint32_t done = 0;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(ns);
    }
}

// wait on the pipeline's critical hand-over (a stage released by its last consumer warp): the refill
// starts the moment this returns, so the probe interval is short (measured with knob WT_TRACE: the 400 ns
// sleep left a released stage idle for 0.6 us, the suspended try_wait for 1.6 us, of a 5 us tile period)
__device__ __forceinline__ void mbar_wait_handover(uint32_t bar, uint32_t parity)
{
    uint32_t done = 0;
    for (;;) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        __nanosleep(40);
    }
}

// lane 0's draw from a shared-memory counter (the C++ atomicAdd carries warp-aggregation code for
// a uniform address that a single active lane does not need)
__device__ __forceinline__ int smem_fetch_add(int *ctr)
{
    int old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(ctr)) : "memory");
    return old;
}

__device__ __forceinline__ double flip_sign(double v, unsigned sgn)
{
    return __hiloint2double(__double2hiint(v) ^ (int)sgn, __double2loint(v));
}

// min(1, a / b) with a / b exactly the IEEE quotient.  A zero numerator (a local extremum: bound ==
// low-order value, a quarter of all cells) would send the compiler's division into its slow path
// on almost every warp; branch-free detour: q = (a == 0 ? 1 : a) / b, and for a == 0 the result
// (+-0) * (1 / b) has the value and sign of (+-0) / b for every b (0 * inf = NaN = 0 / 0 included).
__device__ __forceinline__ double limit_quotient(double a, double b)
{
    const bool z = a == 0.;
    const double q = (z ? 1. : a) / b;
    const double r = z ? a * q : q;
    double o;
    asm("{\n"
        ".reg .pred p;\n"
        "setp.lt.f64 p, %1, 0d3FF0000000000000;\n"
        "selp.f64 %0, %1, 0d3FF0000000000000, p;\n"
        "}"
        : "=d"(o)
        : "d"(r));
    return o;
}


// min(1, x) as one compare + one select (the C++ form compiles to DSETP.MIN + FSEL + SEL + LOP3)
__device__ __forceinline__ double min_one(double x)
{
    double o;
    asm("{\n"
        ".reg .pred p;\n"
        "setp.lt.f64 p, %1, 0d3FF0000000000000;\n"
        "selp.f64 %0, %1, 0d3FF0000000000000, p;\n"
        "}"
        : "=d"(o)
        : "d"(x));
    return o;
}

// ---- hot inner bodies in PTX: ptxas keeps the predicated form (one DSETP...AND + two FSEL per
// bound update, one DSETP + one predicated DADD per sum) where the C++ form was if-converted into
// branches and select chains (2.5x the instructions) ---------------------------------------------

// One edge of phase A for two levels z0, z0+1.  meta = depth | writer << 30 | second << 31.
//   bounds: hi = pick_max(hi, x), lw = pick_min(lw, y) for levels above the edge depth
//   b1 horizontal (reference.cpp:417-423): q = +-h;  p += max(0, q);  m += min(0, q), written for
//   the (idle) FP64 pipe instead of compare + select on the (busy) ALU pipe:
//     max(0, q) = (q + |q|) / 2 and min(0, q) = (q - |q|) / 2 exactly, q = h * s with s = +-1, so
//     p = fma(fma(h, s, |h|), 0.5, p) rounds once, to the same value as p + max(0, q); a
//     non-positive q adds +0, which changes neither p >= +0 nor m <= +0.
__device__ __forceinline__ void wt_edge_a(int z0, int meta, const double2 &x, const double2 &y, const double2 &h,
                                          double &hi0, double &hi1, double &lw0, double &lw1, double &p0,
                                          double &p1, double &m0, double &m1)
{
    asm("{\n"
        ".reg .pred P0, P1, q;\n"
        ".reg .b32 dg, sh, z1;\n"
        ".reg .f64 s, a, t, h0, h1;\n"
        "and.b32 dg, %9, 0xffff;\n"
        "and.b32 sh, %9, 0x80000000;\n"
        "or.b32 sh, sh, 0x3FF00000;\n"
        "mov.b64 s, {0, sh};\n"
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
        // levels below the edge contribute a flux of +0 (ptxas turns a predicated fp64 accumulate into
        // DFMA + two FSEL; masking the flux costs one select pair per level instead of two): h = +0
        // gives t = +0 for the plus sum and -0 or +0 for the minus sum, which change neither
        "selp.f64 h0, %14, 0d0000000000000000, P0;\n"
        "selp.f64 h1, %15, 0d0000000000000000, P1;\n"
        "abs.f64 a, h0;\n"
        "fma.rn.f64 t, h0, s, a;\n"
        "fma.rn.f64 %4, t, 0d3FE0000000000000, %4;\n"
        "neg.f64 a, a;\n"
        "fma.rn.f64 t, h0, s, a;\n"
        "fma.rn.f64 %6, t, 0d3FE0000000000000, %6;\n"
        "abs.f64 a, h1;\n"
        "fma.rn.f64 t, h1, s, a;\n"
        "fma.rn.f64 %5, t, 0d3FE0000000000000, %5;\n"
        "neg.f64 a, a;\n"
        "fma.rn.f64 t, h1, s, a;\n"
        "fma.rn.f64 %7, t, 0d3FE0000000000000, %7;\n"
        "}"
        : "+d"(hi0), "+d"(hi1), "+d"(lw0), "+d"(lw1), "+d"(p0), "+d"(p1), "+d"(m0), "+d"(m1)
        : "r"(z0), "r"(meta), "d"(x.x), "d"(x.y), "d"(y.x), "d"(y.y), "d"(h.x), "d"(h.y));
}

// The same edge with the a1 pass folded in (NCV < 0: no converter warps, no pass over the staged rows): x / y are
// the neighbour's RAW fct_LO / ttf rows and its a1 bounds max(LO, ttf) / min(LO, ttf) (reference.cpp:315-316) are
// taken on the fly: hi = max(hi, LO, ttf), lw = min(lw, LO, ttf) -- exact, max / min are associative.
__device__ __forceinline__ void wt_edge_a_fold(int z0, int meta, const double2 &x, const double2 &y, const double2 &h,
                                               double &hi0, double &hi1, double &lw0, double &lw1, double &p0,
                                               double &p1, double &m0, double &m1)
{
    asm("{\n"
        ".reg .pred P0, P1, q;\n"
        ".reg .b32 dg, sh, z1;\n"
        ".reg .f64 s, a, t, h0, h1;\n"
        "and.b32 dg, %9, 0xffff;\n"
        "and.b32 sh, %9, 0x80000000;\n"
        "or.b32 sh, sh, 0x3FF00000;\n"
        "mov.b64 s, {0, sh};\n"
        "add.s32 z1, %8, 1;\n"
        "setp.lt.s32 P0, %8, dg;\n"
        "setp.lt.s32 P1, z1, dg;\n"
        "setp.lt.and.f64 q, %0, %10, P0;\n"
        "selp.f64 %0, %10, %0, q;\n"
        "setp.lt.and.f64 q, %0, %12, P0;\n"
        "selp.f64 %0, %12, %0, q;\n"
        "setp.lt.and.f64 q, %1, %11, P1;\n"
        "selp.f64 %1, %11, %1, q;\n"
        "setp.lt.and.f64 q, %1, %13, P1;\n"
        "selp.f64 %1, %13, %1, q;\n"
        "setp.lt.and.f64 q, %10, %2, P0;\n"
        "selp.f64 %2, %10, %2, q;\n"
        "setp.lt.and.f64 q, %12, %2, P0;\n"
        "selp.f64 %2, %12, %2, q;\n"
        "setp.lt.and.f64 q, %11, %3, P1;\n"
        "selp.f64 %3, %11, %3, q;\n"
        "setp.lt.and.f64 q, %13, %3, P1;\n"
        "selp.f64 %3, %13, %3, q;\n"
        "selp.f64 h0, %14, 0d0000000000000000, P0;\n"
        "selp.f64 h1, %15, 0d0000000000000000, P1;\n"
        "abs.f64 a, h0;\n"
        "fma.rn.f64 t, h0, s, a;\n"
        "fma.rn.f64 %4, t, 0d3FE0000000000000, %4;\n"
        "neg.f64 a, a;\n"
        "fma.rn.f64 t, h0, s, a;\n"
        "fma.rn.f64 %6, t, 0d3FE0000000000000, %6;\n"
        "abs.f64 a, h1;\n"
        "fma.rn.f64 t, h1, s, a;\n"
        "fma.rn.f64 %5, t, 0d3FE0000000000000, %5;\n"
        "neg.f64 a, a;\n"
        "fma.rn.f64 t, h1, s, a;\n"
        "fma.rn.f64 %7, t, 0d3FE0000000000000, %7;\n"
        "}"
        : "+d"(hi0), "+d"(hi1), "+d"(lw0), "+d"(lw1), "+d"(p0), "+d"(p1), "+d"(m0), "+d"(m1)
        : "r"(z0), "r"(meta), "d"(x.x), "d"(x.y), "d"(y.x), "d"(y.y), "d"(h.x), "d"(h.y));
}

// store two consecutive levels, or only the first when the column ends between them (one
// predicated pair instead of two divergent code paths)
__device__ __forceinline__ void wt_store2(double *p, double a, double b, bool both)
{
    asm volatile(
        "{\n"
        ".reg .pred f;\n"
        "setp.ne.s32 f, %3, 0;\n"
        "@f st.global.v2.f64 [%0], {%1, %2};\n"
        "@!f st.global.f64 [%0], %1;\n"
        "}" ::"l"(p), "d"(a), "d"(b), "r"((int)both)
        : "memory");
}

// b1 vertical of two levels (reference.cpp:397-398): p = max(0, f[z]) + max(0, -f[z+1]),
// m = min(0, f[z]) + min(0, -f[z+1]).  On the FP64 pipe, without compares and selects:
// a = f + |f| = 2 max(0, f) and c = f - |f| = 2 min(0, f) are exact, max(0, -f) = -c / 2 and
// min(0, -f) = -a / 2, so p = (a_z - c_z+1) / 2 and m = (c_z - a_z+1) / 2 round once, at the same
// place and to the same value as the sums of the reference (scaling by 2 is exact).  A flux of -0
// yields m = -0 where the compare-select form yields +0; the sums only feed m*dt*area_inv - eps
// and further additions, where the two zeros are indistinguishable.
__device__ __forceinline__ void wt_b1v(double f0, double f1, double f2, double &p0, double &p1, double &m0, double &m1)
{
    const double a0 = f0 + fabs(f0), c0 = f0 - fabs(f0);
    const double a1 = f1 + fabs(f1), c1 = f1 - fabs(f1);
    const double a2 = f2 + fabs(f2), c2 = f2 - fabs(f2);
    p0 = (a0 - c1) * 0.5;
    m0 = (c0 - a1) * 0.5;
    p1 = (a1 - c2) * 0.5;
    m1 = (c1 - a2) * 0.5;
}

// One edge of phase B for two levels (docs/refactoring.md:246-261 + :303-314).  n1 = edges[2g],
// n2 = edges[2g+1]; h >= 0: ae = min(1, plus[n1], minus[n2]) else min(1, minus[n1], plus[n2]); the own node is
// n2 when `second`.  hl = ae*h (identical on both end nodes); dh +-= hl * (dt/area).
// Written so that the own node's share is hoisted out of the edge loop: with p1 = min(1, plus[own]) and
// m1 = min(1, minus[own]) (per level, computed once per item), u = min(p1, minus[other]) and
// v = min(m1, plus[other]) are the two candidates whatever the roles -- first end: h >= 0 ? u : v; second end:
// h >= 0 ? min(1, plus[other], minus[own]) = v : u -- so ae = ((h >= 0) xor second) ? u : v: three compares and
// three selects per level instead of three and six (min is exact and associative: the same value as the
// listing's min(1, a, b) in any order).
__device__ __forceinline__ void wt_edge_b(int z0, int meta, const double2 &po, const double2 &mo, const double2 &h,
                                          double p10, double p11, double m10, double m11, double ar0, double ar1,
                                          double &dh0, double &dh1, double &hl0, double &hl1)
{
    asm("{\n"
        ".reg .pred P0, P1, S, q, r;\n"
        ".reg .b32 dg, sg, a, b, z1;\n"
        ".reg .f64 u, v, ae, t;\n"
        "and.b32 dg, %5, 0xffff;\n"
        "and.b32 sg, %5, 0x80000000;\n"
        "add.s32 z1, %4, 1;\n"
        "setp.lt.s32 P0, %4, dg;\n"
        "setp.lt.s32 P1, z1, dg;\n"
        "setp.lt.s32 S, %5, 0;\n"
        // level z0
        "setp.lt.f64 r, %8, %12;\n"
        "selp.f64 u, %8, %12, r;\n"
        "setp.lt.f64 r, %6, %14;\n"
        "selp.f64 v, %6, %14, r;\n"
        "setp.ge.xor.f64 q, %10, 0d0000000000000000, S;\n"
        "selp.f64 ae, u, v, q;\n"
        "mul.rn.f64 %2, ae, %10;\n"
        "mov.b64 {a, b}, %2;\n"
        "xor.b32 b, b, sg;\n"
        "mov.b64 t, {a, b};\n"
        "mul.rn.f64 t, t, %16;\n"
        "@P0 add.rn.f64 %0, %0, t;\n"
        // level z0+1
        "setp.lt.f64 r, %9, %13;\n"
        "selp.f64 u, %9, %13, r;\n"
        "setp.lt.f64 r, %7, %15;\n"
        "selp.f64 v, %7, %15, r;\n"
        "setp.ge.xor.f64 q, %11, 0d0000000000000000, S;\n"
        "selp.f64 ae, u, v, q;\n"
        "mul.rn.f64 %3, ae, %11;\n"
        "mov.b64 {a, b}, %3;\n"
        "xor.b32 b, b, sg;\n"
        "mov.b64 t, {a, b};\n"
        "mul.rn.f64 t, t, %17;\n"
        "@P1 add.rn.f64 %1, %1, t;\n"
        "}"
        : "+d"(dh0), "+d"(dh1), "=&d"(hl0), "=&d"(hl1)
        : "r"(z0), "r"(meta), "d"(po.x), "d"(po.y), "d"(mo.x), "d"(mo.y), "d"(h.x), "d"(h.y), "d"(p10), "d"(p11),
          "d"(m10), "d"(m11), "d"(ar0), "d"(ar1));
}

struct WtView {
    int n_copies, n_nodes, n_witems;
    unsigned char *blob, *rowsA, *rowsB, *erows;
    const int4 *hdr, *ent;
    const unsigned short *sched;
    int rows_bytes, erows_bytes, tx_bytes;
    const int2 *copies;
};

__device__ __forceinline__ WtView wt_view(unsigned char *stage)
{
    WtView v;
    v.blob = stage;
    const int4 h0 = reinterpret_cast<const int4 *>(stage)[0];
    const int4 h1 = reinterpret_cast<const int4 *>(stage)[1];
    const int4 h2 = reinterpret_cast<const int4 *>(stage)[2];
    v.n_copies = h0.x;
    v.n_nodes = h0.z;
    v.n_witems = h0.w;
    v.rows_bytes = h2.y;
    v.erows_bytes = h2.z;
    v.tx_bytes = h2.w;
    v.rowsA = stage + h2.x;
    v.rowsB = v.rowsA + h2.y;
    v.erows = v.rowsB + h2.y;
    v.copies = reinterpret_cast<const int2 *>(stage + WT_HDR_BYTES);
    v.hdr = reinterpret_cast<const int4 *>(stage + h1.y);
    v.ent = reinterpret_cast<const int4 *>(stage + h1.z);
    v.sched = reinterpret_cast<const unsigned short *>(stage + h1.w);
    return v;
}

struct WtItem {
    bool act, out;   // lane holds a slot; the slot stores results (not a ghost)
    int z0, nz, cnt, fm, sd, own;
    unsigned grow;   // element offset of (node, z0) in the padded node arrays
    const int4 *en;
    const unsigned char *ra, *rb, *re;
};

__device__ __forceinline__ WtItem wt_item(const WtView &V, int wi, int lane)
{
    WtItem I;
    const unsigned d = V.sched[wi * 32 + lane];
    I.act = d != WT_IDLE;
    I.out = I.act && !(d & 0x8000u);
    const int4 hd = V.hdr[I.act ? (d & 0xffu) : 0u];
    I.z0 = I.act ? (int)((d >> 8) & 0x7fu) * 2 : 0;
    I.nz = I.act ? (hd.y & 0xff) : 0;
    I.cnt = I.act ? (int)((unsigned)hd.w >> 16) : 0;
    I.fm = (hd.y >> 8) & 0xff;
    I.sd = (hd.y >> 16) & 0xff;
    I.own = hd.z;
    I.grow = (unsigned)hd.x + (unsigned)I.z0;
    I.en = V.ent + (hd.w & 0xffff);
    I.ra = V.rowsA + I.z0 * 8;
    I.rb = V.rowsB + I.z0 * 8;
    I.re = V.erows + I.z0 * 8;
    return I;
}

// The own-column values an item needs FIRST (everything else it loads has the whole edge loop to
// arrive): the raw vertical fluxes of levels z0 .. z0+2 and, in phase B, the cell areas.  Loaded
// one item ahead.
struct WtEarly {
    double f0, f1, f2, a0, a1;
};
template <bool PHASE_A, bool ITER = false>
__device__ __forceinline__ WtEarly wt_early(const Arrays &A, const WtView &V, int wi, int lane, const double *g_v, size_t tn)
{
    WtEarly E;
    E.f0 = E.f1 = E.f2 = 0.;
    E.a0 = E.a1 = 1.;
    const unsigned d = V.sched[wi * 32 + lane];
    if (d != WT_IDLE && !(d & 0x8000u)) {
        const int4 hd = V.hdr[d & 0xffu];
        const int z0 = (int)((d >> 8) & 0x7fu) * 2, nz = hd.y & 0xff;
        const unsigned grow = (unsigned)hd.x + (unsigned)z0;
        const double2 ff = __ldg(reinterpret_cast<const double2 *>(g_v + grow));
        if (z0 + 2 <= nz) E.f2 = __ldg(g_v + grow + 2);
        E.f0 = ff.x;
        E.f1 = ff.y;
        if (PHASE_A) {
            // area_inv is loaded at the top of its item and first used after the edge loop, a few hundred cycles
            // later: less than a DRAM access.  Pull it into L2 one item ahead (one probe per 64 bytes)
            if ((z0 & 7) == 0 && (A.flags & 2)) prefetch_l2(A.area_inv + grow);
        }
        if (!PHASE_A) {
            const double2 aa = __ldg(reinterpret_cast<const double2 *>(A.area + grow));
            E.a0 = aa.x;
            E.a1 = aa.y;
            // the operands of c vertical are consumed after the item's edge loop: pull them into L2
            // now (one probe per 64 bytes) instead of holding 24 registers for loads in flight
            if (!ITER && (z0 & 7) == 0 && !(A.flags & 5)) {   // (bit 4: the issuers prefetch the whole tile, opt 128)
                prefetch_l2(A.del_v + tn + grow);
                prefetch_l2(A.del_h + tn + grow);
                prefetch_l2(A.ttf + tn + grow);
                prefetch_l2(A.lo + tn + grow);
                prefetch_l2(A.hnode + grow);
                prefetch_l2(A.hnode_new + grow);
            }
        }
    }
    return E;
}

// ---- phase A: one warp item ------------------------------------------------------------------------
// What an item does for its successor right before its edge loop: broadcast the index lane 0 drew
// at the top of the item (the shared atomic has long returned) and issue the successor's
// first-needed loads, which then have the edge loop and the epilogue to arrive.
template <bool PHASE_A, bool ITER = false>
__device__ __forceinline__ void wt_next(const Arrays &A, const WtView &V, int lane, size_t tn, const double *g_v,
                                        int raw, int &wn, WtEarly &En)
{
    wn = __shfl_sync(0xffffffffu, raw, 0);
    if (wn < V.n_witems) En = wt_early<PHASE_A, ITER>(A, V, wn, lane, g_v, tn);
}

// VLIMIT_ONE: the vertical 3-point stencil over the cluster bounds (vlimit == 1, reference.cpp:380-392);
// else A.vlimit is 2 or 3 (docs/refactoring.md:113-148): the cluster bound of a level is widened (2) or
// narrowed (3) by the node's own a1 maxima of levels z-1 .. z+1, which stand in the staged own row
template <bool VLIMIT_ONE, bool FOLD = false>
__device__ __forceinline__ void wt_item_a(const Arrays &A, const WtView &V, int wi, int lane, size_t tn,
                                          const double *g_lo, const double *g_v, const WtEarly &E, int raw, int &wn,
                                          WtEarly &En)
{
    const WtItem I = wt_item(V, wi, lane);
    const int z0 = I.z0, nz = I.nz;
    double hi0, hi1, lw0, lw1, p0, p1, m0, m1, l0 = 0., l1 = 0., ai0 = 0., ai1 = 0.;
    {
        // own column: needed after the gather only
        if (I.out) {
            const double2 aa = __ldg(reinterpret_cast<const double2 *>(A.area_inv + I.grow));
            ai0 = aa.x; ai1 = aa.y;
            if (!FOLD) {   // (folded: the staged own row still holds the raw fct_LO, see below)
                const double2 ll = __ldg(reinterpret_cast<const double2 *>(g_lo + I.grow));
                l0 = ll.x; l1 = ll.y;
            }
        }
        // cluster bounds start from the (-big, +big) fill of ring elements that already ended
        // (reference.cpp:341-349) and the node's own a1 bounds
        const bool fl0 = z0 >= I.fm, fl1 = z0 + 1 >= I.fm;
        hi0 = fl0 ? -A.big : -CUDART_INF;
        lw0 = fl0 ? A.big : CUDART_INF;
        hi1 = fl1 ? -A.big : -CUDART_INF;
        lw1 = fl1 ? A.big : CUDART_INF;
        const double2 x = *reinterpret_cast<const double2 *>(I.ra + I.own);
        const double2 y = *reinterpret_cast<const double2 *>(I.rb + I.own);
        if (FOLD) {
            l0 = x.x;
            l1 = x.y;
        }
        if (I.act && z0 < I.sd) {
            hi0 = pick_max(hi0, FOLD ? pick_max(x.x, y.x) : x.x);
            lw0 = pick_min(lw0, FOLD ? pick_min(x.x, y.x) : y.x);
        }
        if (I.act && z0 + 1 < I.sd) {
            hi1 = pick_max(hi1, FOLD ? pick_max(x.y, y.y) : x.y);
            lw1 = pick_min(lw1, FOLD ? pick_min(x.y, y.y) : y.y);
        }
        wt_b1v(E.f0, E.f1, E.f2, p0, p1, m0, m1);
    }
    wt_next<true>(A, V, lane, tn, g_v, raw, wn, En);
    // ---- the node's edges in ascending edge id: a2/a3 bounds + b1 horizontal ----
#pragma unroll 2
    for (int k = 0; k < I.cnt; ++k) {
        const int4 e = I.en[k];
        const double2 x = *reinterpret_cast<const double2 *>(I.ra + e.y);
        const double2 y = *reinterpret_cast<const double2 *>(I.rb + e.y);
        const double2 h = *reinterpret_cast<const double2 *>(I.re + e.x);
        if (FOLD) wt_edge_a_fold(z0, e.z, x, y, h, hi0, hi1, lw0, lw1, p0, p1, m0, m1);
        else wt_edge_a(z0, e.z, x, y, h, hi0, hi1, lw0, lw1, p0, p1, m0, m1);
    }
    // ---- vertical 3-point stencil of a3 (reference.cpp:380-392): neighbouring slots are the
    // neighbouring lanes (ghost slots included) ----
    const double pmax = __shfl_up_sync(0xffffffffu, hi1, 1), pmin = __shfl_up_sync(0xffffffffu, lw1, 1);
    const double nmax = __shfl_down_sync(0xffffffffu, hi0, 1), nmin = __shfl_down_sync(0xffffffffu, lw0, 1);
    if (!I.out) return;
    double bm0 = hi0, bn0 = lw0, bm1 = hi1, bn1 = lw1;
    if (VLIMIT_ONE) {
        if (z0 > 0 && z0 < nz - 1) {
            bm0 = pick_max(pick_max(pmax, hi0), hi1);
            bn0 = pick_min(pick_min(pmin, lw0), lw1);
        }
        if (z0 + 1 < nz - 1) {
            bm1 = pick_max(pick_max(hi0, hi1), nmax);
            bn1 = pick_min(pick_min(lw0, lw1), nmin);
        }
    } else {
        // own a1 maxima of levels z0-1 .. z0+2 (the listing takes maxval AND minval from fct_ttf_max)
        const unsigned char *ar = I.ra + I.own;
        const double2 a = *reinterpret_cast<const double2 *>(ar);
        const bool widen = A.vlimit == 2;
        if (z0 > 0 && z0 < nz - 1) {
            const double am = *reinterpret_cast<const double *>(ar - 8);
            const double vmax = pick_max(pick_max(am, a.x), a.y), vmin = pick_min(pick_min(am, a.x), a.y);
            bm0 = widen ? pick_max(hi0, vmax) : pick_min(hi0, vmax);
            bn0 = widen ? pick_min(lw0, vmin) : pick_max(lw0, vmin);
        }
        if (z0 + 1 < nz - 1) {
            const double ap = *reinterpret_cast<const double *>(ar + 16);
            const double vmax = pick_max(pick_max(a.x, a.y), ap), vmin = pick_min(pick_min(a.x, a.y), ap);
            bm1 = widen ? pick_max(hi1, vmax) : pick_min(hi1, vmax);
            bn1 = widen ? pick_min(lw1, vmin) : pick_max(lw1, vmin);
        }
    }
    bm0 -= l0;
    bn0 -= l0;
    bm1 -= l1;
    bn1 -= l1;
    // ---- b2, reference.cpp:432-435 ----
    const double pf0 = limit_quotient(bm0, p0 * A.dt * ai0 + A.eps);
    const double mf0 = limit_quotient(bn0, m0 * A.dt * ai0 - A.eps);
    const double pf1 = limit_quotient(bm1, p1 * A.dt * ai1 + A.eps);
    const double mf1 = limit_quotient(bn1, m1 * A.dt * ai1 - A.eps);
    const size_t off = tn + I.grow;
    const bool both = z0 + 1 < nz;
    wt_store2(A.ttf_max + off, bm0, bm1, both);
    wt_store2(A.ttf_min + off, bn0, bn1, both);
    wt_store2(A.plus + off, pf0, pf1, both);
    wt_store2(A.minus + off, mf0, mf1, both);
}

// ---- phase B: one warp item ------------------------------------------------------------------------
__device__ __forceinline__ void wt_item_b(const Arrays &A, const WtView &V, int wi, int lane, size_t tn,
                                          const double *g_v, double *g_vout, double *g_ho, const WtEarly &E, int raw,
                                          int &wn, WtEarly &En)
{
    const WtItem I = wt_item(V, wi, lane);
    wt_next<false>(A, V, lane, tn, g_v, raw, wn, En);
    if (!I.out) return;   // ghost slots exist for phase A's stencil only
    const int z0 = I.z0, nz = I.nz;
    const size_t off = tn + I.grow;
    const double2 q_dh = *reinterpret_cast<const double2 *>(A.del_h + off);
    // own factors from the staged rows: levels z0-1 .. z0+2
    const unsigned char *pr = I.ra + I.own, *mr = I.rb + I.own;
    const double2 pp = *reinterpret_cast<const double2 *>(pr);
    const double2 mm = *reinterpret_cast<const double2 *>(mr);
    const double p_m1 = *reinterpret_cast<const double *>(pr - 8);
    const double m_m1 = *reinterpret_cast<const double *>(mr - 8);
    const double p_p2 = *reinterpret_cast<const double *>(pr + 16);
    const double m_p2 = *reinterpret_cast<const double *>(mr + 16);
    // ---- b3 vertical, docs/refactoring.md:205-231 (the bottom flux stays) ----
    // Interface z between levels z-1 and z: flux >= 0: ae = min(1, minus[z-1], plus[z]); else min(1, plus[z-1],
    // minus[z]); the surface (z = 0) sees only its own level.  Branch-free: both candidates from the hoisted
    // p1 = min(1, plus), m1 = min(1, minus) of the two own levels, then one select on the sign (a warp holds both
    // signs, so the branchy form ran both sides anyway); ae = 1 leaves a flux untouched exactly.
    const double p10 = min_one(pp.x), p11 = min_one(pp.y), m10 = min_one(mm.x), m11 = min_one(mm.y);
    double fl0, fl1, fl2;
    {
        const bool top = z0 == 0;
        const double up_m = top ? 1. : m_m1, up_p = top ? 1. : p_m1;       // level z0-1 (none above the surface)
        const double ae0 = (E.f0 >= 0.) ? pick_min(p10, up_m) : pick_min(m10, up_p);
        const double ae1 = (E.f1 >= 0.) ? pick_min(p11, mm.x) : pick_min(m11, pp.x);
        const double ae2 = (E.f2 >= 0.) ? pick_min(min_one(p_p2), mm.y) : pick_min(min_one(m_p2), pp.y);
        fl0 = ae0 * E.f0;
        fl1 = ((z0 + 1 < nz) ? ae1 : 1.) * E.f1;
        fl2 = ((z0 + 2 < nz) ? ae2 : 1.) * E.f2;
    }
    // (a cached dt / area array instead of these two divisions was measured neutral: profiles/r2_v15_*)
    const double ar0 = A.dt / E.a0, ar1 = A.dt / E.a1;
    double dh0 = q_dh.x, dh1 = q_dh.y;
    // (p10 .. m11: the own node's share of every edge factor, see wt_edge_b)
    // ---- b3 horizontal + c horizontal over the node's edges, ascending edge id ----
#pragma unroll 2
    for (int k = 0; k < I.cnt; ++k) {
        const int4 e = I.en[k];
        const double2 po = *reinterpret_cast<const double2 *>(I.ra + e.y);
        const double2 mo = *reinterpret_cast<const double2 *>(I.rb + e.y);
        const double2 h = *reinterpret_cast<const double2 *>(I.re + e.x);
        double hl0, hl1;
        wt_edge_b(z0, e.z, po, mo, h, p10, p11, m10, m11, ar0, ar1, dh0, dh1, hl0, hl1);
        const int dg = e.z & 0xffff;
        if ((e.z & 0x40000000) && z0 < dg) wt_store2(g_ho + (unsigned)e.w + z0, hl0, hl1, z0 + 1 < dg);
    }
    // ---- c vertical, docs/refactoring.md:295-300 (operands prefetched into L2 one item ahead) ----
    const double2 q_dv = *reinterpret_cast<const double2 *>(A.del_v + off);
    const double2 q_t = __ldg(reinterpret_cast<const double2 *>(A.ttf + off));
    const double2 q_l = __ldg(reinterpret_cast<const double2 *>(A.lo + off));
    const double2 q_hn = __ldg(reinterpret_cast<const double2 *>(A.hnode + I.grow));
    const double2 q_hw = __ldg(reinterpret_cast<const double2 *>(A.hnode_new + I.grow));
    const double dv0 = q_dv.x - q_t.x * q_hn.x + q_l.x * q_hw.x + (fl0 - fl1) * ar0;
    const double dv1 = q_dv.y - q_t.y * q_hn.y + q_l.y * q_hw.y + (fl1 - fl2) * ar1;
    const bool both = z0 + 1 < nz;
    wt_store2(g_vout + I.grow, fl0, fl1, both);
    wt_store2(A.del_v + off, dv0, dv1, both);
    wt_store2(A.del_h + off, dh0, dh1, both);
}


// a / b exactly as IEEE division; a zero numerator (a fully rejected or absent flux) takes the
// branch-free detour of limit_quotient instead of the division's slow path
__device__ __forceinline__ double div_exact(double a, double b)
{
    const bool z = a == 0.;
    const double q = (z ? 1. : a) / b;
    return z ? a * q : q;
}

// ---- phase B of the iterative branch (iter_yn, docs/refactoring.md:226-290): one warp item ----------
// b3 vertical + b3 horizontal as in wt_item_b, but what leaves the item is the REJECTED part of every
// flux, (1-ae)*flux, into fct_adf_v2 / fct_adf_h2 (md:228-230, :258-260), and the low-order solution
// updated with the limited parts (md:265-287: the vertical term, then the node's edges in ascending
// id; x*dt/area/hnode_new evaluated left to right).  The limited fluxes themselves are not stored:
// the subroutine ends with fct_adf_* = fct_adf_*2.
__device__ __forceinline__ void wt_item_b_iter(const Arrays &A, const WtView &V, int wi, int lane, size_t tn,
                                               const double *g_v, double *g_v2, double *g_h2, const WtEarly &E, int raw,
                                               int &wn, WtEarly &En)
{
    const WtItem I = wt_item(V, wi, lane);
    wt_next<false, true>(A, V, lane, tn, g_v, raw, wn, En);
    if (!I.out) return;
    const int z0 = I.z0, nz = I.nz;
    const size_t off = tn + I.grow;
    double *lo = const_cast<double *>(A.lo) + off;
    const double2 q_l = *reinterpret_cast<const double2 *>(lo);
    const double2 q_hw = __ldg(reinterpret_cast<const double2 *>(A.hnode_new + I.grow));
    const unsigned char *pr = I.ra + I.own, *mr = I.rb + I.own;
    const double2 pp = *reinterpret_cast<const double2 *>(pr);
    const double2 mm = *reinterpret_cast<const double2 *>(mr);
    const double p_m1 = *reinterpret_cast<const double *>(pr - 8);
    const double m_m1 = *reinterpret_cast<const double *>(mr - 8);
    const double p_p2 = *reinterpret_cast<const double *>(pr + 16);
    const double m_p2 = *reinterpret_cast<const double *>(mr + 16);
    // ---- b3 vertical: factors of levels z0, z0+1 and (for the difference of level z0+1) z0+2 ----
    double ae0 = 1., ae1 = 1., ae2 = 1.;
    if (z0 == 0) {
        ae0 = pick_min(ae0, (E.f0 >= 0.) ? pp.x : mm.x);
    } else if (E.f0 >= 0.) {
        ae0 = pick_min(ae0, m_m1);
        ae0 = pick_min(ae0, pp.x);
    } else {
        ae0 = pick_min(ae0, p_m1);
        ae0 = pick_min(ae0, mm.x);
    }
    const bool in1 = z0 + 1 < nz, in2 = z0 + 2 < nz;
    if (in1) {
        if (E.f1 >= 0.) {
            ae1 = pick_min(ae1, mm.x);
            ae1 = pick_min(ae1, pp.y);
        } else {
            ae1 = pick_min(ae1, pp.x);
            ae1 = pick_min(ae1, mm.y);
        }
    }
    if (in2) {
        if (E.f2 >= 0.) {
            ae2 = pick_min(ae2, mm.y);
            ae2 = pick_min(ae2, p_p2);
        } else {
            ae2 = pick_min(ae2, pp.y);
            ae2 = pick_min(ae2, m_p2);
        }
    }
    const double fl0 = ae0 * E.f0, fl1 = in1 ? ae1 * E.f1 : E.f1, fl2 = in2 ? ae2 * E.f2 : E.f2;
    if (z0 > 0) g_v2[I.grow] = (1.0 - ae0) * E.f0;
    if (in1) g_v2[I.grow + 1] = (1.0 - ae1) * E.f1;
    // ---- low-order update, vertical term ----
    double l0 = q_l.x + div_exact(div_exact((fl0 - fl1) * A.dt, E.a0), q_hw.x);
    double l1 = q_l.y + div_exact(div_exact((fl1 - fl2) * A.dt, E.a1), q_hw.y);
    // ---- b3 horizontal + low-order update over the node's edges, ascending edge id ----
    for (int k = 0; k < I.cnt; ++k) {
        const int4 e = I.en[k];
        const double2 po = *reinterpret_cast<const double2 *>(I.ra + e.y);
        const double2 mo = *reinterpret_cast<const double2 *>(I.rb + e.y);
        const double2 h = *reinterpret_cast<const double2 *>(I.re + e.x);
        const int dg = e.z & 0xffff;
        const bool second = e.z < 0, writer = e.z & 0x40000000;
        double rj0 = 0., rj1 = 0.;
        if (z0 < dg) {
            const double p1 = second ? po.x : pp.x, m1 = second ? mo.x : mm.x;
            const double p2 = second ? pp.x : po.x, m2 = second ? mm.x : mo.x;
            double ae = 1.;
            if (h.x >= 0.) {
                ae = pick_min(ae, p1);
                ae = pick_min(ae, m2);
            } else {
                ae = pick_min(ae, m1);
                ae = pick_min(ae, p2);
            }
            rj0 = (1.0 - ae) * h.x;
            const double x = div_exact(div_exact(ae * h.x * A.dt, E.a0), q_hw.x);
            l0 = second ? l0 - x : l0 + x;
        }
        if (z0 + 1 < dg) {
            const double p1 = second ? po.y : pp.y, m1 = second ? mo.y : mm.y;
            const double p2 = second ? pp.y : po.y, m2 = second ? mm.y : mo.y;
            double ae = 1.;
            if (h.y >= 0.) {
                ae = pick_min(ae, p1);
                ae = pick_min(ae, m2);
            } else {
                ae = pick_min(ae, m1);
                ae = pick_min(ae, p2);
            }
            rj1 = (1.0 - ae) * h.y;
            const double x = div_exact(div_exact(ae * h.y * A.dt, E.a1), q_hw.y);
            l1 = second ? l1 - x : l1 + x;
        }
        if (writer && z0 < dg) wt_store2(g_h2 + (unsigned)e.w + z0, rj0, rj1, z0 + 1 < dg);
    }
    wt_store2(lo, l0, l1, in1);
}

// ------------------------------------------------------------------------------------------------
// The persistent kernel.  PHASE_A: a1 + a2 + a3 + b1 vertical + b1 horizontal + b2;
// else: b3 vertical + b3 horizontal + c vertical + c horizontal.
// dynamic smem: WT_SMEM_HEAD + NSTAGE * stage_bytes
// ------------------------------------------------------------------------------------------------
// RC > 0: register re-allocation between the warp roles (setmaxnreg).  The CTA is launched at <= 64
// registers per thread; warps 0..3 form ONE producer warpgroup (fetcher, NPW issuers, in phase A the
// converters; a warp without a role leaves at once) that shrinks to WT_PRODUCER_REGS, and the NWC
// consumer warps (a multiple of 4: whole warpgroups) grow to RC each: 24 consumers at 80 registers
// or 20 at 96 instead of 17..21 at 80 -- the producers no longer pin registers they never use.
// setmaxnreg moves registers INSIDE the CTA's launch allocation (threads x registers of the compiled
// kernel): the launch must already own what the roles hold afterwards.  28 warps x 72 registers = 64 512 =
// 4 producer warps x 24 + 24 consumer warps x 80.
// With RC == 72 the producer side is TWO warpgroups (fetcher, four issuers, the converters) at 32 registers
// and the CTA is launched with 32 warps at 64: 8 x 32 x 32 + 24 x 32 x 72 = 63 488 of 65 536.
constexpr int WT_PRODUCER_WARPS = 4;
constexpr int WT_PRODUCER_REGS = 24;
__host__ __device__ constexpr int wt_producer_warps(int rc) { return rc == 72 ? 8 : 4; }
__host__ __device__ constexpr int wt_producer_regs(int rc) { return rc == 72 ? 32 : 24; }
template <int N>
__device__ __forceinline__ void reg_dealloc()
{
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_alloc()
{
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}

// NCV: phase A's a1 pass over the landed rows.  NCV > 0: that many dedicated converter warps (round 1).
// NCV == 0 (pipeline v2): the pass is cut into WT_CONV_CHUNKS chunks that the CONSUMER warps draw from the
// tile's item counter ahead of its warp items: the first warps to reach a tile convert it, in parallel and at
// full issue priority, instead of two converter warps taking 2.2 us of every 4 us refill (WT_TRACE).
constexpr int WT_CONV_CHUNKS = 16;
// Tile-level L2 prefetch (WarpTilesDev::opt bit 128): the arrays the consumers load straight from global memory,
// over the runs of own columns the blob lists behind its copies, one bulk prefetch per (run, array), issued by the
// issuer warps while the tile is still a tile period away -- instead of one L2 probe per 64 bytes and array issued
// by every consumer item one item ahead (ncu: a fifth of phase B's stall samples wait for exactly those loads).
template <bool PHASE_A, bool ITER>
__device__ __forceinline__ void wt_prefetch_own(const Arrays &A, const int2 *entries, int n_pf, int tr, int idx, int stride)
{
    constexpr int NA = PHASE_A ? 2 : 8;
    for (int u = idx; u < n_pf * NA; u += stride) {
        const int2 r = entries[u / NA];
        const uint32_t sz = (((uint32_t)r.y >> 14) & 0x3fffu) << 4;
        const size_t g = (uint32_t)r.x;
        const double *p;
        switch (u % NA) {
        case 0: p = A.adf_v + tr * A.ts_nodev + g; break;
        case 1: p = (PHASE_A ? A.area_inv : A.area) + g; break;
        case 2: p = A.del_v + tr * A.ts_node + g; break;
        case 3: p = A.del_h + tr * A.ts_node + g; break;
        case 4: p = A.ttf + tr * A.ts_node + g; break;
        case 5: p = A.lo + tr * A.ts_node + g; break;
        case 6: p = A.hnode + g; break;
        default: p = A.hnode_new + g; break;
        }
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(sz) : "memory");
    }
}

template <bool PHASE_A, int NSTAGE, int NWC, int NPW, bool VLIMIT_ONE = true, bool ITER = false, int RC = 0, int NCV = WT_CONVERTERS>
__global__ void __launch_bounds__((RC > 0 ? wt_producer_warps(RC) : NPW + 1 + ((PHASE_A && NCV > 0) ? NCV : 0)) * 32 + NWC * 32, 1)
k_phase_warp(Arrays A, WarpTilesDev T, int ntracers, int stage_bytes, int *sched_ctr)
{
    extern __shared__ __align__(128) unsigned char wt_sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // first consumer warp
    constexpr int NPROD = RC > 0 ? wt_producer_warps(RC) : NPW + 1 + ((PHASE_A && NCV > 0) ? NCV : 0);
    constexpr int PREGS = wt_producer_regs(RC);
    constexpr bool FOLD = PHASE_A && NCV < 0;     // no a1 pass: the consumers read the raw (fct_LO, ttf) rows
    static_assert(!FOLD || VLIMIT_ONE, "the vlimit variants read the own a1 maxima from the converted row");
    constexpr bool CONSUMERS_CONVERT = PHASE_A && NCV == 0;
    constexpr int NCH = CONSUMERS_CONVERT ? WT_CONV_CHUNKS : 0;   // conversion chunks ahead of a tile's warp items
    static_assert(RC == 0 || (NPW + 1 + ((PHASE_A && NCV > 0) ? NCV : 0) <= NPROD && NWC % 4 == 0 &&
                              NPROD * 32 * PREGS + NWC * 32 * RC <= (NPROD + NWC) * 32 * ((65536 / ((NPROD + NWC) * 32)) & ~7) &&
                              RC % 8 == 0),
                  "register budget of the re-allocated roles");
    const uint32_t bar = smem_u32(wt_sm);
    // barriers: [0,NSTAGE) stage empty, [NSTAGE,2N) blob landed, [2N,3N) rows landed, [3N,4N) a1 done,
    // [4N,5N) copy list landed in its slot, [5N,6N) copy list consumed
    auto b_empty = [&](int s) { return bar + 8u * s; };
    auto b_blob = [&](int s) { return bar + 8u * (NSTAGE + s); };
    auto b_rows = [&](int s) { return bar + 8u * (2 * NSTAGE + s); };
    auto b_ready = [&](int s) { return bar + 8u * (3 * NSTAGE + s); };
    auto b_pre = [&](int s) { return bar + 8u * (4 * NSTAGE + s); };
    auto b_prefree = [&](int s) { return bar + 8u * (5 * NSTAGE + s); };
    int *next_item = reinterpret_cast<int *>(wt_sm + 8 * 6 * NSTAGE);
    int *tile_of = next_item + NSTAGE;   // (tile, tracer) index staged in each stage, -1: no more work
    int *tracer_of = tile_of + NSTAGE;   // its tracer (one integer division per tile, not one per warp)
    int *pre_tracer = tracer_of + NSTAGE;   // tracer of the copy list in each slot, -1: no more work
    static_assert(8 * 6 * NSTAGE + 16 * NSTAGE <= WT_SMEM_CTRL && NSTAGE <= 4, "smem head");
    const int total = T.ntiles * ntracers;
    const bool lists_ahead = T.opt & 4;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(b_empty(s), NWC);
            mbar_init(b_blob(s), 1);
            mbar_init(b_rows(s), 1);
            mbar_init(b_ready(s), CONSUMERS_CONVERT ? WT_CONV_CHUNKS : (NCV > 0 ? NCV : 1));
            mbar_init(b_pre(s), 1);
            mbar_init(b_prefree(s), NPW);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (RC > 0) {
        if (warp < NPROD) reg_dealloc<PREGS>();
        else reg_alloc<(RC > 0 ? RC : 80)>();
    }

    if (warp == 0) {
        // ---- blob fetcher: refills a stage as soon as every consumer warp has left it.  Tiles are
        // drawn from a device-wide counter, so the CTAs work through the space-filling curve in
        // order: the tiles in flight stay a compact patch whose halo rows hit L2, and no SM idles at
        // the end.  (sched_ctr == nullptr: static round-robin.) ----
        for (int it = 0;; ++it) {
            const int s = it % NSTAGE;
            int v = 0;   // drawn before the wait: the device-wide atomic is off the refill's critical path
            if (lane == 0) v = sched_ctr ? atomicAdd(sched_ctr, 1) : (int)(blockIdx.x + (unsigned)it * gridDim.x);
            v = __shfl_sync(0xffffffffu, v, 0);
            unsigned b0 = 0, b1 = 0;
            if (lane == 0 && v < total) {
                b0 = __ldg(T.blob_off + v % T.ntiles);
                b1 = __ldg(T.blob_off + v % T.ntiles + 1);
                // the stage is still busy: have the blob wait in L2
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(T.blob + b0), "r"((b1 - b0) * 16u) : "memory");
            }
            if (lists_ahead) {
                // the copy list goes ahead into its slot as soon as the issuers are done with the slot's last list
                mbar_wait_idle(b_prefree(s), ((it / NSTAGE) & 1) ^ 1, T.opt);
                if (lane == 0) {
                    if (v < total) {
                        const uint32_t nb = min((b1 - b0) * 16u, (uint32_t)WT_PRE_BYTES);
                        pre_tracer[s] = v / T.ntiles;
                        mbar_expect_tx(b_pre(s), nb);
                        bulk_g2s(smem_u32(wt_sm + WT_SMEM_CTRL + s * WT_PRE_BYTES), T.blob + b0, nb, b_pre(s));
                    } else {
                        pre_tracer[s] = -1;
                        mbar_arrive(b_pre(s));
                    }
                }
            }
            if (T.opt & 8) mbar_wait_handover(b_empty(s), ((it / NSTAGE) & 1) ^ 1);
            else mbar_wait_idle(b_empty(s), ((it / NSTAGE) & 1) ^ 1, T.opt);
            wt_trace(T, it, 0);
            if (v >= total) {
                if (lane == 0) {
                    tile_of[s] = -1;
                    mbar_arrive(b_blob(s));
                }
                break;
            }
            if (lane == 0) {
                next_item[s] = 0;
                tile_of[s] = v;
                tracer_of[s] = v / T.ntiles;
                mbar_expect_tx(b_blob(s), (b1 - b0) * 16u);
                bulk_g2s(smem_u32(wt_sm + WT_SMEM_HEAD + (size_t)s * stage_bytes), T.blob + b0, (b1 - b0) * 16u, b_blob(s));
            }
            wt_trace(T, it, 1);
            __syncwarp();
        }
        // the last CTA to finish drawing rearms the counter for the next launch
        if (sched_ctr && lane == 0) {
            if (atomicAdd(sched_ctr + 1, 1) == (int)gridDim.x - 1) {
                sched_ctr[0] = 0;
                sched_ctr[1] = 0;
            }
        }
    } else if (warp <= NPW) {
        // ---- copy issuers: one bulk copy per staged row, the list shared by NPW warps ----
        for (int it = 0; lists_ahead; ++it) {
            // list from its own slot: rows pulled into L2 while the stage is still busy, copies issued
            // the moment it is released (next to the blob's copy, not after it)
            const int s = it % NSTAGE;
            mbar_wait_idle(b_pre(s), (it / NSTAGE) & 1, T.opt);
            const int tr = pre_tracer[s];
            if (tr < 0) break;
            const unsigned char *pl = wt_sm + WT_SMEM_CTRL + s * WT_PRE_BYTES;
            const int n_copies = reinterpret_cast<const int4 *>(pl)[0].x;
            const int4 h2 = reinterpret_cast<const int4 *>(pl)[2];
            const int2 *copies = reinterpret_cast<const int2 *>(pl + WT_HDR_BYTES);
            const double *ga = (PHASE_A ? A.lo : A.plus) + tr * A.ts_node;
            const double *gb = (PHASE_A ? A.ttf : A.minus) + tr * A.ts_node;
            const double *ge = A.adf_h_in + tr * A.ts_edge;
            for (int u = (warp - 1) * 32 + lane; u < n_copies; u += NPW * 32) {
                const int2 r = copies[u];
                const uint32_t sz = (((uint32_t)r.y >> 14) & 0x3fffu) << 4, arr = ((uint32_t)r.y >> 28) & 3u;
                const double *src = (arr == 0 ? ga : (arr == 1 ? gb : ge)) + (uint32_t)r.x;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(sz) : "memory");
            }
            if ((T.opt & 384) == 128 && !ITER)   // two tile periods ahead (with the rows) ...
                wt_prefetch_own<PHASE_A, ITER>(A, copies + n_copies, reinterpret_cast<const int4 *>(pl)[3].x, tr, (warp - 1) * 32 + lane, NPW * 32);
            if (T.opt & 8) mbar_wait_handover(b_empty(s), ((it / NSTAGE) & 1) ^ 1);
            else mbar_wait_idle(b_empty(s), ((it / NSTAGE) & 1) ^ 1, T.opt);
            if (warp == 1) wt_trace(T, it, 2);
            if (warp == 1 && lane == 0) mbar_expect_tx(b_rows(s), (uint32_t)h2.w);
            const uint32_t sa = smem_u32(wt_sm + WT_SMEM_HEAD + (size_t)s * stage_bytes + h2.x);
            for (int u = (warp - 1) * 32 + lane; u < n_copies; u += NPW * 32) {
                const int2 r = copies[u];
                const uint32_t so = ((uint32_t)r.y & 0x3fffu) << 4, sz = (((uint32_t)r.y >> 14) & 0x3fffu) << 4;
                const uint32_t arr = ((uint32_t)r.y >> 28) & 3u;
                const double *src = (arr == 0 ? ga : (arr == 1 ? gb : ge)) + (uint32_t)r.x;
                bulk_g2s(sa + so, src, sz, b_rows(s));
            }
            if ((T.opt & 384) == 384 && !ITER)   // ... or (bit 256) one tile period ahead, after the stage's own copies
                wt_prefetch_own<PHASE_A, ITER>(A, copies + n_copies, reinterpret_cast<const int4 *>(pl)[3].x, tr, (warp - 1) * 32 + lane, NPW * 32);
            if (warp == 1) wt_trace(T, it, 3);
            __syncwarp();
            if (lane == 0) mbar_arrive(b_prefree(s));
        }
        for (int it = 0; !lists_ahead; ++it) {
            const int s = it % NSTAGE;
            mbar_wait_idle(b_blob(s), (it / NSTAGE) & 1, T.opt);
            if (tile_of[s] < 0) break;
            const int tr = tracer_of[s];
            if (warp == 1) wt_trace(T, it, 2);
            const WtView V = wt_view(wt_sm + WT_SMEM_HEAD + (size_t)s * stage_bytes);
            // the transaction count may run negative until this arrives; the phase cannot complete before
            if (warp == 1 && lane == 0)
                mbar_expect_tx(b_rows(s), T.diag == 0 ? (uint32_t)V.tx_bytes : (T.diag == 1 ? 2u * (uint32_t)V.rows_bytes : 0u));
            const double *ga = (PHASE_A ? A.lo : A.plus) + tr * A.ts_node;
            const double *gb = (PHASE_A ? A.ttf : A.minus) + tr * A.ts_node;
            const double *ge = A.adf_h_in + tr * A.ts_edge;
            const uint32_t sa = smem_u32(V.rowsA);
            for (int u = (warp - 1) * 32 + lane; u < V.n_copies; u += NPW * 32) {
                const int2 r = V.copies[u];
                const uint32_t so = ((uint32_t)r.y & 0x3fffu) << 4, sz = (((uint32_t)r.y >> 14) & 0x3fffu) << 4;
                const uint32_t arr = ((uint32_t)r.y >> 28) & 3u;
                const double *src = (arr == 0 ? ga : (arr == 1 ? gb : ge)) + (uint32_t)r.x;
                if (T.diag == 0 || (T.diag == 1 && arr != 2)) bulk_g2s(sa + so, src, sz, b_rows(s));
            }
            if ((T.opt & 128) && !ITER)
                wt_prefetch_own<PHASE_A, ITER>(A, V.copies + V.n_copies, reinterpret_cast<const int4 *>(V.blob)[3].x, tr, (warp - 1) * 32 + lane, NPW * 32);
            if (warp == 1) wt_trace(T, it, 3);
            __syncwarp();
        }
    } else if (PHASE_A && NCV > 0 && warp <= NPW + NCV) {
        // ---- phase A: a1 in place on the landed rows, (fct_LO, ttf) -> (max, min), reference.cpp:315-316 ----
        const int cl = (warp - NPW - 1) * 32 + lane;   // lane among the converter warps
        for (int it = 0;; ++it) {
            const int s = it % NSTAGE;
            mbar_wait_idle(b_blob(s), (it / NSTAGE) & 1, T.opt);
            if (tile_of[s] < 0) break;
            const WtView V = wt_view(wt_sm + WT_SMEM_HEAD + (size_t)s * stage_bytes);
            mbar_wait_idle(b_rows(s), (it / NSTAGE) & 1, T.opt);
            if (cl < 32) wt_trace(T, it, 4);
            double2 *pa = reinterpret_cast<double2 *>(V.rowsA), *pb = reinterpret_cast<double2 *>(V.rowsB);
            const int n16 = V.rows_bytes >> 4;
#pragma unroll 4
            for (int g = cl; g < n16; g += 32 * (NCV > 0 ? NCV : 1)) {
                const double2 l = pa[g], t = pb[g];
                pa[g] = make_double2(pick_max(l.x, t.x), pick_max(l.y, t.y));
                pb[g] = make_double2(pick_min(l.x, t.x), pick_min(l.y, t.y));
            }
            fence_proxy_async();   // the next refill of this stage is written by the async proxy
            if (cl < 32) wt_trace(T, it, 5);
            __syncwarp();
            if (lane == 0) mbar_arrive(b_ready(s));
        }
    } else if (warp >= NPROD) {
        // ---- consumers ----
        // (with CONSUMERS_CONVERT the first NCH draws of a tile are its conversion chunks: indices -NCH .. -1)
        auto draw = [&](int s) {   // lane 0 only; broadcast with __shfl_sync when needed
            int wi = 0;
            if (lane == 0) wi = smem_fetch_add(next_item + s) - NCH;
            return wi;
        };
        for (int it = 0;; ++it) {
            const int s = it % NSTAGE;
            mbar_wait(b_blob(s), (it / NSTAGE) & 1);
            if (tile_of[s] < 0) break;
            const int tr = tracer_of[s];
            const bool first_loads_ahead = T.opt & 2;
            if (!first_loads_ahead && !CONSUMERS_CONVERT) mbar_wait((PHASE_A && !FOLD) ? b_ready(s) : b_rows(s), (it / NSTAGE) & 1);
            const WtView V = wt_view(wt_sm + WT_SMEM_HEAD + (size_t)s * stage_bytes);
            const size_t tn = tr * A.ts_node;
            const double *g_v = A.adf_v + tr * A.ts_nodev;
            // one item ahead: lane 0 draws the next index at the top of an item; the item broadcasts
            // it and loads the successor's first-needed values before its own edge loop.  The first
            // item of a tile needs the blob only (schedule + node headers): its loads travel while the
            // warp waits for the tile's rows
            int wi = __shfl_sync(0xffffffffu, draw(s), 0);
            if (CONSUMERS_CONVERT) {
                // a1 in place on the landed rows, (fct_LO, ttf) -> (max, min), reference.cpp:315-316: the warps
                // that reach the tile first take its chunks
                if (wi < 0) mbar_wait(b_rows(s), (it / NSTAGE) & 1);
                if (wi < 0 && warp == NPROD) wt_trace(T, it, 4);
                while (wi < 0) {
                    double2 *pa = reinterpret_cast<double2 *>(V.rowsA), *pb = reinterpret_cast<double2 *>(V.rowsB);
                    const int n16 = V.rows_bytes >> 4, per = (n16 + NCH - 1) / (NCH > 0 ? NCH : 1);
                    const int g0 = (wi + NCH) * per, g1 = min(g0 + per, n16);
#pragma unroll 4
                    for (int g = g0 + lane; g < g1; g += 32) {
                        const double2 l = pa[g], t = pb[g];
                        pa[g] = make_double2(pick_max(l.x, t.x), pick_max(l.y, t.y));
                        pb[g] = make_double2(pick_min(l.x, t.x), pick_min(l.y, t.y));
                    }
                    fence_proxy_async();   // the next refill of this stage is written by the async proxy
                    __syncwarp();
                    if (lane == 0) mbar_arrive(b_ready(s));
                    wi = __shfl_sync(0xffffffffu, draw(s), 0);
                }
            }
            WtEarly E;
            if (wi < V.n_witems) E = wt_early<PHASE_A, ITER>(A, V, wi, lane, g_v, tn);
            if (first_loads_ahead || CONSUMERS_CONVERT) mbar_wait((PHASE_A && !FOLD) ? b_ready(s) : b_rows(s), (it / NSTAGE) & 1);
            if (warp == NPROD) wt_trace(T, it, 6);
            if (warp == NPROD + NWC - 1) wt_trace(T, it, 8);
            while (wi < V.n_witems) {
                WtEarly En;
                int wn = 0;
                const int raw = draw(s);
                if (PHASE_A) wt_item_a<VLIMIT_ONE, FOLD>(A, V, wi, lane, tn, A.lo + tn, g_v, E, raw, wn, En);
                else if (ITER)
                    wt_item_b_iter(A, V, wi, lane, tn, g_v, A.adf_v2 + tr * A.ts_nodev, A.adf_h2 + tr * A.ts_edge, E, raw, wn, En);
                else
                    wt_item_b(A, V, wi, lane, tn, g_v, A.adf_v_out + tr * A.ts_nodev, A.adf_h_out + tr * A.ts_edge, E, raw,
                              wn, En);
                wi = wn;
                E = En;
            }
            if (warp == NPROD) wt_trace(T, it, 7);
            if (warp == NPROD + NWC - 1) wt_trace(T, it, 9);
            __syncwarp();
            if (lane == 0) mbar_arrive(b_empty(s));
        }
    }
}

}   // namespace fct
