// Thin C++ host driver over CUDA streams: the C ABI of include/fesom2-accelerate.h.
// Replaces the reference's src/fesom2-accelerate.cu (allocation, transfers, streams, the three
// *_comm_acc_ orchestration entry points) and adds the device-resident plan / fields / step path.
// No CPU compute path exists in this file: every stage is a kernel launch.
#include <cuda_runtime.h>

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/fesom2-accelerate.h"
#include "fct_internal.h"
#include "fct_kernels.cuh"
#include "fct_plan.h"

namespace fct {

static std::atomic<long long> g_launches{0};
static std::atomic<int> g_fused{0};

bool cuda_ok(cudaError_t e, const char *what)
{
    if (e != cudaSuccess) {
        std::fprintf(stderr, "fesom2-accelerate: CUDA error \"%s\" (%s) in %s\n", cudaGetErrorName(e),
                     cudaGetErrorString(e), what);
        return false;
    }
    return true;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---------------------------------------------------------------------------------------------
// launch geometry: blockDim = (level slots of one column, columns per block)
// ---------------------------------------------------------------------------------------------
struct Geom {
    dim3 block;
    int ny;
    size_t smem;   // stencil scratch of a3 / phase A
};
static Geom geom(int nl, int vec)
{
    Geom g;
    const int lx = std::max(1, (nl - 1 + vec - 1) / vec);
    int ny = std::max(1, 256 / lx);
    g.block = dim3(lx, ny, 1);
    g.ny = ny;
    g.smem = (size_t)ny * 2 * ((size_t)lx * vec + 2) * sizeof(double);
    return g;
}

typedef void (*kern_t)(Arrays, MeshDev, const int *, int, int);

template <int VEC>
static kern_t kernel_of(int stage)
{
    switch (stage) {
    case ST_A1: return k_a1<VEC>;
    case ST_A2: return k_a2<VEC>;
    case ST_A3: return k_a3<VEC>;
    case ST_B1V: return k_b1v<VEC>;
    case ST_B1H: return k_b1h<VEC>;
    case ST_B2: return k_b2<VEC>;
    case ST_B3V: return k_b3v<VEC>;
    case ST_B3H: return k_b3h<VEC>;
    case ST_CV: return k_cv<VEC>;
    case ST_CH: return k_ch<VEC>;
    case ST_PHASE_A: return k_phaseA<VEC>;
    case ST_PHASE_B: return k_phaseB<VEC>;
    case ST_B1H_ATOMIC: return k_b1h_atomic<VEC>;
    case ST_CH_ATOMIC: return k_ch_atomic<VEC>;
    case ST_A3_VLIMIT2:
    case ST_A3_VLIMIT3: return k_a3_vlimit<VEC>;
    case ST_B3V_ITER: return k_b3v_iter<VEC>;
    case ST_B3H_ITER: return k_b3h_iter<VEC>;
    case ST_LO_UPDATE: return k_lo_update<VEC>;
    }
    return nullptr;
}

// Launch one stage over `count` items (nodes / elements / edges) starting at `first` of `list`
// (or of the identity when list is null) for `ntracers` tracers.
bool launch_stage(int stage, int vec, const Arrays &A, const MeshDev &M, const int *list, int first,
                  int count, int ntracers, cudaStream_t s)
{
    if (count <= 0 || ntracers <= 0) return true;
    kern_t k = (vec == 2) ? kernel_of<2>(stage) : kernel_of<1>(stage);
    if (!k) return false;
    const Geom g = geom(A.nl, vec);
    const bool a3v = stage == ST_A3_VLIMIT2 || stage == ST_A3_VLIMIT3;
    const size_t smem = (stage == ST_A3 || stage == ST_PHASE_A || a3v) ? g.smem : 0;
    dim3 grid((count + g.ny - 1) / g.ny, ntracers, 1);
    if (a3v) {
        Arrays B = A;
        B.vlimit = stage == ST_A3_VLIMIT2 ? 2 : 3;
        k<<<grid, g.block, smem, s>>>(B, M, list, first, count);
    } else {
        k<<<grid, g.block, smem, s>>>(A, M, list, first, count);
    }
    count_launch(1);
    return cuda_ok(cudaGetLastError(), "kernel launch");
}

// ---------------------------------------------------------------------------------------------
// handles
// ---------------------------------------------------------------------------------------------
static const unsigned HANDLE_MAGIC = 0x46435431u;   // "FCT1"

static gpuMemory *new_handle(void *host, size_t bytes, bool create_event)
{
    gpuMemory *h = new (std::nothrow) gpuMemory;
    if (!h) return nullptr;
    std::memset(h, 0, sizeof(*h));
    h->host_pointer = host;
    h->size = bytes;
    h->magic = HANDLE_MAGIC;
    if (!cuda_ok(cudaMalloc(&h->device_pointer, bytes ? bytes : 1), "cudaMalloc")) {
        delete h;
        return nullptr;
    }
    if (create_event) {
        cudaEvent_t ev;
        if (!cuda_ok(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate")) {
            cudaFree(h->device_pointer);
            delete h;
            return nullptr;
        }
        h->event = (void *)ev;
        h->has_event = 1;
    }
    return h;
}

static inline gpuMemory *H(void **p) { return p ? static_cast<gpuMemory *>(*p) : nullptr; }
static inline cudaStream_t S(void **s)
{
    return (s && *s) ? *static_cast<cudaStream_t *>(*s) : (cudaStream_t)0;
}
template <class T>
static inline T *dev(void **p)
{
    gpuMemory *h = H(p);
    return h ? static_cast<T *>(h->device_pointer) : nullptr;
}

static bool h2d(gpuMemory *h, bool sync, cudaStream_t s, bool record)
{
    if (!h || !h->host_pointer) return false;
    if (sync) return cuda_ok(cudaMemcpy(h->device_pointer, h->host_pointer, h->size, cudaMemcpyHostToDevice), "H2D");
    bool ok = cuda_ok(cudaMemcpyAsync(h->device_pointer, h->host_pointer, h->size, cudaMemcpyHostToDevice, s), "H2D async");
    if (record && h->has_event) {
        ok = cuda_ok(cudaEventRecord((cudaEvent_t)h->event, s), "event record") && ok;
        h->event_recorded = 1;
    }
    return ok;
}
static bool d2h(gpuMemory *h, bool sync, cudaStream_t s)
{
    if (!h || !h->host_pointer) return false;
    if (sync) return cuda_ok(cudaMemcpy(h->host_pointer, h->device_pointer, h->size, cudaMemcpyDeviceToHost), "D2H");
    return cuda_ok(cudaMemcpyAsync(h->host_pointer, h->device_pointer, h->size, cudaMemcpyDeviceToHost, s), "D2H async");
}
static void wait_upload(gpuMemory *h, cudaStream_t s)
{
    if (h && h->has_event && h->event_recorded) cudaStreamWaitEvent(s, (cudaEvent_t)h->event, 0);
}

// ---------------------------------------------------------------------------------------------
// plans
// ---------------------------------------------------------------------------------------------
template <class T>
static T *upload_vec(Plan *p, const std::vector<T> &v)
{
    T *d = nullptr;
    const size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    if (!cuda_ok(cudaMalloc(&d, bytes), "cudaMalloc(plan)")) return nullptr;
    p->owned.push_back(d);
    if (!v.empty() && !cuda_ok(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice), "H2D(plan)"))
        return nullptr;
    return d;
}

static bool upload_derived(Plan *p, const DerivedHost &d)
{
    p->dev.nbr_off = upload_vec(p, d.nbr_off);
    p->dev.nbr = upload_vec(p, d.nbr);
    p->dev.fillmin = upload_vec(p, d.fillmin);
    p->dev.edg_off = upload_vec(p, d.edg_off);
    p->dev.edg = upload_vec(p, d.edg);
    p->d_boundary = upload_vec(p, d.boundary);
    p->d_interior = upload_vec(p, d.interior);
    p->boundary_flag.assign((size_t)p->N, 0);
    for (int n : d.boundary) p->boundary_flag[(size_t)n] = 1;
    p->n_boundary = (int)d.boundary.size();
    p->n_interior = (int)d.interior.size();
    return p->dev.nbr_off && p->dev.nbr && p->dev.fillmin && p->dev.edg_off && p->dev.edg &&
           p->d_boundary && p->d_interior;
}

// ---------------------------------------------------------------------------------------------
// tile-staged fused kernels: plan tables, launch
// ---------------------------------------------------------------------------------------------
// tuning knobs: fct_ale_tune_("NAME", value) overrides the environment variable FCT_NAME
static std::map<std::string, int> g_tune;
static std::mutex g_tune_mutex;
static int env_int(const char *name, int dflt)
{
    {
        std::lock_guard<std::mutex> lock(g_tune_mutex);
        auto it = g_tune.find(name);
        if (it != g_tune.end()) return it->second;
    }
    const char *v = std::getenv(name);
    return (v && *v) ? std::atoi(v) : dflt;
}
int tune_int(const char *name, int dflt) { return env_int(name, dflt); }
void set_tune(const char *name, int value)
{
    std::lock_guard<std::mutex> lock(g_tune_mutex);
    g_tune[std::string("FCT_") + name] = value;
}

static const size_t TILE_SMEM_CAP = 113 * 1024;   // at least two CTAs per SM

static bool upload_tileset(Plan *p, const TileSetHost &h, int TN, int TE, TileDev &T)
{
    T.row_off = upload_vec(p, h.row_off);
    T.rows = upload_vec(p, h.rows);
    T.hdr = upload_vec(p, h.hdr);
    T.work_off = upload_vec(p, h.work_off);
    T.ent = upload_vec(p, h.ent);
    T.TN = TN;
    T.TE = TE;
    T.max_rows = h.max_rows;
    T.ntiles = h.ntiles;
    return T.row_off && T.rows && T.hdr && T.work_off && T.ent;
}

static void build_plan_tiles(Plan *p, const DerivedHost &d, const int *nlev_n)
{
    p->tiles_ok = false;
    if (env_int("FCT_TILE", 1) == 0) return;
    const int verbose = env_int("FCT_VERBOSE", 0);
    int TE = 1;
    for (int n = 0; n < p->N; ++n) TE = std::max(TE, d.edg_off[n + 1] - d.edg_off[n]);
    if (TE > 16) {
        if (verbose) std::fprintf(stderr, "fesom2-accelerate: a node has %d edges: using the untiled fused kernels\n", TE);
        return;
    }
    const std::vector<int> *lists[3] = {nullptr, &d.boundary, &d.interior};
    const int nsets = p->H > 0 ? 3 : 1;
    // the two phases want different tile sizes: phase A is issue-bound and prefers many small
    // CTAs per SM, phase B amortises its staging over larger patches (measured, profiles/)
    for (int ph = 0; ph < 2; ++ph) {
        const int TN = std::max(1, env_int(ph == 0 ? "FCT_TILE_NODES_A" : "FCT_TILE_NODES_B", env_int("FCT_TILE_NODES", ph == 0 ? 12 : 24)));
        const int iters = std::max(1, env_int(ph == 0 ? "FCT_TILE_ITERS_A" : "FCT_TILE_ITERS_B", env_int("FCT_TILE_ITERS", ph == 0 ? 1 : 2)));
        TileSetHost h[3];
        for (int s = 0; s < nsets; ++s) {
            if (!build_tileset(d, nlev_n, p->N, p->N + p->H, lists[s], TN, TE, 2, TILE_THREADS, iters * TILE_THREADS, h[s])) {
                if (verbose) std::fprintf(stderr, "fesom2-accelerate: mesh is not a plain triangulation: using the untiled fused kernels\n");
                return;
            }
            const size_t smem = tile_smem_bytes(ph == 0, p->pitch, TN, TE, h[s].max_rows);
            if (smem > TILE_SMEM_CAP) {
                if (verbose)
                    std::fprintf(stderr, "fesom2-accelerate: tile needs %d rows (%zu B smem): using the untiled fused kernels\n",
                                 h[s].max_rows, smem);
                return;
            }
        }
        for (int s = 0; s < nsets; ++s)
            if (!upload_tileset(p, h[s], TN, TE, p->tiles[ph][s])) return;
        if (verbose)
            std::fprintf(stderr, "fesom2-accelerate: phase %c: %d tiles of <= %d nodes, <= %d staged rows, %d entries/node, %zu B smem\n",
                         ph == 0 ? 'A' : 'B', h[0].ntiles, TN, h[0].max_rows, TE, tile_smem_bytes(ph == 0, p->pitch, TN, TE, h[0].max_rows));
    }
    p->tiles_ok = true;
}

// ---- warp-item kernels: plan tables, launch ---------------------------------------------------
// Ring depth (knob WT_STAGES, 0 / unset: automatic).  Measured, interleaved in thermal steady state
// (profiles/r1_v14_ab_ring_depth.log): deep columns leave a third of the ring only ~20 nodes per tile,
// and two larger stages with the copy lists travelling ahead of their blobs (WT_OPT bit 4) beat three
// stages by 2 % per step on the nl = 70 and nl = 80 meshes, while on nl = 48 (34 nodes per tile)
// three stages stay ahead by 2 %.  The padded layout cannot send its lists ahead (too many copies).
static int wt_stages(int nl, bool packed)
{
    const int v = env_int("FCT_WT_STAGES", 0);
    if (v > 0) return std::min(std::max(v, 2), 4);
    return (packed && nl >= 60) ? 2 : 3;
}
// one stage of the ring: an equal share of the 227 KB a CTA may own
static int wt_stage_cap(int stages) { return ((WT_SMEM_MAX - WT_SMEM_HEAD) / stages) & ~127; }

static void build_plan_wtiles(Plan *p, const DerivedHost &d, const int *nlev_n)
{
    p->wtiles_ok = false;
    if (env_int("FCT_WTILE", 1) == 0) return;
    const int verbose = env_int("FCT_VERBOSE", 0);
    const int cap_knob = env_int("FCT_WT_SMEM", 0);   // 0: default
    int TN = env_int("FCT_WT_NODES", 0);
    TN = TN <= 0 ? 96 : std::min(TN, 255);
    const std::vector<int> *lists[3] = {nullptr, &d.boundary, &d.interior};
    const int nsets = p->H > 0 ? 3 : 1;
    for (int layout = 0; layout < 2; ++layout) {
    const bool packed = layout == 1;
    const int cap = cap_knob <= 0 ? wt_stage_cap(wt_stages(p->nl, packed)) : std::min(std::max(cap_knob, 8 * 1024), wt_stage_cap(2));
    if (packed) {
        if (!packed_columns(d, nlev_n, p->N + p->H, p->G, p->ncol, p->ecol)) {
            // fct_ale_fields_create_packed_ then answers istat = 1; the padded layout stays available
            std::fprintf(stderr, "fesom2-accelerate: packed column offsets of this mesh exceed 32 bits: no packed level storage\n");
            p->ncol.clear();
            p->ecol.clear();
            return;
        }
        p->d_ncol = upload_vec(p, p->ncol);
        p->d_ecol = upload_vec(p, p->ecol);
        if (!p->d_ncol || !p->d_ecol) return;
    }
    WarpTilesHost h[3];
    for (int s = 0; s < nsets; ++s) {
        if (!build_warptiles(d, nlev_n, p->N, p->N + p->H, p->G, p->pitch, packed ? p->ncol.data() : nullptr,
                             packed ? p->ecol.data() : nullptr, lists[s], TN, cap, h[s])) {
            if (verbose) std::fprintf(stderr, "fesom2-accelerate: mesh is not eligible for the warp-item kernels\n");
            return;
        }
    }
    for (int s = 0; s < nsets; ++s) {
        WarpTilesDev &T = packed ? p->wtiles_pk[s] : p->wtiles[s];
        T.blob = upload_vec(p, h[s].blob);
        T.blob_off = upload_vec(p, h[s].blob_off);
        T.ntiles = h[s].ntiles;
        T.smem_bytes = h[s].smem_bytes;
        T.max_copies = 0;
        for (int t = 0; t < h[s].ntiles; ++t)   // copies + L2 prefetch entries: what must fit the slot that travels ahead
            T.max_copies = std::max(T.max_copies, (int)h[s].blob[h[s].blob_off[t]].x + (int)h[s].blob[h[s].blob_off[t] + 3].x);
        if (!T.blob || !T.blob_off) return;
        if (verbose && h[s].ntiles > 0)
            std::fprintf(stderr,
                         "fesom2-accelerate: %s warp tiles[%d]: %d tiles, %.1f nodes/tile, %.2f staged rows/node, %.2f staged edge rows/node "
                         "(%.2f edge uses/node), %.1f bulk copies/tile, lane fill %.1f%%, %.0f B plan/node, %d B smem/stage\n",
                         packed ? "packed" : "padded", s, h[s].ntiles, (double)h[s].nodes / h[s].ntiles,
                         (double)h[s].staged_rows / h[s].nodes, (double)h[s].staged_erows / h[s].nodes,
                         (double)h[s].edge_uses / h[s].nodes, (double)h[s].copies / h[s].ntiles,
                         100.0 * h[s].slots / std::max<long long>(h[s].lanes, 1), 16.0 * h[s].blob.size() / h[s].nodes,
                         h[s].smem_bytes);
    }
    if (packed) p->wtiles_pk_ok = true;
    else p->wtiles_ok = true;
    }
}

// Per-DEVICE launch state (a process may drive several GPUs): the opt-in to > 48 KB of dynamic shared
// memory is a per-device function attribute, and so are the SM count and the occupancy.
static std::mutex g_dev_mutex;
static int current_device()
{
    int dev_id = 0;
    cudaGetDevice(&dev_id);
    return dev_id;
}
static bool ensure_smem_attr(const void *fn, size_t smem, bool *first = nullptr)
{
    static std::map<std::pair<int, const void *>, size_t> done;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    size_t &have = done[std::make_pair(current_device(), fn)];
    if (first) *first = false;
    if (smem <= have) return true;
    if (!cuda_ok(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "smem attribute")) return false;
    have = smem;
    if (first) *first = true;
    return true;
}
static int device_sms()
{
    static std::map<int, int> sms;
    std::lock_guard<std::mutex> lock(g_dev_mutex);
    const int dev_id = current_device();
    int &n = sms[dev_id];
    if (n == 0) {
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev_id);
        n = std::max(n, 1);
    }
    return n;
}

static long long *g_trace = nullptr;   // knob WT_TRACE (single device, profiling runs only)

typedef void (*warp_kern_t)(Arrays, WarpTilesDev, int, int, int *);
struct WarpVariant {
    warp_kern_t fn;
    bool phase_a;
    int stages, consumers, issuers;
    int regs = 0;   // > 0: roles re-allocate registers (setmaxnreg): 4 producer warps + consumers at `regs` each
    int conv = 2;   // phase A: dedicated converter warps (2, 3, 4); 0: the consumer warps run the a1 pass in chunks;
                    // -1: a1 folded into the consumers' edge loop (the default shape since round 2)
};
// total warps (1 fetcher + issuers + [converter] + consumers) a multiple of 4: see fct_warp_kernels.cuh
#define WT_VA(S, C, I) {k_phase_warp<true, S, C, I>, true, S, C, I}
#define WT_VB(S, C, I) {k_phase_warp<false, S, C, I>, false, S, C, I}
// register re-allocation between the roles: one producer warpgroup, whole consumer warpgroups
#define WT_RA(S, C, I, R) {k_phase_warp<true, S, C, I, true, false, R>, true, S, C, I, R}
#define WT_RB(S, C, I, R) {k_phase_warp<false, S, C, I, true, false, R>, false, S, C, I, R}
// pipeline v2: the consumers convert the landed rows (no converter warps)
#define WT_A2(S, C, I, R) {k_phase_warp<true, S, C, I, true, false, R, 0>, true, S, C, I, R, 0}
// four converter warps instead of two (the a1 pass is 2.2 us of a 4 us refill with two, knob WT_TRACE)
#define WT_A4(S, C, I) {k_phase_warp<true, S, C, I, true, false, 0, 4>, true, S, C, I, 0, 4}
// 32 warps: two producer warpgroups at 32 registers (fetcher, 4 issuers, 2 or 3 converters) + 24 consumers at 72
#define WT_A72(S, I, V) {k_phase_warp<true, S, 24, I, true, false, 72, V>, true, S, 24, I, 72, V}
#define WT_B72(S, I) {k_phase_warp<false, S, 24, I, true, false, 72>, false, S, 24, I, 72}
// the a1 pass folded into the consumers' edge loop: no converter warps, no pass over the staged rows (conv = -1)
#define WT_AF(S, C, I) {k_phase_warp<true, S, C, I, true, false, 0, -1>, true, S, C, I, 0, -1}
static const WarpVariant g_wvariants[] = {
    WT_AF(2, 19, 4), WT_AF(3, 19, 4), WT_AF(2, 21, 2), WT_AF(3, 21, 2), WT_AF(4, 19, 4), WT_AF(4, 21, 2),
    WT_AF(2, 24, 3), WT_AF(3, 24, 3), WT_AF(2, 23, 4), WT_AF(3, 23, 4),   // 28 warps at 72 registers
    WT_A72(2, 4, 2), WT_A72(3, 4, 2), WT_A72(2, 4, 3), WT_A72(3, 4, 3), WT_B72(2, 4), WT_B72(3, 4),
    WT_A4(2, 15, 4), WT_A4(3, 15, 4), WT_A4(2, 17, 2), WT_A4(3, 17, 2), WT_A4(2, 19, 4), WT_A4(3, 19, 4), WT_A4(4, 17, 2),
    WT_A2(2, 19, 4, 0), WT_A2(3, 19, 4, 0), WT_A2(2, 21, 2, 0), WT_A2(3, 21, 2, 0),
    WT_A2(2, 24, 3, 80), WT_A2(3, 24, 3, 80), WT_A2(2, 24, 1, 80), WT_A2(3, 24, 1, 80),
    WT_A2(2, 24, 3, 0), WT_A2(3, 24, 3, 0), WT_VB(2, 24, 3), WT_VB(3, 24, 3),   // 28 warps at 72 registers, no re-allocation
    WT_RB(2, 24, 3, 80), WT_RB(3, 24, 3, 80),
    WT_RA(2, 24, 1, 80), WT_RA(3, 24, 1, 80), WT_RA(2, 20, 1, 88), WT_RA(3, 20, 1, 88),
    WT_RB(2, 24, 1, 80), WT_RB(3, 24, 1, 80), WT_RB(2, 20, 1, 88), WT_RB(3, 20, 1, 88),
    WT_VA(3, 17, 4), WT_VA(3, 13, 4), WT_VA(3, 15, 6), WT_VA(3, 9, 4), WT_VA(3, 19, 2),
    WT_VA(2, 17, 4), WT_VA(2, 13, 4), WT_VA(2, 15, 6), WT_VA(2, 19, 2), WT_VA(4, 17, 4), WT_VA(4, 15, 6),
    WT_VB(3, 19, 4), WT_VB(3, 15, 4), WT_VB(3, 11, 4), WT_VB(3, 17, 6), WT_VB(3, 21, 2),
    WT_VB(2, 19, 4), WT_VB(2, 15, 4), WT_VB(2, 11, 4), WT_VB(2, 21, 2), WT_VB(4, 19, 4), WT_VB(4, 17, 6),
};
// phase A with vlimit 2 / 3 (docs/refactoring.md:113-148): the default shapes of the packed and of
// the padded layout for every ring depth
#define WT_VL(S, C, I) {k_phase_warp<true, S, C, I, false>, true, S, C, I}
// phase B of the iterative branch (docs/refactoring.md:226-290)
#define WT_IT(S, C, I) {k_phase_warp<false, S, C, I, true, true>, false, S, C, I}
static const WarpVariant g_wvariants_it[] = {
    WT_IT(3, 21, 2), WT_IT(3, 19, 4), WT_IT(2, 19, 4), WT_IT(4, 19, 4),
};
static const WarpVariant g_wvariants_vl[] = {
    WT_VL(3, 19, 2), WT_VL(3, 17, 4), WT_VL(2, 19, 2), WT_VL(2, 17, 4), WT_VL(4, 17, 4),
};


// which: 0 all owned nodes, 1 boundary list, 2 interior list
bool launch_warp(int stage, const Arrays &A, const Plan *p, int which, int ntracers, cudaStream_t s)
{
    const bool isA = stage == ST_PHASE_A, iter = stage == ST_PHASE_B_ITER;
    const bool packed = A.pitchL == 0;   // Arrays of a packed Fields object carry no pitch
    if (iter && (!A.adf_v2 || !A.adf_h2)) return false;
    if (packed ? !p->wtiles_pk_ok : !p->wtiles_ok) {
        std::fprintf(stderr, "fesom2-accelerate: this plan has no warp-item tiles for this layout\n");
        return false;
    }
    WarpTilesDev T = packed ? p->wtiles_pk[which] : p->wtiles[which];
    T.diag = env_int("FCT_WT_DIAG", 0);
    T.opt = env_int("FCT_WT_OPT", -1);   // < 0: automatic, see below
    if (T.ntiles <= 0) return true;
    if (!packed && (A.pitchL != p->pitch || A.pitchV != p->pitch || A.pitchH != p->pitch)) {
        std::fprintf(stderr, "fesom2-accelerate: the warp-item kernels need the plan's padded pitch\n");
        return false;
    }
    const int stage_bytes = (T.smem_bytes + 127) & ~127;
    // as deep a ring as the tiles of this plan admit
    int stages = wt_stages(p->nl, packed);
    while (stages > 2 && WT_SMEM_HEAD + (size_t)stages * stage_bytes > (size_t)WT_SMEM_MAX) --stages;
    // measured (profiles/r1_v11_sched_options_sweep_mid.log, r1_v14_ab_ring_depth.log): first loads before
    // the rows wait (2) is +10 % on phase A; lists ahead (4) pays with a two-stage ring only; suspended
    // producers (1) are neutral
    // round 2 (profiles/r2_v21_*, r2_v23_*): the issuers pull the own columns of the arrays the consumers load from global
    // memory into L2 tile by tile (128): phase A +1..3 %, phase B +3 % with three stages; with two stages (lists ahead)
    // phase B wants it ONE tile period ahead, after the stage's own copies (256): +7 %, two periods ahead: -1 %
    if (T.opt < 0) T.opt = 2 | (stages == 2 ? 4 : 0) | 128 | ((!isA && stages == 2) ? 256 : 0);
    if (T.max_copies > WT_PRE_MAX_COPIES || T.diag != 0) T.opt &= ~4;
    if (WT_SMEM_HEAD + (size_t)stages * stage_bytes > (size_t)WT_SMEM_MAX) {
        std::fprintf(stderr, "fesom2-accelerate: warp tiles of %d B do not fit two stages\n", stage_bytes);
        return false;
    }
    // consumer / issuer warps: the closest compiled variant (0: default)
    int nwc = env_int(isA ? "FCT_WT_WARPS_A" : "FCT_WT_WARPS_B", 0), npw = env_int("FCT_WT_ISSUERS", 0);
    // 24 warps in all at 80 registers; the packed layout needs a sixth of the bulk copies: two issuers
    // (with the lists ahead the issuers also pull the rows into L2: four again, measured
    //  profiles/r1_v14_ab_ring_depth.log)
    // knob WT_REGS: 0 = every warp at 80 registers (24 warps in all); 80 / 88 = the roles re-allocate
    // registers (setmaxnreg): 4 producer warps at 24 + 24 consumers at 80 (28 warps launched at 72) or
    // 20 consumers at 88 (24 warps launched at 80)
    const bool plain = !iter && (!isA || A.vlimit == 1 || A.vlimit == 0);
    int regs = plain ? env_int("FCT_WT_REGS", 0) : 0;
    if (regs != 0 && regs != 80 && regs != 88 && regs != 72) regs = 80;
    // knob WT_CONV (phase A, vlimit 1): -1 = a1 folded into the edge loop, 2 / 3 / 4 = that many converter warps
    // (round 1: 2), 0 = the consumer warps run the a1 pass in chunks
    // default since round 2: -1, the a1 pass folded into the consumers' edge loop (no converter warps, 20 % fewer
    // shared-memory wavefronts): +7 % on phase A, interleaved, on nl = 48 and nl = 70 (profiles/r2_v10_*)
    int conv = (isA && plain) ? env_int("FCT_WT_CONV", -1) : 2;
    if (conv != 0 && conv != 4 && conv != 3 && conv != -1) conv = 2;
    if (regs == 72) {
        npw = 4;
        nwc = 24;
    } else if (regs > 0) {
        npw = npw <= 0 ? ((isA && conv != 0) ? 1 : 3) : npw;   // 4 producer warps: fetcher + issuers (+ 2 converters)
        nwc = nwc <= 0 ? (regs == 80 ? 24 : 20) : nwc;
    } else {
        npw = npw <= 0 ? ((packed && stages != 2) ? 2 : 4) : npw;
        nwc = nwc <= 0 ? (isA ? 23 - std::max(conv, 0) - npw : 23 - npw) : nwc;
    }
    constexpr int NV1 = sizeof(g_wvariants) / sizeof(g_wvariants[0]);
    constexpr int NVL = sizeof(g_wvariants_vl) / sizeof(g_wvariants_vl[0]);
    constexpr int NVI = sizeof(g_wvariants_it) / sizeof(g_wvariants_it[0]);
    const bool vl = isA && A.vlimit != 1 && A.vlimit != 0;
    const WarpVariant *table = vl ? g_wvariants_vl : (iter ? g_wvariants_it : g_wvariants);
    int vi = -1, best = 1 << 30;
    // the closest compiled shape; a knob combination nobody compiled (e.g. three converter warps without the
    // 32-warp CTA) falls back to the same ring depth and register scheme with any a1 scheme
    for (int pass = 0; pass < 2 && vi < 0; ++pass)
        for (int i = 0; i < (vl ? NVL : (iter ? NVI : NV1)); ++i) {
            if (table[i].stages != stages || table[i].phase_a != isA || table[i].regs != regs) continue;
            if (pass == 0 && isA && plain && table[i].conv != conv) continue;
            const int dist = 4 * std::abs(table[i].consumers - nwc) + std::abs(table[i].issuers - npw);
            if (dist < best) {
                best = dist;
                vi = i;
            }
        }
    if (vi < 0) {
        std::fprintf(stderr, "fesom2-accelerate: no warp-item kernel compiled for %d stages with WT_REGS %d\n", stages, regs);
        return false;
    }
    const WarpVariant &v = table[vi];
    const size_t smem = WT_SMEM_HEAD + (size_t)stages * stage_bytes;
    bool first = false;
    if (!ensure_smem_attr(reinterpret_cast<const void *>(v.fn), smem, &first)) return false;
    if (first && env_int("FCT_VERBOSE", 0))
        std::fprintf(stderr, "fesom2-accelerate: warp kernel %c: %d stages of %d B, %d consumer + %d issuer%s warps%s, opt %d\n",
                     isA ? 'A' : 'B', stages, stage_bytes, v.consumers, v.issuers, (isA && v.conv > 0) ? (v.conv == 4 ? " + 4 converter" : (v.conv == 3 ? " + 3 converter" : " + 2 converter")) : "",
                     v.regs > 0 ? ", registers re-allocated" : "", T.opt);
    const int sms = device_sms();
    const long long total = (long long)T.ntiles * ntracers;
    if (total >= (1LL << 30)) {
        std::fprintf(stderr, "fesom2-accelerate: too many (tile, tracer) pairs for one launch\n");
        return false;
    }
    // device-wide tile counters: per device a ring of self-rearming pairs, one per launch in flight
    static std::map<int, int *> ctr_rings;
    static std::atomic<unsigned> ctr_next{0};
    constexpr unsigned CTR_SLOTS = 256;
    int *ctr = nullptr;
    if (env_int("FCT_WT_DYNAMIC", 1)) {
        static std::mutex ring_mutex;
        std::lock_guard<std::mutex> lock(ring_mutex);
        int dev_id = 0;
        cudaGetDevice(&dev_id);
        int *&ring = ctr_rings[dev_id];
        if (!ring) {
            // (the launch streams are non-blocking: wait for the null stream's memset once)
            if (!cuda_ok(cudaMalloc(&ring, CTR_SLOTS * 2 * sizeof(int)), "cudaMalloc(counters)") ||
                !cuda_ok(cudaMemset(ring, 0, CTR_SLOTS * 2 * sizeof(int)), "cudaMemset(counters)") ||
                !cuda_ok(cudaStreamSynchronize(0), "cudaMemset(counters)")) {
                ring = nullptr;
                return false;
            }
        }
        ctr = ring + 2 * (ctr_next.fetch_add(1) % CTR_SLOTS);
    }
    // profiling aid (knob WT_TRACE 1 = phase A, 2 = phase B): pipeline time stamps of CTA 0, read back with fct_ale_trace_read_
    T.trace = nullptr;
    if (env_int("FCT_WT_TRACE", 0) == (isA ? 1 : 2)) {
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        if (!g_trace && !cuda_ok(cudaMalloc(&g_trace, sizeof(long long) * WT_TRACE_SLOTS * WT_TRACE_ITERS), "cudaMalloc(trace)")) return false;
        cudaMemsetAsync(g_trace, 0, sizeof(long long) * WT_TRACE_SLOTS * WT_TRACE_ITERS, s);
        T.trace = g_trace;
    }
    dim3 grid((unsigned)std::min<long long>(total, sms), 1, 1);
    Arrays Aw = A;
    if (T.opt & 128) Aw.flags |= 4;   // tile-level L2 prefetch by the issuers: the consumers skip their per-item probes
    const int warps = v.regs > 0 ? wt_producer_warps(v.regs) + v.consumers : v.issuers + 1 + (isA ? std::max(v.conv, 0) : 0) + v.consumers;
    v.fn<<<grid, warps * 32, smem, s>>>(Aw, T, ntracers, stage_bytes, ctr);
    count_launch(1);
    return cuda_ok(cudaGetLastError(), "warp kernel launch");
}

typedef void (*tile_kern_t)(Arrays, TileDev);

// Variants of the tile kernels: HB edge-flux rows loaded ahead per item, MINB resident CTAs the
// register allocation is bounded for.  FCT_TILE_VARIANT_A / _B select one (tuning knob).
struct TileVariant {
    tile_kern_t fn;
    const char *name;
};
static const TileVariant g_variants_a[] = {
    {k_phaseA_tile<2, 0, 4>, "hb0-min4"}, {k_phaseA_tile<2, 8, 3>, "hb8-min3"}, {k_phaseA_tile<2, 4, 3>, "hb4-min3"},
    {k_phaseA_tile<2, 8, 2>, "hb8-min2"}, {k_phaseA_tile<2, 0, 3>, "hb0-min3"}, {k_phaseA_tile<2, 6, 4>, "hb6-min4"},
};
static const TileVariant g_variants_b[] = {
    {k_phaseB_tile<2, 4, 3>, "hb4-min3"}, {k_phaseB_tile<2, 0, 3>, "hb0-min3"}, {k_phaseB_tile<2, 8, 2>, "hb8-min2"},
    {k_phaseB_tile<2, 8, 3>, "hb8-min3"}, {k_phaseB_tile<2, 4, 2>, "hb4-min2"}, {k_phaseB_tile<2, 6, 3>, "hb6-min3"},
};

// which: 0 all owned nodes, 1 boundary list, 2 interior list
bool launch_tile(int stage, const Arrays &A, const Plan *p, int which, int ntracers, cudaStream_t s)
{
    const bool isA = stage == ST_PHASE_A;
    TileDev T = p->tiles[isA ? 0 : 1][which];
    if (T.ntiles <= 0) return true;
    constexpr int NV = 6;
    const int vi = std::min(std::max(env_int(isA ? "FCT_TILE_VARIANT_A" : "FCT_TILE_VARIANT_B", 0), 0), NV - 1);
    const TileVariant &v = isA ? g_variants_a[vi] : g_variants_b[vi];
    const size_t smem = tile_smem_bytes(isA, A.pitchL, T.TN, T.TE, T.max_rows);
    if (!ensure_smem_attr(reinterpret_cast<const void *>(v.fn), smem)) return false;
    // tables of the tile one "resident wave" ahead are prefetched into L2; occupancy per (device, kernel, smem)
    int resident = 0;
    {
        static std::map<std::tuple<int, const void *, size_t>, int> occ;
        const int sms = device_sms();
        std::lock_guard<std::mutex> lock(g_dev_mutex);
        int &r = occ[std::make_tuple(current_device(), reinterpret_cast<const void *>(v.fn), smem)];
        if (r == 0) {
            int per_sm = 0;
            if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v.fn, TILE_THREADS, smem) != cudaSuccess) per_sm = 2;
            r = std::max(per_sm, 1) * sms;
            if (env_int("FCT_VERBOSE", 0))
                std::fprintf(stderr, "fesom2-accelerate: phase %c tile kernel %s: %zu B smem, %d CTAs/SM\n",
                             isA ? 'A' : 'B', v.name, smem, per_sm);
        }
        resident = r;
    }
    T.ahead = env_int("FCT_TILE_AHEAD", resident);
    dim3 grid(T.ntiles, ntracers, 1);
    v.fn<<<grid, TILE_THREADS, smem, s>>>(A, T);
    count_launch(1);
    return cuda_ok(cudaGetLastError(), "tile kernel launch");
}

void destroy_plan(Plan *p)
{
    if (!p) return;
    for (void *d : p->owned) cudaFree(d);
    p->magic = 0;
    delete p;
}

Plan *create_plan_host(int N, int H, int E, int G, int nl, const int *nlev_n, const int *nlev_e,
                       const int *elem_nodes, const int *nie_num, const int *nie, int nie_dim,
                       const int *edges, const int *edge_tri)
{
    DerivedHost d;
    if (!build_derived(N, H, E, G, nl, nlev_e, elem_nodes, nie_num, nie, nie_dim, edges, edge_tri, d))
        return nullptr;
    Plan *p = new (std::nothrow) Plan;
    if (!p) return nullptr;
    p->N = N; p->H = H; p->E = E; p->G = G; p->nl = nl; p->nie_dim = nie_dim;
    p->pitch = (nl + 7) & ~7;   // rows start on 64-byte boundaries: no DRAM sector is shared by two rows
    p->owns_mesh = true;
    auto up = [&](const int *src, size_t n) -> const int * {
        std::vector<int> v(src, src + n);
        return upload_vec(p, v);
    };
    p->dev.nlev_n = up(nlev_n, (size_t)N + H);
    p->dev.nlev_e = up(nlev_e, (size_t)E);
    p->dev.elem_nodes = up(elem_nodes, (size_t)3 * E);
    p->dev.nie_num = up(nie_num, (size_t)N);
    p->dev.nie = up(nie, (size_t)N * nie_dim);
    p->dev.nie_dim = nie_dim;
    p->dev.edges = up(edges, (size_t)2 * G);
    p->dev.edge_tri = up(edge_tri, (size_t)2 * G);
    if (!p->dev.nlev_n || !p->dev.nlev_e || !p->dev.elem_nodes || !p->dev.nie_num || !p->dev.nie ||
        !p->dev.edges || !p->dev.edge_tri || !upload_derived(p, d)) {
        destroy_plan(p);
        return nullptr;
    }
    build_plan_tiles(p, d, nlev_n);
    build_plan_wtiles(p, d, nlev_n);
    return p;
}

// Plans of the handle-based (legacy) path are built on first use from the device copies the
// caller made with transfer_mesh_ and cached by the identity of those buffers.
typedef std::tuple<const void *, const void *, const void *, const void *, const void *, int, int, int, int, int> PlanKey;
static std::map<PlanKey, Plan *> g_plan_cache;
static std::mutex g_plan_mutex;

static bool fetch(std::vector<int> &dst, const int *d, size_t n)
{
    dst.resize(n);
    return n == 0 || cuda_ok(cudaMemcpy(dst.data(), d, n * sizeof(int), cudaMemcpyDeviceToHost), "D2H(mesh)");
}

// nie / nie_num / elem_nodes may be null (inter/post/c calls only know the edge arrays): then the
// ring lists are left empty, which is all b3h / c_h need.
static Plan *plan_for_handles(int N, int H, int E, int G, int nl, const int *d_nlev_e,
                              const int *d_elem_nodes, const int *d_nie_num, const int *d_nie,
                              int nie_dim, const int *d_edges, const int *d_edge_tri)
{
    std::lock_guard<std::mutex> lock(g_plan_mutex);
    PlanKey key(d_nlev_e, d_elem_nodes, d_nie, d_edges, d_edge_tri, N, H, E, G, nl);
    auto it = g_plan_cache.find(key);
    if (it != g_plan_cache.end()) return it->second;
    if (!d_nie) {
        // an edge-only request is served by any cached plan of the same edge tables
        for (auto &kv : g_plan_cache) {
            const PlanKey &k = kv.first;
            if (std::get<0>(k) == d_nlev_e && std::get<3>(k) == d_edges && std::get<4>(k) == d_edge_tri &&
                std::get<5>(k) == N && std::get<6>(k) >= H && std::get<8>(k) == G && std::get<9>(k) == nl)
                return kv.second;
        }
    }
    std::vector<int> nlev_e, en, num, nie, edges, etri;
    if (!fetch(edges, d_edges, (size_t)2 * G) || !fetch(etri, d_edge_tri, (size_t)2 * G)) return nullptr;
    if (E <= 0) {
        // element count unknown to this call: size the depth array from the edge table
        int emax = 0;
        for (int v : etri) emax = std::max(emax, v);
        E = emax;
    }
    if (!fetch(nlev_e, d_nlev_e, (size_t)E)) return nullptr;
    DerivedHost d;
    const bool rings = d_elem_nodes && d_nie_num && d_nie;
    if (rings) {
        if (!fetch(en, d_elem_nodes, (size_t)3 * E) || !fetch(num, d_nie_num, (size_t)N) ||
            !fetch(nie, d_nie, (size_t)N * nie_dim))
            return nullptr;
        if (!build_derived(N, H, E, G, nl, nlev_e.data(), en.data(), num.data(), nie.data(), nie_dim,
                           edges.data(), etri.data(), d))
            return nullptr;
    } else {
        if (!build_derived(N, H, E, G, nl, nlev_e.data(), nullptr, nullptr, nullptr, 0, edges.data(),
                           etri.data(), d))
            return nullptr;
    }
    Plan *p = new (std::nothrow) Plan;
    if (!p) return nullptr;
    p->N = N; p->H = H; p->E = E; p->G = G; p->nl = nl; p->nie_dim = nie_dim;
    p->pitch = nl;
    p->owns_mesh = false;
    if (!upload_derived(p, d)) {
        destroy_plan(p);
        return nullptr;
    }
    g_plan_cache[key] = p;
    return p;
}

}   // namespace fct

using namespace fct;

// =============================================================================================
// Part 1: the reference's ABI
// =============================================================================================
extern "C" {

void set_mpi_rank_(int *rank, int *total_ranks)
{
    const int tot = (total_ranks && *total_ranks > 0) ? *total_ranks : 1;
    const int rank_on_node = (rank ? *rank : 0) % tot;
    int count = 0;
    if (!cuda_ok(cudaGetDeviceCount(&count), "cudaGetDeviceCount") || count < 1) {
        std::fprintf(stderr, "fesom2-accelerate: no CUDA device for rank %d\n", rank ? *rank : 0);
        return;
    }
    cuda_ok(cudaSetDevice(rank_on_node % count), "cudaSetDevice");
}

void transfer_mesh_(void **ret, int *host_ptr, int *size, int *istat)
{
    gpuMemory *h = new_handle((void *)host_ptr, (size_t)(*size) * sizeof(int), false);
    if (h && h2d(h, true, 0, false)) {
        *ret = h;
        *istat = 0;
    } else {
        if (h) {
            cudaFree(h->device_pointer);
            delete h;
        }
        *ret = nullptr;
        *istat = 1;
    }
}

void alloc_var_(void **ret, real_type *host_ptr, int *size, bool *create_event, int *istat)
{
    gpuMemory *h = new_handle((void *)host_ptr, (size_t)(*size) * sizeof(real_type), create_event && *create_event);
    *istat = h ? 0 : 1;
    *ret = h;
}

void reserve_var_(void **ret, int *size, bool *create_event, int *istat)
{
    gpuMemory *h = new_handle(nullptr, (size_t)(*size) * sizeof(real_type), create_event && *create_event);
    *istat = h ? 0 : 1;
    *ret = h;
}

void allocate_pinned_doubles_(void **hostptr, int *size, int *istat)
{
    *istat = 0;
    if (!cuda_ok(cudaMallocHost(hostptr, sizeof(double) * (size_t)(*size)), "cudaMallocHost")) {
        std::fprintf(stderr, "fesom2-accelerate: page-locked allocation failed, using malloc\n");
        *hostptr = std::malloc(sizeof(double) * (size_t)(*size));
        *istat = 1;
    }
}

void transfer_var_(void **mem, real_type *host_ptr)
{
    gpuMemory *h = H(mem);
    if (!h) return;
    h->host_pointer = (void *)host_ptr;
    h2d(h, true, 0, false);
}

void transfer_var_async_(void **mem, real_type *host_ptr, void **stream, bool *record_event)
{
    gpuMemory *h = H(mem);
    if (!h) return;
    h->host_pointer = (void *)host_ptr;
    h2d(h, false, S(stream), record_event && *record_event);
}

void make_stream_(void **stream, int *istat)
{
    *istat = 0;
    cudaStream_t *s = new cudaStream_t;
    // a blocking stream, as the reference creates it (src/fesom2-accelerate.cu:174): it serialises with
    // the legacy default stream, so the synchronous calls of the ABI (transfer_var_, transfer_mesh_,
    // transfer_var_back_: plain cudaMemcpy on the null stream) wait for the kernels and copies queued
    // on it and vice versa -- the implicit ordering an unmodified caller relies on
    if (!cuda_ok(cudaStreamCreate(s), "cudaStreamCreate")) {
        std::fprintf(stderr, "fesom2-accelerate: stream creation failed, returning the default stream\n");
        *s = (cudaStream_t)0;
        *istat = 1;
    }
    *stream = s;
}

void await_stream_(void **s, int *istat)
{
    *istat = 0;
    if (!cuda_ok(cudaStreamSynchronize(S(s)), "cudaStreamSynchronize")) *istat = 1;
}

// ---- the three orchestration calls ----------------------------------------------------------

static Arrays dense_arrays(int nl)
{
    Arrays A;
    std::memset(&A, 0, sizeof(A));
    A.nl = nl;
    A.pitchL = nl - 1;
    A.pitchV = nl;
    A.pitchH = nl - 1;
    A.pitchU = nl - 1;
    return A;
}

void fct_ale_pre_comm_acc_(int *alg_state, void **s, void **fct_ttf_max, void **fct_ttf_min,
                           void **fct_plus, void **fct_minus, void **ttf, void **fct_LO,
                           void **fct_adf_v, void **fct_adf_h, void **UV_rhs, void **area_inv,
                           int *myDim_nod2D, int *eDim_nod2D, int *myDim_elem2D, int *myDim_edge2D,
                           int *nl, void **nlevels_nod2D, void **nlevels_elem2D, void **elem2D_nodes,
                           void **nod_in_elem2D_num, void **nod_in_elem2D, int *nod_in_elem2D_dim,
                           void **nod2D_edges, void **elem2D_edges, int *vlimit, real_type *flux_eps,
                           real_type *bignumber, real_type *dt)
{
    *alg_state = 0;
    cudaStream_t st = S(s);
    const int N = *myDim_nod2D, Hn = *eDim_nod2D, E = *myDim_elem2D, G = *myDim_edge2D;
    const int vl = vlimit ? *vlimit : 1;
    if (vl < 1 || vl > 3) {
        std::fprintf(stderr, "fesom2-accelerate: vlimit = %d (1, 2 or 3)\n", vl);
        return;
    }
    Plan *p = plan_for_handles(N, Hn, E, G, *nl, dev<int>(nlevels_elem2D), dev<int>(elem2D_nodes),
                               dev<int>(nod_in_elem2D_num), dev<int>(nod_in_elem2D), *nod_in_elem2D_dim,
                               dev<int>(nod2D_edges), dev<int>(elem2D_edges));
    if (!p) return;
    h2d(H(fct_LO), false, st, false);

    Arrays A = dense_arrays(*nl);
    A.ttf = dev<double>(ttf);
    A.lo = dev<double>(fct_LO);
    A.adf_v = dev<double>(fct_adf_v);
    A.adf_v_out = A.adf_v;
    A.adf_h_in = dev<double>(fct_adf_h);
    A.adf_h_out = dev<double>(fct_adf_h);
    A.ttf_max = dev<double>(fct_ttf_max);
    A.ttf_min = dev<double>(fct_ttf_min);
    A.plus = dev<double>(fct_plus);
    A.minus = dev<double>(fct_minus);
    A.uv_rhs = dev<double2>(UV_rhs);
    A.area_inv = dev<double>(area_inv);
    A.dt = *dt;
    A.eps = *flux_eps;
    A.big = *bignumber;
    MeshDev M = p->dev;
    M.nlev_n = dev<int>(nlevels_nod2D);
    M.nlev_e = dev<int>(nlevels_elem2D);
    M.elem_nodes = dev<int>(elem2D_nodes);
    M.nie = dev<int>(nod_in_elem2D);
    M.nie_num = dev<int>(nod_in_elem2D_num);
    M.nie_dim = *nod_in_elem2D_dim;
    M.edges = dev<int>(nod2D_edges);
    M.edge_tri = dev<int>(elem2D_edges);

    wait_upload(H(ttf), st);
    bool ok = true;
    if (g_fused.load() && vl == 1) {   // the fused phase implements vlimit 1; 2 and 3 run stage by stage
        // a1 on the halo rows keeps fct_ttf_max/min identical to the staged run there
        ok = ok && launch_stage(ST_A1, 1, A, M, nullptr, N, Hn, 1, st);
        wait_upload(H(fct_adf_v), st);
        wait_upload(H(fct_adf_h), st);
        ok = ok && launch_stage(ST_PHASE_A, 1, A, M, nullptr, 0, N, 1, st);
    } else {
        ok = ok && launch_stage(ST_A1, 1, A, M, nullptr, 0, N + Hn, 1, st);
        if (ok) *alg_state = 1;
        ok = ok && launch_stage(ST_A2, 1, A, M, nullptr, 0, E, 1, st);
        if (ok) *alg_state = 2;
        ok = ok && launch_stage(vl == 1 ? ST_A3 : (vl == 2 ? ST_A3_VLIMIT2 : ST_A3_VLIMIT3), 1, A, M, nullptr, 0, N, 1, st);
        if (ok) *alg_state = 3;
        wait_upload(H(fct_adf_v), st);
        ok = ok && launch_stage(ST_B1V, 1, A, M, nullptr, 0, N, 1, st);
        if (ok) *alg_state = 4;
        wait_upload(H(fct_adf_h), st);
        ok = ok && launch_stage(ST_B1H, 1, A, M, nullptr, 0, N, 1, st);
        if (ok) *alg_state = 5;
        ok = ok && launch_stage(ST_B2, 1, A, M, nullptr, 0, N, 1, st);
    }
    if (!ok) return;
    *alg_state = 6;
    if (!d2h(H(fct_plus), false, st) || !d2h(H(fct_minus), false, st)) *alg_state = 0;
}

void fct_ale_inter_comm_acc_(int *alg_state, void **s, void **fct_plus, void **fct_minus,
                             void **fct_adf_v, int *myDim_nod2D, int *nl, void **nlevels_nod2D)
{
    cudaStream_t st = S(s);
    Arrays A = dense_arrays(*nl);
    A.adf_v = dev<double>(fct_adf_v);
    A.adf_v_out = A.adf_v;
    A.plus = dev<double>(fct_plus);
    A.minus = dev<double>(fct_minus);
    MeshDev M;
    std::memset(&M, 0, sizeof(M));
    M.nlev_n = dev<int>(nlevels_nod2D);
    if (!launch_stage(ST_B3V, 1, A, M, nullptr, 0, *myDim_nod2D, 1, st)) return;
    *alg_state = 7;
    if (!d2h(H(fct_adf_v), false, st)) *alg_state = 0;
}

void fct_ale_post_comm_acc_(int *alg_state, void **s, void **fct_plus, void **fct_minus,
                            void **fct_adf_h, int *myDim_edge2D, int *nl, void **nlevels_elem2D,
                            int *nod_in_elem2D_dim, void **nod2D_edges, void **elem2D_edges)
{
    (void)nod_in_elem2D_dim;
    cudaStream_t st = S(s);
    h2d(H(fct_plus), false, st, false);
    h2d(H(fct_minus), false, st, false);
    Arrays A = dense_arrays(*nl);
    A.adf_h_in = dev<double>(fct_adf_h);
    A.adf_h_out = dev<double>(fct_adf_h);
    A.plus = dev<double>(fct_plus);
    A.minus = dev<double>(fct_minus);
    MeshDev M;
    std::memset(&M, 0, sizeof(M));
    M.nlev_e = dev<int>(nlevels_elem2D);
    M.edges = dev<int>(nod2D_edges);
    M.edge_tri = dev<int>(elem2D_edges);
    if (!launch_stage(ST_B3H, 1, A, M, nullptr, 0, *myDim_edge2D, 1, st)) return;
    *alg_state = 8;
    if (!d2h(H(fct_adf_h), false, st)) *alg_state = 0;
}

// ---- legacy single-kernel entry points --------------------------------------------------------
static void legacy_a1_a2(bool do_a1, bool do_a2, int maxLevels, int nNodes, int nElements,
                         gpuMemory *nlev_n, gpuMemory *nlev_e, gpuMemory *en, gpuMemory *tmax,
                         gpuMemory *tmin, gpuMemory *lo, gpuMemory *ttf, gpuMemory *uv, bool sync,
                         cudaStream_t st)
{
    Arrays A = dense_arrays(maxLevels + 1);
    A.big = 1.0e3;   // kernels/fct_ale_a2.cu:21 hard-codes it; so does this legacy entry point
    MeshDev M;
    std::memset(&M, 0, sizeof(M));
    if (do_a1) {
        if (!h2d(lo, sync, st, false) || !h2d(ttf, sync, st, false)) return;
        A.lo = (const double *)lo->device_pointer;
        A.ttf = (const double *)ttf->device_pointer;
        M.nlev_n = (const int *)nlev_n->device_pointer;
    } else {
        if (!h2d(tmax, sync, st, false) || !h2d(tmin, sync, st, false)) return;
    }
    A.ttf_max = (double *)tmax->device_pointer;
    A.ttf_min = (double *)tmin->device_pointer;
    if (do_a1 && !launch_stage(ST_A1, 1, A, M, nullptr, 0, nNodes, 1, st)) return;
    if (do_a2) {
        M.nlev_e = (const int *)nlev_e->device_pointer;
        M.elem_nodes = (const int *)en->device_pointer;
        A.uv_rhs = (double2 *)uv->device_pointer;
        if (!launch_stage(ST_A2, 1, A, M, nullptr, 0, nElements, 1, st)) return;
        d2h(uv, sync, st);
    } else {
        if (!d2h(tmax, sync, st)) return;
        d2h(tmin, sync, st);
    }
}

void fct_ale_a1_accelerated(const int maxLevels, const int nNodes, struct gpuMemory *nLevels_nod2D,
                            struct gpuMemory *fct_ttf_max, struct gpuMemory *fct_ttf_min,
                            struct gpuMemory *fct_low_order, struct gpuMemory *ttf, bool synchronous,
                            void *stream)
{
    legacy_a1_a2(true, false, maxLevels, nNodes, 0, nLevels_nod2D, nullptr, nullptr, fct_ttf_max,
                 fct_ttf_min, fct_low_order, ttf, nullptr, synchronous, (cudaStream_t)stream);
}

void fct_ale_a2_accelerated(const int maxLevels, const int nElements, struct gpuMemory *nLevels_elem,
                            struct gpuMemory *elementNodes, struct gpuMemory *UV_rhs,
                            struct gpuMemory *fct_ttf_max, struct gpuMemory *fct_ttf_min,
                            bool synchronous, void *stream)
{
    legacy_a1_a2(false, true, maxLevels, 0, nElements, nullptr, nLevels_elem, elementNodes, fct_ttf_max,
                 fct_ttf_min, nullptr, nullptr, UV_rhs, synchronous, (cudaStream_t)stream);
}

void fct_ale_a1_a2_accelerated(const int maxLevels, const int nNodes, const int nElements,
                               struct gpuMemory *nLevels_nod2D, struct gpuMemory *nLevels_elem,
                               struct gpuMemory *elementNodes, struct gpuMemory *fct_ttf_max,
                               struct gpuMemory *fct_ttf_min, struct gpuMemory *fct_low_order,
                               struct gpuMemory *ttf, struct gpuMemory *UV_rhs, bool synchronous,
                               void *stream)
{
    legacy_a1_a2(true, true, maxLevels, nNodes, nElements, nLevels_nod2D, nLevels_elem, elementNodes,
                 fct_ttf_max, fct_ttf_min, fct_low_order, ttf, UV_rhs, synchronous, (cudaStream_t)stream);
}

// ---- host-array entry points with the reference oracle's names: GPU-computed -------------------
}   // extern "C"
struct Scratch {
    std::vector<void *> bufs;
    bool ok = true;
    template <class T>
    T *up(const T *host, size_t n)
    {
        T *d = nullptr;
        if (!ok) return nullptr;
        if (!cuda_ok(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(T)), "cudaMalloc(scratch)")) {
            ok = false;
            return nullptr;
        }
        bufs.push_back(d);
        if (host && n && !cuda_ok(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice), "H2D(scratch)")) ok = false;
        return d;
    }
    template <class T>
    void down(T *host, const T *d, size_t n)
    {
        if (ok && n && !cuda_ok(cudaMemcpy(host, d, n * sizeof(T), cudaMemcpyDeviceToHost), "D2H(scratch)")) ok = false;
    }
    ~Scratch()
    {
        for (void *b : bufs) cudaFree(b);
    }
};
extern "C" {

void fct_ale_a1_reference_(int *nNodes, int *nLevels_nod2D, int *nl, real_type *fct_ttf_max,
                           real_type *fct_ttf_min, real_type *fct_low_order, real_type *ttf)
{
    const size_t n = (size_t)*nNodes, L = (size_t)*nl - 1;
    Scratch sc;
    Arrays A = dense_arrays(*nl);
    MeshDev M;
    std::memset(&M, 0, sizeof(M));
    M.nlev_n = sc.up(nLevels_nod2D, n);
    A.lo = sc.up(fct_low_order, n * L);
    A.ttf = sc.up(ttf, n * L);
    A.ttf_max = sc.up(fct_ttf_max, n * L);
    A.ttf_min = sc.up(fct_ttf_min, n * L);
    if (!sc.ok || !launch_stage(ST_A1, 1, A, M, nullptr, 0, (int)n, 1, 0)) return;
    sc.down(fct_ttf_max, A.ttf_max, n * L);
    sc.down(fct_ttf_min, A.ttf_min, n * L);
}

// number of node rows the a2 gather can touch = highest node id of the element table
static int max_id(const int *v, size_t n)
{
    int m = 0;
    for (size_t i = 0; i < n; ++i) m = std::max(m, v[i]);
    return m;
}

void fct_ale_a2_reference_(int *nElements, int *nLevels_elem2D, int *nl, real_type *UV_rhs,
                           int *elem2D_nodes, real_type *fct_ttf_max, real_type *fct_ttf_min,
                           real_type *bignumber)
{
    const size_t E = (size_t)*nElements, L = (size_t)*nl - 1;
    const size_t n = (size_t)max_id(elem2D_nodes, 3 * E);
    Scratch sc;
    Arrays A = dense_arrays(*nl);
    A.big = *bignumber;
    MeshDev M;
    std::memset(&M, 0, sizeof(M));
    M.nlev_e = sc.up(nLevels_elem2D, E);
    M.elem_nodes = sc.up(elem2D_nodes, 3 * E);
    A.ttf_max = sc.up(fct_ttf_max, n * L);
    A.ttf_min = sc.up(fct_ttf_min, n * L);
    A.uv_rhs = (double2 *)sc.up(UV_rhs, E * L * 2);
    if (!sc.ok || !launch_stage(ST_A2, 1, A, M, nullptr, 0, (int)E, 1, 0)) return;
    sc.down(UV_rhs, (double *)A.uv_rhs, E * L * 2);
}

void fct_ale_a3_reference_(int *nNodes2D, int *nLevels_nod2D, int *nl, real_type *fct_ttf_max,
                           real_type *fct_ttf_min, real_type *fct_LO, real_type *UV_rhs,
                           real_type *fct_plus, real_type *fct_minus, real_type *fct_adf_v,
                           int *nod_in_elem2D, int *nod_in_elem2D_num, int *nod_in_elem2D_dim)
{
    const size_t n = (size_t)*nNodes2D, L = (size_t)*nl - 1, dim = (size_t)*nod_in_elem2D_dim;
    size_t E = 0;
    for (size_t i = 0; i < n; ++i)
        for (int k = 0; k < nod_in_elem2D_num[i]; ++k) E = std::max(E, (size_t)nod_in_elem2D[i * dim + k]);
    Scratch sc;
    Arrays A = dense_arrays(*nl);
    MeshDev M;
    std::memset(&M, 0, sizeof(M));
    M.nlev_n = sc.up(nLevels_nod2D, n);
    M.nie = sc.up(nod_in_elem2D, n * dim);
    M.nie_num = sc.up(nod_in_elem2D_num, n);
    M.nie_dim = (int)dim;
    A.lo = sc.up(fct_LO, n * L);
    A.uv_rhs = (double2 *)sc.up(UV_rhs, E * L * 2);
    A.ttf_max = sc.up(fct_ttf_max, n * L);
    A.ttf_min = sc.up(fct_ttf_min, n * L);
    A.plus = sc.up(fct_plus, n * L);
    A.minus = sc.up(fct_minus, n * L);
    A.adf_v = sc.up(fct_adf_v, n * (L + 1));
    if (!sc.ok || !launch_stage(ST_A3, 1, A, M, nullptr, 0, (int)n, 1, 0) ||
        !launch_stage(ST_B1V, 1, A, M, nullptr, 0, (int)n, 1, 0))
        return;
    sc.down(fct_ttf_max, A.ttf_max, n * L);
    sc.down(fct_ttf_min, A.ttf_min, n * L);
    sc.down(fct_plus, A.plus, n * L);
    sc.down(fct_minus, A.minus, n * L);
}

void fct_ale_a4_reference_(int *nNodes2D, int *nLevels_nod2D, int *nLevels_elem2D, int *nl,
                           int *nEdges2D, real_type *fct_plus, real_type *fct_minus,
                           real_type *fct_adf_h, real_type *area_inv, real_type *fct_ttf_max,
                           real_type *fct_ttf_min, int *edges, int *edge_tri, real_type *flux_eps,
                           real_type *dt)
{
    const size_t n = (size_t)*nNodes2D, L = (size_t)*nl - 1, G = (size_t)*nEdges2D;
    // b1 horizontal also adds into the halo rows the edges reach (reference.cpp:417-423)
    const size_t nt = std::max(n, (size_t)max_id(edges, 2 * G));
    const size_t E = (size_t)max_id(edge_tri, 2 * G);
    DerivedHost d;
    if (!build_derived((int)nt, 0, (int)E, (int)G, *nl, nLevels_elem2D, nullptr, nullptr, nullptr, 0, edges, edge_tri, d))
        return;
    Scratch sc;
    Arrays A = dense_arrays(*nl);
    A.dt = *dt;
    A.eps = *flux_eps;
    MeshDev M;
    std::memset(&M, 0, sizeof(M));
    // every row an edge reaches takes part in b1h; rows past nNodes2D have no depth entry in the
    // caller's array, so they inherit the deepest edge that touches them
    std::vector<int> nlev(nt, 0);
    std::memcpy(nlev.data(), nLevels_nod2D, n * sizeof(int));
    for (size_t i = n; i < nt; ++i)
        for (int k = d.edg_off[i]; k < d.edg_off[i + 1]; ++k) nlev[i] = std::max(nlev[i], FCT_META_DEPTH(d.edg[k].z) + 1);
    M.nlev_n = sc.up(nlev.data(), nt);
    M.edg_off = sc.up(d.edg_off.data(), d.edg_off.size());
    M.edg = sc.up(d.edg.data(), d.edg.size());
    A.adf_h_in = sc.up(fct_adf_h, G * L);
    A.area_inv = sc.up(area_inv, n * (L + 1));
    A.ttf_max = sc.up(fct_ttf_max, n * L);
    A.ttf_min = sc.up(fct_ttf_min, n * L);
    A.plus = sc.up(fct_plus, nt * L);
    A.minus = sc.up(fct_minus, nt * L);
    if (!sc.ok || !launch_stage(ST_B1H, 1, A, M, nullptr, 0, (int)nt, 1, 0) ||
        !launch_stage(ST_B2, 1, A, M, nullptr, 0, (int)n, 1, 0))
        return;
    sc.down(fct_plus, A.plus, nt * L);
    sc.down(fct_minus, A.minus, nt * L);
}

void fct_ale_pre_comm_(int *alg_state, real_type *fct_ttf_max, real_type *fct_ttf_min,
                       real_type *fct_plus, real_type *fct_minus, real_type *ttf, real_type *fct_LO,
                       real_type *fct_adf_v, real_type *fct_adf_h, real_type *UV_rhs,
                       real_type *area_inv, int *myDim_nod2D, int *eDim_nod2D, int *myDim_elem2D,
                       int *myDim_edge2D, int *nl, int *nlevels_nod2D, int *nlevels_elem2D,
                       int *elem2D_nodes, int *nod_in_elem2D_num, int *nod_in_elem2D,
                       int *nod_in_elem2D_dim, int *nod2D_edges, int *elem2D_edges, int *vlimit,
                       real_type *flux_eps, real_type *bignumber, real_type *dt)
{
    // composition of src/reference.cpp:289-304, each part GPU-computed
    *alg_state = 0;
    int nNodes = *myDim_nod2D + *eDim_nod2D;
    fct_ale_a1_reference_(&nNodes, nlevels_nod2D, nl, fct_ttf_max, fct_ttf_min, fct_LO, ttf);
    *alg_state = 1;
    fct_ale_a2_reference_(myDim_elem2D, nlevels_elem2D, nl, UV_rhs, elem2D_nodes, fct_ttf_max, fct_ttf_min, bignumber);
    *alg_state = 2;
    if (*vlimit == 1) {
        fct_ale_a3_reference_(myDim_nod2D, nlevels_nod2D, nl, fct_ttf_max, fct_ttf_min, fct_LO, UV_rhs,
                              fct_plus, fct_minus, fct_adf_v, nod_in_elem2D, nod_in_elem2D_num, nod_in_elem2D_dim);
        *alg_state = 4;
        fct_ale_a4_reference_(myDim_nod2D, nlevels_nod2D, nlevels_elem2D, nl, myDim_edge2D, fct_plus, fct_minus,
                              fct_adf_h, area_inv, fct_ttf_max, fct_ttf_min, nod2D_edges, elem2D_edges, flux_eps, dt);
        *alg_state = 5;
    }
}

// =============================================================================================
// Part 2: new ABI
// =============================================================================================

void fct_ale_c_acc_(int *alg_state, void **s, void **del_ttf_advvert, void **del_ttf_advhoriz,
                    void **ttf, void **fct_LO, void **hnode, void **hnode_new, void **fct_adf_v,
                    void **fct_adf_h, void **area, int *myDim_nod2D, int *myDim_edge2D, int *nl,
                    void **nlevels_nod2D, void **nlevels_elem2D, void **nod2D_edges,
                    void **elem2D_edges, real_type *dt)
{
    cudaStream_t st = S(s);
    const int N = *myDim_nod2D, G = *myDim_edge2D;
    Plan *p = nullptr;
    {
        // halo count is not an argument of this call: derive the node-row count from the handle size
        const gpuMemory *dh = H(del_ttf_advhoriz);
        const int rows = dh ? (int)(dh->size / sizeof(double) / (size_t)(*nl - 1)) : N;
        p = plan_for_handles(N, std::max(rows - N, 0), 0, G, *nl, dev<int>(nlevels_elem2D), nullptr, nullptr,
                             nullptr, 0, dev<int>(nod2D_edges), dev<int>(elem2D_edges));
    }
    if (!p) return;
    for (void **v : {area, hnode, hnode_new, del_ttf_advvert, del_ttf_advhoriz}) {
        gpuMemory *h = H(v);
        if (h && h->host_pointer) h2d(h, false, st, false);
    }
    Arrays A = dense_arrays(*nl);
    A.ttf = dev<double>(ttf);
    A.lo = dev<double>(fct_LO);
    A.adf_v = dev<double>(fct_adf_v);
    A.adf_v_out = A.adf_v;
    A.adf_h_in = dev<double>(fct_adf_h);
    A.adf_h_out = dev<double>(fct_adf_h);
    A.del_v = dev<double>(del_ttf_advvert);
    A.del_h = dev<double>(del_ttf_advhoriz);
    A.area = dev<double>(area);
    A.hnode = dev<double>(hnode);
    A.hnode_new = dev<double>(hnode_new);
    A.dt = *dt;
    MeshDev M = p->dev;
    M.nlev_n = dev<int>(nlevels_nod2D);
    M.nlev_e = dev<int>(nlevels_elem2D);
    M.edges = dev<int>(nod2D_edges);
    M.edge_tri = dev<int>(elem2D_edges);
    if (!launch_stage(ST_CV, 1, A, M, nullptr, 0, N, 1, st)) return;
    *alg_state = 9;
    if (!launch_stage(ST_CH, 1, A, M, nullptr, 0, N, 1, st)) return;
    *alg_state = 10;
    if (!d2h(H(del_ttf_advvert), false, st) || !d2h(H(del_ttf_advhoriz), false, st)) *alg_state = 0;
}

void transfer_var_back_(void **mem, real_type *host_ptr)
{
    gpuMemory *h = H(mem);
    if (!h) return;
    h->host_pointer = (void *)host_ptr;
    d2h(h, true, 0);
}

void transfer_var_back_async_(void **mem, real_type *host_ptr, void **stream)
{
    gpuMemory *h = H(mem);
    if (!h) return;
    h->host_pointer = (void *)host_ptr;
    d2h(h, false, S(stream));
}

void free_var_(void **mem, int *istat)
{
    gpuMemory *h = H(mem);
    *istat = 1;
    if (!h || h->magic != HANDLE_MAGIC) return;
    {
        // drop cached plans that borrow this buffer
        std::lock_guard<std::mutex> lock(g_plan_mutex);
        for (auto it = g_plan_cache.begin(); it != g_plan_cache.end();) {
            const void *d = h->device_pointer;
            const PlanKey &k = it->first;
            if (std::get<0>(k) == d || std::get<1>(k) == d || std::get<2>(k) == d || std::get<3>(k) == d ||
                std::get<4>(k) == d) {
                destroy_plan(it->second);
                it = g_plan_cache.erase(it);
            } else {
                ++it;
            }
        }
    }
    bool ok = cuda_ok(cudaFree(h->device_pointer), "cudaFree");
    if (h->has_event) ok = cuda_ok(cudaEventDestroy((cudaEvent_t)h->event), "cudaEventDestroy") && ok;
    h->magic = 0;
    delete h;
    *mem = nullptr;
    *istat = ok ? 0 : 1;
}

void free_pinned_doubles_(void **hostptr, int *istat)
{
    *istat = 0;
    if (!hostptr || !*hostptr) return;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, *hostptr) == cudaSuccess && at.type == cudaMemoryTypeHost) {
        if (!cuda_ok(cudaFreeHost(*hostptr), "cudaFreeHost")) *istat = 1;
    } else {
        cudaGetLastError();
        std::free(*hostptr);   // the malloc fallback of allocate_pinned_doubles_
    }
    *hostptr = nullptr;
}

void free_stream_(void **stream, int *istat)
{
    *istat = 0;
    if (!stream || !*stream) return;
    cudaStream_t *s = static_cast<cudaStream_t *>(*stream);
    if (*s && !cuda_ok(cudaStreamDestroy(*s), "cudaStreamDestroy")) *istat = 1;
    delete s;
    *stream = nullptr;
}

void fct_ale_set_fused_(int *fused) { g_fused.store(fused && *fused ? 1 : 0); }

void fct_ale_tune_(const char *name, int *value)
{
    if (name && value) set_tune(name, *value);
}

void fct_ale_launch_count_(long long *count) { *count = g_launches.load(); }

void fct_ale_trace_read_(long long *stamps, int *capacity, int *slots, int *istat)
{
    *istat = 1;
    *slots = WT_TRACE_SLOTS;
    if (!g_trace || !stamps) return;
    const size_t n = std::min<size_t>((size_t)std::max(*capacity, 0), (size_t)WT_TRACE_SLOTS * WT_TRACE_ITERS);
    if (!cuda_ok(cudaDeviceSynchronize(), "trace sync") ||
        !cuda_ok(cudaMemcpy(stamps, g_trace, n * sizeof(long long), cudaMemcpyDeviceToHost), "trace read"))
        return;
    *istat = 0;
}

void fct_ale_device_info_(char *name64, int *cc_major, int *cc_minor, int *sm_count, int *istat)
{
    *istat = 1;
    int devid = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&devid) != cudaSuccess || cudaGetDeviceProperties(&prop, devid) != cudaSuccess) {
        cudaGetLastError();
        if (name64) name64[0] = 0;
        return;
    }
    if (name64) {
        std::strncpy(name64, prop.name, 63);
        name64[63] = 0;
    }
    *cc_major = prop.major;
    *cc_minor = prop.minor;
    *sm_count = prop.multiProcessorCount;
    *istat = 0;
}

void fct_ale_event_create_(void **event, int *istat)
{
    cudaEvent_t e = nullptr;
    *istat = cuda_ok(cudaEventCreate(&e), "cudaEventCreate") ? 0 : 1;
    *event = (void *)e;
}

void fct_ale_event_record_(void **event, void **stream, int *istat)
{
    *istat = cuda_ok(cudaEventRecord((cudaEvent_t)*event, S(stream)), "cudaEventRecord") ? 0 : 1;
}

void fct_ale_stream_wait_event_(void **stream, void **event, int *istat)
{
    *istat = cuda_ok(cudaStreamWaitEvent(S(stream), (cudaEvent_t)*event, 0), "cudaStreamWaitEvent") ? 0 : 1;
}

void fct_ale_event_elapsed_ms_(void **start, void **stop, real_type *ms, int *istat)
{
    float t = 0.f;
    *istat = 1;
    if (!cuda_ok(cudaEventSynchronize((cudaEvent_t)*stop), "cudaEventSynchronize")) return;
    if (!cuda_ok(cudaEventElapsedTime(&t, (cudaEvent_t)*start, (cudaEvent_t)*stop), "cudaEventElapsedTime")) return;
    *ms = (double)t;
    *istat = 0;
}

void fct_ale_event_destroy_(void **event, int *istat)
{
    *istat = cuda_ok(cudaEventDestroy((cudaEvent_t)*event), "cudaEventDestroy") ? 0 : 1;
    *event = nullptr;
}

void fct_ale_mem_info_(long long *free_bytes, long long *total_bytes, int *istat)
{
    size_t f = 0, t = 0;
    *istat = cuda_ok(cudaMemGetInfo(&f, &t), "cudaMemGetInfo") ? 0 : 1;
    *free_bytes = (long long)f;
    *total_bytes = (long long)t;
}

}   // extern "C"
