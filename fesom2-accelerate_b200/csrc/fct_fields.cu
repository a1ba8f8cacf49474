// Device-resident path: plan / fields / step (Part 2 of include/fesom2-accelerate.h).
// Two storage layouts: padded rows (64-byte pitch: every column starts on a DRAM sector, each thread
// moves one double2 per array; all kernels) and the packed level storage of the fast path (active
// levels only, columns back to back; the warp-item kernels).  A step is ten stage launches (mode 0)
// or two fused launches (modes 1-3), optionally split into boundary / interior node sets around the
// NVLink halo exchange; fct_ale_step_general_ adds the vlimit 2 / 3 and iterative branches of the
// subroutine.  Host copies go through one contiguous PCIe copy + a repack kernel.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>

#include "../../include/fesom2-accelerate.h"
#include "fct_internal.h"

namespace fct {

enum RowKind { ROW_NODE, ROW_EDGE, ROW_ELEM };
struct FieldMeta {
    RowKind kind;
    int width_minus;    // dense host row width = nl - width_minus   (UV_rhs: 2*(nl-1))
    bool per_tracer;
};
static FieldMeta meta_of(int id)
{
    switch (id) {
    case FCT_ADF_V:
    case FCT_ADF_V2:
    case FCT_ADF_V_OUT: return {ROW_NODE, 0, true};
    case FCT_AREA:
    case FCT_AREA_INV: return {ROW_NODE, 0, false};
    case FCT_HNODE:
    case FCT_HNODE_NEW: return {ROW_NODE, 1, false};
    case FCT_ADF_H:
    case FCT_ADF_H2:
    case FCT_ADF_H_OUT: return {ROW_EDGE, 1, true};
    case FCT_UV_RHS: return {ROW_ELEM, 1, true};
    default: return {ROW_NODE, 1, true};
    }
}

static inline Plan *P_(void **p)
{
    Plan *q = p ? static_cast<Plan *>(*p) : nullptr;
    return (q && q->magic == PLAN_MAGIC) ? q : nullptr;
}
static inline Fields *F_(void **p)
{
    Fields *q = p ? static_cast<Fields *>(*p) : nullptr;
    return (q && q->magic == FIELDS_MAGIC) ? q : nullptr;
}
static inline cudaStream_t S_(void **s)
{
    return (s && *s) ? *static_cast<cudaStream_t *>(*s) : (cudaStream_t)0;
}

static size_t field_rows(const Fields *f, RowKind k)
{
    return k == ROW_NODE ? f->rows : (k == ROW_EDGE ? (size_t)f->plan->G : (size_t)f->plan->E);
}

static Arrays arrays_of(const Fields *f, int mode, double dt, double eps, double big)
{
    Arrays A;
    std::memset(&A, 0, sizeof(A));
    A.ttf = f->buf[FCT_TTF];
    A.lo = f->buf[FCT_LO];
    A.adf_v = f->buf[FCT_ADF_V];
    A.adf_v_out = (mode >= 1) ? f->buf[FCT_ADF_V_OUT] : f->buf[FCT_ADF_V];
    A.adf_h_in = f->buf[FCT_ADF_H];
    A.adf_h_out = (mode >= 1) ? f->buf[FCT_ADF_H_OUT] : f->buf[FCT_ADF_H];
    A.ttf_max = f->buf[FCT_TTF_MAX];
    A.ttf_min = f->buf[FCT_TTF_MIN];
    A.plus = f->buf[FCT_PLUS];
    A.minus = f->buf[FCT_MINUS];
    A.del_v = f->buf[FCT_DEL_V];
    A.del_h = f->buf[FCT_DEL_H];
    A.uv_rhs = reinterpret_cast<double2 *>(f->buf[FCT_UV_RHS]);
    A.adf_v2 = f->buf[FCT_ADF_V2];
    A.adf_h2 = f->buf[FCT_ADF_H2];
    A.vlimit = 1;
    A.ts_node = f->ts_node;
    A.ts_nodev = f->ts_node;
    A.ts_edge = f->ts_edge;
    A.ts_uv = f->ts_uv;
    A.area = f->buf[FCT_AREA];
    A.area_inv = f->buf[FCT_AREA_INV];
    A.hnode = f->buf[FCT_HNODE];
    A.hnode_new = f->buf[FCT_HNODE_NEW];
    A.pitchL = A.pitchV = A.pitchH = A.pitchU = f->P;
    A.nl = f->plan->nl;
    A.dt = dt;
    A.eps = eps;
    A.big = big;
    A.flags = tune_int("FCT_WT_FLAGS", 0);
    return A;
}

// dense host rows (width W) <-> padded device rows (pitch P).  A strided cudaMemcpy2D of 400..600-byte
// rows runs at a few GB/s over PCIe; one contiguous copy plus this repack runs at link speed.
__global__ void k_repack(double *__restrict__ dst, const double *__restrict__ src, size_t rows, int W, int dpitch, int spitch)
{
    const size_t n = rows * (size_t)W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / (size_t)W;
        const int c = (int)(i - r * (size_t)W);
        dst[r * (size_t)dpitch + c] = src[r * (size_t)spitch + c];
    }
}

// dense host rows (width W) <-> packed columns: row r owns device slots [col[r], col[r+1]); only the
// levels that have a slot travel, a download fills the rest of the dense row with zeros
__global__ void k_pack_columns(double *__restrict__ dst, const double *__restrict__ src, const unsigned *__restrict__ col,
                               size_t rows, int W)
{
    const size_t n = rows * (size_t)W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / (size_t)W;
        const unsigned c = (unsigned)(i - r * (size_t)W);
        const unsigned c0 = __ldg(col + r), c1 = __ldg(col + r + 1);
        if (c < c1 - c0) dst[c0 + c] = src[i];
    }
}
__global__ void k_unpack_columns(double *__restrict__ dst, const double *__restrict__ src, const unsigned *__restrict__ col,
                                 size_t rows, int W)
{
    const size_t n = rows * (size_t)W;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t r = i / (size_t)W;
        const unsigned c = (unsigned)(i - r * (size_t)W);
        const unsigned c0 = __ldg(col + r), c1 = __ldg(col + r + 1);
        dst[i] = (c < c1 - c0) ? src[c0 + c] : 0.;
    }
}

// Measured alternative (knob FCT_DIRECT_COPY=1): the same two movements straight between PAGE-LOCKED
// host memory and the packed columns, one warp per row: the SMs read / write the host array over
// PCIe themselves (mapped, zero-copy), so only the levels that have a slot cross the link on the
// way up (the inactive 30 % of a dense array stay behind) and no dense staging copy exists in HBM.
// Measured on B200 (gpurun_out/s3_e2e_copy.log): SM-issued PCIe reads stop at 35 GB/s of useful
// bytes against 54 GB/s for the copy engine's dense copy (37 GB/s useful), so the staged copy stays
// the default; the direct path is for callers short of HBM.
__global__ void k_pack_columns_direct(double *__restrict__ dst, const double *__restrict__ host,
                                      const unsigned *__restrict__ col, size_t rows, int W)
{
    const int lane = threadIdx.x & 31;
    const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t r = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nw) {
        const unsigned c0 = __ldg(col + r);
        const int n = min((int)(__ldg(col + r + 1) - c0), W);
        const double *src = host + r * (size_t)W;
        for (int c = lane; c < n; c += 32) dst[c0 + c] = __ldcs(src + c);
    }
}
__global__ void k_unpack_columns_direct(double *__restrict__ host, const double *__restrict__ src,
                                        const unsigned *__restrict__ col, size_t rows, int W)
{
    const int lane = threadIdx.x & 31;
    const size_t nw = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t r = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nw) {
        const unsigned c0 = __ldg(col + r);
        const int n = min((int)(__ldg(col + r + 1) - c0), W);
        double *dst = host + r * (size_t)W;
        for (int c = lane; c < W; c += 32) __stcs(dst + c, c < n ? src[c0 + c] : 0.);
    }
}

// device alias of a page-locked (cudaMallocHost / cudaHostRegister) host pointer, else null
static double *mapped_alias(const void *host)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? static_cast<double *>(at.devicePointer) : nullptr;
}

static bool run_stage(Fields *f, const Arrays &A, int stage, const int *list, int first, int count, cudaStream_t s)
{
    return launch_stage(stage, 2, A, f->plan->dev, list, first, count, f->T, s);
}

// The two fused phases of one step (modes 1, 2, 3 of fct_ale_step_), overlapped with the halo exchange
// when there is one.  A.vlimit selects the a3 variant of the warp-item phase A.
static void fused_step(Fields *f, Halo *h, cudaStream_t s, const Arrays &A, int mode, int *alg_state, bool iter = false)
{
    const Plan *p = f->plan;
    const int N = p->N;
    const bool warped = (f->packed ? p->wtiles_pk_ok : p->wtiles_ok) && mode == 1;
    const bool tiled = p->tiles_ok && (mode == 3 || (mode == 1 && !warped));
    if ((A.vlimit != 1 || iter) && !warped) {
        std::fprintf(stderr, "fesom2-accelerate: vlimit 2 / 3 and the iterative branch are fused in the warp-item kernels only\n");
        return;
    }
    const int stage_b = iter ? ST_PHASE_B_ITER : ST_PHASE_B;
    // one fused phase over a node set: 0 all owned, 1 boundary, 2 interior
    auto phase = [&](int stage, int which) -> bool {
        if (warped) return launch_warp(stage, A, p, which, f->T, s);
        if (tiled) return launch_tile(stage, A, p, which, f->T, s);
        if (which == 0) return run_stage(f, A, stage, nullptr, 0, N, s);
        return which == 1 ? run_stage(f, A, stage, p->d_boundary, 0, p->n_boundary, s)
                          : run_stage(f, A, stage, p->d_interior, 0, p->n_interior, s);
    };
    if (!h) {
        if (!phase(ST_PHASE_A, 0)) return;
        *alg_state = 6;
        if (!phase(stage_b, 0)) return;
        *alg_state = 10;
        return;
    }
    // Overlapped schedule: the boundary nodes' factors are computed first and travel over NVLink
    // on the halo's own stream while the interior nodes run phase A and phase B; the boundary
    // nodes' phase B (the only consumer of remote factors) comes last.
    cudaStream_t c = halo_comm_stream(h);
    if (!phase(ST_PHASE_A, 1)) return;
    if (!cuda_ok(cudaEventRecord(halo_event(h, 0), s), "event") ||
        !cuda_ok(cudaStreamWaitEvent(c, halo_event(h, 0), 0), "wait"))
        return;
    // knob HALO_SKIP (timing experiment only, wrong halo rows): the same launches without the exchange,
    // so that the exposed part of the communication can be measured as a difference of step times
    const bool skip = tune_int("FCT_HALO_SKIP", 0) != 0;
    if (!cuda_ok(cudaEventRecord(halo_timing_event(h, 0), c), "event")) return;
    if (!skip && !halo_exchange(f, h, c)) return;
    if (!cuda_ok(cudaEventRecord(halo_timing_event(h, 1), c), "event")) return;
    if (!cuda_ok(cudaEventRecord(halo_event(h, 1), c), "event")) return;
    if (!phase(ST_PHASE_A, 2)) return;
    *alg_state = 6;
    if (!phase(stage_b, 2)) return;
    if (!cuda_ok(cudaStreamWaitEvent(s, halo_event(h, 1), 0), "wait")) return;
    if (!phase(stage_b, 1)) return;
    *alg_state = 10;
}

}   // namespace fct

using namespace fct;

extern "C" {

void fct_ale_plan_create_(void **plan, int *myDim_nod2D, int *eDim_nod2D, int *myDim_elem2D,
                          int *myDim_edge2D, int *nl, int *nlevels_nod2D, int *nlevels_elem2D,
                          int *elem2D_nodes, int *nod_in_elem2D_num, int *nod_in_elem2D,
                          int *nod_in_elem2D_dim, int *edges, int *edge_tri, int *istat)
{
    *plan = nullptr;
    *istat = 1;
    int ndev = 0;
    if (!cuda_ok(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev < 1) return;
    if (*nl < 3 || *nl > 0xffff) {
        std::fprintf(stderr, "fesom2-accelerate: nl = %d out of range\n", *nl);
        return;
    }
    Plan *p = create_plan_host(*myDim_nod2D, *eDim_nod2D, *myDim_elem2D, *myDim_edge2D, *nl, nlevels_nod2D,
                               nlevels_elem2D, elem2D_nodes, nod_in_elem2D_num, nod_in_elem2D,
                               *nod_in_elem2D_dim, edges, edge_tri);
    if (!p) return;
    *plan = p;
    *istat = 0;
}

void fct_ale_plan_destroy_(void **plan, int *istat)
{
    Plan *p = P_(plan);
    *istat = p ? 0 : 1;
    if (p) destroy_plan(p);
    if (plan) *plan = nullptr;
}

void fct_ale_plan_kernels_(void **plan, int *warp_tiles, int *staged_tiles, int *packed_tiles)
{
    Plan *p = P_(plan);
    *warp_tiles = (p && p->wtiles_ok) ? 1 : 0;
    *staged_tiles = (p && p->tiles_ok) ? 1 : 0;
    *packed_tiles = (p && p->wtiles_pk_ok) ? 1 : 0;
}

void fct_ale_plan_pitch_(void **plan, int *pitch)
{
    Plan *p = P_(plan);
    *pitch = p ? p->pitch : 0;
}

static size_t field_doubles(const Fields *f, int id)
{
    const FieldMeta m = meta_of(id);
    size_t n = field_rows(f, m.kind) * f->P * (m.per_tracer ? f->T : 1) * (id == FCT_UV_RHS ? 2 : 1);
    if (f->packed) n = (m.kind == ROW_EDGE ? f->ts_edge : f->ts_node) * (m.per_tracer ? f->T : 1);
    return n ? n : 1;
}

// buffers of the iterative branch (rejected flux parts), allocated at their first use
static bool ensure_iter_buffers(Fields *f)
{
    for (int id : {(int)FCT_ADF_V2, (int)FCT_ADF_H2}) {
        if (f->buf[id]) continue;
        const size_t n = field_doubles(f, id);
        if (!cuda_ok(cudaMalloc(&f->buf[id], n * sizeof(double)), "cudaMalloc(iter buffers)") ||
            !cuda_ok(cudaMemset(f->buf[id], 0, n * sizeof(double)), "cudaMemset(iter buffers)"))
            return false;
        // the caller's streams are non-blocking: they do not wait for the null stream's memset
        if (!cuda_ok(cudaStreamSynchronize(0), "cudaMemset(iter buffers)")) return false;
    }
    return true;
}

static void fields_create(void **fields, void **plan, int *ntracers, bool with_uv, bool packed, int *istat)
{
    *fields = nullptr;
    *istat = 1;
    Plan *p = P_(plan);
    if (!p || *ntracers < 1) return;
    if (packed && !p->wtiles_pk_ok) {
        std::fprintf(stderr, "fesom2-accelerate: the packed level storage needs a plan with warp-item tiles\n");
        return;
    }
    Fields *f = new (std::nothrow) Fields;
    if (!f) return;
    f->plan = p;
    f->T = *ntracers;
    f->packed = packed;
    f->P = packed ? 0 : p->pitch;
    f->rows = (size_t)p->N + p->H;
    f->ts_node = packed ? (size_t)p->ncol.back() : f->rows * f->P;
    f->ts_edge = packed ? (size_t)p->ecol.back() : (size_t)p->G * f->P;
    f->ts_uv = (size_t)p->E * f->P;
    bool ok = true;
    for (int id = 0; id < FCT_FIELD_COUNT && ok; ++id) {
        if (id == FCT_UV_RHS && (!with_uv || packed)) continue;
        if (id == FCT_ADF_V2 || id == FCT_ADF_H2) continue;   // on demand: ensure_iter_buffers
        const size_t n = field_doubles(f, id);
        ok = cuda_ok(cudaMalloc(&f->buf[id], n * sizeof(double)), "cudaMalloc(fields)") &&
             cuda_ok(cudaMemset(f->buf[id], 0, n * sizeof(double)), "cudaMemset(fields)");
    }
    // the callers' streams are non-blocking: make sure no memset of the null stream is still pending
    ok = ok && cuda_ok(cudaStreamSynchronize(0), "cudaMemset(fields)");
    if (!ok) {
        for (double *b : f->buf)
            if (b) cudaFree(b);
        delete f;
        return;
    }
    *fields = f;
    *istat = 0;
}

void fct_ale_fields_create_(void **fields, void **plan, int *ntracers, int *with_uv_rhs, int *istat)
{
    fields_create(fields, plan, ntracers, with_uv_rhs && *with_uv_rhs, false, istat);
}

void fct_ale_fields_create_packed_(void **fields, void **plan, int *ntracers, int *istat)
{
    fields_create(fields, plan, ntracers, false, true, istat);
}

void fct_ale_fields_destroy_(void **fields, int *istat)
{
    Fields *f = F_(fields);
    *istat = f ? 0 : 1;
    if (!f) return;
    for (double *b : f->buf)
        if (b) cudaFree(b);
    for (double *b : f->stage)
        if (b) cudaFree(b);
    f->magic = 0;
    delete f;
    *fields = nullptr;
}

static void field_copy(void **fields, int *field, int *tracer, real_type *host, void **stream, int *istat, bool up)
{
    *istat = 1;
    Fields *f = F_(fields);
    if (f && (*field == FCT_ADF_V2 || *field == FCT_ADF_H2) && !ensure_iter_buffers(f)) return;
    if (!f || *field < 0 || *field >= FCT_FIELD_COUNT || !f->buf[*field] || !host) return;
    const FieldMeta m = meta_of(*field);
    const int t = m.per_tracer ? *tracer : 0;
    if (t < 0 || t >= f->T) return;
    const size_t rows = field_rows(f, m.kind);
    const size_t unit = (*field == FCT_UV_RHS) ? 2 : 1;
    const size_t width = (size_t)(f->plan->nl - m.width_minus) * unit * sizeof(double);
    const size_t pitch = (size_t)f->P * unit * sizeof(double);
    double *d = f->buf[*field] + (f->packed ? (size_t)t * (m.kind == ROW_EDGE ? f->ts_edge : f->ts_node)
                                            : (size_t)t * rows * f->P * unit);
    if (rows == 0) {
        *istat = 0;
        return;
    }
    cudaStream_t st = S_(stream);
    if (f->packed && tune_int("FCT_DIRECT_COPY", 0) != 0) {
        if (double *alias = mapped_alias(host)) {
            const unsigned *col = m.kind == ROW_EDGE ? f->plan->d_ecol : f->plan->d_ncol;
            const int W = (int)(width / sizeof(double));
            const int blocks = (int)std::min<size_t>((rows + 7) / 8, (size_t)148 * 16);
            if (up) k_pack_columns_direct<<<blocks, 256, 0, st>>>(d, alias, col, rows, W);
            else k_unpack_columns_direct<<<blocks, 256, 0, st>>>(alias, d, col, rows, W);
            count_launch(1);
            *istat = cuda_ok(cudaGetLastError(), "direct pack") ? 0 : 1;
            return;
        }
    }
    if (!f->packed && width == pitch) {
        cudaError_t e = up ? cudaMemcpyAsync(d, host, rows * pitch, cudaMemcpyHostToDevice, st)
                           : cudaMemcpyAsync(host, d, rows * pitch, cudaMemcpyDeviceToHost, st);
        *istat = cuda_ok(e, up ? "field upload" : "field download") ? 0 : 1;
        return;
    }
    // one contiguous PCIe copy through a dense staging buffer + a repack kernel.  One buffer per
    // direction, stream ordered: all uploads of a Fields object belong on one stream, all its
    // downloads on one (possibly other) stream; ordering between the two is the caller's (events)
    const size_t W = width / sizeof(double), Pd = pitch / sizeof(double), need = rows * W;
    const int dir = up ? 0 : 1;
    if (need > f->stage_doubles[dir]) {
        cudaStreamSynchronize(st);
        if (f->stage[dir]) cudaFree(f->stage[dir]);
        f->stage[dir] = nullptr;
        f->stage_doubles[dir] = 0;
        if (!cuda_ok(cudaMalloc(&f->stage[dir], need * sizeof(double)), "cudaMalloc(staging)")) return;
        f->stage_doubles[dir] = need;
    }
    double *stage = f->stage[dir];
    const int threads = 256;
    const int blocks = (int)std::min<size_t>((need + threads - 1) / threads, (size_t)148 * 16);
    bool ok;
    const unsigned *col = m.kind == ROW_EDGE ? f->plan->d_ecol : f->plan->d_ncol;
    if (up) {
        ok = cuda_ok(cudaMemcpyAsync(stage, host, need * sizeof(double), cudaMemcpyHostToDevice, st), "field upload");
        if (ok) {
            if (f->packed) k_pack_columns<<<blocks, threads, 0, st>>>(d, stage, col, rows, (int)W);
            else k_repack<<<blocks, threads, 0, st>>>(d, stage, rows, (int)W, (int)Pd, (int)W);
            count_launch(1);
            ok = cuda_ok(cudaGetLastError(), "repack");
        }
    } else {
        if (f->packed) k_unpack_columns<<<blocks, threads, 0, st>>>(stage, d, col, rows, (int)W);
        else k_repack<<<blocks, threads, 0, st>>>(stage, d, rows, (int)W, (int)W, (int)Pd);
        count_launch(1);
        ok = cuda_ok(cudaGetLastError(), "repack") &&
             cuda_ok(cudaMemcpyAsync(host, stage, need * sizeof(double), cudaMemcpyDeviceToHost, st), "field download");
    }
    *istat = ok ? 0 : 1;
}

// ---- host arrays already in the packed level storage (a caller that keeps its columns packed) ----
// One contiguous copy straight into / out of the device array: no staging buffer, no repack kernel, and
// only the slots that exist cross the link (about 70 % of the dense array).
static void field_copy_packed(void **fields, int *field, int *tracer, real_type *host, void **stream, int *istat, bool up)
{
    *istat = 1;
    Fields *f = F_(fields);
    if (f && (*field == FCT_ADF_V2 || *field == FCT_ADF_H2) && !ensure_iter_buffers(f)) return;
    if (!f || !f->packed || *field < 0 || *field >= FCT_FIELD_COUNT || *field == FCT_UV_RHS || !f->buf[*field] || !host) {
        if (f && !f->packed) std::fprintf(stderr, "fesom2-accelerate: packed host copies need fields in the packed level storage\n");
        return;
    }
    const FieldMeta m = meta_of(*field);
    const int t = m.per_tracer ? *tracer : 0;
    if (t < 0 || t >= f->T) return;
    const size_t n = m.kind == ROW_EDGE ? f->ts_edge : f->ts_node;
    double *d = f->buf[*field] + (size_t)t * n;
    cudaStream_t st = S_(stream);
    const cudaError_t e = up ? cudaMemcpyAsync(d, host, n * sizeof(double), cudaMemcpyHostToDevice, st)
                             : cudaMemcpyAsync(host, d, n * sizeof(double), cudaMemcpyDeviceToHost, st);
    *istat = cuda_ok(e, up ? "packed field upload" : "packed field download") ? 0 : 1;
}

void fct_ale_field_upload_packed_(void **fields, int *field, int *tracer, real_type *host_packed, void **stream, int *istat)
{
    field_copy_packed(fields, field, tracer, host_packed, stream, istat, true);
}

void fct_ale_field_download_packed_(void **fields, int *field, int *tracer, real_type *host_packed, void **stream, int *istat)
{
    field_copy_packed(fields, field, tracer, host_packed, stream, istat, false);
}

void fct_ale_plan_packed_size_(void **plan, long long *node_doubles, long long *edge_doubles, int *istat)
{
    Plan *p = P_(plan);
    *istat = 1;
    *node_doubles = *edge_doubles = 0;
    if (!p || !p->wtiles_pk_ok || p->ncol.empty() || p->ecol.empty()) return;
    *node_doubles = (long long)p->ncol.back();
    *edge_doubles = (long long)p->ecol.back();
    *istat = 0;
}

void fct_ale_plan_packed_columns_(void **plan, int *kind, unsigned *columns, int *istat)
{
    Plan *p = P_(plan);
    *istat = 1;
    if (!p || !p->wtiles_pk_ok || !columns) return;
    const std::vector<unsigned> &col = *kind == 1 ? p->ecol : p->ncol;
    std::memcpy(columns, col.data(), col.size() * sizeof(unsigned));
    *istat = 0;
}

void fct_ale_field_link_bytes_(void **fields, int *field, real_type *host, int *upload, long long *bytes)
{
    *bytes = 0;
    Fields *f = F_(fields);
    if (!f || *field < 0 || *field >= FCT_FIELD_COUNT || !f->buf[*field] || !host) return;
    const FieldMeta m = meta_of(*field);
    const size_t rows = field_rows(f, m.kind);
    const size_t W = (size_t)(f->plan->nl - m.width_minus) * (*field == FCT_UV_RHS ? 2 : 1);
    *bytes = (long long)(rows * W * sizeof(double));
    if (!(f->packed && *upload && tune_int("FCT_DIRECT_COPY", 0) != 0 && mapped_alias(host))) return;
    const std::vector<unsigned> &col = m.kind == ROW_EDGE ? f->plan->ecol : f->plan->ncol;
    long long n = 0;
    for (size_t r = 0; r < rows; ++r) n += (long long)std::min<size_t>(col[r + 1] - col[r], W);
    *bytes = n * (long long)sizeof(double);
}

void fct_ale_field_upload_(void **fields, int *field, int *tracer, real_type *host, void **stream, int *istat)
{
    field_copy(fields, field, tracer, host, stream, istat, true);
}

void fct_ale_field_download_(void **fields, int *field, int *tracer, real_type *host, void **stream, int *istat)
{
    field_copy(fields, field, tracer, host, stream, istat, false);
}

void fct_ale_stage_(void **fields, void **stream, int *stage, real_type *dt, real_type *flux_eps,
                    real_type *bignumber, int *istat)
{
    *istat = 1;
    Fields *f = F_(fields);
    if (!f) return;
    const Plan *p = f->plan;
    const int st = *stage;
    if (st < 0 || st > ST_LAST) return;
    if (st >= ST_B3V_ITER && !f->packed && !ensure_iter_buffers(f)) return;
    const Arrays A = arrays_of(f, (st >= ST_PHASE_A && st < ST_B1H_ATOMIC) ? 1 : 0, *dt, *flux_eps, *bignumber);
    if (f->packed && !(st >= ST_PHASE_A_WARP && st <= 23)) {
        std::fprintf(stderr, "fesom2-accelerate: packed fields run the warp-item kernels only (stages 18-23)\n");
        return;
    }
    if (st >= ST_PHASE_A_WARP && st <= 23) {
        // 18/19: all owned nodes; 20/21: phase A on the boundary / interior tiles; 22/23: phase B
        if (f->packed ? !p->wtiles_pk_ok : !p->wtiles_ok) {
            std::fprintf(stderr, "fesom2-accelerate: this plan has no warp-item tiles\n");
            return;
        }
        const int which = st <= ST_PHASE_B_WARP ? 0 : 1 + ((st - 20) & 1);
        const int phase = st <= ST_PHASE_B_WARP ? (st == ST_PHASE_A_WARP ? ST_PHASE_A : ST_PHASE_B)
                                                : (st < 22 ? ST_PHASE_A : ST_PHASE_B);
        if (which != 0 && p->H == 0) {
            std::fprintf(stderr, "fesom2-accelerate: boundary / interior tiles exist on partitioned plans only\n");
            return;
        }
        if (launch_warp(phase, A, p, which, f->T, S_(stream))) *istat = 0;
        return;
    }
    if (st >= ST_PHASE_A_TILE && st <= 17) {
        // 12/13: all owned nodes; 14/15: phase A on the boundary / interior list; 16/17: phase B
        if (!p->tiles_ok) {
            std::fprintf(stderr, "fesom2-accelerate: this plan has no tiles\n");
            return;
        }
        const int which = st <= ST_PHASE_B_TILE ? 0 : 1 + ((st - 14) & 1);
        const int phase = st <= ST_PHASE_B_TILE ? (st == ST_PHASE_A_TILE ? ST_PHASE_A : ST_PHASE_B)
                                                : (st < 16 ? ST_PHASE_A : ST_PHASE_B);
        if (which != 0 && p->H == 0) {
            std::fprintf(stderr, "fesom2-accelerate: boundary / interior tiles exist on partitioned plans only\n");
            return;
        }
        if (launch_tile(phase, A, p, which, f->T, S_(stream))) *istat = 0;
        return;
    }
    int count = p->N;
    if (st == ST_A1) count = p->N + p->H;
    else if (st == ST_A2) count = p->E;
    else if (st == ST_B3H || st == ST_B1H_ATOMIC || st == ST_CH_ATOMIC || st == ST_B3H_ITER) count = p->G;
    if (st == ST_A2 || st == ST_A3 || st == ST_A3_VLIMIT2 || st == ST_A3_VLIMIT3) {
        if (!f->buf[FCT_UV_RHS]) {
            std::fprintf(stderr, "fesom2-accelerate: stage %d needs fields created with UV_rhs\n", st);
            return;
        }
    }
    if (run_stage(f, A, st, nullptr, 0, count, S_(stream))) *istat = 0;
}

void fct_ale_step_(void **fields, void **halo, void **stream, int *mode, real_type *dt,
                   real_type *flux_eps, real_type *bignumber, int *alg_state)
{
    *alg_state = 0;
    Fields *f = F_(fields);
    if (!f) return;
    const Plan *p = f->plan;
    cudaStream_t s = S_(stream);
    Halo *h = (halo && *halo) ? static_cast<Halo *>(*halo) : nullptr;
    if (h && !halo_valid(h)) return;
    const Arrays A = arrays_of(f, *mode, *dt, *flux_eps, *bignumber);
    const int N = p->N;
    if (f->packed && *mode != 1) {
        std::fprintf(stderr, "fesom2-accelerate: packed fields run mode 1 (the warp-item kernels) only\n");
        return;
    }
    if (*mode == 0) {
        if (!f->buf[FCT_UV_RHS]) {
            std::fprintf(stderr, "fesom2-accelerate: staged mode needs fields created with UV_rhs\n");
            return;
        }
        static const int pre[6] = {ST_A1, ST_A2, ST_A3, ST_B1V, ST_B1H, ST_B2};
        const int cnt[6] = {N + p->H, p->E, N, N, N, N};
        for (int i = 0; i < 6; ++i) {
            if (!run_stage(f, A, pre[i], nullptr, 0, cnt[i], s)) return;
            *alg_state = i + 1;
        }
        if (h && !halo_exchange(f, h, s)) return;
        static const int post[4] = {ST_B3V, ST_B3H, ST_CV, ST_CH};
        const int cnt2[4] = {N, p->G, N, N};
        for (int i = 0; i < 4; ++i) {
            if (!run_stage(f, A, post[i], nullptr, 0, cnt2[i], s)) return;
            *alg_state = 7 + i;
        }
        return;
    }
    fused_step(f, h, s, A, *mode, alg_state);
}

void fct_ale_step_general_(void **fields, void **halo, void **stream, int *vlimit, int *iter_yn, real_type *dt,
                            real_type *flux_eps, real_type *bignumber, int *alg_state)
{
    *alg_state = 0;
    Fields *f = F_(fields);
    if (!f) return;
    const Plan *p = f->plan;
    cudaStream_t s = S_(stream);
    Halo *h = (halo && *halo) ? static_cast<Halo *>(*halo) : nullptr;
    if (h && !halo_valid(h)) return;
    const int vl = *vlimit, N = p->N;
    const bool iter = *iter_yn != 0;
    if (vl < 1 || vl > 3) {
        std::fprintf(stderr, "fesom2-accelerate: vlimit = %d (1, 2 or 3)\n", vl);
        return;
    }
    if (f->packed) {
        // the fast path's own layout: fused warp-item phases, phase A in its vlimit variant
        if (iter && !ensure_iter_buffers(f)) return;
        Arrays A = arrays_of(f, 1, *dt, *flux_eps, *bignumber);
        A.vlimit = vl;
        fused_step(f, h, s, A, 1, alg_state, iter);
        if (!iter || *alg_state != 10) return;
        // fct_adf_* = fct_adf_*2 (md:288-289), then the fct_LO halo rows for the next pass
        *alg_state = 9;
        if (!cuda_ok(cudaMemcpyAsync(f->buf[FCT_ADF_V], f->buf[FCT_ADF_V2], field_doubles(f, FCT_ADF_V) * sizeof(double),
                                     cudaMemcpyDeviceToDevice, s), "fct_adf_v = fct_adf_v2") ||
            !cuda_ok(cudaMemcpyAsync(f->buf[FCT_ADF_H], f->buf[FCT_ADF_H2], field_doubles(f, FCT_ADF_H) * sizeof(double),
                                     cudaMemcpyDeviceToDevice, s), "fct_adf_h = fct_adf_h2"))
            return;
        if (h && !halo_exchange_field(f, h, s, FCT_LO)) return;
        *alg_state = 10;
        return;
    }
    if (!f->buf[FCT_UV_RHS]) {
        std::fprintf(stderr, "fesom2-accelerate: the general step runs the stage kernels on padded fields: create them with UV_rhs\n");
        return;
    }
    if (iter && !ensure_iter_buffers(f)) return;
    const Arrays A = arrays_of(f, 0, *dt, *flux_eps, *bignumber);
    const int a3 = vl == 1 ? ST_A3 : (vl == 2 ? ST_A3_VLIMIT2 : ST_A3_VLIMIT3);
    const int pre[6] = {ST_A1, ST_A2, a3, ST_B1V, ST_B1H, ST_B2};
    const int cnt[6] = {N + p->H, p->E, N, N, N, N};
    for (int i = 0; i < 6; ++i) {
        if (!run_stage(f, A, pre[i], nullptr, 0, cnt[i], s)) return;
        *alg_state = i + 1;
    }
    if (h && !halo_exchange(f, h, s)) return;
    if (!iter) {
        static const int post[4] = {ST_B3V, ST_B3H, ST_CV, ST_CH};
        const int cnt2[4] = {N, p->G, N, N};
        for (int i = 0; i < 4; ++i) {
            if (!run_stage(f, A, post[i], nullptr, 0, cnt2[i], s)) return;
            *alg_state = 7 + i;
        }
        return;
    }
    // docs/refactoring.md:226-290: limit, keep the rejected parts, update the low-order solution,
    // hand the rejected parts to the next pass; fct_LO halo rows travel to the neighbours
    if (!run_stage(f, A, ST_B3V_ITER, nullptr, 0, N, s)) return;
    *alg_state = 7;
    if (!run_stage(f, A, ST_B3H_ITER, nullptr, 0, p->G, s)) return;
    *alg_state = 8;
    if (!run_stage(f, A, ST_LO_UPDATE, nullptr, 0, N, s)) return;
    *alg_state = 9;
    if (!cuda_ok(cudaMemcpyAsync(f->buf[FCT_ADF_V], f->buf[FCT_ADF_V2], field_doubles(f, FCT_ADF_V) * sizeof(double),
                                 cudaMemcpyDeviceToDevice, s), "fct_adf_v = fct_adf_v2") ||
        !cuda_ok(cudaMemcpyAsync(f->buf[FCT_ADF_H], f->buf[FCT_ADF_H2], field_doubles(f, FCT_ADF_H) * sizeof(double),
                                 cudaMemcpyDeviceToDevice, s), "fct_adf_h = fct_adf_h2"))
        return;
    if (h && !halo_exchange_field(f, h, s, FCT_LO)) return;
    *alg_state = 10;
}

}   // extern "C"
