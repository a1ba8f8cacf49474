// Halo exchange of fct_plus / fct_minus between mesh partitions (docs/refactoring.md:200, :235:
// exchange_nod / exchange_nod_end, which the reference stages through the host and MPI).
// Here both arrays stay on the device: one pack kernel gathers the boundary rows of all peers,
// grouped ncclSend / ncclRecv move them over NVLink, and the receives land directly in the halo
// rows (halo nodes are numbered grouped by owner, so no unpack is needed).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../include/fesom2-accelerate.h"
#include "fct_internal.h"

namespace fct {

struct Halo {
    unsigned magic = HALO_MAGIC;
    Plan *plan = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    std::vector<int> peers, send_off, send_cnt, recv_first, recv_cnt;
    int total_send = 0;
    int *d_send_nodes = nullptr;
    // packed level storage: first slot of every send column in the send buffer, [total_send + 1]
    std::vector<unsigned> send_col;
    unsigned *d_send_col = nullptr;
    double *sendbuf = nullptr;
    size_t sendbuf_doubles = 0;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // timing events around the exchange of the overlapped step (fct_ale_halo_comm_ms_)
    cudaEvent_t tev[2] = {nullptr, nullptr};
    bool timed = false;
};

// NCCL is bound at first use, not at link time: a host process that already carries an NCCL (for
// instance torch's bundled one, same SONAME as the system library) keeps a single copy, and the
// library loads on machines without NCCL as long as no halo is created.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
static NcclApi load_nccl()
{
    NcclApi a;
    void *h = nullptr;
    auto sym = [&](const char *name) -> void * {
        void *p = dlsym(RTLD_DEFAULT, name);
        if (!p) {
            if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
            if (h) p = dlsym(h, name);
        }
        return p;
    };
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
    a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
    a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
    a.Send = reinterpret_cast<decltype(a.Send)>(sym("ncclSend"));
    a.Recv = reinterpret_cast<decltype(a.Recv)>(sym("ncclRecv"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
    a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.GroupStart && a.GroupEnd && a.Send && a.Recv &&
           a.GetErrorString;
    if (!a.ok) std::fprintf(stderr, "fesom2-accelerate: NCCL (libnccl.so.2) is not available: no multi-GPU halo exchange\n");
    return a;
}
static const NcclApi &nccl()
{
    static const NcclApi a = load_nccl();
    return a;
}
#define ncclGetUniqueId nccl().GetUniqueId
#define ncclCommInitRank nccl().CommInitRank
#define ncclCommDestroy nccl().CommDestroy
#define ncclGroupStart nccl().GroupStart
#define ncclGroupEnd nccl().GroupEnd
#define ncclSend nccl().Send
#define ncclRecv nccl().Recv
#define ncclGetErrorString nccl().GetErrorString

bool halo_valid(Halo *h) { return h && h->magic == HALO_MAGIC; }
cudaStream_t halo_comm_stream(Halo *h) { return h->comm_stream; }
cudaEvent_t halo_event(Halo *h, int which) { return h->ev[which]; }
cudaEvent_t halo_timing_event(Halo *h, int which)
{
    h->timed = true;
    return h->tev[which];
}

static bool nccl_ok(ncclResult_t r, const char *what)
{
    if (r != ncclSuccess) {
        std::fprintf(stderr, "fesom2-accelerate: NCCL error \"%s\" in %s\n", ncclGetErrorString(r), what);
        return false;
    }
    return true;
}

// sendbuf[((t*2 + a) * total + i) * P + z] = (a ? minus : plus)[t][send_nodes[i]][z]
__global__ void k_halo_pack(const double *__restrict__ plus, const double *__restrict__ minus,
                            const int *__restrict__ send_nodes, int total, int P, size_t ts_node,
                            double *__restrict__ sendbuf)
{
    const int i = blockIdx.x * blockDim.y + threadIdx.y;
    if (i >= total) return;
    const int t = blockIdx.y;
    const int z = threadIdx.x * 2;
    if (z >= P) return;
    const size_t src = t * ts_node + (size_t)__ldg(send_nodes + i) * P + z;
    const double2 p = *reinterpret_cast<const double2 *>(plus + src);
    const double2 m = *reinterpret_cast<const double2 *>(minus + src);
    *reinterpret_cast<double2 *>(sendbuf + ((size_t)(t * 2 + 0) * total + i) * P + z) = p;
    *reinterpret_cast<double2 *>(sendbuf + ((size_t)(t * 2 + 1) * total + i) * P + z) = m;
}

// packed level storage: one block per send column
__global__ void k_halo_pack_columns(const double *__restrict__ plus, const double *__restrict__ minus,
                                    const int *__restrict__ send_nodes, const unsigned *__restrict__ send_col,
                                    const unsigned *__restrict__ ncol, size_t total_doubles, size_t ts_node,
                                    double *__restrict__ sendbuf)
{
    const int i = blockIdx.x, t = blockIdx.y;
    const int n = __ldg(send_nodes + i);
    const unsigned c0 = __ldg(ncol + n), cap = __ldg(ncol + n + 1) - c0, d0 = __ldg(send_col + i);
    for (unsigned c = threadIdx.x; c < cap; c += blockDim.x) {
        sendbuf[(size_t)(t * 2 + 0) * total_doubles + d0 + c] = plus[t * ts_node + c0 + c];
        sendbuf[(size_t)(t * 2 + 1) * total_doubles + d0 + c] = minus[t * ts_node + c0 + c];
    }
}

static bool halo_exchange_packed(Fields *f, Halo *h, cudaStream_t s, double *plus, double *minus, int narr)
{
    const Plan *p = f->plan;
    const int T = f->T;
    if (h->send_col.empty() || !h->d_send_col) {
        std::fprintf(stderr, "fesom2-accelerate: this halo was created without the packed column table\n");
        return false;
    }
    const size_t total = h->send_col.back();
    const size_t need = (size_t)2 * T * total;
    if (need > h->sendbuf_doubles) {
        if (h->sendbuf) cudaFree(h->sendbuf);
        h->sendbuf = nullptr;
        h->sendbuf_doubles = 0;
        if (!cuda_ok(cudaMalloc(&h->sendbuf, std::max<size_t>(need, 1) * sizeof(double)), "cudaMalloc(halo)")) return false;
        h->sendbuf_doubles = need;
    }
    if (h->total_send > 0) {
        dim3 grid(h->total_send, T);
        k_halo_pack_columns<<<grid, 64, 0, s>>>(plus, minus, h->d_send_nodes, h->d_send_col, p->d_ncol, total, f->ts_node, h->sendbuf);
        count_launch(1);
        if (!cuda_ok(cudaGetLastError(), "halo pack")) return false;
    }
    if (!nccl_ok(ncclGroupStart(), "ncclGroupStart")) return false;
    bool ok = true;
    for (size_t k = 0; k < h->peers.size() && ok; ++k) {
        const int peer = h->peers[k];
        const size_t s0 = h->send_col[h->send_off[k]], sn = h->send_col[h->send_off[k] + h->send_cnt[k]] - s0;
        const size_t r0 = p->ncol[h->recv_first[k]], rn = p->ncol[h->recv_first[k] + h->recv_cnt[k]] - r0;
        for (int t = 0; t < T && ok; ++t) {
            for (int a = 0; a < narr && ok; ++a) {
                if (sn > 0)
                    ok = nccl_ok(ncclSend(h->sendbuf + (size_t)(t * 2 + a) * total + s0, sn, ncclDouble, peer, h->comm, s), "ncclSend");
                if (ok && rn > 0)
                    ok = nccl_ok(ncclRecv((a ? minus : plus) + t * f->ts_node + r0, rn, ncclDouble, peer, h->comm, s), "ncclRecv");
            }
        }
    }
    ok = nccl_ok(ncclGroupEnd(), "ncclGroupEnd") && ok;
    return ok;
}

// owned-boundary rows -> the neighbours' halo rows of `narr` (1 or 2) node arrays of f
static bool halo_exchange_arrays(Fields *f, Halo *h, cudaStream_t s, double *plus, double *minus, int narr)
{
    if (!halo_valid(h) || h->plan != f->plan) {
        std::fprintf(stderr, "fesom2-accelerate: halo does not belong to these fields\n");
        return false;
    }
    if (f->packed) return halo_exchange_packed(f, h, s, plus, minus, narr);
    const int P = f->P, T = f->T;
    const size_t need = (size_t)2 * T * h->total_send * P;
    if (need > h->sendbuf_doubles) {
        if (h->sendbuf) cudaFree(h->sendbuf);
        h->sendbuf = nullptr;
        h->sendbuf_doubles = 0;
        if (!cuda_ok(cudaMalloc(&h->sendbuf, need * sizeof(double)), "cudaMalloc(halo)")) return false;
        h->sendbuf_doubles = need;
    }
    if (h->total_send > 0) {
        const int lx = P / 2;
        const int ny = 256 / lx > 0 ? 256 / lx : 1;
        dim3 block(lx, ny), grid((h->total_send + ny - 1) / ny, T);
        k_halo_pack<<<grid, block, 0, s>>>(plus, minus, h->d_send_nodes, h->total_send, P, f->ts_node, h->sendbuf);
        count_launch(1);
        if (!cuda_ok(cudaGetLastError(), "halo pack")) return false;
    }
    if (!nccl_ok(ncclGroupStart(), "ncclGroupStart")) return false;
    bool ok = true;
    for (size_t k = 0; k < h->peers.size() && ok; ++k) {
        const int peer = h->peers[k];
        for (int t = 0; t < T && ok; ++t) {
            for (int a = 0; a < narr && ok; ++a) {
                if (h->send_cnt[k] > 0)
                    ok = nccl_ok(ncclSend(h->sendbuf + ((size_t)(t * 2 + a) * h->total_send + h->send_off[k]) * P,
                                          (size_t)h->send_cnt[k] * P, ncclDouble, peer, h->comm, s), "ncclSend");
                if (ok && h->recv_cnt[k] > 0)
                    ok = nccl_ok(ncclRecv((a ? minus : plus) + t * f->ts_node + (size_t)h->recv_first[k] * P,
                                          (size_t)h->recv_cnt[k] * P, ncclDouble, peer, h->comm, s), "ncclRecv");
            }
        }
    }
    ok = nccl_ok(ncclGroupEnd(), "ncclGroupEnd") && ok;
    return ok;
}

bool halo_exchange(Fields *f, Halo *h, cudaStream_t s)
{
    return halo_exchange_arrays(f, h, s, f->buf[FCT_PLUS], f->buf[FCT_MINUS], 2);
}

bool halo_exchange_field(Fields *f, Halo *h, cudaStream_t s, int field)
{
    double *a = f->buf[field];
    return a && halo_exchange_arrays(f, h, s, a, a, 1);
}

}   // namespace fct

using namespace fct;

extern "C" {

void fct_ale_comm_unique_id_(char *id128, int *istat)
{
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    *istat = 1;
    if (!nccl().ok) return;
    *istat = nccl_ok(ncclGetUniqueId(&id), "ncclGetUniqueId") ? 0 : 1;
    if (*istat == 0) std::memcpy(id128, &id, 128);
}

void fct_ale_halo_create_(void **halo, void **plan, char *id128, int *rank, int *nranks, int *npeers,
                          int *peer_ranks, int *send_counts, int *send_nodes, int *recv_first,
                          int *recv_counts, int *istat)
{
    *halo = nullptr;
    *istat = 1;
    Plan *p = plan ? static_cast<Plan *>(*plan) : nullptr;
    if (!p || p->magic != PLAN_MAGIC || !nccl().ok) return;
    Halo *h = new (std::nothrow) Halo;
    if (!h) return;
    h->plan = p;
    h->rank = *rank;
    h->nranks = *nranks;
    int off = 0;
    for (int k = 0; k < *npeers; ++k) {
        h->peers.push_back(peer_ranks[k]);
        h->send_off.push_back(off);
        h->send_cnt.push_back(send_counts[k]);
        h->recv_first.push_back(recv_first[k]);
        h->recv_cnt.push_back(recv_counts[k]);
        off += send_counts[k];
    }
    h->total_send = off;
    // every id is checked before anything indexes with it: send nodes are owned BOUNDARY nodes of the
    // plan (the overlapped step sends right after phase A of the boundary set: any other node's
    // factors would leave stale), receive ranges lie inside the halo rows [N, N + H]
    for (int i = 0; i < off; ++i) {
        const int n = send_nodes[i];
        if (n < 0 || n >= p->N) {
            std::fprintf(stderr, "fesom2-accelerate: halo send node %d is not an owned node\n", n);
            delete h;
            return;
        }
        if (!p->boundary_flag.empty() && !p->boundary_flag[(size_t)n]) {
            std::fprintf(stderr, "fesom2-accelerate: halo send node %d has no halo neighbour in the plan's mesh\n", n);
            delete h;
            return;
        }
    }
    for (int k = 0; k < *npeers; ++k) {
        const long long r0 = recv_first[k], rn = recv_counts[k];
        if (peer_ranks[k] < 0 || peer_ranks[k] >= *nranks || peer_ranks[k] == *rank || send_counts[k] < 0 || rn < 0 ||
            (rn > 0 && (r0 < p->N || r0 + rn > (long long)p->N + p->H))) {
            std::fprintf(stderr, "fesom2-accelerate: halo peer %d: receive range [%lld, %lld) outside the halo rows [%d, %d) or bad rank / count\n",
                         peer_ranks[k], r0, r0 + rn, p->N, p->N + p->H);
            delete h;
            return;
        }
    }
    bool ok = cuda_ok(cudaMalloc(&h->d_send_nodes, (size_t)(off > 0 ? off : 1) * sizeof(int)), "cudaMalloc(halo)");
    if (ok && off > 0)
        ok = cuda_ok(cudaMemcpy(h->d_send_nodes, send_nodes, (size_t)off * sizeof(int), cudaMemcpyHostToDevice), "H2D(halo)");
    if (ok && !p->ncol.empty()) {
        h->send_col.assign((size_t)off + 1, 0u);
        for (int i = 0; i < off && ok; ++i) {
            const int n = send_nodes[i];
            h->send_col[i + 1] = h->send_col[i] + (p->ncol[n + 1] - p->ncol[n]);
        }
        ok = ok && cuda_ok(cudaMalloc(&h->d_send_col, h->send_col.size() * sizeof(unsigned)), "cudaMalloc(halo)") &&
             cuda_ok(cudaMemcpy(h->d_send_col, h->send_col.data(), h->send_col.size() * sizeof(unsigned), cudaMemcpyHostToDevice), "H2D(halo)");
    }
    ok = ok && cuda_ok(cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking), "stream");
    ok = ok && cuda_ok(cudaEventCreateWithFlags(&h->ev[0], cudaEventDisableTiming), "event");
    ok = ok && cuda_ok(cudaEventCreateWithFlags(&h->ev[1], cudaEventDisableTiming), "event");
    ok = ok && cuda_ok(cudaEventCreate(&h->tev[0]), "event") && cuda_ok(cudaEventCreate(&h->tev[1]), "event");
    if (ok) {
        ncclUniqueId id;
        std::memcpy(&id, id128, 128);
        ok = nccl_ok(ncclCommInitRank(&h->comm, *nranks, id, *rank), "ncclCommInitRank");
    }
    if (!ok) {
        int st;
        void *hp = h;
        fct_ale_halo_destroy_(&hp, &st);
        return;
    }
    *halo = h;
    *istat = 0;
}

void fct_ale_halo_destroy_(void **halo, int *istat)
{
    Halo *h = (halo && *halo) ? static_cast<Halo *>(*halo) : nullptr;
    *istat = 1;
    if (!halo_valid(h)) return;
    if (h->comm) ncclCommDestroy(h->comm);
    if (h->d_send_nodes) cudaFree(h->d_send_nodes);
    if (h->d_send_col) cudaFree(h->d_send_col);
    if (h->sendbuf) cudaFree(h->sendbuf);
    if (h->comm_stream) cudaStreamDestroy(h->comm_stream);
    for (auto &e : h->ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : h->tev)
        if (e) cudaEventDestroy(e);
    h->magic = 0;
    delete h;
    *halo = nullptr;
    *istat = 0;
}

// Device time of the last overlapped step's exchange on the halo's own stream (pack kernel + grouped
// send / recv, including the wait for the slowest peer).  Synchronises on that exchange.
void fct_ale_halo_comm_ms_(void **halo, real_type *ms, int *istat)
{
    Halo *h = (halo && *halo) ? static_cast<Halo *>(*halo) : nullptr;
    *istat = 1;
    *ms = 0.;
    if (!halo_valid(h) || !h->timed) return;
    float t = 0.f;
    if (!cuda_ok(cudaEventSynchronize(h->tev[1]), "cudaEventSynchronize") ||
        !cuda_ok(cudaEventElapsedTime(&t, h->tev[0], h->tev[1]), "cudaEventElapsedTime"))
        return;
    *ms = (double)t;
    *istat = 0;
}

void fct_ale_halo_exchange_(void **fields, void **halo, void **stream, int *istat)
{
    Fields *f = fields ? static_cast<Fields *>(*fields) : nullptr;
    Halo *h = halo ? static_cast<Halo *>(*halo) : nullptr;
    *istat = 1;
    if (!f || f->magic != FIELDS_MAGIC || !halo_valid(h)) return;
    cudaStream_t s = (stream && *stream) ? *static_cast<cudaStream_t *>(*stream) : (cudaStream_t)0;
    if (halo_exchange(f, h, s)) *istat = 0;
}

void fct_ale_halo_exchange_field_(void **fields, void **halo, void **stream, int *field, int *istat)
{
    Fields *f = fields ? static_cast<Fields *>(*fields) : nullptr;
    Halo *h = halo ? static_cast<Halo *>(*halo) : nullptr;
    *istat = 1;
    if (!f || f->magic != FIELDS_MAGIC || !halo_valid(h)) return;
    // node arrays of pitch nl-1 only (the rows the pack kernels move)
    const int id = *field;
    const bool node_L = id == FCT_TTF || id == FCT_LO || id == FCT_DEL_V || id == FCT_DEL_H || id == FCT_TTF_MAX ||
                        id == FCT_TTF_MIN || id == FCT_PLUS || id == FCT_MINUS;
    if (!node_L) return;
    cudaStream_t s = (stream && *stream) ? *static_cast<cudaStream_t *>(*stream) : (cudaStream_t)0;
    if (halo_exchange_field(f, h, s, id)) *istat = 0;
}

}   // extern "C"
