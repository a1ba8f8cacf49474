// stress2rhs -- EVP sea-ice dynamics: divergence of the stress tensor into the rhs vectors.
// the other routine the reference analysed (docs/refactoring.md:404-461)
// and restated on the CPU (src/reference.cpp:440-480); the reference has no GPU kernel for it.
//
// The element loop of the reference scatters three contributions per element into its corner nodes
// (U_rhs_ice[node] -= ...).  Here, as for b1 horizontal / c horizontal of fct_ale, the scatter is a
// deterministic node-centric gather: the inspector inverts elem2D_nodes into a CSR list of
// (element, corner) pairs per node in ascending element order -- the order in which the sequential
// loop reaches the node -- so the sums are bit-identical to src/reference.cpp and need no atomics.
// Index expressions are the reference's own, element stride elem2D_nodes_size for the connectivity
// and the "corner * 6 + element" addressing of gradient_sca (src/reference.cpp:460-461).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <new>
#include <vector>

#include "../../include/fesom2-accelerate.h"
#include "fct_internal.h"

namespace fct {

static const unsigned STRESS_MAGIC = 0x53324852u;   // "S2HR"

struct StressPlan {
    unsigned magic = STRESS_MAGIC;
    int N = 0, E = 0;
    int *d_off = nullptr;   // [N+1]
    int *d_ent = nullptr;   // element * 4 + corner, ascending
};

struct StressArrays {
    double *U, *V;
    const double *ice_strength, *elem_area, *s11, *s12, *s22, *grad, *metric, *inv_areamass, *rhs_a, *rhs_m;
};

// one thread per node; the element scalars of a ring are shared by the neighbouring nodes' threads
// through L1 / L2 (consecutive nodes of a curve-ordered mesh share most of their ring)
__global__ void __launch_bounds__(256) k_stress2rhs(StressArrays A, const int *__restrict__ off, const int *__restrict__ ent, int N)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double one_third = 1.0 / 3.0;
    double u = 0.0, v = 0.0;
    const int b = __ldg(off + n), e = __ldg(off + n + 1);
    for (int k = b; k < e; ++k) {
        const int w = __ldg(ent + k);
        const int el = w >> 2, c = w & 3;
        if (__ldg(A.ice_strength + el) > 0.0) {
            const double ar = __ldg(A.elem_area + el), s11 = __ldg(A.s11 + el), s12 = __ldg(A.s12 + el), s22 = __ldg(A.s22 + el);
            const double g0 = __ldg(A.grad + (size_t)c * 6 + el), g1 = __ldg(A.grad + (size_t)(c + 3) * 6 + el);
            const double mf = __ldg(A.metric + el);
            u = u - ar * ((s11 * g0) + (s12 * g1) + (s12 * one_third * mf));
            v = v - ar * ((s12 * g0) + (s22 * g1) - (s11 * one_third * mf));
        }
    }
    const double ia = __ldg(A.inv_areamass + n);
    if (ia > 0.0) {
        u = (u * ia) + __ldg(A.rhs_a + n);
        v = (v * ia) + __ldg(A.rhs_m + n);
    } else {
        u = 0.0;
        v = 0.0;
    }
    A.U[n] = u;
    A.V[n] = v;
}

static inline StressPlan *SP(void **p)
{
    StressPlan *q = p ? static_cast<StressPlan *>(*p) : nullptr;
    return (q && q->magic == STRESS_MAGIC) ? q : nullptr;
}

static StressPlan *stress_plan_create(int N, int E, int size, const int *elem_nodes)
{
    if (N < 0 || E < 0 || size < E || !elem_nodes) return nullptr;
    std::vector<int> off((size_t)N + 1, 0);
    for (int c = 0; c < 3; ++c)
        for (int el = 0; el < E; ++el) {
            const int n = elem_nodes[(size_t)c * size + el];   // 0-based, src/reference.cpp:457
            if (n < 0) {
                std::fprintf(stderr, "fesom2-accelerate: stress2rhs: negative node id\n");
                return nullptr;
            }
            if (n < N) ++off[(size_t)n + 1];
        }
    for (int n = 0; n < N; ++n) off[(size_t)n + 1] += off[n];
    std::vector<int> ent((size_t)off[N]), cur(off.begin(), off.end() - 1);
    // ascending element, then corner: the order of the reference's loops
    for (int el = 0; el < E; ++el)
        for (int c = 0; c < 3; ++c) {
            const int n = elem_nodes[(size_t)c * size + el];
            if (n < N) ent[(size_t)cur[n]++] = el * 4 + c;
        }
    StressPlan *p = new (std::nothrow) StressPlan;
    if (!p) return nullptr;
    p->N = N;
    p->E = E;
    bool ok = cuda_ok(cudaMalloc(&p->d_off, off.size() * sizeof(int)), "cudaMalloc(stress plan)") &&
              cuda_ok(cudaMalloc(&p->d_ent, std::max<size_t>(ent.size(), 1) * sizeof(int)), "cudaMalloc(stress plan)") &&
              cuda_ok(cudaMemcpy(p->d_off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice), "H2D(stress plan)") &&
              (ent.empty() || cuda_ok(cudaMemcpy(p->d_ent, ent.data(), ent.size() * sizeof(int), cudaMemcpyHostToDevice), "H2D(stress plan)"));
    if (!ok) {
        if (p->d_off) cudaFree(p->d_off);
        if (p->d_ent) cudaFree(p->d_ent);
        delete p;
        return nullptr;
    }
    return p;
}

static void stress_plan_destroy(StressPlan *p)
{
    cudaFree(p->d_off);
    cudaFree(p->d_ent);
    p->magic = 0;
    delete p;
}

static bool stress_launch(const StressPlan *p, const StressArrays &A, cudaStream_t s)
{
    if (p->N == 0) return true;
    k_stress2rhs<<<(p->N + 255) / 256, 256, 0, s>>>(A, p->d_off, p->d_ent, p->N);
    count_launch(1);
    return cuda_ok(cudaGetLastError(), "stress2rhs launch");
}

template <class T>
static T *dptr(void **h)
{
    gpuMemory *g = h ? static_cast<gpuMemory *>(*h) : nullptr;
    return g ? static_cast<T *>(g->device_pointer) : nullptr;
}

}   // namespace fct

using namespace fct;

extern "C" {

void stress2rhs_plan_create_(void **plan, int *myDim_nod2D, int *myDim_elem2D, int *elem2D_nodes_size,
                             int *elem2D_nodes, int *istat)
{
    *plan = nullptr;
    *istat = 1;
    int ndev = 0;
    if (!cuda_ok(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev < 1) return;
    StressPlan *p = stress_plan_create(*myDim_nod2D, *myDim_elem2D, *elem2D_nodes_size, elem2D_nodes);
    if (!p) return;
    *plan = p;
    *istat = 0;
}

void stress2rhs_plan_destroy_(void **plan, int *istat)
{
    StressPlan *p = SP(plan);
    *istat = p ? 0 : 1;
    if (p) stress_plan_destroy(p);
    if (plan) *plan = nullptr;
}

void stress2rhs_acc_(void **plan, void **s, void **U_rhs_ice, void **V_rhs_ice, void **ice_strength,
                     void **elem_area, void **sigma11, void **sigma12, void **sigma22, void **gradient_sca,
                     void **metric_factor, void **inv_areamass, void **rhs_a, void **rhs_m, int *istat)
{
    *istat = 1;
    StressPlan *p = SP(plan);
    if (!p) return;
    StressArrays A;
    A.U = dptr<double>(U_rhs_ice);
    A.V = dptr<double>(V_rhs_ice);
    A.ice_strength = dptr<double>(ice_strength);
    A.elem_area = dptr<double>(elem_area);
    A.s11 = dptr<double>(sigma11);
    A.s12 = dptr<double>(sigma12);
    A.s22 = dptr<double>(sigma22);
    A.grad = dptr<double>(gradient_sca);
    A.metric = dptr<double>(metric_factor);
    A.inv_areamass = dptr<double>(inv_areamass);
    A.rhs_a = dptr<double>(rhs_a);
    A.rhs_m = dptr<double>(rhs_m);
    if (!A.U || !A.V || !A.ice_strength || !A.elem_area || !A.s11 || !A.s12 || !A.s22 || !A.grad || !A.metric ||
        !A.inv_areamass || !A.rhs_a || !A.rhs_m)
        return;
    cudaStream_t st = (s && *s) ? *static_cast<cudaStream_t *>(*s) : (cudaStream_t)0;
    if (stress_launch(p, A, st)) *istat = 0;
}

void stress2rhs_(int *myDim_nod2D, int *myDim_elem2D, int *elem2D_nodes_size, real_type *U_rhs_ice,
                 real_type *V_rhs_ice, real_type *ice_strength, int *elem2D_nodes, real_type *elem_area,
                 real_type *sigma11, real_type *sigma12, real_type *sigma22, real_type *gradient_sca,
                 real_type *metric_factor, real_type *inv_areamass, real_type *rhs_a, real_type *rhs_m, int *istat)
{
    *istat = 1;
    const int N = *myDim_nod2D, E = *myDim_elem2D;
    int ndev = 0;
    if (!cuda_ok(cudaGetDeviceCount(&ndev), "cudaGetDeviceCount") || ndev < 1) return;
    StressPlan *p = stress_plan_create(N, E, *elem2D_nodes_size, elem2D_nodes);
    if (!p) return;
    std::vector<void *> bufs;
    bool ok = true;
    auto up = [&](const double *h, size_t n) -> double * {
        double *d = nullptr;
        if (!ok) return nullptr;
        ok = cuda_ok(cudaMalloc(&d, std::max<size_t>(n, 1) * sizeof(double)), "cudaMalloc(stress2rhs)");
        if (!ok) return nullptr;
        bufs.push_back(d);
        if (h && n) ok = cuda_ok(cudaMemcpy(d, h, n * sizeof(double), cudaMemcpyHostToDevice), "H2D(stress2rhs)");
        return d;
    };
    StressArrays A;
    A.U = up(nullptr, N);
    A.V = up(nullptr, N);
    A.ice_strength = up(ice_strength, E);
    A.elem_area = up(elem_area, E);
    A.s11 = up(sigma11, E);
    A.s12 = up(sigma12, E);
    A.s22 = up(sigma22, E);
    A.grad = up(gradient_sca, (size_t)E + 30);   // highest index the reference reads: 5*6 + E-1
    A.metric = up(metric_factor, E);
    A.inv_areamass = up(inv_areamass, N);
    A.rhs_a = up(rhs_a, N);
    A.rhs_m = up(rhs_m, N);
    ok = ok && stress_launch(p, A, 0) &&
         cuda_ok(cudaMemcpy(U_rhs_ice, A.U, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost), "D2H(stress2rhs)") &&
         cuda_ok(cudaMemcpy(V_rhs_ice, A.V, (size_t)N * sizeof(double), cudaMemcpyDeviceToHost), "D2H(stress2rhs)");
    for (void *b : bufs) cudaFree(b);
    stress_plan_destroy(p);
    if (ok) *istat = 0;
}

}   // extern "C"
