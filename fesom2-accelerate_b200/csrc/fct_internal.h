// Declarations shared by the translation units of libfesom2-accelerate.so (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include "fct_kernels.cuh"
#include "fct_plan.h"

namespace fct {

enum Stage {
    ST_A1 = 0, ST_A2 = 1, ST_A3 = 2, ST_B1V = 3, ST_B1H = 4, ST_B2 = 5, ST_B3V = 6, ST_B3H = 7,
    ST_CV = 8, ST_CH = 9, ST_PHASE_A = 10, ST_PHASE_B = 11, ST_PHASE_A_TILE = 12, ST_PHASE_B_TILE = 13,
    ST_PHASE_A_WARP = 18, ST_PHASE_B_WARP = 19,
    ST_B1H_ATOMIC = 24, ST_CH_ATOMIC = 25,   // measured alternatives (fp64 atomics), never on the product path
    // vlimit 2 / 3 and the iterative branch (docs/refactoring.md:113-148, :226-290)
    ST_A3_VLIMIT2 = 26, ST_A3_VLIMIT3 = 27, ST_B3V_ITER = 28, ST_B3H_ITER = 29, ST_LO_UPDATE = 30, ST_LAST = 30,
    ST_PHASE_B_ITER = 31   // internal: the warp-item phase B of the iterative branch (launch_warp only)
};

bool cuda_ok(cudaError_t e, const char *what);
void count_launch(int n);
// tuning knob FCT_<...> (fct_ale_tune_ or the environment), full name with the FCT_ prefix
int tune_int(const char *name, int dflt);
bool launch_stage(int stage, int vec, const Arrays &A, const MeshDev &M, const int *list, int first,
                  int count, int ntracers, cudaStream_t s);
// tile-staged fused phase (stage = ST_PHASE_A / ST_PHASE_B) over node set `which`:
// 0 all owned nodes, 1 boundary list, 2 interior list
bool launch_tile(int stage, const Arrays &A, const Plan *p, int which, int ntracers, cudaStream_t s);
// warp-item fused phase (fct_warp_kernels.cuh), same arguments
bool launch_warp(int stage, const Arrays &A, const Plan *p, int which, int ntracers, cudaStream_t s);
Plan *create_plan_host(int N, int H, int E, int G, int nl, const int *nlev_n, const int *nlev_e,
                       const int *elem_nodes, const int *nie_num, const int *nie, int nie_dim,
                       const int *edges, const int *edge_tri);
void destroy_plan(Plan *p);

static const int FCT_FIELD_COUNT_INTERNAL = 19;   // == FCT_FIELD_COUNT of the public header
static const unsigned FIELDS_MAGIC = 0x464c4453u;
static const unsigned HALO_MAGIC = 0x48414c4fu;
static const unsigned PLAN_MAGIC = 0x504c414eu;

struct Fields {
    unsigned magic = FIELDS_MAGIC;
    Plan *plan = nullptr;
    int T = 0;         // tracers
    int P = 0;         // row pitch (doubles); 0: packed level storage (columns back to back)
    bool packed = false;
    size_t rows = 0;   // N + H
    double *buf[FCT_FIELD_COUNT_INTERNAL] = {nullptr};
    size_t ts_node = 0, ts_edge = 0, ts_uv = 0;
    // dense staging buffers of field_upload_ / field_download_, one per direction so that an upload
    // stream and a download stream can run concurrently (PCIe is full duplex)
    double *stage[2] = {nullptr, nullptr};
    size_t stage_doubles[2] = {0, 0};
};

struct Halo;
// pack + send/recv of fct_plus / fct_minus on stream s (all tracers of f)
bool halo_exchange(Fields *f, Halo *h, cudaStream_t s);
// the same for one per-tracer node array (fct_LO between two passes of the iterative branch, ttf, ...)
bool halo_exchange_field(Fields *f, Halo *h, cudaStream_t s, int field);
// streams / events used to overlap the exchange with interior work
cudaStream_t halo_comm_stream(Halo *h);
cudaEvent_t halo_event(Halo *h, int which);
cudaEvent_t halo_timing_event(Halo *h, int which);   // timing-enabled pair around the exchange
bool halo_valid(Halo *h);

}   // namespace fct
