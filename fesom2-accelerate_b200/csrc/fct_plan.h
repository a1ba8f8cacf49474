// Host-side plan of one (local) mesh: the uploaded connectivity plus the gather lists the
// node-centric kernels walk.  Built once per mesh (inspector), used by every step (executor).
#pragma once
#include <cuda_runtime.h>
#include <vector>

#include "fct_kernels.cuh"

namespace fct {

struct DerivedHost {
    std::vector<int> nbr_off;    // [N+1]
    std::vector<int2> nbr;       // unique nodes of the ring elements of each owned node, self first
    std::vector<int> fillmin;    // [N]
    std::vector<int> edg_off;    // [N+H+1]
    std::vector<int4> edg;       // per node: incident edges, ascending edge id
    std::vector<int> boundary;   // owned nodes with a halo neighbour (ascending)
    std::vector<int> interior;   // the other owned nodes (ascending)
};

// All connectivity 1-based as in the ABI.  Returns false on malformed input.
bool build_derived(int N, int H, int E, int G, int nl, const int *nlev_e, const int *elem_nodes,
                   const int *nie_num, const int *nie, int nie_dim, const int *edges,
                   const int *edge_tri, DerivedHost &out);

struct Plan {
    unsigned magic = 0x504c414eu;
    int N = 0, H = 0, E = 0, G = 0, nl = 0, nie_dim = 0;
    int pitch = 0;          // padded row pitch of the device-resident path
    bool owns_mesh = false; // raw connectivity arrays allocated by the plan (plan_create) or borrowed (handles)
    MeshDev dev{};          // device pointers
    int *d_boundary = nullptr, *d_interior = nullptr;
    int n_boundary = 0, n_interior = 0;
    std::vector<void *> owned;   // device allocations to free
};

}   // namespace fct
