// Host-side plan of one (local) mesh: the uploaded connectivity plus the gather lists the
// node-centric kernels walk.  Built once per mesh (inspector), used by every step (executor).
#pragma once
#include <cuda_runtime.h>
#include <vector>

#include "fct_kernels.cuh"
#include "fct_tile_kernels.cuh"
#include "fct_warp_kernels.cuh"

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

// Per-tile tables of the tile-staged kernels (see fct_tile_kernels.cuh), for one node list.
struct TileSetHost {
    std::vector<int> row_off;
    std::vector<int2> rows;
    std::vector<int4> hdr;
    std::vector<int> work_off;
    std::vector<int4> ent;
    int ntiles = 0, max_rows = 0;
};
// Greedy tiling of `list` (nullptr: the identity 0..N-1): a tile closes after TN nodes or when its
// work items (slots of `vec` levels, padded so a node never straddles a block iteration of
// `threads` items) would exceed `budget`.  Returns false when the mesh is not a triangulation in
// the sense the merged gather list needs (ring neighbours == edge neighbours with equal depths) or
// a node has more than TE edges: the caller then keeps the general (untiled) kernels.
bool build_tileset(const DerivedHost &d, const int *nlev_n, int N, int NT, const std::vector<int> *list,
                   int TN, int TE, int vec, int threads, int budget, TileSetHost &out);

// Per-tile blobs of the warp-item kernels (see fct_warp_kernels.cuh), for one node list.
struct WarpTilesHost {
    std::vector<uint4> blob;
    std::vector<unsigned> blob_off;   // 16-byte units
    int ntiles = 0, smem_bytes = 0;
    // statistics (FCT_VERBOSE)
    long long nodes = 0, staged_rows = 0, staged_erows = 0, edge_uses = 0, slots = 0, lanes = 0, copies = 0;
};
// Greedy tiling of `list` (nullptr: the identity 0..N-1): a tile closes after TN nodes or when its
// shared-memory footprint (blob + two staged node-row regions + the edge-row region) would exceed
// smem_cap (the size of one stage of the kernels' ring).
// Returns false when the mesh is not a plain triangulation (ring neighbours == edge neighbours
// with equal depths, every edge at most as deep as both of its end nodes) or an offset does not
// fit its field: the caller then keeps the other kernels.
// ncol / ecol: column offsets of the packed level storage (packed_columns), or null for the padded
// layout (row = index * P).
bool build_warptiles(const DerivedHost &d, const int *nlev_n, int N, int NT, int G, int P, const unsigned *ncol,
                     const unsigned *ecol, const std::vector<int> *list, int TN, int smem_cap, WarpTilesHost &out);
// false: an offset would not fit 32 bits
bool packed_columns(const DerivedHost &d, const int *nlev_n, int NT, int G, std::vector<unsigned> &ncol,
                    std::vector<unsigned> &ecol);

struct Plan {
    unsigned magic = 0x504c414eu;
    int N = 0, H = 0, E = 0, G = 0, nl = 0, nie_dim = 0;
    int pitch = 0;          // padded row pitch of the device-resident path
    bool owns_mesh = false; // raw connectivity arrays allocated by the plan (plan_create) or borrowed (handles)
    MeshDev dev{};          // device pointers
    int *d_boundary = nullptr, *d_interior = nullptr;
    int n_boundary = 0, n_interior = 0;
    // tile-staged fused kernels: [phase A / B][0 all owned nodes, 1 boundary list, 2 interior list]
    TileDev tiles[2][3] = {};
    bool tiles_ok = false;
    // warp-item kernels: one tile set serves both phases; [0 all owned, 1 boundary, 2 interior]
    WarpTilesDev wtiles[3] = {};
    bool wtiles_ok = false;
    // packed level storage (fast path only): columns back to back, active levels only
    std::vector<unsigned> ncol, ecol;             // host: [N+H+1], [G+1] column offsets in doubles
    const unsigned *d_ncol = nullptr, *d_ecol = nullptr;
    WarpTilesDev wtiles_pk[3] = {};
    bool wtiles_pk_ok = false;
    std::vector<unsigned char> boundary_flag;   // host: [N] 1 = owned node with a halo neighbour (halo validation)
    std::vector<void *> owned;   // device allocations to free
};

}   // namespace fct
