// Inspector: derive the node-centric gather lists from FESOM2's connectivity.
//
//  * node -> edge list in ascending edge id (both roles of a degenerate n1 == n2 edge, first role
//    first): walking it reproduces the summation order of the sequential edge loops of
//    src/reference.cpp:406-425 and docs/refactoring.md:303-314, so the gathers that replace the
//    fp64 atomicAdd scatter of kernels/fct_ale_b1_horizontal.cu / fct_ale_c_horizontal.cu are
//    deterministic AND bit-identical to the CPU reference.
//  * node -> unique nodes of its ring elements, each with the depth of the deepest ring element
//    that holds it, plus the first level at which some ring element is already below its bottom
//    (there a2 stores (-big, +big), src/reference.cpp:341-349).  max/min over that list equals a2
//    followed by the ring reduction of a3 (src/reference.cpp:358-378) without UV_rhs.
#include "fct_plan.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace fct {

bool build_derived(int N, int H, int E, int G, int nl, const int *nlev_e, const int *elem_nodes,
                   const int *nie_num, const int *nie, int nie_dim, const int *edges,
                   const int *edge_tri, DerivedHost &out)
{
    const int NT = N + H;
    (void)nl;
    // ---- edges ----
    out.edg_off.assign((size_t)NT + 1, 0);
    for (int g = 0; g < G; ++g) {
        const int a = edges[2 * g] - 1, b = edges[2 * g + 1] - 1;
        if (a < 0 || a >= NT || b < 0 || b >= NT) {
            std::fprintf(stderr, "fct plan: edge %d has a node outside [1,%d]\n", g + 1, NT);
            return false;
        }
        out.edg_off[a + 1]++;
        out.edg_off[b + 1]++;
    }
    for (int n = 0; n < NT; ++n) out.edg_off[n + 1] += out.edg_off[n];
    out.edg.resize((size_t)out.edg_off[NT]);
    std::vector<int> cur(out.edg_off.begin(), out.edg_off.end() - 1);
    std::vector<char> is_boundary((size_t)N, 0);
    for (int g = 0; g < G; ++g) {
        const int a = edges[2 * g] - 1, b = edges[2 * g + 1] - 1;
        const int el = edge_tri[2 * g] - 1, er = edge_tri[2 * g + 1] - 1;
        if (el < 0 || el >= E || er >= E) {
            std::fprintf(stderr, "fct plan: edge %d has an element outside [1,%d]\n", g + 1, E);
            return false;
        }
        if (a >= N && b >= N) {
            // myDim_edge2D holds the edges that touch an owned node; an edge between two
            // halo nodes would have no node to store its limited flux in the fused phase B
            std::fprintf(stderr, "fct plan: edge %d joins two halo nodes (no owned end): not a FESOM myDim_edge2D edge\n", g + 1);
            return false;
        }
        const int d1 = nlev_e[el] - 1;
        const int d2 = (er >= 0) ? nlev_e[er] - 1 : 0;
        const int depth = std::max(std::max(d1, d2), 0);
        // the first end node stores the limited flux when it is owned, else the second one does
        const int writer_first = (a < N) ? 1 : 0;
        out.edg[(size_t)cur[a]++] = make_int4(g, b, depth | (0 << 16) | (writer_first << 17), 0);
        out.edg[(size_t)cur[b]++] = make_int4(g, a, depth | (1 << 16) | ((1 - writer_first) << 17), 0);
        if (a < N && b >= N) is_boundary[a] = 1;
        if (b < N && a >= N) is_boundary[b] = 1;
    }
    out.boundary.clear();
    out.interior.clear();
    for (int n = 0; n < N; ++n) (is_boundary[n] ? out.boundary : out.interior).push_back(n);

    // ---- ring neighbours ----
    out.nbr_off.assign((size_t)N + 1, 0);
    out.fillmin.assign((size_t)N, 0);
    out.nbr.clear();
    out.nbr.reserve((size_t)N * 8);
    std::vector<int2> loc;
    const bool rings = nie_num && nie && elem_nodes;   // edge-only plans leave the ring lists empty
    for (int n = 0; rings && n < N; ++n) {
        loc.clear();
        loc.push_back(make_int2(n, 0));   // self first; depth 0 = "not part of any ring element"
        int fm = 1 << 30;
        const int cnt = nie_num[n];
        if (cnt < 1 || cnt > nie_dim) {
            std::fprintf(stderr, "fct plan: node %d has %d ring elements (dim %d)\n", n + 1, cnt, nie_dim);
            return false;
        }
        for (int k = 0; k < cnt; ++k) {
            const int e = nie[(size_t)n * nie_dim + k] - 1;
            if (e < 0 || e >= E) {
                std::fprintf(stderr, "fct plan: node %d ring element outside [1,%d]\n", n + 1, E);
                return false;
            }
            const int d = nlev_e[e] - 1;
            fm = std::min(fm, d);
            for (int j = 0; j < 3; ++j) {
                const int m = elem_nodes[3 * e + j] - 1;
                if (m < 0 || m >= NT) return false;
                bool found = false;
                for (auto &x : loc)
                    if (x.x == m) {
                        x.y = std::max(x.y, d);
                        found = true;
                        break;
                    }
                if (!found) loc.push_back(make_int2(m, d));
            }
        }
        out.fillmin[n] = std::max(fm, 0);
        std::sort(loc.begin() + 1, loc.end(), [](const int2 &p, const int2 &q) { return p.x < q.x; });
        out.nbr.insert(out.nbr.end(), loc.begin(), loc.end());
        out.nbr_off[n + 1] = (int)out.nbr.size();
    }
    return true;
}

bool build_tileset(const DerivedHost &d, const int *nlev_n, int N, int NT, const std::vector<int> *list,
                   int TN, int TE, int vec, int threads, int budget, TileSetHost &out)
{
    const int count = list ? (int)list->size() : N;
    out = TileSetHost();
    out.row_off.assign(1, 0);
    std::vector<int> stamp((size_t)NT, -1), local((size_t)NT, 0);
    std::vector<int> tile_nodes, tile_wo;
    std::vector<std::pair<int, int>> a, b;
    int pos = 0, tile = 0;
    while (pos < count) {
        // ---- pick the nodes of this tile ----
        tile_nodes.clear();
        tile_wo.clear();
        int items = 0;
        while (pos < count && (int)tile_nodes.size() < TN) {
            const int n = list ? (*list)[pos] : pos;
            const int nz = std::max(nlev_n[n] - 1, 0);
            const int ch = (nz + vec - 1) / vec;
            if (ch > threads) return false;
            int start = items;
            if (ch > 0 && start / threads != (start + ch - 1) / threads) start = (start / threads + 1) * threads;
            if (start + ch > budget && !tile_nodes.empty()) break;
            tile_nodes.push_back(n);
            tile_wo.push_back(start);
            items = start + ch;
            ++pos;
        }
        const int nt = (int)tile_nodes.size();
        for (int i = 0; i < nt; ++i) {
            stamp[tile_nodes[i]] = tile;
            local[tile_nodes[i]] = i;
            out.rows.push_back(make_int2(tile_nodes[i], std::max(nlev_n[tile_nodes[i]] - 1, 0)));
        }
        // ---- headers, gather lists, halo rows ----
        int nrows = nt;
        for (int i = 0; i < TN; ++i) {
            if (i >= nt) {
                out.hdr.push_back(make_int4(-1, 0, 0, 0));
                out.work_off.push_back(items);
                for (int k = 0; k < TE; ++k) out.ent.push_back(make_int4(0, 0, 0, 0));
                continue;
            }
            const int n = tile_nodes[i];
            const int nb0 = d.nbr_off[n], nb1 = d.nbr_off[n + 1];
            const int e0 = d.edg_off[n], e1 = d.edg_off[n + 1];
            const int cnt = e1 - e0;
            if (cnt > TE || nb1 - nb0 < 1 || d.nbr[nb0].x != n) return false;
            // triangulation check: {ring neighbours} == {other ends of the edges}, same depths
            a.clear();
            b.clear();
            for (int k = nb0 + 1; k < nb1; ++k) a.emplace_back(d.nbr[k].x, d.nbr[k].y);
            for (int k = e0; k < e1; ++k) {
                if (d.edg[k].y == n) return false;
                b.emplace_back(d.edg[k].y, FCT_META_DEPTH(d.edg[k].z));
            }
            std::sort(a.begin(), a.end());
            std::sort(b.begin(), b.end());
            if (a != b) return false;
            for (size_t k = 1; k < b.size(); ++k)
                if (b[k].first == b[k - 1].first) return false;
            out.hdr.push_back(make_int4(n, std::max(nlev_n[n] - 1, 0), d.fillmin[n], (d.nbr[nb0].y & 0xffff) | (cnt << 16)));
            out.work_off.push_back(tile_wo[i]);
            for (int k = 0; k < TE; ++k) {
                if (k >= cnt) {
                    out.ent.push_back(make_int4(0, 0, 0, 0));
                    continue;
                }
                const int4 e = d.edg[e0 + k];
                const int m = e.y;
                if (stamp[m] != tile) {
                    stamp[m] = tile;
                    local[m] = nrows++;
                    out.rows.push_back(make_int2(m, std::max(nlev_n[m] - 1, 0)));
                }
                out.ent.push_back(make_int4(e.x, local[m], e.z, m));
            }
        }
        out.work_off.push_back(items);
        out.max_rows = std::max(out.max_rows, nrows);
        out.row_off.push_back((int)out.rows.size());
        ++tile;
    }
    out.ntiles = tile;
    return true;
}


// ---- warp-item tiles -------------------------------------------------------------------------------
namespace {
struct StagedRow {
    unsigned goff;   // global element offset of the row
    int soff;        // byte offset inside its staging region
    int bytes;       // bytes copied
};
inline int pad16(int b) { return (b + 15) & ~15; }
inline int row_bytes(int levels) { return ((std::max(levels, 0) + 1) & ~1) * 8; }
}   // namespace

// Append the level-pair slots of tile-local node `ln` to a warp-item schedule of `len` lanes.  A
// column that does not fit the rest of the current item is cut, with a ghost slot on either side
// of the cut (phase A's stencil needs the neighbouring cluster bound); a cut is only made when at
// least two real slots fit in front of it.  `v` may be null (dry run for the footprint).
static void schedule_node(std::vector<unsigned short> *v, size_t &len, int ln, int s)
{
    auto put = [&](unsigned short d) {
        if (v) v->push_back(d);
        ++len;
    };
    int i = 0;
    while (i < s) {
        const int fill = (int)(len % 32);
        const int room = 32 - fill;
        const int lead = i > 0 ? 1 : 0;   // continuing a cut column: ghost of slot i-1 first
        if (s - i + lead <= room) {
            if (lead) put((unsigned short)(ln | ((i - 1) << 8) | 0x8000));
            for (; i < s; ++i) put((unsigned short)(ln | (i << 8)));
            break;
        }
        const int real = room - lead - 1;   // slots in front of the trailing ghost
        if (real < 2) {
            for (int k = 0; k < room; ++k) put((unsigned short)WT_IDLE);
            continue;
        }
        if (lead) put((unsigned short)(ln | ((i - 1) << 8) | 0x8000));
        for (int k = 0; k < real; ++k, ++i) put((unsigned short)(ln | (i << 8)));
        put((unsigned short)(ln | (i << 8) | 0x8000));
    }
}

// Column offsets (in doubles) of the packed level storage: node n owns [ncol[n], ncol[n+1]), an even
// number of slots that holds its nlev-1 active levels plus the bottom interface (fct_adf_v, area);
// edge g owns [ecol[g], ecol[g+1]), its active levels rounded up to even.
bool packed_columns(const DerivedHost &d, const int *nlev_n, int NT, int G, std::vector<unsigned> &ncol,
                    std::vector<unsigned> &ecol)
{
    // 32-bit element offsets: what the blobs, the copy lists and the kernels carry.  A mesh whose
    // packed node or edge array does not fit them is refused (the caller falls back / reports istat 1)
    const unsigned long long LIMIT = 0xffffffffull - 4096ull;
    unsigned long long acc = 0;
    ncol.assign((size_t)NT + 1, 0u);
    for (int n = 0; n < NT; ++n) {
        acc += (unsigned long long)((std::max(nlev_n[n] - 1, 0) + 2) & ~1);
        if (acc > LIMIT) return false;
        ncol[n + 1] = (unsigned)acc;
    }
    std::vector<int> depth((size_t)std::max(G, 1), 0);
    for (size_t k = 0; k < d.edg.size(); ++k) depth[d.edg[k].x] = FCT_META_DEPTH(d.edg[k].z);
    ecol.assign((size_t)G + 1, 0u);
    acc = 0;
    for (int g = 0; g < G; ++g) {
        acc += (unsigned long long)((depth[g] + 1) & ~1);
        if (acc > LIMIT) return false;
        ecol[g + 1] = (unsigned)acc;
    }
    return true;
}

bool build_warptiles(const DerivedHost &d, const int *nlev_n, int N, int NT, int G, int P, const unsigned *ncol,
                     const unsigned *ecol, const std::vector<int> *list, int TN, int smem_cap, WarpTilesHost &out)
{
    const int count = list ? (int)list->size() : N;
    const int W = 32;
    const bool packed = ncol && ecol;
    out = WarpTilesHost();
    out.blob_off.assign(1, 0u);
    if (P > 256 || (P & 1) || TN < 1 || TN > 255) return false;
    if (!packed && ((long long)NT * P >= (1LL << 32) || (long long)G * P >= (1LL << 32))) return false;
    if (d.nbr_off.size() != (size_t)N + 1) return false;
    const int slack = P * 8 + 32;   // masked lanes read up to one row past the last staged byte
    // global element offset / staged bytes of a node column and of an edge row.  Padded layout: the
    // active levels (even count); packed layout: the whole column slot, so that consecutive columns
    // are contiguous in global AND shared memory and their copies merge into one.
    auto ngoff = [&](int m) { return packed ? ncol[m] : (unsigned)((long long)m * P); };
    auto egoff = [&](int g) { return packed ? ecol[g] : (unsigned)((long long)g * P); };
    auto nbytes = [&](int m) { return packed ? (int)(ncol[m + 1] - ncol[m]) * 8 : row_bytes(nlev_n[m] - 1); };
    auto ebytes = [&](int g, int dg) { return packed ? (int)(ecol[g + 1] - ecol[g]) * 8 : row_bytes(dg); };

    std::vector<int> nstamp((size_t)NT, -1), nsoff((size_t)NT, 0), estamp((size_t)std::max(G, 1), -1), esoff((size_t)std::max(G, 1), 0);
    std::vector<int> tile_nodes, halo_nodes, tile_edges, edge_depth((size_t)std::max(G, 1), 0), fresh_n, fresh_e;
    std::vector<int4> hdr, ent;
    std::vector<unsigned short> sched;
    std::vector<std::pair<int, int>> a, b;
    std::vector<char> checked((size_t)N, 0);
    struct Copy {
        unsigned goff;
        int soff, bytes, arr;
    };
    std::vector<Copy> copies;

    auto need_bytes = [&](size_t nrows, size_t nerows, size_t nn, size_t nent, size_t nsched, int rb, int eb) {
        const size_t blob = WT_HDR_BYTES + pad16((int)(2 * nrows + nerows + WT_MAX_PREFETCH) * 8) + nn * 16 + nent * 16 +
                            pad16((int)((nsched + W - 1) / W * W) * 2);
        return 16 + blob + 2 * (size_t)rb + (size_t)eb + slack + 256;   // + the schedule's re-ordering margin
    };

    int pos = 0, tile = 0;
    while (pos < count) {
        // ---- 1. pick the nodes of the tile: as many as fit the stage ----
        tile_nodes.clear();
        halo_nodes.clear();
        tile_edges.clear();
        int rb = 0, eb = 0, nn = 0;
        size_t nent = 0, nsched = 0;
        while (pos < count && nn < TN) {
            const int n = list ? (*list)[pos] : pos;
            if (n < 0 || n >= N) return false;
            const int nz = std::max(nlev_n[n] - 1, 0);
            const int e0 = d.edg_off[n], e1 = d.edg_off[n + 1];
            const int cnt = e1 - e0;
            if (nz > 254 || cnt > 255) return false;
            if (!checked[n]) {
                // triangulation check: {ring neighbours} == {other ends of the edges}, same depths,
                // and no edge deeper than either of its end columns (what the reference's own
                // loops assume: reference.cpp:412-423 index the node rows down to the edge depth)
                const int nb0 = d.nbr_off[n], nb1 = d.nbr_off[n + 1];
                if (nb1 - nb0 < 1 || d.nbr[nb0].x != n) return false;
                a.clear();
                b.clear();
                for (int k = nb0 + 1; k < nb1; ++k) a.emplace_back(d.nbr[k].x, d.nbr[k].y);
                for (int k = e0; k < e1; ++k) {
                    const int m = d.edg[k].y, dg = FCT_META_DEPTH(d.edg[k].z);
                    if (m == n || m < 0 || m >= NT) return false;
                    if (dg > nz || dg > std::max(nlev_n[m] - 1, 0)) return false;
                    b.emplace_back(m, dg);
                }
                std::sort(a.begin(), a.end());
                std::sort(b.begin(), b.end());
                if (a != b) return false;
                for (size_t k = 1; k < b.size(); ++k)
                    if (b[k].first == b[k - 1].first) return false;
                if ((d.nbr[nb0].y) > nz) return false;
                checked[n] = 1;
            }
            // stamp what this node adds; undone if the node does not fit any more
            fresh_n.clear();
            fresh_e.clear();
            int add_rb = 0, add_eb = 0;
            auto see_node = [&](int m) {
                if (nstamp[m] != tile) {
                    nstamp[m] = tile;
                    fresh_n.push_back(m);
                    add_rb += nbytes(m);
                }
            };
            see_node(n);
            for (int k = e0; k < e1; ++k) {
                see_node(d.edg[k].y);
                const int g = d.edg[k].x;
                if (estamp[g] != tile) {
                    estamp[g] = tile;
                    fresh_e.push_back(g);
                    edge_depth[g] = FCT_META_DEPTH(d.edg[k].z);
                    add_eb += ebytes(g, edge_depth[g]);
                }
            }
            const int s = (nz + 1) / 2;
            size_t ns = nsched;
            schedule_node(nullptr, ns, nn, s);
            const size_t nrows = tile_nodes.size() + halo_nodes.size() + fresh_n.size();
            const size_t need = need_bytes(nrows, tile_edges.size() + fresh_e.size(), nn + 1, nent + cnt, ns, rb + add_rb, eb + add_eb);
            if ((int)need > smem_cap || nent + cnt > 65535) {
                for (int m : fresh_n) nstamp[m] = -1;
                for (int g : fresh_e) estamp[g] = -1;
                if (nn == 0) return false;
                break;
            }
            for (int m : fresh_n)
                if (m != n) halo_nodes.push_back(m);
            for (int g : fresh_e) tile_edges.push_back(g);
            tile_nodes.push_back(n);
            rb += add_rb;
            eb += add_eb;
            nent += cnt;
            nsched = ns;
            ++nn;
            ++pos;
        }
        // a node first seen as a neighbour may have joined the tile later: it is an own row
        {
            std::vector<int> keep;
            for (int m : halo_nodes) {
                bool own = false;
                if (m < N)
                    for (int t : tile_nodes)
                        if (t == m) {
                            own = true;
                            break;
                        }
                if (!own) keep.push_back(m);
            }
            halo_nodes.swap(keep);
        }
        // ---- 2. lay the stage out: own columns first, in node order, then the halo columns; edge
        //      rows in ascending edge id.  Runs that are contiguous in global memory are then also
        //      contiguous in shared memory and travel as one bulk copy. ----
        copies.clear();
        auto add_copy = [&](unsigned goff, int soff, int bytes, int arr) {
            if (bytes <= 0) return;
            if (!copies.empty()) {
                Copy &c = copies.back();
                if (c.arr == arr && c.goff + (unsigned)(c.bytes / 8) == goff && c.soff + c.bytes == soff && c.bytes + bytes <= (1 << 17)) {
                    c.bytes += bytes;
                    return;
                }
            }
            copies.push_back({goff, soff, bytes, arr});
        };
        int off = 0;
        std::sort(halo_nodes.begin(), halo_nodes.end());
        for (int pass = 0; pass < 2; ++pass)
            for (int m : (pass == 0 ? tile_nodes : halo_nodes)) {
                nsoff[m] = off;
                off += nbytes(m);
            }
        if (off != rb) return false;
        for (int arr = 0; arr < 2; ++arr)
            for (int pass = 0; pass < 2; ++pass)
                for (int m : (pass == 0 ? tile_nodes : halo_nodes)) add_copy(ngoff(m), nsoff[m] + arr * rb, nbytes(m), arr);
        std::sort(tile_edges.begin(), tile_edges.end());
        off = 0;
        for (int g : tile_edges) {
            esoff[g] = off;
            off += ebytes(g, edge_depth[g]);
        }
        if (off != eb) return false;
        for (int g : tile_edges) add_copy(egoff(g), esoff[g] + 2 * rb, ebytes(g, edge_depth[g]), 2);
        // ---- 3. node headers, edge entries, schedule ----
        hdr.clear();
        ent.clear();
        sched.clear();
        for (int i = 0; i < nn; ++i) {
            const int n = tile_nodes[i];
            const int nz = std::max(nlev_n[n] - 1, 0);
            const int e0 = d.edg_off[n], e1 = d.edg_off[n + 1];
            hdr.push_back(make_int4((int)ngoff(n), nz | (std::min(std::max(d.fillmin[n], 0), 255) << 8) | ((d.nbr[d.nbr_off[n]].y & 0xff) << 16),
                                    nsoff[n], (int)ent.size() | ((e1 - e0) << 16)));
            for (int k = e0; k < e1; ++k) {
                const int4 e = d.edg[k];
                const unsigned meta = (unsigned)FCT_META_DEPTH(e.z) | (FCT_META_WRITER(e.z) ? 0x40000000u : 0u) |
                                      (FCT_META_SECOND(e.z) ? 0x80000000u : 0u);
                ent.push_back(make_int4(esoff[e.x], nsoff[e.y], (int)meta, (int)egoff(e.x)));
            }
            out.slots += (nz + 1) / 2;
            out.edge_uses += e1 - e0;
        }
        {
            // the lanes of a warp item walk their nodes' edge lists in lockstep: schedule nodes of
            // equal degree next to each other so that no lane idles through another node's longer list
            std::vector<int> order((size_t)nn);
            for (int i = 0; i < nn; ++i) order[i] = i;
            std::stable_sort(order.begin(), order.end(), [&](int x, int y) {
                return d.edg_off[tile_nodes[x] + 1] - d.edg_off[tile_nodes[x]] > d.edg_off[tile_nodes[y] + 1] - d.edg_off[tile_nodes[y]];
            });
            for (int i : order) {
                size_t len = sched.size();
                schedule_node(&sched, len, i, (std::max(nlev_n[tile_nodes[i]] - 1, 0) + 1) / 2);
            }
        }
        sched.resize((sched.size() + W - 1) / W * W, (unsigned short)WT_IDLE);
        // ---- 4. assemble the blob ----
        // spread the copies over the issuer warps by size: largest first, dealt round-robin by the
        // strided loop of the kernel
        std::stable_sort(copies.begin(), copies.end(), [](const Copy &x, const Copy &y) { return x.bytes > y.bytes; });
        const int off_rows = WT_HDR_BYTES;
        const int n_copies = (int)copies.size();
        // L2 prefetch entries behind the copies: the runs of OWN columns of the tile (consecutive owned nodes are
        // contiguous in both layouts).  The kernels pull every array the consumers load straight from global memory
        // (fct_adf_v, area, the c-vertical operands, ...) into L2 over these ranges one tile ahead.
        std::vector<Copy> pfs;
        for (int i = 0; i < nn;) {
            int j = i;
            while (j + 1 < nn && tile_nodes[j + 1] == tile_nodes[j] + 1) ++j;
            const unsigned g0 = ngoff(tile_nodes[i]);
            long long bytes = (long long)(ngoff(tile_nodes[j]) - g0) * 8 + (packed ? nbytes(tile_nodes[j]) : P * 8);
            bytes = std::min<long long>(bytes, 0x3fff * 16);
            if ((int)pfs.size() < WT_MAX_PREFETCH && bytes >= 16) pfs.push_back({g0, 0, (int)(bytes & ~15LL), 3});
            i = j + 1;
        }
        const int n_pf = (int)pfs.size();
        const int off_hdr = off_rows + pad16((n_copies + n_pf) * 8);
        const int off_ent = off_hdr + (int)hdr.size() * 16;
        const int off_sched = off_ent + (int)ent.size() * 16;
        const int blob_bytes = off_sched + pad16((int)sched.size() * 2);
        long long tx = 0;
        for (auto &c : copies) tx += c.bytes;
        if (tx >= (1 << 20) || 2 * rb + eb >= (1 << 18)) return false;   // mbarrier transaction-count range, 14-bit offsets
        std::vector<unsigned char> buf((size_t)blob_bytes, 0);
        int *h = reinterpret_cast<int *>(buf.data());
        h[0] = n_copies;
        h[1] = (int)tile_edges.size();
        h[2] = nn;
        h[3] = (int)sched.size() / W;
        h[4] = (int)(tile_nodes.size() + halo_nodes.size());
        h[5] = off_hdr;
        h[6] = off_ent;
        h[7] = off_sched;
        h[8] = blob_bytes;
        h[9] = rb;
        h[10] = eb;
        h[11] = (int)tx;
        h[12] = n_pf;
        {
            int2 *t = reinterpret_cast<int2 *>(buf.data() + off_rows);
            for (int k = 0; k < n_copies; ++k)
                t[k] = make_int2((int)copies[k].goff, (copies[k].soff >> 4) | ((copies[k].bytes >> 4) << 14) | (copies[k].arr << 28));
            for (int k = 0; k < n_pf; ++k)
                t[n_copies + k] = make_int2((int)pfs[k].goff, ((pfs[k].bytes >> 4) << 14) | (3 << 28));
        }
        if (!hdr.empty()) std::memcpy(buf.data() + off_hdr, hdr.data(), hdr.size() * 16);
        if (!ent.empty()) std::memcpy(buf.data() + off_ent, ent.data(), ent.size() * 16);
        if (!sched.empty()) std::memcpy(buf.data() + off_sched, sched.data(), sched.size() * 2);
        const size_t at = out.blob.size();
        out.blob.resize(at + (size_t)blob_bytes / 16);
        std::memcpy(out.blob.data() + at, buf.data(), (size_t)blob_bytes);
        out.blob_off.push_back((unsigned)out.blob.size());
        out.smem_bytes = std::max(out.smem_bytes, 16 + blob_bytes + 2 * rb + eb + slack);
        out.nodes += nn;
        out.staged_rows += (long long)(tile_nodes.size() + halo_nodes.size());
        out.staged_erows += (long long)tile_edges.size();
        out.copies += n_copies;
        out.lanes += (long long)sched.size();
        ++tile;
    }
    out.ntiles = tile;
    return true;
}

}   // namespace fct

// Host-only introspection of the inspector (no CUDA device needed): builds the gather lists and the
// warp-item tile blobs of one node set and copies them out, so that the tables the kernels consume
// can be checked on a CPU box (tests/test_warp_plan.py interprets them with numpy).
extern "C" void fct_ale_plan_inspect_(int *myDim_nod2D, int *eDim_nod2D, int *myDim_elem2D, int *myDim_edge2D,
                                      int *nl, int *nlevels_nod2D, int *nlevels_elem2D, int *elem2D_nodes,
                                      int *nod_in_elem2D_num, int *nod_in_elem2D, int *nod_in_elem2D_dim,
                                      int *edges, int *edge_tri, int *tile_nodes, int *smem_cap,
                                      int *which, int *packed, long long *blob_capacity, unsigned *blob, int *tiles_capacity,
                                      unsigned *blob_off, int *ntiles, int *smem_bytes, int *istat)
{
    using namespace fct;
    *istat = 1;
    *ntiles = 0;
    *smem_bytes = 0;
    DerivedHost d;
    const int N = *myDim_nod2D, H = *eDim_nod2D;
    if (!build_derived(N, H, *myDim_elem2D, *myDim_edge2D, *nl, nlevels_elem2D, elem2D_nodes, nod_in_elem2D_num,
                       nod_in_elem2D, *nod_in_elem2D_dim, edges, edge_tri, d))
        return;
    const std::vector<int> *list = *which == 1 ? &d.boundary : (*which == 2 ? &d.interior : nullptr);
    WarpTilesHost h;
    const int P = (*nl + 7) & ~7;
    // *packed != 0: the packed level storage (columns hold their slots back to back)
    std::vector<unsigned> ncol, ecol;
    if (*packed && !packed_columns(d, nlevels_nod2D, N + H, *myDim_edge2D, ncol, ecol)) {
        *istat = 2;   // packed offsets would not fit 32 bits
        return;
    }
    if (!build_warptiles(d, nlevels_nod2D, N, N + H, *myDim_edge2D, P, *packed ? ncol.data() : nullptr,
                         *packed ? ecol.data() : nullptr, list, *tile_nodes, *smem_cap, h)) {
        *istat = 2;   // mesh not eligible
        return;
    }
    *ntiles = h.ntiles;
    *smem_bytes = h.smem_bytes;
    if ((long long)h.blob.size() * 4 > *blob_capacity || h.ntiles + 1 > *tiles_capacity) {
        *istat = 3;   // caller's buffers too small
        return;
    }
    std::memcpy(blob, h.blob.data(), h.blob.size() * 16);
    std::memcpy(blob_off, h.blob_off.data(), h.blob_off.size() * sizeof(unsigned));
    *istat = 0;
}
