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

}   // namespace fct
