/*
 * TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded CPU restatement of FESOM2's fct_ale limiter chain, used only as the
 * checker by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under fesom2-accelerate_b200/ may link, load or call it.
 *
 * Parity pinning:
 *   a1, a2, a3(+b1 vertical), a4 (= b1 horizontal + b2) are pinned bit-for-bit against the
 *   reference's own src/reference.cpp compiled unmodified into oracle/_ref/libref.so
 *   (tests/test_oracle.py).  b3 vertical/horizontal and c vertical/horizontal have NO working C++
 *   in the reference (src/reference.cpp:11-287 is a zero-stride skeleton); they restate the
 *   Fortran listing docs/refactoring.md:204-263 and :292-314 and are pinned against golden vectors
 *   produced by the reference's numpy reference() functions (tests/golden/make_golden.py).
 *
 * Layout (src/reference.cpp:309-334, :396, :419, :431): L = nl-1; node fields [node*L + z];
 * fct_adf_v / area / area_inv [node*nl + z]; fct_adf_h [edge*L + z]; UV_rhs [(elem*L + z)*2 + k];
 * connectivity is 1-based int32.
 *
 * Build: gcc -O2 -ffp-contract=off (no FMA contraction, so the operation order written here is
 * the rounding order).
 */
#include <stddef.h>

/* std::max / std::min semantics of the reference (first argument wins ties, no NaN handling) */
static inline double pick_max(double a, double b) { return (a < b) ? b : a; }
static inline double pick_min(double a, double b) { return (b < a) ? b : a; }

/* a1: src/reference.cpp:306-319, docs/refactoring.md:47-52 */
void oracle_a1(int n_nodes, const int *nlev_nod, int nl, double *ttf_max, double *ttf_min,
               const double *lo, const double *ttf)
{
    const size_t L = (size_t)nl - 1;
    for (int n = 0; n < n_nodes; ++n) {
        const double *l = lo + n * L, *t = ttf + n * L;
        double *hi = ttf_max + n * L, *lw = ttf_min + n * L;
        const int nz = nlev_nod[n] - 1;
        for (int z = 0; z < nz; ++z) {
            hi[z] = pick_max(l[z], t[z]);
            lw[z] = pick_min(l[z], t[z]);
        }
    }
}

/* a2: src/reference.cpp:321-351, docs/refactoring.md:58-70 */
void oracle_a2(int n_elem, const int *nlev_elem, int nl, double *uv_rhs, const int *elem_nodes,
               const double *ttf_max, const double *ttf_min, double bignumber)
{
    const size_t L = (size_t)nl - 1;
    for (int e = 0; e < n_elem; ++e) {
        const double *mx0 = ttf_max + (size_t)(elem_nodes[3 * e + 0] - 1) * L;
        const double *mx1 = ttf_max + (size_t)(elem_nodes[3 * e + 1] - 1) * L;
        const double *mx2 = ttf_max + (size_t)(elem_nodes[3 * e + 2] - 1) * L;
        const double *mn0 = ttf_min + (size_t)(elem_nodes[3 * e + 0] - 1) * L;
        const double *mn1 = ttf_min + (size_t)(elem_nodes[3 * e + 1] - 1) * L;
        const double *mn2 = ttf_min + (size_t)(elem_nodes[3 * e + 2] - 1) * L;
        double *uv = uv_rhs + (size_t)e * L * 2;
        const int nz = nlev_elem[e] - 1;
        int z = 0;
        for (; z < nz; ++z) {
            uv[2 * z] = pick_max(pick_max(mx0[z], mx1[z]), mx2[z]);
            uv[2 * z + 1] = pick_min(pick_min(mn0[z], mn1[z]), mn2[z]);
        }
        /* nlevels(elem) <= nl-1: levels nlevels(elem)..nl-1 (1-based) take the neutral bounds */
        if (nlev_elem[e] <= nl - 1) {
            for (z = nlev_elem[e] - 1; z < (int)L; ++z) {
                uv[2 * z] = -bignumber;
                uv[2 * z + 1] = bignumber;
            }
        }
    }
}

/* a3 (vlimit == 1), bounds part only: src/reference.cpp:353-392, docs/refactoring.md:77-108.
 * scratch: 2*(nl-1) doubles. */
void oracle_a3(int n_nodes, const int *nlev_nod, int nl, double *ttf_max, double *ttf_min,
               const double *lo, const double *uv_rhs, const int *nod_in_elem, const int *nod_in_elem_num,
               int ring_dim, double *scratch)
{
    const size_t L = (size_t)nl - 1;
    double *tv_max = scratch, *tv_min = scratch + L;
    for (int n = 0; n < n_nodes; ++n) {
        const int nz = nlev_nod[n] - 1;
        const int *ring = nod_in_elem + (size_t)n * ring_dim;
        const int cnt = nod_in_elem_num[n];
        const double *uv = uv_rhs + (size_t)(ring[0] - 1) * L * 2;
        for (int z = 0; z < nz; ++z) {
            tv_max[z] = uv[2 * z];
            tv_min[z] = uv[2 * z + 1];
        }
        for (int k = 1; k < cnt; ++k) {
            uv = uv_rhs + (size_t)(ring[k] - 1) * L * 2;
            for (int z = 0; z < nz; ++z) {
                tv_max[z] = pick_max(tv_max[z], uv[2 * z]);
                tv_min[z] = pick_min(tv_min[z], uv[2 * z + 1]);
            }
        }
        const double *l = lo + n * L;
        double *hi = ttf_max + n * L, *lw = ttf_min + n * L;
        hi[0] = tv_max[0] - l[0];
        lw[0] = tv_min[0] - l[0];
        for (int z = 1; z < nz - 1; ++z) {
            hi[z] = pick_max(pick_max(tv_max[z - 1], tv_max[z]), tv_max[z + 1]) - l[z];
            lw[z] = pick_min(pick_min(tv_min[z - 1], tv_min[z]), tv_min[z + 1]) - l[z];
        }
        hi[nz - 1] = tv_max[nz - 1] - l[nz - 1];
        lw[nz - 1] = tv_min[nz - 1] - l[nz - 1];
    }
}

/* b1 vertical: src/reference.cpp:393-399, docs/refactoring.md:156-169 (assigns: the zeroing loop
 * of md:156-161 is folded in, as in the reference) */
void oracle_b1_vertical(int n_nodes, const int *nlev_nod, int nl, double *plus, double *minus,
                        const double *adf_v)
{
    const size_t L = (size_t)nl - 1;
    for (int n = 0; n < n_nodes; ++n) {
        const double *v = adf_v + (size_t)n * nl;
        double *p = plus + n * L, *m = minus + n * L;
        const int nz = nlev_nod[n] - 1;
        for (int z = 0; z < nz; ++z) {
            p[z] = pick_max(0., v[z]) + pick_max(0., -v[z + 1]);
            m[z] = pick_min(0., v[z]) + pick_min(0., -v[z + 1]);
        }
    }
}

static inline int edge_depth(const int *edge_tri, const int *nlev_elem, int g)
{
    const int el = edge_tri[2 * g] - 1, er = edge_tri[2 * g + 1] - 1;
    const int d1 = nlev_elem[el] - 1;
    const int d2 = (er >= 0) ? nlev_elem[er] - 1 : 0;
    return d1 > d2 ? d1 : d2;
}

/* b1 horizontal: src/reference.cpp:406-425, docs/refactoring.md:172-186 */
void oracle_b1_horizontal(int n_edges, int nl, const int *nlev_elem, const int *edges,
                          const int *edge_tri, const double *adf_h, double *plus, double *minus)
{
    const size_t L = (size_t)nl - 1;
    for (int g = 0; g < n_edges; ++g) {
        const size_t a = (size_t)(edges[2 * g] - 1) * L, b = (size_t)(edges[2 * g + 1] - 1) * L;
        const double *h = adf_h + (size_t)g * L;
        const int nz = edge_depth(edge_tri, nlev_elem, g);
        for (int z = 0; z < nz; ++z) {
            const double f = h[z];
            plus[a + z] += pick_max(0., f);
            minus[a + z] += pick_min(0., f);
            plus[b + z] += pick_max(0., -f);
            minus[b + z] += pick_min(0., -f);
        }
    }
}

/* b2: src/reference.cpp:426-437 (multiplies by area_inv; the Fortran md:190-197 divides by area) */
void oracle_b2(int n_nodes, const int *nlev_nod, int nl, double *plus, double *minus,
               const double *ttf_max, const double *ttf_min, const double *area_inv, double dt,
               double flux_eps)
{
    const size_t L = (size_t)nl - 1;
    for (int n = 0; n < n_nodes; ++n) {
        const double *ai = area_inv + (size_t)n * nl;
        const int nz = nlev_nod[n] - 1;
        for (int z = 0; z < nz; ++z) {
            const size_t i = n * L + z;
            double flux = plus[i] * dt * ai[z] + flux_eps;
            plus[i] = pick_min(1., ttf_max[i] / flux);
            flux = minus[i] * dt * ai[z] - flux_eps;
            minus[i] = pick_min(1., ttf_min[i] / flux);
        }
    }
}

/* b3 vertical: docs/refactoring.md:205-233 (iter_yn = .false.); numpy twin
 * kernels/fct_ale_b3_vertical.py:163-181 (with the stride-nl correction for fct_adf_v) */
void oracle_b3_vertical(int n_nodes, const int *nlev_nod, int nl, double *adf_v, const double *plus,
                        const double *minus)
{
    const size_t L = (size_t)nl - 1;
    for (int n = 0; n < n_nodes; ++n) {
        double *v = adf_v + (size_t)n * nl;
        const double *p = plus + n * L, *m = minus + n * L;
        double ae = 1.;
        if (v[0] >= 0.) ae = pick_min(ae, p[0]);
        else ae = pick_min(ae, m[0]);
        v[0] = ae * v[0];
        const int nz = nlev_nod[n] - 1;
        for (int z = 1; z < nz; ++z) {
            ae = 1.;
            if (v[z] >= 0.) {
                ae = pick_min(ae, m[z - 1]);
                ae = pick_min(ae, p[z]);
            } else {
                ae = pick_min(ae, p[z - 1]);
                ae = pick_min(ae, m[z]);
            }
            v[z] = ae * v[z];
        }
        /* the bottom flux v[nz] stays as it is (md:232) */
    }
}

/* b3 horizontal: docs/refactoring.md:238-263; numpy twin kernels/fct_ale_b3_horizontal.py:76-101 */
void oracle_b3_horizontal(int n_edges, int nl, const int *nlev_elem, const int *edges,
                          const int *edge_tri, double *adf_h, const double *plus, const double *minus)
{
    const size_t L = (size_t)nl - 1;
    for (int g = 0; g < n_edges; ++g) {
        const size_t a = (size_t)(edges[2 * g] - 1) * L, b = (size_t)(edges[2 * g + 1] - 1) * L;
        double *h = adf_h + (size_t)g * L;
        const int nz = edge_depth(edge_tri, nlev_elem, g);
        for (int z = 0; z < nz; ++z) {
            double ae = 1.;
            if (h[z] >= 0.) {
                ae = pick_min(ae, plus[a + z]);
                ae = pick_min(ae, minus[b + z]);
            } else {
                ae = pick_min(ae, minus[a + z]);
                ae = pick_min(ae, plus[b + z]);
            }
            h[z] = ae * h[z];
        }
    }
}

/* c vertical: docs/refactoring.md:295-300.  The flux term is grouped x * (dt / area) exactly like
 * both executable statements of this stage in the reference -- kernels/fct_ale_c_vertical.cu:12
 * and the numpy twin kernels/fct_ale_c_vertical.py:41-44 -- so the golden vectors of the twin
 * are reproduced bit for bit.  The Fortran listing writes x*dt/area = (x*dt)/area, which differs
 * from this by at most a few ulp (well inside the 1e-12 relative bar of BASELINE.json). */
void oracle_c_vertical(int n_nodes, const int *nlev_nod, int nl, double *del_v, const double *ttf,
                       const double *hnode, const double *lo, const double *hnode_new,
                       const double *adf_v, const double *area, double dt)
{
    const size_t L = (size_t)nl - 1;
    for (int n = 0; n < n_nodes; ++n) {
        const double *v = adf_v + (size_t)n * nl, *ar = area + (size_t)n * nl;
        const int nz = nlev_nod[n] - 1;
        for (int z = 0; z < nz; ++z) {
            const size_t i = n * L + z;
            del_v[i] = del_v[i] - (ttf[i] * hnode[i]) + (lo[i] * hnode_new[i])
                       + ((v[z] - v[z + 1]) * (dt / ar[z]));
        }
    }
}

/* c horizontal: docs/refactoring.md:303-314; grouping h * (dt / area) as in
 * kernels/fct_ale_c_horizontal.cu:25-26 and the numpy twin kernels/fct_ale_c_horizontal.py:53-71 */
void oracle_c_horizontal(int n_edges, int nl, const int *nlev_elem, const int *edges,
                         const int *edge_tri, const double *adf_h, const double *area, double *del_h,
                         double dt)
{
    const size_t L = (size_t)nl - 1;
    for (int g = 0; g < n_edges; ++g) {
        const size_t n1 = (size_t)(edges[2 * g] - 1), n2 = (size_t)(edges[2 * g + 1] - 1);
        const double *h = adf_h + (size_t)g * L;
        const int nz = edge_depth(edge_tri, nlev_elem, g);
        for (int z = 0; z < nz; ++z) {
            del_h[n1 * L + z] = del_h[n1 * L + z] + (h[z] * (dt / area[n1 * nl + z]));
            del_h[n2 * L + z] = del_h[n2 * L + z] - (h[z] * (dt / area[n2 * nl + z]));
        }
    }
}

/* pre_comm composition: src/reference.cpp:289-304 (a1 over owned+halo, the rest over owned) */
void oracle_pre_comm(int my_nod, int e_nod, int my_elem, int my_edge, int nl, const int *nlev_nod,
                     const int *nlev_elem, const int *elem_nodes, const int *nod_in_elem_num,
                     const int *nod_in_elem, int ring_dim, const int *edges, const int *edge_tri,
                     double *ttf_max, double *ttf_min, double *plus, double *minus, const double *ttf,
                     const double *lo, const double *adf_v, const double *adf_h, double *uv_rhs,
                     const double *area_inv, double flux_eps, double bignumber, double dt,
                     double *scratch)
{
    oracle_a1(my_nod + e_nod, nlev_nod, nl, ttf_max, ttf_min, lo, ttf);
    oracle_a2(my_elem, nlev_elem, nl, uv_rhs, elem_nodes, ttf_max, ttf_min, bignumber);
    oracle_a3(my_nod, nlev_nod, nl, ttf_max, ttf_min, lo, uv_rhs, nod_in_elem, nod_in_elem_num,
              ring_dim, scratch);
    oracle_b1_vertical(my_nod, nlev_nod, nl, plus, minus, adf_v);
    oracle_b1_horizontal(my_edge, nl, nlev_elem, edges, edge_tri, adf_h, plus, minus);
    oracle_b2(my_nod, nlev_nod, nl, plus, minus, ttf_max, ttf_min, area_inv, dt, flux_eps);
}

/* everything after the halo exchange of fct_plus / fct_minus: md:204-314 with iter_yn = .false. */
void oracle_post_comm(int my_nod, int my_edge, int nl, const int *nlev_nod, const int *nlev_elem,
                      const int *edges, const int *edge_tri, const double *plus, const double *minus,
                      double *adf_v, double *adf_h, const double *ttf, const double *lo,
                      const double *hnode, const double *hnode_new, const double *area,
                      double *del_v, double *del_h, double dt)
{
    oracle_b3_vertical(my_nod, nlev_nod, nl, adf_v, plus, minus);
    oracle_b3_horizontal(my_edge, nl, nlev_elem, edges, edge_tri, adf_h, plus, minus);
    oracle_c_vertical(my_nod, nlev_nod, nl, del_v, ttf, hnode, lo, hnode_new, adf_v, area, dt);
    oracle_c_horizontal(my_edge, nl, nlev_elem, edges, edge_tri, adf_h, area, del_h, dt);
}

/* ------------------------------------------------------------------------------------------------
 * the branches of the Fortran listing the reference never made
 * executable (src/reference.cpp:51-96 are TODO stubs, kernels/fct_ale_a3.py:152-155 is `pass`).
 * PARITY UNPINNED: no golden vector, test or runnable reference code exists for them; these
 * functions follow docs/refactoring.md line by line, quirks included.
 * ------------------------------------------------------------------------------------------------ */

/* a3 with vlimit == 2 (docs/refactoring.md:113-129) or vlimit == 3 (md:131-148).  On entry
 * ttf_max / ttf_min hold the a1 bounds.  The listing takes BOTH the maxval and the minval of the
 * vertical neighbourhood from fct_ttf_max (md:120-121, md:139-140); restated as written.
 * scratch: 2*(nl-1) doubles. */
void oracle_a3_vlimit(int vlimit, int n_nodes, const int *nlev_nod, int nl, double *ttf_max,
                      double *ttf_min, const double *lo, const double *uv_rhs, const int *nod_in_elem,
                      const int *nod_in_elem_num, int ring_dim, double *scratch)
{
    const size_t L = (size_t)nl - 1;
    double *tv_max = scratch, *tv_min = scratch + L;
    for (int n = 0; n < n_nodes; ++n) {
        const int nz = nlev_nod[n] - 1;
        const int *ring = nod_in_elem + (size_t)n * ring_dim;
        const int cnt = nod_in_elem_num[n];
        const double *uv = uv_rhs + (size_t)(ring[0] - 1) * L * 2;
        for (int z = 0; z < nz; ++z) {
            tv_max[z] = uv[2 * z];
            tv_min[z] = uv[2 * z + 1];
        }
        for (int k = 1; k < cnt; ++k) {
            uv = uv_rhs + (size_t)(ring[k] - 1) * L * 2;
            for (int z = 0; z < nz; ++z) {
                tv_max[z] = pick_max(tv_max[z], uv[2 * z]);
                tv_min[z] = pick_min(tv_min[z], uv[2 * z + 1]);
            }
        }
        double *hi = ttf_max + n * L, *lw = ttf_min + n * L;
        for (int z = 1; z < nz - 1; ++z) {   /* nz = 2 .. nlevels_nod2D(n)-2, 1-based */
            const double vmax = pick_max(pick_max(hi[z - 1], hi[z]), hi[z + 1]);
            const double vmin = pick_min(pick_min(hi[z - 1], hi[z]), hi[z + 1]);
            if (vlimit == 2) {
                tv_max[z] = pick_max(tv_max[z], vmax);
                tv_min[z] = pick_min(tv_min[z], vmin);
            } else {
                tv_max[z] = pick_min(tv_max[z], vmax);
                tv_min[z] = pick_max(tv_min[z], vmin);
            }
        }
        const double *l = lo + n * L;
        for (int z = 0; z < nz; ++z) {
            hi[z] = tv_max[z] - l[z];
            lw[z] = tv_min[z] - l[z];
        }
    }
}

/* b3 vertical with iter_yn (md:205-233): the rejected part of every flux below the surface goes
 * to adf_v2 (md:228-230; the surface level nz = 1 has no such statement, adf_v2 keeps its value). */
void oracle_b3_vertical_iter(int n_nodes, const int *nlev_nod, int nl, double *adf_v, double *adf_v2,
                             const double *plus, const double *minus)
{
    const size_t L = (size_t)nl - 1;
    for (int n = 0; n < n_nodes; ++n) {
        double *v = adf_v + (size_t)n * nl, *v2 = adf_v2 + (size_t)n * nl;
        const double *p = plus + n * L, *m = minus + n * L;
        double ae = 1.;
        if (v[0] >= 0.) ae = pick_min(ae, p[0]);
        else ae = pick_min(ae, m[0]);
        v[0] = ae * v[0];
        const int nz = nlev_nod[n] - 1;
        for (int z = 1; z < nz; ++z) {
            ae = 1.;
            if (v[z] >= 0.) {
                ae = pick_min(ae, m[z - 1]);
                ae = pick_min(ae, p[z]);
            } else {
                ae = pick_min(ae, p[z - 1]);
                ae = pick_min(ae, m[z]);
            }
            v2[z] = (1.0 - ae) * v[z];
            v[z] = ae * v[z];
        }
    }
}

/* b3 horizontal with iter_yn (md:238-263) */
void oracle_b3_horizontal_iter(int n_edges, int nl, const int *nlev_elem, const int *edges,
                               const int *edge_tri, double *adf_h, double *adf_h2, const double *plus,
                               const double *minus)
{
    const size_t L = (size_t)nl - 1;
    for (int g = 0; g < n_edges; ++g) {
        const size_t a = (size_t)(edges[2 * g] - 1) * L, b = (size_t)(edges[2 * g + 1] - 1) * L;
        double *h = adf_h + (size_t)g * L, *h2 = adf_h2 + (size_t)g * L;
        const int nz = edge_depth(edge_tri, nlev_elem, g);
        for (int z = 0; z < nz; ++z) {
            double ae = 1.;
            if (h[z] >= 0.) {
                ae = pick_min(ae, plus[a + z]);
                ae = pick_min(ae, minus[b + z]);
            } else {
                ae = pick_min(ae, minus[a + z]);
                ae = pick_min(ae, plus[b + z]);
            }
            h2[z] = (1.0 - ae) * h[z];
            h[z] = ae * h[z];
        }
    }
}

/* "c. Update the LO" of the iterative branch (md:265-287), operation order of the listing:
 * x*dt/area/hnode_new = ((x*dt)/area)/hnode_new.  Every end node of a local edge is updated,
 * halo nodes included, as the edge loop of the listing does. */
void oracle_lo_update(int n_nodes, int n_edges, int nl, const int *nlev_nod, const int *nlev_elem,
                      const int *edges, const int *edge_tri, double *lo, const double *adf_v,
                      const double *adf_h, const double *area, const double *hnode_new, double dt)
{
    const size_t L = (size_t)nl - 1;
    for (int n = 0; n < n_nodes; ++n) {
        const double *v = adf_v + (size_t)n * nl, *ar = area + (size_t)n * nl;
        const int nz = nlev_nod[n] - 1;
        for (int z = 0; z < nz; ++z) {
            const size_t i = n * L + z;
            lo[i] = lo[i] + (v[z] - v[z + 1]) * dt / ar[z] / hnode_new[i];
        }
    }
    for (int g = 0; g < n_edges; ++g) {
        const size_t n1 = (size_t)(edges[2 * g] - 1), n2 = (size_t)(edges[2 * g + 1] - 1);
        const double *h = adf_h + (size_t)g * L;
        const int nz = edge_depth(edge_tri, nlev_elem, g);
        for (int z = 0; z < nz; ++z) {
            lo[n1 * L + z] = lo[n1 * L + z] + h[z] * dt / area[n1 * nl + z] / hnode_new[n1 * L + z];
            lo[n2 * L + z] = lo[n2 * L + z] - h[z] * dt / area[n2 * nl + z] / hnode_new[n2 * L + z];
        }
    }
}

/* ------------------------------------------------------------------------------------------------
 * stress2rhs: src/reference.cpp:440-480, docs/refactoring.md:409-461.
 * Same index expressions as the reference: 0-based elem2D_nodes with element stride
 * elem2D_nodes_size (reference.cpp:457) and gradient_sca addressed "corner*6 + element"
 * (reference.cpp:460-461).  Pinned against the reference's own compiled function
 * (tests/test_oracle.py, tests/golden/ref_cpp_stress2rhs.npz).
 * ------------------------------------------------------------------------------------------------ */
void oracle_stress2rhs(int n_nodes, int n_elems, int en_size, double *u_rhs, double *v_rhs,
                       const double *ice_strength, const int *elem_nodes, const double *elem_area,
                       const double *sigma11, const double *sigma12, const double *sigma22,
                       const double *gradient_sca, const double *metric_factor,
                       const double *inv_areamass, const double *rhs_a, const double *rhs_m)
{
    const double one_third = 1.0 / 3.0;
    for (int n = 0; n < n_nodes; ++n) {
        u_rhs[n] = 0.0;
        v_rhs[n] = 0.0;
    }
    for (int e = 0; e < n_elems; ++e) {
        if (ice_strength[e] > 0.0) {
            for (int k = 0; k < 3; ++k) {
                const int node = elem_nodes[(size_t)k * en_size + e];
                u_rhs[node] -= elem_area[e] * ((sigma11[e] * gradient_sca[k * 6 + e])
                                               + (sigma12[e] * gradient_sca[(k + 3) * 6 + e])
                                               + (sigma12[e] * one_third * metric_factor[e]));
                v_rhs[node] -= elem_area[e] * ((sigma12[e] * gradient_sca[k * 6 + e])
                                               + (sigma22[e] * gradient_sca[(k + 3) * 6 + e])
                                               - (sigma11[e] * one_third * metric_factor[e]));
            }
        }
    }
    for (int n = 0; n < n_nodes; ++n) {
        if (inv_areamass[n] > 0.0) {
            u_rhs[n] = (u_rhs[n] * inv_areamass[n]) + rhs_a[n];
            v_rhs[n] = (v_rhs[n] * inv_areamass[n]) + rhs_m[n];
        } else {
            u_rhs[n] = 0.0;
            v_rhs[n] = 0.0;
        }
    }
}
