/*
 * fesom2-accelerate.h -- C ABI of the B200-native fct_ale tracer limiter.
 *
 * Drop-in boundary.  Part 1 re-declares, binary-compatibly, every symbol FESOM2's Fortran binds
 * through ISO_C_BINDING in the reference library (each entry cites the reference interface it
 * replaces as <reference file>:<line>).  Part 2 is new ABI that the reference lacks (SURVEY.md
 * section 8b): stage c on the device, release calls, and a device-resident plan / fields / step /
 * halo interface used by the fused two-kernel path and by the multi-GPU runs.
 *
 * Conventions, identical to the reference (include/fesom2-accelerate.h:128-236 there):
 *   - extern "C", gfortran-style lower-case names with a trailing underscore;
 *   - every argument by pointer, scalars included; opaque handles travel as void** (the address
 *     of a Fortran type(c_ptr));
 *   - real_type is double; connectivity is 1-based int32;
 *   - status through int* istat (0 ok, 1 failed / fell back) and int* alg_state (last completed
 *     stage); errors are also printed on stderr; no call throws or aborts;
 *   - the *_acc_ calls are asynchronous on the caller's stream: await_stream_ before touching the
 *     host copies.
 *   - There is no CPU fallback anywhere: without a CUDA device every compute call reports failure.
 *
 * Layout of the arrays behind the handles (reference src/reference.cpp:309-334, :396, :419):
 *   L = nl-1.  node fields [node*L + z]; fct_adf_v, area, area_inv [node*nl + z];
 *   fct_adf_h [edge*L + z]; UV_rhs [(elem*L + z)*2 + {0:max,1:min}].
 */
#ifndef FESOM2_ACCELERATE_B200_H
#define FESOM2_ACCELERATE_B200_H

#include <stddef.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

typedef double real_type; /* reference include/fesom2-accelerate.h:10 */

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* Part 1: the reference's symbols                                                             */
/* ------------------------------------------------------------------------------------------ */

/* Leading members are the reference's struct gpuMemory (include/fesom2-accelerate.h:19-28), so
 * the legacy gpuMemory* entry points stay binary compatible; the tail is ours. */
struct gpuMemory {
    void *host_pointer;
    void *device_pointer;
    size_t size;      /* bytes */
    void *event;      /* cudaEvent_t; valid only when has_event != 0 */
    int has_event;
    int event_recorded;
    unsigned magic;
};

/* replaces set_mpi_rank_  (include/fesom2-accelerate.h:146, src/fesom2-accelerate.cu:206) */
void set_mpi_rank_(int *rank, int *total_ranks);
/* replaces transfer_mesh_ (include :147, src :114): allocate *size int32 and upload synchronously */
void transfer_mesh_(void **ret, int *host_ptr, int *size, int *istat);
/* replaces alloc_var_ (include :148, src :129): *size doubles bound to host_ptr, no upload */
void alloc_var_(void **ret, real_type *host_ptr, int *size, bool *create_event, int *istat);
/* replaces reserve_var_ (include :149, src :136): device-only buffer */
void reserve_var_(void **ret, int *size, bool *create_event, int *istat);
/* replaces allocate_pinned_doubles_ (include :150, src :143); falls back to malloc with istat=1 */
void allocate_pinned_doubles_(void **hostptr, int *size, int *istat);
/* replaces transfer_var_ (include :151, src :156): rebind host pointer, synchronous upload */
void transfer_var_(void **mem, real_type *host_ptr);
/* replaces transfer_var_async_ (include :152, src :163) */
void transfer_var_async_(void **mem, real_type *host_ptr, void **stream, bool *record_event);
/* replace make_stream_ / await_stream_ (include :153-154, src :170, :188) */
void make_stream_(void **stream, int *istat);
void await_stream_(void **s, int *istat);

/* replaces fct_ale_pre_comm_acc_ (include :156, src :258-340): uploads fct_LO, waits on the upload
 * events of ttf / fct_adf_v / fct_adf_h, runs a1..b2, downloads fct_plus / fct_minus
 * asynchronously; *alg_state = 6. */
void fct_ale_pre_comm_acc_(int *alg_state, void **s, void **fct_ttf_max, void **fct_ttf_min,
                           void **fct_plus, void **fct_minus, void **ttf, void **fct_LO,
                           void **fct_adf_v, void **fct_adf_h, void **UV_rhs, void **area_inv,
                           int *myDim_nod2D, int *eDim_nod2D, int *myDim_elem2D, int *myDim_edge2D,
                           int *nl, void **nlevels_nod2D, void **nlevels_elem2D, void **elem2D_nodes,
                           void **nod_in_elem2D_num, void **nod_in_elem2D, int *nod_in_elem2D_dim,
                           void **nod2D_edges, void **elem2D_edges, int *vlimit, real_type *flux_eps,
                           real_type *bignumber, real_type *dt);
/* replaces fct_ale_inter_comm_acc_ (include :157, src :342-356): b3 vertical, downloads fct_adf_v;
 * *alg_state = 7 */
void fct_ale_inter_comm_acc_(int *alg_state, void **s, void **fct_plus, void **fct_minus,
                             void **fct_adf_v, int *myDim_nod2D, int *nl, void **nlevels_nod2D);
/* replaces fct_ale_post_comm_acc_ (include :158, src :358-379): re-uploads the halo-updated
 * fct_plus / fct_minus, b3 horizontal, downloads fct_adf_h; *alg_state = 8.  The 8th argument is
 * the ELEMENT depth array, as in the reference's definition (src :369). */
void fct_ale_post_comm_acc_(int *alg_state, void **s, void **fct_plus, void **fct_minus,
                            void **fct_adf_h, int *myDim_edge2D, int *nl, void **nlevels_elem2D,
                            int *nod_in_elem2D_dim, void **nod2D_edges, void **elem2D_edges);

/* replace fct_ale_a1_accelerated / a2 / a1_a2 (include :173, :189, :209; src :42, :69, :91):
 * synchronous (or stream-async) copies around single launches.  The reference gives the trailing
 * two parameters C++ default values; C callers pass them explicitly. */
void fct_ale_a1_accelerated(const int maxLevels, const int nNodes, struct gpuMemory *nLevels_nod2D,
                            struct gpuMemory *fct_ttf_max, struct gpuMemory *fct_ttf_min,
                            struct gpuMemory *fct_low_order, struct gpuMemory *ttf, bool synchronous,
                            void *stream);
void fct_ale_a2_accelerated(const int maxLevels, const int nElements, struct gpuMemory *nLevels_elem,
                            struct gpuMemory *elementNodes, struct gpuMemory *UV_rhs,
                            struct gpuMemory *fct_ttf_max, struct gpuMemory *fct_ttf_min,
                            bool synchronous, void *stream);
void fct_ale_a1_a2_accelerated(const int maxLevels, const int nNodes, const int nElements,
                               struct gpuMemory *nLevels_nod2D, struct gpuMemory *nLevels_elem,
                               struct gpuMemory *elementNodes, struct gpuMemory *fct_ttf_max,
                               struct gpuMemory *fct_ttf_min, struct gpuMemory *fct_low_order,
                               struct gpuMemory *ttf, struct gpuMemory *UV_rhs, bool synchronous,
                               void *stream);

/* replace fct_ale_a{1,2,3,4}_reference_ and fct_ale_pre_comm_ (include :142, :225-235;
 * src/reference.cpp:289-438).  Same names, arguments (HOST arrays) and results, but computed on
 * the GPU: each call uploads its inputs, runs the corresponding stage kernels and downloads the
 * outputs synchronously.  a3 includes b1 vertical and a4 is b1 horizontal + b2, as there. */
void fct_ale_a1_reference_(int *nNodes, int *nLevels_nod2D, int *nl, real_type *fct_ttf_max,
                           real_type *fct_ttf_min, real_type *fct_low_order, real_type *ttf);
void fct_ale_a2_reference_(int *nElements, int *nLevels_elem2D, int *nl, real_type *UV_rhs,
                           int *elem2D_nodes, real_type *fct_ttf_max, real_type *fct_ttf_min,
                           real_type *bignumber);
void fct_ale_a3_reference_(int *nNodes2D, int *nLevels_nod2D, int *nl, real_type *fct_ttf_max,
                           real_type *fct_ttf_min, real_type *fct_LO, real_type *UV_rhs,
                           real_type *fct_plus, real_type *fct_minus, real_type *fct_adf_v,
                           int *nod_in_elem2D, int *nod_in_elem2D_num, int *nod_in_elem2D_dim);
void fct_ale_a4_reference_(int *nNodes2D, int *nLevels_nod2D, int *nLevels_elem2D, int *nl,
                           int *nEdges2D, real_type *fct_plus, real_type *fct_minus,
                           real_type *fct_adf_h, real_type *area_inv, real_type *fct_ttf_max,
                           real_type *fct_ttf_min, int *edges, int *edge_tri, real_type *flux_eps,
                           real_type *dt);
void fct_ale_pre_comm_(int *alg_state, real_type *fct_ttf_max, real_type *fct_ttf_min,
                       real_type *fct_plus, real_type *fct_minus, real_type *ttf, real_type *fct_LO,
                       real_type *fct_adf_v, real_type *fct_adf_h, real_type *UV_rhs,
                       real_type *area_inv, int *myDim_nod2D, int *eDim_nod2D, int *myDim_elem2D,
                       int *myDim_edge2D, int *nl, int *nlevels_nod2D, int *nlevels_elem2D,
                       int *elem2D_nodes, int *nod_in_elem2D_num, int *nod_in_elem2D,
                       int *nod_in_elem2D_dim, int *nod2D_edges, int *elem2D_edges, int *vlimit,
                       real_type *flux_eps, real_type *bignumber, real_type *dt);

/* ------------------------------------------------------------------------------------------ */
/* Part 2: new ABI (absent from the reference; SURVEY.md section 8b "New ABI")                  */
/* ------------------------------------------------------------------------------------------ */

/* Stage c on the device (the reference leaves docs/refactoring.md:292-314 to the Fortran CPU
 * code; its kernels/fct_ale_c_*.cu are built but never launched).  Handle-based like the calls
 * above; area / hnode / hnode_new / del_* are uploaded from their bound host pointers first,
 * del_ttf_advvert / del_ttf_advhoriz are downloaded asynchronously after; *alg_state = 10. */
void fct_ale_c_acc_(int *alg_state, void **s, void **del_ttf_advvert, void **del_ttf_advhoriz,
                    void **ttf, void **fct_LO, void **hnode, void **hnode_new, void **fct_adf_v,
                    void **fct_adf_h, void **area, int *myDim_nod2D, int *myDim_edge2D, int *nl,
                    void **nlevels_nod2D, void **nlevels_elem2D, void **nod2D_edges,
                    void **elem2D_edges, real_type *dt);

/* synchronous / asynchronous download of a variable into host_ptr (rebinds the host pointer) */
void transfer_var_back_(void **mem, real_type *host_ptr);
void transfer_var_back_async_(void **mem, real_type *host_ptr, void **stream);
/* release calls the reference never had */
void free_var_(void **mem, int *istat);
void free_pinned_doubles_(void **hostptr, int *istat);
void free_stream_(void **stream, int *istat);
/* 0: one kernel per reference stage (default, materialises UV_rhs); 1: fused phase kernels */
void fct_ale_set_fused_(int *fused);
/* tuning knob: same effect as the environment variable FCT_<name> (NUL-terminated name), e.g.
 * "TILE" 0/1, "TILE_NODES", "TILE_ITERS" (read when a plan is created), "TILE_VARIANT_A",
 * "TILE_VARIANT_B", "TILE_AHEAD" (read at every launch).  Warp-item kernels: "WT_STAGES", "WT_NODES",
 * "WT_SMEM" (plan), "WT_OPT" (bit mask of scheduling options), "WT_ISSUERS", "WT_WARPS_A", "WT_WARPS_B",
 * "WT_CONV" (phase A: -1 a1 folded into the edge loop = default, 2 / 3 / 4 converter warps, 0 consumers run
 * the a1 pass), "WT_REGS" (0 | 72 | 80 | 88: register re-allocation between the warp roles, setmaxnreg),
 * "WT_TRACE", "HALO_SKIP", "DIRECT_COPY" (launch).  Results never depend on them (measured alternatives,
 * DESIGN.md 3.3 / 3.4) -- except the timing experiments "WT_DIAG" and "HALO_SKIP", which skip data movement. */
void fct_ale_tune_(const char *name, int *value);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
void fct_ale_launch_count_(long long *count);
/* name and compute capability of the bound device; istat=1 when there is none */
void fct_ale_device_info_(char *name64, int *cc_major, int *cc_minor, int *sm_count, int *istat);

/* CUDA-event timing on the caller's stream (bench.py, the harness): create / record / elapsed */
void fct_ale_event_create_(void **event, int *istat);
void fct_ale_event_record_(void **event, void **stream, int *istat);
/* make *stream wait for *event (ordering between an upload stream and a download stream) */
void fct_ale_stream_wait_event_(void **stream, void **event, int *istat);
void fct_ale_event_elapsed_ms_(void **start, void **stop, real_type *ms, int *istat);
void fct_ale_event_destroy_(void **event, int *istat);
/* free / total bytes of device memory on the bound device */
void fct_ale_mem_info_(long long *free_bytes, long long *total_bytes, int *istat);

/* ---- device-resident path: plan (mesh + derived gather lists), fields, step ---------------- */

/* Build the per-mesh plan from HOST connectivity (1-based, as the Fortran holds it): uploads the
 * mesh, derives the node->neighbour list (a1+a2+a3 without UV_rhs), the node->edge gather list in
 * ascending edge order (deterministic replacement of the b1h / c_h atomics) and the
 * boundary / interior node split used to overlap the halo exchange. */
void fct_ale_plan_create_(void **plan, int *myDim_nod2D, int *eDim_nod2D, int *myDim_elem2D,
                          int *myDim_edge2D, int *nl, int *nlevels_nod2D, int *nlevels_elem2D,
                          int *elem2D_nodes, int *nod_in_elem2D_num, int *nod_in_elem2D,
                          int *nod_in_elem2D_dim, int *edges, int *edge_tri, int *istat);
void fct_ale_plan_destroy_(void **plan, int *istat);
/* Host-only introspection of the inspector (no CUDA device needed): the warp-item tile tables the
 * fused kernels consume for node set *which (0 all owned, 1 boundary, 2 interior) in the padded
 * (*packed == 0) or packed (*packed != 0) level storage, as 32-bit words
 * (layout: fesom2-accelerate_b200/csrc/fct_warp_kernels.cuh).  *istat: 0 ok, 1 malformed mesh,
 * 2 mesh not eligible for the warp-item kernels, 3 buffers too small. */
void fct_ale_plan_inspect_(int *myDim_nod2D, int *eDim_nod2D, int *myDim_elem2D, int *myDim_edge2D,
                           int *nl, int *nlevels_nod2D, int *nlevels_elem2D, int *elem2D_nodes,
                           int *nod_in_elem2D_num, int *nod_in_elem2D, int *nod_in_elem2D_dim,
                           int *edges, int *edge_tri, int *tile_nodes, int *smem_cap,
                           int *which, int *packed, long long *blob_capacity, unsigned *blob, int *tiles_capacity,
                           unsigned *blob_off, int *ntiles, int *smem_bytes, int *istat);
/* which fused kernels *mode 1 of fct_ale_step_ will run on this plan: the persistent TMA-staged
 * warp-item kernels (*warp_tiles), else the tile-staged ones (*staged_tiles), else the untiled;
 * *packed_tiles: fct_ale_fields_create_packed_ is available */
void fct_ale_plan_kernels_(void **plan, int *warp_tiles, int *staged_tiles, int *packed_tiles);
/* pitch (in doubles) of every padded device row of this plan: nl rounded up to a multiple of 8
 * (64 bytes), so that no DRAM sector is shared by two rows */
void fct_ale_plan_pitch_(void **plan, int *pitch);

/* Device arrays for a batch of *ntracers tracers on a plan, rows padded to the plan pitch.
 * *with_uv_rhs != 0 also allocates UV_rhs (needed by the staged mode only). */
void fct_ale_fields_create_(void **fields, void **plan, int *ntracers, int *with_uv_rhs, int *istat);
/* The same in the PACKED level storage, the layout of the fast path: every column holds only its
 * active levels (+ the bottom interface), columns back to back, so no DRAM granule is wasted on
 * the tail of a row, the footprint shrinks by the inactive share (about 30 %), and the own columns
 * of a tile travel as one bulk copy per array.  Packed fields support fct_ale_step_ mode 1 (plans
 * of plain triangulations), upload / download and the halo exchange; a download fills the levels
 * that have no slot with zeros. */
void fct_ale_fields_create_packed_(void **fields, void **plan, int *ntracers, int *istat);
void fct_ale_fields_destroy_(void **fields, int *istat);

/* Field identifiers for upload / download */
enum fct_field_id {
    FCT_TTF = 0, FCT_LO = 1, FCT_ADF_V = 2, FCT_ADF_H = 3, FCT_AREA = 4, FCT_AREA_INV = 5,
    FCT_HNODE = 6, FCT_HNODE_NEW = 7, FCT_DEL_V = 8, FCT_DEL_H = 9, FCT_TTF_MAX = 10,
    FCT_TTF_MIN = 11, FCT_PLUS = 12, FCT_MINUS = 13, FCT_UV_RHS = 14, FCT_ADF_H_OUT = 15,
    FCT_ADF_V_OUT = 16,
    /* rejected flux parts of the iterative branch (docs/refactoring.md:228-230, :258-260),
     * allocated at first use */
    FCT_ADF_V2 = 17, FCT_ADF_H2 = 18, FCT_FIELD_COUNT = 19
};
/* dense host array (the Fortran layout above) <-> padded device rows of tracer *tracer
 * (mesh-static fields area / area_inv / hnode / hnode_new ignore *tracer).  Asynchronous on
 * *stream; host memory should be pinned for the copy to overlap. */
/* The copy is one contiguous PCIe copy through a dense staging buffer plus a repack kernel.
 * Measured alternative for packed fields and PAGE-LOCKED host arrays (allocate_pinned_doubles_,
 * cudaHostRegister), tuning knob "DIRECT_COPY" 1: one kernel reads / writes the host array over
 * PCIe itself (mapped memory), so an upload moves only the levels that have a slot (about 70 % of
 * a dense array) and no staging buffer exists; SM-issued PCIe reads reach 35 GB/s against the copy
 * engine's 54 GB/s on B200, so it saves memory, not time. */
void fct_ale_field_upload_(void **fields, int *field, int *tracer, real_type *host, void **stream,
                           int *istat);
void fct_ale_field_download_(void **fields, int *field, int *tracer, real_type *host, void **stream,
                             int *istat);

/* Host arrays ALREADY in the packed level storage (fields of fct_ale_fields_create_packed_ only): a caller
 * that keeps its columns packed moves one contiguous block per array, no staging buffer, no repack kernel,
 * and only the slots that exist cross the link (about 70 % of the dense array).  Layout of a packed host
 * array: row r (node or edge) owns the doubles [col[r], col[r+1]) with col from
 * fct_ale_plan_packed_columns_ (*kind 0: nodes, [myDim_nod2D + eDim_nod2D + 1] entries; 1: edges,
 * [myDim_edge2D + 1]); level z of the row is element col[r] + z; the total length is
 * fct_ale_plan_packed_size_.  Node rows hold nlevels-1 active levels plus the bottom interface, rounded up
 * to an even count; edge rows their active levels rounded up to even. */
void fct_ale_plan_packed_size_(void **plan, long long *node_doubles, long long *edge_doubles, int *istat);
void fct_ale_plan_packed_columns_(void **plan, int *kind, unsigned *columns, int *istat);
void fct_ale_field_upload_packed_(void **fields, int *field, int *tracer, real_type *host_packed, void **stream,
                                  int *istat);
void fct_ale_field_download_packed_(void **fields, int *field, int *tracer, real_type *host_packed, void **stream,
                                    int *istat);

/* bytes that cross the host link when `field` is uploaded from (*upload != 0) or downloaded to
 * `host`: the dense array, or only the slots of the packed storage on the direct path */
void fct_ale_field_link_bytes_(void **fields, int *field, real_type *host, int *upload, long long *bytes);

/* One fct_ale step a1..c over all tracers of `fields`, everything resident on the device.
 *   *mode 0: ten stage kernels (one per reference kernel);  *mode 1: two fused phase kernels, the
 *   fastest variant the plan supports (TMA-staged warp-item kernels on plain triangulations, else
 *   the tile-staged, else the untiled ones);  *mode 2: the untiled fused kernels;  *mode 3: the
 *   tile-staged fused kernels (modes 2 and 3 are measured alternatives).
 * When `halo` is non-null the fct_plus / fct_minus halo exchange runs between b2 and b3
 * horizontal over NVLink, overlapped with the interior nodes' phase B work.
 * In the fused modes (1, 2, 3) the limited fluxes are written to the FCT_ADF_H_OUT / FCT_ADF_V_OUT
 * buffers (an in-place update would race with the neighbouring threads that still read the raw
 * values); mode 0 updates FCT_ADF_H / FCT_ADF_V in place like the reference.  *alg_state = 10 on
 * success. */
void fct_ale_step_(void **fields, void **halo, void **stream, int *mode, real_type *dt,
                   real_type *flux_eps, real_type *bignumber, int *alg_state);
/* The whole subroutine of docs/refactoring.md:13-315 with the branches the reference never made
 * executable (src/reference.cpp:51-96 are stubs): *vlimit 1, 2 or 3
 * (md:77-148) and *iter_yn (md:226-290: b3 keeps the rejected part of every flux in FCT_ADF_V2 /
 * FCT_ADF_H2, the limited fluxes update fct_LO, then fct_adf_* = fct_adf_*2 and -- with a halo --
 * the fct_LO halo rows are exchanged for the next pass).  Padded fields (created with UV_rhs): stage
 * kernels, in place like mode 0 of fct_ale_step_.  Packed fields: the fused fast path, two launches
 * like mode 1 (plain pass: limited fluxes in FCT_ADF_*_OUT), phase A in its vlimit variant and, for
 * an iterative pass, phase B in its iterative variant.  *alg_state = 10 on success. */
void fct_ale_step_general_(void **fields, void **halo, void **stream, int *vlimit, int *iter_yn, real_type *dt,
                            real_type *flux_eps, real_type *bignumber, int *alg_state);
/* single stage of the staged mode (per-stage ncu sweep): 0 a1, 1 a2, 2 a3, 3 b1v, 4 b1h, 5 b2,
 * 6 b3v, 7 b3h, 8 c_v, 9 c_h; 10 fused phase A, 11 fused phase B; 12 / 13 the tile-staged
 * fused phases (14-17: their boundary / interior subsets); 18 / 19 the warp-item fused phases
 * (20 / 21: phase A on the boundary / interior tiles, 22 / 23: phase B);
 * 24 / 25: b1h / c_h as the reference's edge-centric fp64 atomicAdd scatter
 * (kernels/fct_ale_b1_horizontal.cu:24-27, fct_ale_c_horizontal.cu:25-26) -- a measured
 * alternative only: not bit-reproducible, never launched by fct_ale_step_ or the *_acc_ calls;
 * 26 / 27: a3 with vlimit 2 / 3; 28 / 29: b3 vertical / horizontal with iter_yn; 30: the low-order
 * update of the iterative branch (docs/refactoring.md:265-287) */
void fct_ale_stage_(void **fields, void **stream, int *stage, real_type *dt, real_type *flux_eps,
                    real_type *bignumber, int *istat);

/* ---- multi-GPU halo exchange of fct_plus / fct_minus (docs/refactoring.md:200, :235) ------- */

/* 128-byte NCCL unique id, created on one rank and handed to the others by the caller (MPI in
 * FESOM2, torch.distributed's store in bench.py). */
void fct_ale_comm_unique_id_(char *id128, int *istat);
/* Per-rank halo descriptor.  send_nodes: concatenated 0-based local owned node ids, send_counts[p]
 * of them for peer_ranks[p], in the order of the receiver's halo numbering; the rows received from
 * peer p land in local nodes [recv_first[p], recv_first[p]+recv_counts[p]) (halo nodes are grouped
 * by owner).  Every id is validated (istat = 1): send nodes must be owned nodes with a halo
 * neighbour in the plan's mesh, receive ranges must lie inside the halo rows. */
void fct_ale_halo_create_(void **halo, void **plan, char *id128, int *rank, int *nranks, int *npeers,
                          int *peer_ranks, int *send_counts, int *send_nodes, int *recv_first,
                          int *recv_counts, int *istat);
void fct_ale_halo_destroy_(void **halo, int *istat);
/* Device time (ms, CUDA events on the halo's own stream) of the exchange inside the last
 * fct_ale_step_ / fct_ale_step_general_ that used this halo: pack kernel + grouped send / recv,
 * including the wait for the slowest peer.  Synchronises on that exchange.  Tuning knob
 * "HALO_SKIP" 1 runs the same launches WITHOUT the exchange (timing experiment, stale halo rows):
 * step time minus that = the part of the communication the overlap does not hide. */
void fct_ale_halo_comm_ms_(void **halo, real_type *ms, int *istat);
/* Profiling aid of the warp-item kernels (tuning knob "WT_TRACE" 1: phase A, 2: phase B): SM-clock stamps of
 * the pipeline events of CTA 0 of the last traced launch (*slots per tile iteration: stage seen empty, blob
 * issued, row copies started / issued, rows landed, a1 done, first / last consumer warp enters / leaves the
 * tile).  tools/trace_pipeline.py turns them into the per-tile refill and compute times. */
void fct_ale_trace_read_(long long *stamps, int *capacity, int *slots, int *istat);
/* the exchange alone (tests, timing) */
void fct_ale_halo_exchange_(void **fields, void **halo, void **stream, int *istat);
/* exchange_nod of ONE per-tracer node array of width nl-1 (*field: FCT_LO between two passes of
 * the iterative branch, FCT_TTF, ...): owned-boundary rows -> the neighbours' halo rows */
void fct_ale_halo_exchange_field_(void **fields, void **halo, void **stream, int *field, int *istat);

/* ---- stress2rhs: divergence of the sea-ice stress tensor into the rhs vectors ---------------- */
/* SURVEY.md section 8(f) row 4.  Replaces the reference's CPU restatement src/reference.cpp:440-480
 * (docs/refactoring.md:409-461; the reference has no GPU kernel for it): the element -> node
 * scatter runs as a deterministic node-centric gather in ascending element order, bit-identical to
 * the sequential loop.  Index conventions are reference.cpp's: elem2D_nodes 0-based with element
 * stride *elem2D_nodes_size ([3][size]), gradient_sca addressed corner*6 + element (at least
 * myDim_elem2D + 30 entries); nodes >= myDim_nod2D (halo corners) are not written. */
void stress2rhs_plan_create_(void **plan, int *myDim_nod2D, int *myDim_elem2D, int *elem2D_nodes_size,
                             int *elem2D_nodes, int *istat);
void stress2rhs_plan_destroy_(void **plan, int *istat);
/* device-resident: every array a handle of alloc_var_ / reserve_var_; asynchronous on *s */
void stress2rhs_acc_(void **plan, void **s, void **U_rhs_ice, void **V_rhs_ice, void **ice_strength,
                     void **elem_area, void **sigma11, void **sigma12, void **sigma22, void **gradient_sca,
                     void **metric_factor, void **inv_areamass, void **rhs_a, void **rhs_m, int *istat);
/* host arrays in, host arrays out (synchronous; argument order of reference.cpp:440) */
void stress2rhs_(int *myDim_nod2D, int *myDim_elem2D, int *elem2D_nodes_size, real_type *U_rhs_ice,
                 real_type *V_rhs_ice, real_type *ice_strength, int *elem2D_nodes, real_type *elem_area,
                 real_type *sigma11, real_type *sigma12, real_type *sigma22, real_type *gradient_sca,
                 real_type *metric_factor, real_type *inv_areamass, real_type *rhs_a, real_type *rhs_m, int *istat);

#ifdef __cplusplus
}
#endif
#endif /* FESOM2_ACCELERATE_B200_H */
