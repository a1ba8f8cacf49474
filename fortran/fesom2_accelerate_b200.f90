! ISO_C_BINDING interfaces of the NEW entry points of libfesom2-accelerate.so for B200
! (Part 2 of include/fesom2-accelerate.h).  The entry points the reference already had (Part 1:
! set_mpi_rank_ ... fct_ale_post_comm_acc_, /root/reference/include/fesom2-accelerate.h:146-158) keep
! their names and signatures, so the interface blocks of the ESiWACE-S1/fesom2 fork stay as they are.
!
! Conventions (the reference's): every argument by reference, opaque handles are type(c_ptr)
! variables whose ADDRESS is passed (C side: void**), status in istat (0 ok) / alg_state.
! SURVEY.md section 8(f) row 3.  No Fortran compiler exists in the build image of this repository:
! the module is checked against the header by tests/test_abi_symbols.py (every bind(C) name must be a
! symbol the header declares and the library exports), not by compilation.
module fesom2_accelerate_b200
  use iso_c_binding
  implicit none

  ! enum fct_field_id
  integer(c_int), parameter :: FCT_TTF = 0, FCT_LO = 1, FCT_ADF_V = 2, FCT_ADF_H = 3, FCT_AREA = 4,      &
       FCT_AREA_INV = 5, FCT_HNODE = 6, FCT_HNODE_NEW = 7, FCT_DEL_V = 8, FCT_DEL_H = 9,                 &
       FCT_TTF_MAX = 10, FCT_TTF_MIN = 11, FCT_PLUS = 12, FCT_MINUS = 13, FCT_UV_RHS = 14,               &
       FCT_ADF_H_OUT = 15, FCT_ADF_V_OUT = 16, FCT_ADF_V2 = 17, FCT_ADF_H2 = 18

  interface
    ! ---- stage c on the device (the reference leaves docs/refactoring.md:292-314 to the CPU) ----
    subroutine fct_ale_c_acc(alg_state, stream, del_ttf_advvert, del_ttf_advhoriz, ttf, fct_LO,      &
                             hnode, hnode_new, fct_adf_v, fct_adf_h, area, myDim_nod2D,              &
                             myDim_edge2D, nl, nlevels_nod2D, nlevels_elem2D, edges, edge_tri, dt)   &
                             bind(C, name="fct_ale_c_acc_")
      import :: c_int, c_ptr, c_double
      integer(c_int) :: alg_state, myDim_nod2D, myDim_edge2D, nl
      type(c_ptr)    :: stream, del_ttf_advvert, del_ttf_advhoriz, ttf, fct_LO, hnode, hnode_new,    &
                        fct_adf_v, fct_adf_h, area, nlevels_nod2D, nlevels_elem2D, edges, edge_tri
      real(c_double) :: dt
    end subroutine

    ! ---- handle ABI: downloads and release calls the reference never had ----
    subroutine transfer_var_back(mem, host) bind(C, name="transfer_var_back_")
      import :: c_ptr, c_double
      type(c_ptr) :: mem
      real(c_double) :: host(*)
    end subroutine
    subroutine transfer_var_back_async(mem, host, stream) bind(C, name="transfer_var_back_async_")
      import :: c_ptr, c_double
      type(c_ptr) :: mem, stream
      real(c_double) :: host(*)
    end subroutine
    subroutine free_var(mem, istat) bind(C, name="free_var_")
      import :: c_ptr, c_int
      type(c_ptr) :: mem
      integer(c_int) :: istat
    end subroutine
    subroutine free_pinned_doubles(hostptr, istat) bind(C, name="free_pinned_doubles_")
      import :: c_ptr, c_int
      type(c_ptr) :: hostptr
      integer(c_int) :: istat
    end subroutine
    subroutine free_stream(stream, istat) bind(C, name="free_stream_")
      import :: c_ptr, c_int
      type(c_ptr) :: stream
      integer(c_int) :: istat
    end subroutine
    ! 0: one kernel per reference stage; 1: fused phase kernels inside the *_comm_acc_ calls
    subroutine fct_ale_set_fused(fused) bind(C, name="fct_ale_set_fused_")
      import :: c_int
      integer(c_int) :: fused
    end subroutine

    ! ---- device-resident path: plan (once per mesh partition), fields, step ----
    subroutine fct_ale_plan_create(plan, myDim_nod2D, eDim_nod2D, myDim_elem2D, myDim_edge2D, nl,    &
                                   nlevels_nod2D, nlevels_elem2D, elem2D_nodes, nod_in_elem2D_num,   &
                                   nod_in_elem2D, nod_in_elem2D_dim, edges, edge_tri, istat)         &
                                   bind(C, name="fct_ale_plan_create_")
      import :: c_int, c_ptr
      type(c_ptr)    :: plan
      integer(c_int) :: myDim_nod2D, eDim_nod2D, myDim_elem2D, myDim_edge2D, nl, nod_in_elem2D_dim, istat
      integer(c_int) :: nlevels_nod2D(*), nlevels_elem2D(*), elem2D_nodes(*), nod_in_elem2D_num(*),  &
                        nod_in_elem2D(*), edges(*), edge_tri(*)
    end subroutine
    subroutine fct_ale_plan_destroy(plan, istat) bind(C, name="fct_ale_plan_destroy_")
      import :: c_int, c_ptr
      type(c_ptr) :: plan
      integer(c_int) :: istat
    end subroutine
    subroutine fct_ale_plan_kernels(plan, warp_tiles, staged_tiles, packed_tiles) bind(C, name="fct_ale_plan_kernels_")
      import :: c_int, c_ptr
      type(c_ptr) :: plan
      integer(c_int) :: warp_tiles, staged_tiles, packed_tiles
    end subroutine
    subroutine fct_ale_plan_pitch(plan, pitch) bind(C, name="fct_ale_plan_pitch_")
      import :: c_int, c_ptr
      type(c_ptr) :: plan
      integer(c_int) :: pitch
    end subroutine
    subroutine fct_ale_fields_create(fields, plan, ntracers, with_uv_rhs, istat) bind(C, name="fct_ale_fields_create_")
      import :: c_int, c_ptr
      type(c_ptr) :: fields, plan
      integer(c_int) :: ntracers, with_uv_rhs, istat
    end subroutine
    ! the fast path's own layout: active levels only, columns back to back (fct_ale_step mode 1 only)
    subroutine fct_ale_fields_create_packed(fields, plan, ntracers, istat) bind(C, name="fct_ale_fields_create_packed_")
      import :: c_int, c_ptr
      type(c_ptr) :: fields, plan
      integer(c_int) :: ntracers, istat
    end subroutine
    subroutine fct_ale_fields_destroy(fields, istat) bind(C, name="fct_ale_fields_destroy_")
      import :: c_int, c_ptr
      type(c_ptr) :: fields
      integer(c_int) :: istat
    end subroutine
    ! host arrays in the model's own layout, e.g. ttf(nl-1, myDim_nod2D+eDim_nod2D); tracer is 0-based
    subroutine fct_ale_field_upload(fields, field, tracer, host, stream, istat) bind(C, name="fct_ale_field_upload_")
      import :: c_int, c_ptr, c_double
      type(c_ptr) :: fields, stream
      integer(c_int) :: field, tracer, istat
      real(c_double) :: host(*)
    end subroutine
    subroutine fct_ale_field_download(fields, field, tracer, host, stream, istat) bind(C, name="fct_ale_field_download_")
      import :: c_int, c_ptr, c_double
      type(c_ptr) :: fields, stream
      integer(c_int) :: field, tracer, istat
      real(c_double) :: host(*)
    end subroutine
    subroutine fct_ale_field_link_bytes(fields, field, host, upload, bytes) bind(C, name="fct_ale_field_link_bytes_")
      import :: c_int, c_ptr, c_double, c_long_long
      type(c_ptr) :: fields
      integer(c_int) :: field, upload
      real(c_double) :: host(*)
      integer(c_long_long) :: bytes
    end subroutine
    ! one whole fct_ale step a1..c, all tracers; halo exchange over NVLink when halo /= c_null_ptr
    ! (pass a type(c_ptr) variable holding c_null_ptr for "no halo")
    subroutine fct_ale_step(fields, halo, stream, mode, dt, flux_eps, bignumber, alg_state) bind(C, name="fct_ale_step_")
      import :: c_int, c_ptr, c_double
      type(c_ptr) :: fields, halo, stream
      integer(c_int) :: mode, alg_state
      real(c_double) :: dt, flux_eps, bignumber
    end subroutine
    ! the subroutine with its vlimit (1, 2, 3) and iter_yn (0 / 1) branches, docs/refactoring.md:13-315
    subroutine fct_ale_step_general(fields, halo, stream, vlimit, iter_yn, dt, flux_eps, bignumber, alg_state) &
                                    bind(C, name="fct_ale_step_general_")
      import :: c_int, c_ptr, c_double
      type(c_ptr) :: fields, halo, stream
      integer(c_int) :: vlimit, iter_yn, alg_state
      real(c_double) :: dt, flux_eps, bignumber
    end subroutine
    subroutine fct_ale_stage(fields, stream, stage, dt, flux_eps, bignumber, istat) bind(C, name="fct_ale_stage_")
      import :: c_int, c_ptr, c_double
      type(c_ptr) :: fields, stream
      integer(c_int) :: stage, istat
      real(c_double) :: dt, flux_eps, bignumber
    end subroutine

    ! ---- multi-GPU: NCCL communicator hand-over and the halo of fct_plus / fct_minus ----
    ! rank 0 creates the id, MPI_Bcast(id128, 128, MPI_BYTE, 0, MPI_COMM_FESOM) hands it to the others
    subroutine fct_ale_comm_unique_id(id128, istat) bind(C, name="fct_ale_comm_unique_id_")
      import :: c_char, c_int
      character(kind=c_char) :: id128(128)
      integer(c_int) :: istat
    end subroutine
    ! send_nodes = com_nod2D%slist - 1; recv_first / recv_counts follow com_nod2D%rptr (halo nodes are
    ! grouped by owner in FESOM2's numbering, which lets the receives land in place)
    subroutine fct_ale_halo_create(halo, plan, id128, rank, nranks, npeers, peer_ranks, send_counts, &
                                   send_nodes, recv_first, recv_counts, istat) bind(C, name="fct_ale_halo_create_")
      import :: c_char, c_int, c_ptr
      type(c_ptr) :: halo, plan
      character(kind=c_char) :: id128(128)
      integer(c_int) :: rank, nranks, npeers, istat
      integer(c_int) :: peer_ranks(*), send_counts(*), send_nodes(*), recv_first(*), recv_counts(*)
    end subroutine
    subroutine fct_ale_halo_destroy(halo, istat) bind(C, name="fct_ale_halo_destroy_")
      import :: c_int, c_ptr
      type(c_ptr) :: halo
      integer(c_int) :: istat
    end subroutine
    subroutine fct_ale_halo_exchange(fields, halo, stream, istat) bind(C, name="fct_ale_halo_exchange_")
      import :: c_int, c_ptr
      type(c_ptr) :: fields, halo, stream
      integer(c_int) :: istat
    end subroutine
    ! host arrays already in the packed level storage: one contiguous copy, no repack
    subroutine fct_ale_plan_packed_size(plan, node_doubles, edge_doubles, istat) bind(C, name="fct_ale_plan_packed_size_")
      import :: c_int, c_long_long, c_ptr
      type(c_ptr) :: plan
      integer(c_long_long) :: node_doubles, edge_doubles
      integer(c_int) :: istat
    end subroutine
    subroutine fct_ale_plan_packed_columns(plan, kind, columns, istat) bind(C, name="fct_ale_plan_packed_columns_")
      import :: c_int, c_ptr
      type(c_ptr) :: plan
      integer(c_int) :: kind, columns(*), istat
    end subroutine
    subroutine fct_ale_field_upload_packed(fields, field, tracer, host_packed, stream, istat) &
                                           bind(C, name="fct_ale_field_upload_packed_")
      import :: c_int, c_double, c_ptr
      type(c_ptr) :: fields, stream
      integer(c_int) :: field, tracer, istat
      real(c_double) :: host_packed(*)
    end subroutine
    subroutine fct_ale_field_download_packed(fields, field, tracer, host_packed, stream, istat) &
                                             bind(C, name="fct_ale_field_download_packed_")
      import :: c_int, c_double, c_ptr
      type(c_ptr) :: fields, stream
      integer(c_int) :: field, tracer, istat
      real(c_double) :: host_packed(*)
    end subroutine
    subroutine fct_ale_trace_read(stamps, capacity, slots, istat) bind(C, name="fct_ale_trace_read_")
      import :: c_int, c_long_long
      integer(c_long_long) :: stamps(*)
      integer(c_int) :: capacity, slots, istat
    end subroutine
    ! device time (ms) of the exchange inside the last overlapped step that used this halo
    subroutine fct_ale_halo_comm_ms(halo, ms, istat) bind(C, name="fct_ale_halo_comm_ms_")
      import :: c_int, c_double, c_ptr
      type(c_ptr) :: halo
      real(c_double) :: ms
      integer(c_int) :: istat
    end subroutine
    ! exchange_nod of one per-tracer node array of width nl-1 (FCT_LO, FCT_TTF, ...)
    subroutine fct_ale_halo_exchange_field(fields, halo, stream, field, istat) bind(C, name="fct_ale_halo_exchange_field_")
      import :: c_int, c_ptr
      type(c_ptr) :: fields, halo, stream
      integer(c_int) :: field, istat
    end subroutine

    ! ---- stress2rhs (EVP sea ice; replaces src/reference.cpp:440-480) ----
    ! elem2D_nodes: 0-based, (elem2D_nodes_size, 3) seen from Fortran, as the reference indexes it
    subroutine stress2rhs_plan_create(plan, myDim_nod2D, myDim_elem2D, elem2D_nodes_size, elem2D_nodes, istat) &
                                      bind(C, name="stress2rhs_plan_create_")
      import :: c_int, c_ptr
      type(c_ptr) :: plan
      integer(c_int) :: myDim_nod2D, myDim_elem2D, elem2D_nodes_size, istat
      integer(c_int) :: elem2D_nodes(*)
    end subroutine
    subroutine stress2rhs_plan_destroy(plan, istat) bind(C, name="stress2rhs_plan_destroy_")
      import :: c_int, c_ptr
      type(c_ptr) :: plan
      integer(c_int) :: istat
    end subroutine
    subroutine stress2rhs_acc(plan, stream, U_rhs_ice, V_rhs_ice, ice_strength, elem_area, sigma11, sigma12, &
                              sigma22, gradient_sca, metric_factor, inv_areamass, rhs_a, rhs_m, istat)        &
                              bind(C, name="stress2rhs_acc_")
      import :: c_int, c_ptr
      type(c_ptr) :: plan, stream, U_rhs_ice, V_rhs_ice, ice_strength, elem_area, sigma11, sigma12, sigma22, &
                     gradient_sca, metric_factor, inv_areamass, rhs_a, rhs_m
      integer(c_int) :: istat
    end subroutine
    subroutine stress2rhs_host(myDim_nod2D, myDim_elem2D, elem2D_nodes_size, U_rhs_ice, V_rhs_ice,          &
                               ice_strength, elem2D_nodes, elem_area, sigma11, sigma12, sigma22,             &
                               gradient_sca, metric_factor, inv_areamass, rhs_a, rhs_m, istat)               &
                               bind(C, name="stress2rhs_")
      import :: c_int, c_double
      integer(c_int) :: myDim_nod2D, myDim_elem2D, elem2D_nodes_size, istat
      integer(c_int) :: elem2D_nodes(*)
      real(c_double) :: U_rhs_ice(*), V_rhs_ice(*), ice_strength(*), elem_area(*), sigma11(*), sigma12(*),   &
                        sigma22(*), gradient_sca(*), metric_factor(*), inv_areamass(*), rhs_a(*), rhs_m(*)
    end subroutine

    ! ---- events, introspection, tuning ----
    subroutine fct_ale_event_create(event, istat) bind(C, name="fct_ale_event_create_")
      import :: c_int, c_ptr
      type(c_ptr) :: event
      integer(c_int) :: istat
    end subroutine
    subroutine fct_ale_event_record(event, stream, istat) bind(C, name="fct_ale_event_record_")
      import :: c_int, c_ptr
      type(c_ptr) :: event, stream
      integer(c_int) :: istat
    end subroutine
    subroutine fct_ale_stream_wait_event(stream, event, istat) bind(C, name="fct_ale_stream_wait_event_")
      import :: c_int, c_ptr
      type(c_ptr) :: stream, event
      integer(c_int) :: istat
    end subroutine
    subroutine fct_ale_event_elapsed_ms(start, stop, ms, istat) bind(C, name="fct_ale_event_elapsed_ms_")
      import :: c_int, c_ptr, c_double
      type(c_ptr) :: start, stop
      real(c_double) :: ms
      integer(c_int) :: istat
    end subroutine
    subroutine fct_ale_event_destroy(event, istat) bind(C, name="fct_ale_event_destroy_")
      import :: c_int, c_ptr
      type(c_ptr) :: event
      integer(c_int) :: istat
    end subroutine
    subroutine fct_ale_mem_info(free_bytes, total_bytes, istat) bind(C, name="fct_ale_mem_info_")
      import :: c_int, c_long_long
      integer(c_long_long) :: free_bytes, total_bytes
      integer(c_int) :: istat
    end subroutine
    subroutine fct_ale_device_info(name64, cc_major, cc_minor, sm_count, istat) bind(C, name="fct_ale_device_info_")
      import :: c_char, c_int
      character(kind=c_char) :: name64(64)
      integer(c_int) :: cc_major, cc_minor, sm_count, istat
    end subroutine
    subroutine fct_ale_launch_count(count) bind(C, name="fct_ale_launch_count_")
      import :: c_long_long
      integer(c_long_long) :: count
    end subroutine
    ! name: NUL-terminated, e.g. "WT_OPT"//c_null_char
    subroutine fct_ale_tune(name, value) bind(C, name="fct_ale_tune_")
      import :: c_char, c_int
      character(kind=c_char) :: name(*)
      integer(c_int) :: value
    end subroutine
  end interface

contains

  ! The body of FESOM2's fct_ale (docs/refactoring.md:13-315) for tracer fields that stay on the GPU:
  ! upload what changed on the host, one call for a1 .. c, download the advective tendencies.
  ! vlimit == 1 and no iteration: the fused fast path (mode 1); otherwise the general step.
  subroutine fct_ale_device(fields, halo, stream, tracer, ttf, fct_LO, fct_adf_v, fct_adf_h,         &
                            del_ttf_advvert, del_ttf_advhoriz, vlimit, iter_yn, dt, flux_eps,        &
                            bignumber, ok)
    type(c_ptr), intent(inout) :: fields, halo, stream
    integer(c_int), intent(in) :: tracer, vlimit
    logical, intent(in)        :: iter_yn
    real(c_double), intent(inout) :: ttf(*), fct_LO(*), fct_adf_v(*), fct_adf_h(*),                  &
                                     del_ttf_advvert(*), del_ttf_advhoriz(*)
    real(c_double), intent(in) :: dt, flux_eps, bignumber
    logical, intent(out)       :: ok
    integer(c_int) :: istat, alg_state, mode, it
    mode = 1
    it = merge(1, 0, iter_yn)
    call fct_ale_field_upload(fields, FCT_TTF, tracer, ttf, stream, istat)
    call fct_ale_field_upload(fields, FCT_LO, tracer, fct_LO, stream, istat)
    call fct_ale_field_upload(fields, FCT_ADF_V, tracer, fct_adf_v, stream, istat)
    call fct_ale_field_upload(fields, FCT_ADF_H, tracer, fct_adf_h, stream, istat)
    call fct_ale_field_upload(fields, FCT_DEL_V, tracer, del_ttf_advvert, stream, istat)
    call fct_ale_field_upload(fields, FCT_DEL_H, tracer, del_ttf_advhoriz, stream, istat)
    if (vlimit == 1 .and. .not. iter_yn) then
      call fct_ale_step(fields, halo, stream, mode, dt, flux_eps, bignumber, alg_state)
    else
      call fct_ale_step_general(fields, halo, stream, vlimit, it, dt, flux_eps, bignumber, alg_state)
    end if
    ok = alg_state == 10
    if (.not. ok) return
    if (iter_yn) then
      ! next pass: the updated low-order solution and the rejected flux parts (now fct_adf_*)
      call fct_ale_field_download(fields, FCT_LO, tracer, fct_LO, stream, istat)
      call fct_ale_field_download(fields, FCT_ADF_V, tracer, fct_adf_v, stream, istat)
      call fct_ale_field_download(fields, FCT_ADF_H, tracer, fct_adf_h, stream, istat)
    else
      call fct_ale_field_download(fields, FCT_DEL_V, tracer, del_ttf_advvert, stream, istat)
      call fct_ale_field_download(fields, FCT_DEL_H, tracer, del_ttf_advhoriz, stream, istat)
    end if
    ! await_stream_ (Part 1 of the header) before the host reads the arrays
  end subroutine fct_ale_device

end module fesom2_accelerate_b200
