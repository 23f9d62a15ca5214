!! decomp_2d_b200.f90 -- the remaining touch points of the GPU build, as the maintainer would patch them.
!!
!! SOURCE ONLY.  (1) module d2d_b200_state: the library context, created where the reference calls
!! decomp_2d_nccl_init (src/decomp_2d_init_fin.f90:154-156) and destroyed where it calls
!! decomp_2d_nccl_fin (:214-216).  (2) the body that replaces the `#if defined(_GPU)` branches of the
!! four transpose workers (src/transpose_x_to_y.f90:71-135 and friends): one call, the library packs,
!! exchanges and unpacks.  work1_r_d / work2_r_d (src/decomp_2d_cumpi.f90) are no longer needed.
module d2d_b200_state

   use iso_c_binding
   use d2d_b200_capi
   use decomp_2d_mpi
   use mpi

   implicit none

   type(c_ptr), save, public :: d2d_ctx = c_null_ptr

contains

   ! replaces decomp_2d_nccl_init (src/decomp_2d_nccl.f90:151-193): rank 0 creates the unique id,
   ! MPI_Bcast distributes it (:185), every rank joins.  The device is the one bind.sh selected.
   subroutine d2d_b200_init(p_row, p_col, device)
      integer, intent(in) :: p_row, p_col, device
      integer(c_signed_char) :: id(128)
      integer :: ierror
      if (nrank == 0) call d2d_check(d2d_get_unique_id(id), __FILE__, __LINE__)
      call MPI_BCAST(id, 128, MPI_BYTE, 0, decomp_2d_comm, ierror)
      if (ierror /= 0) call decomp_2d_abort(__FILE__, __LINE__, ierror, "MPI_BCAST")
      call d2d_check(d2d_ctx_create(d2d_ctx, id, int(nproc, c_int), int(nrank, c_int), int(p_row, c_int), &
                                    int(p_col, c_int), int(device, c_int)), __FILE__, __LINE__)
   end subroutine d2d_b200_init

   ! NCCL-free variant: the library only needs an all-gather of a few hundred bytes per rank at context / plan creation
   ! (its CUDA-IPC handles); MPI_Allgather on decomp_2d_comm provides it -- the reference bootstraps NCCL through MPI the
   ! same way (src/decomp_2d_nccl.f90:181-191).  Build with -DD2D_BOOTSTRAP_MPI to select it.
   function d2d_b200_allgather(user, send, recv, nbytes) bind(C) result(ierr)
      type(c_ptr), value :: user, send, recv
      integer(c_int64_t), value :: nbytes
      integer(c_int) :: ierr
      integer(c_signed_char), pointer :: s(:), r(:)
      integer :: code
      call c_f_pointer(send, s, [nbytes])
      call c_f_pointer(recv, r, [nbytes * nproc])
      call MPI_ALLGATHER(s, int(nbytes), MPI_BYTE, r, int(nbytes), MPI_BYTE, decomp_2d_comm, code)
      ierr = int(code, c_int)
   end function d2d_b200_allgather

   subroutine d2d_b200_init_mpi(p_row, p_col, device)
      integer, intent(in) :: p_row, p_col, device
      call d2d_check(d2d_ctx_create_bootstrap(d2d_ctx, int(nproc, c_int), int(nrank, c_int), int(p_row, c_int), &
                                              int(p_col, c_int), int(device, c_int), c_funloc(d2d_b200_allgather), c_null_ptr), &
                     __FILE__, __LINE__)
#ifdef EVEN
      call d2d_check(d2d_ctx_set_even(d2d_ctx, 1_c_int), __FILE__, __LINE__)
#endif
   end subroutine d2d_b200_init_mpi

   ! replaces the body of halo_exchange_{real,complex} (src/halo.f90:311-399) on the device: `out` was allocated by
   ! update_halo with the halo extents (src/halo_common.f90:19-23); periodic_x/y/z are the module variables of decomp_2d_mpi
   subroutine d2d_b200_update_halo(handle, ipencil, level, is_complex, in, out)
      type(c_ptr), intent(in) :: handle, in, out
      integer, intent(in) :: ipencil, level
      logical, intent(in) :: is_complex
      integer(c_int) :: dtype, per(3)
#ifdef DOUBLE_PREC
      dtype = D2D_F64
#else
      dtype = D2D_F32
#endif
      per = [merge(1_c_int, 0_c_int, periodic_x), merge(1_c_int, 0_c_int, periodic_y), merge(1_c_int, 0_c_int, periodic_z)]
      call d2d_check(d2d_halo_update(d2d_ctx, handle, int(ipencil - 1, c_int), int(level, c_int), dtype, &
                                     merge(1_c_int, 0_c_int, is_complex), per, in, out), __FILE__, __LINE__)
   end subroutine d2d_b200_update_halo

   subroutine d2d_b200_fin()
      if (c_associated(d2d_ctx)) call d2d_check(d2d_ctx_destroy(d2d_ctx), __FILE__, __LINE__)
      d2d_ctx = c_null_ptr
   end subroutine d2d_b200_fin

   ! One routine serves the 16 workers transpose_{x_to_y,y_to_z,z_to_y,y_to_x}_{real,complex}
   ! (x both precisions); `handle` is the d2d_decomp* cached in the decomp_info (one new component
   ! `type(c_ptr) :: b200 = c_null_ptr`, filled by decomp_info_init with d2d_decomp_create and
   ! released by decomp_info_finalize with d2d_decomp_destroy).
   subroutine d2d_b200_transpose(direction, handle, is_complex, src, dst)
      integer(c_int), intent(in) :: direction
      type(c_ptr), intent(in) :: handle
      logical, intent(in) :: is_complex
      type(c_ptr), intent(in) :: src, dst       ! device addresses (c_devloc / host_data use_device)
      integer(c_int) :: dtype
#ifdef DOUBLE_PREC
      dtype = D2D_F64
#else
      dtype = D2D_F32
#endif
      call d2d_check(d2d_transpose(d2d_ctx, handle, direction, dtype, merge(1_c_int, 0_c_int, is_complex), src, dst), &
                     __FILE__, __LINE__)
   end subroutine d2d_b200_transpose

end module d2d_b200_state
