!! d2d_b200_capi.f90 -- ISO_C_BINDING interfaces of libd2dfft_b200.so (include/d2d_b200.h).
!!
!! SOURCE ONLY: this image has no Fortran compiler, so this file has never been compiled here.
!! It is the binding a 2DECOMP&FFT maintainer adds to src/ (see INTEGRATION.md); every interface is
!! a 1:1 transcription of a prototype of include/d2d_b200.h (same order, same argument meaning).
!! Device arrays are passed as type(c_ptr) obtained with c_devloc (CUDA Fortran) or, under OpenACC,
!! with `!$acc host_data use_device(a)` + c_loc(a).
module d2d_b200_capi

   use iso_c_binding

   implicit none

   public

   integer(c_int), parameter :: D2D_F32 = 0, D2D_F64 = 1
   integer(c_int), parameter :: D2D_X_TO_Y = 0, D2D_Y_TO_Z = 1, D2D_Z_TO_Y = 2, D2D_Y_TO_X = 3
   integer(c_int), parameter :: D2D_MEMCPY_H2D = 1, D2D_MEMCPY_D2H = 2, D2D_MEMCPY_D2D = 3

   interface

      ! ---- communicator / context (replaces decomp_2d_nccl_init / _fin, src/decomp_2d_nccl.f90:151-211)
      function d2d_get_unique_id(id) bind(C, name="d2d_get_unique_id") result(ierr)
         import :: c_int, c_signed_char
         integer(c_signed_char), intent(out) :: id(128)
         integer(c_int) :: ierr
      end function d2d_get_unique_id

      function d2d_ctx_create(ctx, id, nranks, rank, p_row, p_col, device) bind(C, name="d2d_ctx_create") result(ierr)
         import :: c_int, c_ptr, c_signed_char
         type(c_ptr), intent(out) :: ctx
         integer(c_signed_char), intent(in) :: id(128)
         integer(c_int), value :: nranks, rank, p_row, p_col, device
         integer(c_int) :: ierr
      end function d2d_ctx_create

      ! the same without NCCL: the library all-gathers its CUDA-IPC handles through a callback (MPI_Allgather on
      ! decomp_2d_comm, see d2d_b200_allgather in decomp_2d_b200.f90); the data plane is its own peer-memory exchange
      function d2d_ctx_create_bootstrap(ctx, nranks, rank, p_row, p_col, device, allgather, user) &
         bind(C, name="d2d_ctx_create_bootstrap") result(ierr)
         import :: c_int, c_ptr, c_funptr
         type(c_ptr), intent(out) :: ctx
         integer(c_int), value :: nranks, rank, p_row, p_col, device
         type(c_funptr), value :: allgather   ! int (*)(void *user, const void *send, void *recv, int64_t bytes)
         type(c_ptr), value :: user
         integer(c_int) :: ierr
      end function d2d_ctx_create_bootstrap

      ! EVEN builds (padded MPI_ALLTOALL, src/decomp_2d.f90:1186-1204): same pencils, padded buffer layout
      function d2d_ctx_set_even(ctx, even) bind(C, name="d2d_ctx_set_even") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: even
         integer(c_int) :: ierr
      end function d2d_ctx_set_even

      function d2d_decomp_even(decomp, x1count, y1count, y2count, z2count, even) bind(C, name="d2d_decomp_even") result(ierr)
         import :: c_int, c_ptr, c_int64_t
         type(c_ptr), value :: decomp
         integer(c_int64_t), intent(out) :: x1count, y1count, y2count, z2count
         integer(c_int), intent(out) :: even
         integer(c_int) :: ierr
      end function d2d_decomp_even

      ! update_halo + halo_exchange (src/halo.f90:101-198, 311-399): `out` is the pencil with `level` ghost layers
      function d2d_halo_update(ctx, decomp, pencil, level, dtype, is_complex, periodic, in, out) &
         bind(C, name="d2d_halo_update") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx, decomp
         integer(c_int), value :: pencil, level, dtype, is_complex
         integer(c_int), intent(in) :: periodic(3)
         type(c_ptr), value :: in, out
         integer(c_int) :: ierr
      end function d2d_halo_update

      ! decomp_2d_fft_3d on HOST arrays (upload, transform, download pipelined inside the library)
      function d2d_fft_3d_r2c_host(plan, in_r, out_c) bind(C, name="d2d_fft_3d_r2c_host") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: plan, in_r, out_c
         integer(c_int) :: ierr
      end function d2d_fft_3d_r2c_host

      function d2d_fft_3d_c2r_host(plan, in_c, out_r) bind(C, name="d2d_fft_3d_c2r_host") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: plan, in_c, out_r
         integer(c_int) :: ierr
      end function d2d_fft_3d_c2r_host

      function d2d_fft_3d_c2c_host(plan, in, out, isign) bind(C, name="d2d_fft_3d_c2c_host") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: plan, in, out
         integer(c_int), value :: isign
         integer(c_int) :: ierr
      end function d2d_fft_3d_c2c_host

      function d2d_ctx_destroy(ctx) bind(C, name="d2d_ctx_destroy") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int) :: ierr
      end function d2d_ctx_destroy

      function d2d_ctx_sync(ctx) bind(C, name="d2d_ctx_sync") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int) :: ierr
      end function d2d_ctx_sync

      function d2d_ctx_set_blocking(ctx, blocking) bind(C, name="d2d_ctx_set_blocking") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: blocking
         integer(c_int) :: ierr
      end function d2d_ctx_set_blocking

      function d2d_best_2d_grid(nproc, p_row, p_col) bind(C, name="d2d_best_2d_grid") result(ierr)
         import :: c_int
         integer(c_int), value :: nproc
         integer(c_int), intent(out) :: p_row, p_col
         integer(c_int) :: ierr
      end function d2d_best_2d_grid

      ! ---- decomposition (decomp_info_init, src/decomp_2d.f90:382-490)
      function d2d_decomp_create(ctx, nx, ny, nz, decomp) bind(C, name="d2d_decomp_create") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: nx, ny, nz
         type(c_ptr), intent(out) :: decomp
         integer(c_int) :: ierr
      end function d2d_decomp_create

      function d2d_decomp_destroy(decomp) bind(C, name="d2d_decomp_destroy") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: decomp
         integer(c_int) :: ierr
      end function d2d_decomp_destroy

      function d2d_decomp_query(decomp, xst, xen, xsz, yst, yen, ysz, zst, zen, zsz) &
         bind(C, name="d2d_decomp_query") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: decomp
         integer(c_int), intent(out), dimension(3) :: xst, xen, xsz, yst, yen, ysz, zst, zen, zsz
         integer(c_int) :: ierr
      end function d2d_decomp_query

      ! ---- transposes (src/transpose_*.f90, long variants; direction = D2D_X_TO_Y ...)
      function d2d_transpose(ctx, decomp, direction, dtype, is_complex, src, dst) &
         bind(C, name="d2d_transpose") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx, decomp
         integer(c_int), value :: direction, dtype, is_complex
         type(c_ptr), value :: src, dst      ! device pointers
         integer(c_int) :: ierr
      end function d2d_transpose

      ! ---- FFT plans and 3-D transforms (src/fft_cufft.f90:263-483, 676-1170)
      function d2d_fft_plan_create(ctx, fmt, nx, ny, nz, dtype, inplace, skip, plan) &
         bind(C, name="d2d_fft_plan_create") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ctx
         integer(c_int), value :: fmt, nx, ny, nz, dtype, inplace
         integer(c_int), intent(in) :: skip(3)
         type(c_ptr), intent(out) :: plan
         integer(c_int) :: ierr
      end function d2d_fft_plan_create

      function d2d_fft_plan_destroy(plan) bind(C, name="d2d_fft_plan_destroy") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: plan
         integer(c_int) :: ierr
      end function d2d_fft_plan_destroy

      function d2d_fft_3d_c2c(plan, in, out, isign) bind(C, name="d2d_fft_3d_c2c") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: plan, in, out
         integer(c_int), value :: isign
         integer(c_int) :: ierr
      end function d2d_fft_3d_c2c

      function d2d_fft_3d_r2c(plan, in_r, out_c) bind(C, name="d2d_fft_3d_r2c") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: plan, in_r, out_c
         integer(c_int) :: ierr
      end function d2d_fft_3d_r2c

      function d2d_fft_3d_c2r(plan, in_c, out_r) bind(C, name="d2d_fft_3d_c2r") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: plan, in_c, out_r
         integer(c_int) :: ierr
      end function d2d_fft_3d_c2r

      ! ---- memory (alloc_dev.f90, block_gpu.f90:90,144,224,229, decomp_pool.f90:202-270)
      function d2d_dev_alloc(ptr, bytes) bind(C, name="d2d_dev_alloc") result(ierr)
         import :: c_int, c_ptr, c_int64_t
         type(c_ptr), intent(out) :: ptr
         integer(c_int64_t), value :: bytes
         integer(c_int) :: ierr
      end function d2d_dev_alloc

      function d2d_dev_free(ptr) bind(C, name="d2d_dev_free") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ptr
         integer(c_int) :: ierr
      end function d2d_dev_free

      function d2d_host_alloc_pinned(ptr, bytes) bind(C, name="d2d_host_alloc_pinned") result(ierr)
         import :: c_int, c_ptr, c_int64_t
         type(c_ptr), intent(out) :: ptr
         integer(c_int64_t), value :: bytes
         integer(c_int) :: ierr
      end function d2d_host_alloc_pinned

      function d2d_host_free(ptr) bind(C, name="d2d_host_free") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), value :: ptr
         integer(c_int) :: ierr
      end function d2d_host_free

      function d2d_host_get_device_pointer(dev_ptr, host_ptr) bind(C, name="d2d_host_get_device_pointer") result(ierr)
         import :: c_int, c_ptr
         type(c_ptr), intent(out) :: dev_ptr
         type(c_ptr), value :: host_ptr
         integer(c_int) :: ierr
      end function d2d_host_get_device_pointer

      function d2d_memcpy(dst, src, bytes, kind) bind(C, name="d2d_memcpy") result(ierr)
         import :: c_int, c_ptr, c_int64_t
         type(c_ptr), value :: dst, src
         integer(c_int64_t), value :: bytes
         integer(c_int), value :: kind
         integer(c_int) :: ierr
      end function d2d_memcpy

      function d2d_last_error() bind(C, name="d2d_last_error") result(msg)
         import :: c_ptr
         type(c_ptr) :: msg     ! NUL-terminated; copy with c_f_pointer before the next library call
      end function d2d_last_error

   end interface

contains

   ! status /= 0  ->  decomp_2d_abort(file, line, status, message), the reference's error convention
   ! (src/decomp_2d_mpi.f90:145-191); the library itself never aborts or prints.
   subroutine d2d_check(ierr, file, line)
      use decomp_2d_mpi, only: decomp_2d_abort
      integer(c_int), intent(in) :: ierr
      character(len=*), intent(in) :: file
      integer, intent(in) :: line
      character(kind=c_char), pointer :: cmsg(:)
      character(len=512) :: msg
      integer :: i
      if (ierr == 0) return
      call c_f_pointer(d2d_last_error(), cmsg, [512])
      msg = ' '
      do i = 1, 512
         if (cmsg(i) == c_null_char) exit
         msg(i:i) = cmsg(i)
      end do
      call decomp_2d_abort(file, line, int(ierr), trim(msg))
   end subroutine d2d_check

end module d2d_b200_capi
