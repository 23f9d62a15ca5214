!! fft_b200.f90 -- drop-in FFT backend file for 2DECOMP&FFT: defines `module decomp_2d_fft` exactly
!! like src/fft_cufft.f90 does (one backend file = one definition of the module, selected in
!! src/CMakeLists.txt:36-53), but every device action goes through libd2dfft_b200.so.
!!
!! SOURCE ONLY (no Fortran compiler in this image).  What changes against src/fft_cufft.f90:
!!   - the 12 cuFFT plans, the cuFFT work area and wk2_c2c / wk2_r2c / wk13 (fft_cufft.f90:37-61,
!!     263-431) become ONE opaque plan handle; the library owns its stage buffers;
!!   - fft_3d_c2c / fft_3d_r2c / fft_3d_c2r (fft_cufft.f90:676-790, 795-934, 939-1170) become one
!!     call each; stage sequencing, transposes and exchanges happen inside the library;
!!   - no per-call alloc_x(wk1) / deallocate (fft_cufft.f90:699, 745, 964, 1026).
!! Everything in fft_common.f90 (decomp_2d_fft_init overloads, engines, get_size ...) is reused as is.
module decomp_2d_fft

   use decomp_2d
   use decomp_2d_constants
   use decomp_2d_mpi
   use decomp_2d_profiler
   use iso_c_binding
   use d2d_b200_capi
   use d2d_b200_state, only: d2d_ctx          ! the context created by decomp_2d_init (see decomp_2d_b200.f90)

   implicit none

   private

   integer, parameter, public :: D2D_FFT_BACKEND = D2D_FFT_BACKEND_CUFFT   ! reuse the GPU backend id

   type decomp_2d_fft_engine
      type(c_ptr), private :: plan = c_null_ptr        ! d2d_fft_plan*
      integer, private :: format
      logical, private :: initialised = .false.
      integer, private :: nx_fft, ny_fft, nz_fft
      type(decomp_info), pointer, public :: ph => null()
      type(decomp_info), private :: ph_target
      type(decomp_info), public :: sp
      logical, private :: inplace
      logical, private :: skip_x_c2c, skip_y_c2c, skip_z_c2c
   contains
      procedure, public :: init => decomp_2d_fft_engine_init
      procedure, public :: fin => decomp_2d_fft_engine_fin
      procedure, public :: use_it => decomp_2d_fft_engine_use_it
      generic, public :: fft => c2c, r2c, c2r
      procedure, private :: c2c => decomp_2d_fft_engine_fft_c2c
      procedure, private :: r2c => decomp_2d_fft_engine_fft_r2c
      procedure, private :: c2r => decomp_2d_fft_engine_fft_c2r
   end type decomp_2d_fft_engine

   type(c_ptr), save :: cur_plan = c_null_ptr

#include "fft_common.f90"

   subroutine init_fft_engine(engine)
      type(decomp_2d_fft_engine), target, intent(inout) :: engine
      integer(c_int) :: skip(3), dtype
      skip = 0
      if (engine%skip_x_c2c) skip(1) = 1
      if (engine%skip_y_c2c) skip(2) = 1
      if (engine%skip_z_c2c) skip(3) = 1
#ifdef DOUBLE_PREC
      dtype = D2D_F64
#else
      dtype = D2D_F32
#endif
      call d2d_check(d2d_fft_plan_create(d2d_ctx, int(engine%format, c_int), engine%nx_fft, engine%ny_fft, &
                                         engine%nz_fft, dtype, merge(1_c_int, 0_c_int, engine%inplace), skip, &
                                         engine%plan), __FILE__, __LINE__)
      call decomp_2d_fft_log("d2d_b200")
   end subroutine init_fft_engine

   subroutine finalize_fft_engine(engine)
      type(decomp_2d_fft_engine), optional :: engine
      if (present(engine)) then
         if (c_associated(engine%plan)) call d2d_check(d2d_fft_plan_destroy(engine%plan), __FILE__, __LINE__)
         engine%plan = c_null_ptr
      else
         cur_plan = c_null_ptr
      end if
   end subroutine finalize_fft_engine

   subroutine use_fft_engine(engine)
      type(decomp_2d_fft_engine), target, intent(in) :: engine
      cur_plan = engine%plan
   end subroutine use_fft_engine

   ! in / out live on the device (the examples wrap the calls in `!$acc data copyin(in) copy(out)`)
   subroutine fft_3d_c2c(in, out, isign)
      complex(mytype), dimension(:, :, :), intent(INOUT), target :: in
      complex(mytype), dimension(:, :, :), intent(OUT), target :: out
      integer, intent(IN) :: isign
      if (decomp_profiler_fft) call decomp_profiler_start("fft_c2c")
      !$acc host_data use_device(in, out)
      call d2d_check(d2d_fft_3d_c2c(cur_plan, c_loc(in), c_loc(out), int(isign, c_int)), __FILE__, __LINE__)
      !$acc end host_data
      if (decomp_profiler_fft) call decomp_profiler_end("fft_c2c")
   end subroutine fft_3d_c2c

   subroutine fft_3d_r2c(in_r, out_c)
      real(mytype), dimension(:, :, :), intent(IN), target :: in_r
      complex(mytype), dimension(:, :, :), intent(OUT), target :: out_c
      if (decomp_profiler_fft) call decomp_profiler_start("fft_r2c")
      !$acc host_data use_device(in_r, out_c)
      call d2d_check(d2d_fft_3d_r2c(cur_plan, c_loc(in_r), c_loc(out_c)), __FILE__, __LINE__)
      !$acc end host_data
      if (decomp_profiler_fft) call decomp_profiler_end("fft_r2c")
   end subroutine fft_3d_r2c

   subroutine fft_3d_c2r(in_c, out_r)
      complex(mytype), dimension(:, :, :), intent(INOUT), target :: in_c
      real(mytype), dimension(:, :, :), intent(OUT), target :: out_r
      if (decomp_profiler_fft) call decomp_profiler_start("fft_c2r")
      !$acc host_data use_device(in_c, out_r)
      call d2d_check(d2d_fft_3d_c2r(cur_plan, c_loc(in_c), c_loc(out_r)), __FILE__, __LINE__)
      !$acc end host_data
      if (decomp_profiler_fft) call decomp_profiler_end("fft_c2r")
   end subroutine fft_3d_c2r

end module decomp_2d_fft
