/*
 * d2d_b200.h -- C ABI of libd2dfft_b200.so: the B200-native (sm_100a) replacement of the GPU pieces
 * of 2DECOMP&FFT's distributed 3-D FFT hot path.
 *
 * The reference has no FFI layer: its GPU backend is CUDA-Fortran written inline.  The entry points
 * below are exactly the places where the reference touches `cudafor`, `cufft` and `nccl`; each
 * one cites the reference code it replaces (paths relative to the reference tree).  A Fortran
 * ISO_C_BINDING shim (2decomp-fft_b200/fortran/) keeps the public Fortran API
 * (decomp_2d_init, decomp_info, transpose_*, decomp_2d_fft_init/_3d/_finalize) on top of it.
 *
 * Conventions
 *   - every function returns an int status, 0 = success; the library never aborts or prints.  The
 *     caller maps non-zero to decomp_2d_abort (src/decomp_2d_mpi.f90:145-191).  d2d_last_error()
 *     gives the message of the last failure on the calling thread.
 *   - all array arguments are DEVICE pointers unless the name ends in _host; arrays are
 *     Fortran-ordered pencils with the extents given by the decomp (xsz / ysz / zsz).
 *   - sizes / counts / displacements are 64-bit (the reference uses default integers and
 *     overflows at 1024^3 on one rank).
 *   - one host thread drives one context (= one rank = one GPU), as in the reference
 *     (README.md:78-80).  Work is enqueued on the context's stream; with `blocking` on (default,
 *     like the reference's host-synchronous transposes, src/decomp_2d_nccl.f90:249) every call
 *     returns after the device finished.
 */
#ifndef D2D_B200_H
#define D2D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct d2d_ctx d2d_ctx;           /* rank context: device, streams, communicator, work buffers */
typedef struct d2d_group d2d_group;       /* in-process rank group (thread-per-rank transport)          */
typedef struct d2d_decomp d2d_decomp;     /* decomp_info (src/info.f90:19-47)                           */
typedef struct d2d_fft_plan d2d_fft_plan; /* decomp_2d_fft_engine (src/fft_cufft.f90:37-61)             */

/* src/decomp_2d_constants.f90:15-32 (mytype is a run-time choice here), :86-87, :92-93 */
enum { D2D_F32 = 0, D2D_F64 = 1 };
enum { D2D_FFT_FORWARD = -1, D2D_FFT_BACKWARD = 1 };
enum { D2D_PHYSICAL_IN_X = 1, D2D_PHYSICAL_IN_Z = 3 };
enum { D2D_X_TO_Y = 0, D2D_Y_TO_Z = 1, D2D_Z_TO_Y = 2, D2D_Y_TO_X = 3 };
enum { D2D_MEMCPY_H2D = 1, D2D_MEMCPY_D2H = 2, D2D_MEMCPY_D2D = 3 };
/* transports of the all-to-all */
enum { D2D_TRANSPORT_NONE = 0, D2D_TRANSPORT_NCCL = 1, D2D_TRANSPORT_LOCAL = 2, D2D_TRANSPORT_BOOT = 3 };

/* ---- communicator / context --------------------------------------------------------------------
 * replaces decomp_2d_nccl_init/_fin (src/decomp_2d_nccl.f90:151-211) and the rank/coord bookkeeping
 * of decomp_2d_init_ref (src/decomp_2d_init_fin.f90:95-123): rank r has coord (r / p_col, r % p_col);
 * the COL communicator (x<->y) joins the p_row ranks sharing coord(2), ROW (y<->z) the p_col ranks
 * sharing coord(1).  The unique id replaces the MPI_Bcast of ncclUniqueId (decomp_2d_nccl.f90:185). */
int d2d_get_unique_id(unsigned char id[128]);
int d2d_ctx_create(d2d_ctx **ctx, const unsigned char id[128], int nranks, int rank, int p_row, int p_col, int device);
/* The same without NCCL: the caller supplies a blocking all-gather of `bytes` host bytes per rank over the ranks of the
 * job (MPI_Allgather on decomp_2d_comm in the Fortran shim -- the reference bootstraps NCCL through MPI the same way,
 * src/decomp_2d_nccl.f90:181-191; torch.distributed in the Python mirror).  It is used at context / plan creation only
 * (exchange of CUDA-IPC handles); the data plane is the library's own peer-memory exchange (copy engines and peer stores
 * over NVLink with stream-ordered flags).  Returns non-zero from the callback to report failure. */
typedef int (*d2d_allgather_fn)(void *user, const void *send, void *recv, int64_t bytes);
int d2d_ctx_create_bootstrap(d2d_ctx **ctx, int nranks, int rank, int p_row, int p_col, int device, d2d_allgather_fn allgather, void *user);
/* thread-per-rank mode inside one process (any number of ranks per device): the exchange is a
 * device-to-device copy between the rank buffers.  Used where NCCL cannot run (several ranks on one
 * GPU) and for single-process multi-GPU drivers. */
int d2d_group_create(d2d_group **grp, int nranks);
int d2d_group_destroy(d2d_group *grp);
/* a rank of the group failed: every rank blocked in (or later entering) an exchange of this group returns an error instead
 * of waiting for it -- the in-process analogue of decomp_2d_abort -> MPI_ABORT (src/decomp_2d_mpi.f90:166-191) */
int d2d_group_abort(d2d_group *grp);
int d2d_ctx_create_in_group(d2d_ctx **ctx, d2d_group *grp, int rank, int p_row, int p_col, int device);
int d2d_ctx_destroy(d2d_ctx *ctx);
int d2d_ctx_sync(d2d_ctx *ctx);                    /* cudaStreamSynchronize of the context's streams */
int d2d_ctx_set_blocking(d2d_ctx *ctx, int blocking);
/* EVEN builds of the reference (-DEVEN: padded MPI_ALLTOALL instead of MPI_ALLTOALLV, src/transpose_*.f90 `#ifdef EVEN`
 * branches, counts from src/decomp_2d.f90:1186-1204): the bare transposes lay their send / receive buffers out with one
 * padded count per communicator (segment m at m * count) and exchange equal-sized messages.  The pencils that come out are
 * the same; what changes is the buffer footprint (src/decomp_2d.f90:443-446).  Default off, like the reference. */
int d2d_ctx_set_even(d2d_ctx *ctx, int even);
void *d2d_ctx_stream(d2d_ctx *ctx);                /* cudaStream_t */
int d2d_ctx_info(const d2d_ctx *ctx, int *nranks, int *rank, int dims[2], int coord[2], int *transport);
int64_t d2d_ctx_launch_count(const d2d_ctx *ctx);  /* kernels launched by this context so far */
/* per-stage device timers (CUDA events on the launching stream).  Labels follow the reference's
 * profiler regions (src/profiler_caliper.f90; "transp_x_y", "fft_r2c", ...) plus one per kernel. */
int d2d_ctx_profile(d2d_ctx *ctx, int enable);
int d2d_ctx_profile_count(d2d_ctx *ctx);
int d2d_ctx_profile_get(d2d_ctx *ctx, int i, char label[64], double *total_ms, int64_t *calls, double *bytes);
int d2d_ctx_profile_reset(d2d_ctx *ctx);
/* best_2d_grid (src/decomp_2d_init_fin.f90:270-300) */
int d2d_best_2d_grid(int nproc, int *p_row, int *p_col);

/* ---- decomposition -----------------------------------------------------------------------------
 * decomp_info_init / decomp_info_finalize (src/decomp_2d.f90:382-515), partition (:1016-1064),
 * distribute (:1070-1105), prepare_buffer (:1138-1183).  Starts/ends are 1-based like the reference. */
int d2d_decomp_create(d2d_ctx *ctx, int nx, int ny, int nz, d2d_decomp **decomp);
/* the same record for an arbitrary rank of a p_row x p_col grid; needs no device (host arithmetic only) */
int d2d_decomp_create_for_rank(int nx, int ny, int nz, int p_row, int p_col, int rank, d2d_decomp **decomp);
int d2d_decomp_destroy(d2d_decomp *decomp);
int d2d_decomp_query(const d2d_decomp *decomp, int xst[3], int xen[3], int xsz[3], int yst[3], int yen[3], int ysz[3],
                     int zst[3], int zen[3], int zsz[3]);
int d2d_decomp_dist(const d2d_decomp *decomp, int *x1dist, int *y1dist, int *y2dist, int *z2dist);
int d2d_decomp_counts(const d2d_decomp *decomp, int64_t *x1cnts, int64_t *y1cnts, int64_t *y2cnts, int64_t *z2cnts,
                      int64_t *x1disp, int64_t *y1disp, int64_t *y2disp, int64_t *z2disp);

/* x1count / y1count / y2count / z2count of an EVEN build (src/decomp_2d.f90:1197-1203) and decomp%even (:448-454) */
int d2d_decomp_even(const d2d_decomp *decomp, int64_t *x1count, int64_t *y1count, int64_t *y2count, int64_t *z2count, int *even);

/* ---- transposes --------------------------------------------------------------------------------
 * transpose_{x_to_y,y_to_z,z_to_y,y_to_x}_{real,complex} (src/transpose_*.f90 long variants):
 * pack -> all-to-all(v) -> unpack; dims==1 is a copy.  Bit-exact data movement. */
int d2d_transpose(d2d_ctx *ctx, const d2d_decomp *decomp, int direction, int dtype, int is_complex, const void *src, void *dst);
int d2d_transpose_x_to_y(d2d_ctx *ctx, const d2d_decomp *decomp, int dtype, int is_complex, const void *src, void *dst);
int d2d_transpose_y_to_z(d2d_ctx *ctx, const d2d_decomp *decomp, int dtype, int is_complex, const void *src, void *dst);
int d2d_transpose_z_to_y(d2d_ctx *ctx, const d2d_decomp *decomp, int dtype, int is_complex, const void *src, void *dst);
int d2d_transpose_y_to_x(d2d_ctx *ctx, const d2d_decomp *decomp, int dtype, int is_complex, const void *src, void *dst);

/* ---- halo cells ---------------------------------------------------------------------------------
 * update_halo (src/halo.f90:101-198, src/halo_common.f90) + halo_exchange (src/halo.f90:311-399,
 * src/halo_exchange_{x,y,z}_body.f90): `in` is a pencil of `decomp` (pencil = 0 X, 1 Y, 2 Z), `out` the same pencil with
 * `level` ghost layers on both sides of its two decomposed axes -- X: (n1, n2+2L, n3+2L), Y: (n1+2L, n2, n3+2L),
 * Z: (n1+2L, n2+2L, n3).  The interior is copied, then the ghost layers are filled from the neighbouring pencils in two
 * exchanges (the second one carries the corners).  periodic[3] = periodic_bc of decomp_2d_init (NULL: none); ghost layers
 * beyond a non-periodic boundary are left untouched.  Collective over the ranks of the context. */
int d2d_halo_update(d2d_ctx *ctx, const d2d_decomp *decomp, int pencil, int level, int dtype, int is_complex, const int periodic[3],
                    const void *in, void *out);

/* ---- FFT ---------------------------------------------------------------------------------------
 * plan = decomp_2d_fft_engine_init + init_fft_engine (src/fft_common.f90:141-242,
 * src/fft_cufft.f90:263-431); `sp` is the decomp of (nx/2+1,ny,nz) for PHYSICAL_IN_X and of
 * (nx,ny,nz/2+1) for PHYSICAL_IN_Z (fft_common.f90:210-216).  skip = opt_skip_XYZ_c2c.
 * inplace: c2c / c2r may overwrite their input (fft_cufft.f90:696-706, 961-971). */
int d2d_fft_plan_create(d2d_ctx *ctx, int format, int nx, int ny, int nz, int dtype, int inplace, const int skip[3], d2d_fft_plan **plan);
int d2d_fft_plan_destroy(d2d_fft_plan *plan);
int d2d_fft_plan_ph(const d2d_fft_plan *plan, const d2d_decomp **ph);
int d2d_fft_plan_sp(const d2d_fft_plan *plan, const d2d_decomp **sp);
/* decomp_2d_fft_get_size (src/fft_common.f90:311-327): spectral-side pencil of sp, 1-based */
int d2d_fft_get_size(const d2d_fft_plan *plan, int istart[3], int iend[3], int isize[3]);
/* decomp_2d_fft_3d: fft_3d_c2c / fft_3d_r2c / fft_3d_c2r (src/fft_cufft.f90:676-790, 795-934, 939-1170) */
int d2d_fft_3d_c2c(d2d_fft_plan *plan, void *in, void *out, int isign);
int d2d_fft_3d_r2c(d2d_fft_plan *plan, const void *in_r, void *out_c);
int d2d_fft_3d_c2r(d2d_fft_plan *plan, void *in_c, void *out_r);
/* the same with HOST arrays (what the reference's examples do with `!$acc data copyin(in) copy(out)`,
 * examples/fft_physical_z/fft_r2c_z.f90): H2D of the input, transform, D2H of the output */
int d2d_fft_3d_r2c_host(d2d_fft_plan *plan, const void *in_r_host, void *out_c_host);
int d2d_fft_3d_c2r_host(d2d_fft_plan *plan, const void *in_c_host, void *out_r_host);
int d2d_fft_3d_c2c_host(d2d_fft_plan *plan, const void *in_host, void *out_host, int isign);

/* batched 1-D transforms on one local (n1,n2,n3) array: c2c_1m_{x,y,z}, r2c_1m_{x,z}, c2r_1m_{x,z}
 * (src/fft_cufft.f90:489-671).  axis = 0,1,2.  For r2c (n1,n2,n3) is the REAL shape and the complex
 * array has n/2+1 along `axis`; for c2r (n1,n2,n3) is the real OUTPUT shape. */
int d2d_fft_c2c_1m(d2d_ctx *ctx, int dtype, int axis, int n1, int n2, int n3, const void *in, void *out, int isign);
int d2d_fft_r2c_1m(d2d_ctx *ctx, int dtype, int axis, int n1, int n2, int n3, const void *in_r, void *out_c);
int d2d_fft_c2r_1m(d2d_ctx *ctx, int dtype, int axis, int n1, int n2, int n3, const void *in_c, void *out_r);

/* ---- memory ------------------------------------------------------------------------------------
 * replaces alloc_dev.f90 (device allocatables), block_gpu.f90:90,144,224,229 (cudaHostAlloc /
 * cudaFreeHost), decomp_pool.f90:202-270 (cudaHostGetDevicePointer) and the cudaMemcpy calls. */
int d2d_dev_alloc(void **ptr, int64_t bytes);
int d2d_dev_free(void *ptr);
int d2d_host_alloc_pinned(void **ptr, int64_t bytes);
int d2d_host_free(void *ptr);
int d2d_host_get_device_pointer(void **dev_ptr, void *host_ptr);
int d2d_memcpy(void *dst, const void *src, int64_t bytes, int kind);
int d2d_memcpy_async(d2d_ctx *ctx, void *dst, const void *src, int64_t bytes, int kind);

const char *d2d_last_error(void);
const char *d2d_version(void);
/* number of compiled FFT kernel instantiations, and a description of entry i (diagnostics) */
int d2d_fft_kernel_count(void);
int d2d_fft_kernel_describe(int i, char *buf, int buflen);
/* Host-only view of the private wire layout used inside decomp_2d_fft_3d (diagnostics / CPU tests):
 * the piece map of the stage on `pencil` (0 x, 1 y, 2 z) for the link towards pencil `other`, as a
 * producer (consumer = 0) or consumer (1).  Element e of line (a,b) lives at element offset
 * off[m] + (e - e0[m]) se[m] + a sa[m] + b sb[m] of the peers' buffer (in_self[m] = 0: the receive
 * buffer for a consumer, the send buffer for a producer) or of the buffer holding this rank's own
 * block (in_self[m] = 1: always the producer's send buffer).  na/nb = batch extents of the stage;
 * cnt/disp = this side's block sizes / displacements (blocks are padded to `padq` elements along
 * their unit-stride axis, so these are NOT the reference's x1cnts... tables). */
int d2d_debug_link_map(const d2d_decomp *decomp, int pencil, int other, int consumer, int padq, int *np, int e0[9], int64_t off[8],
                       int in_self[8], int64_t se[8], int64_t sa[8], int64_t sb[8], int *na, int *nb, int64_t cnt[8], int64_t disp[8]);
/* Chunk [f0, f1) of a link along its free axis (the axis neither stage of the link transforms: x for Z<->Y, z for Y<->X;
 * batch axis a of the stage when axis_is_a = 1, else b; extent nf): the contiguous element sub-range off[m], cnt[m] of every
 * block m that a producer restricted to the chunk writes / a consumer restricted to it reads (chunk-wise overlap of the
 * exchange with its neighbouring stages; host arithmetic only). */
int d2d_debug_link_chunk(const d2d_decomp *decomp, int pencil, int other, int padq, int f0, int f1, int *np, int *axis_is_a, int *nf,
                         int64_t off[8], int64_t cnt[8]);
/* the user's dense pencil seen with the same (a,b) convention */
int d2d_debug_user_map(const d2d_decomp *decomp, int pencil, int64_t *se, int64_t *sa, int64_t *sb, int *n, int *na, int *nb);

#ifdef __cplusplus
}
#endif
#endif /* D2D_B200_H */
