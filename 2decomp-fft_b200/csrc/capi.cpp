// capi.cpp -- the extern "C" surface declared in include/d2d_b200.h.  Every entry point converts
// exceptions into a status code + thread-local message; nothing here aborts or prints.
#include <cstdio>
#include <cstring>

#include "common.h"
#include "fft_registry.h"

namespace d2d {
struct Plan;
Plan *plan_create(Ctx *ctx, int format, int nx, int ny, int nz, int dtype, int inplace, const int skip[3]);
void plan_destroy(Plan *p);
const d2d_decomp *plan_ph(const Plan *p);
const d2d_decomp *plan_sp(const Plan *p);
void fft_3d_c2c(Plan *p, void *in, void *out, int isign);
void fft_3d_r2c(Plan *p, const void *in_r, void *out_c);
void fft_3d_c2r(Plan *p, void *in_c, void *out_r);
void fft_3d_host(Plan *p, int which, const void *in_h, void *out_h, int isign);
void plan_get_size(const Plan *p, int istart[3], int iend[3], int isize[3]);
Ctx *plan_ctx(Plan *p);
void fft_1m(Ctx *ctx, int dtype, int mode, int axis, int n1, int n2, int n3, const void *in, void *out, int isign);
void halo_update(Ctx *ctx, const Decomp &d, int pencil, int level, int es, const int periodic[3], const void *in, void *out);
} // namespace d2d

using namespace d2d;

struct d2d_group {
   d2d::Group *g;
   int nranks;
};

#define D2D_TRY try {
#define D2D_CATCH                                                                                                      \
   }                                                                                                                   \
   catch (const d2d::Error &e)                                                                                         \
   {                                                                                                                   \
      d2d::set_last_error(e.what());                                                                                   \
      return e.code ? e.code : 1;                                                                                      \
   }                                                                                                                   \
   catch (const std::exception &e)                                                                                     \
   {                                                                                                                   \
      d2d::set_last_error(e.what());                                                                                   \
      return 1;                                                                                                        \
   }                                                                                                                   \
   return 0;

static void ctx_common_init(d2d_ctx *h, int nranks, int rank, int p_row, int p_col, int device)
{
   D2D_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "invalid rank / nranks");
   D2D_REQUIRE(p_row >= 1 && p_col >= 1 && p_row * p_col == nranks,
               "Invalid 2D processor grid - nproc /= p_row*p_col"); // src/decomp_2d_init_fin.f90:79-82
   D2D_REQUIRE(p_row <= kMaxP && p_col <= kMaxP, "process grid side larger than 8 is not supported");
   Ctx &c = h->c;
   c.nranks = nranks; c.rank = rank; c.p_row = p_row; c.p_col = p_col;
   c.c1 = rank / p_col; c.c2 = rank % p_col; c.device = device;
   D2D_CHECK_CUDA(cudaSetDevice(device));
   D2D_CHECK_CUDA(cudaStreamCreate(&c.stream)); // blocking stream: ordered against the legacy default stream
}

extern "C" {

const char *d2d_last_error(void) { return d2d::get_last_error().c_str(); }
const char *d2d_version(void) { return "d2d_b200 0.1 (sm_100a)"; }

int d2d_get_unique_id(unsigned char id[128])
{
   D2D_TRY
   nccl_unique_id(id);
   D2D_CATCH
}

int d2d_ctx_create(d2d_ctx **ctx, const unsigned char id[128], int nranks, int rank, int p_row, int p_col, int device)
{
   D2D_TRY
   std::unique_ptr<d2d_ctx> h(new d2d_ctx());
   ctx_common_init(h.get(), nranks, rank, p_row, p_col, device);
   if (nranks > 1) {
      D2D_REQUIRE(id != nullptr, "a unique id is required for nranks > 1");
      h->c.tr.reset(make_nccl_transport(id, nranks, rank));
   }
   *ctx = h.release();
   D2D_CATCH
}

int d2d_ctx_create_bootstrap(d2d_ctx **ctx, int nranks, int rank, int p_row, int p_col, int device, d2d_allgather_fn allgather, void *user)
{
   D2D_TRY
   std::unique_ptr<d2d_ctx> h(new d2d_ctx());
   ctx_common_init(h.get(), nranks, rank, p_row, p_col, device);
   if (nranks > 1) {
      D2D_REQUIRE(allgather != nullptr, "an all-gather callback is required for nranks > 1");
      h->c.tr.reset(make_boot_transport(allgather, user, nranks, rank));
   }
   *ctx = h.release();
   D2D_CATCH
}

int d2d_group_create(d2d_group **grp, int nranks)
{
   D2D_TRY
   D2D_REQUIRE(nranks >= 1, "invalid nranks");
   *grp = new d2d_group{group_create(nranks), nranks};
   D2D_CATCH
}
int d2d_group_destroy(d2d_group *grp)
{
   D2D_TRY
   if (grp) { group_destroy(grp->g); delete grp; }
   D2D_CATCH
}
int d2d_group_abort(d2d_group *grp)
{
   D2D_TRY
   D2D_REQUIRE(grp != nullptr, "null group");
   group_abort(grp->g);
   D2D_CATCH
}
int d2d_ctx_create_in_group(d2d_ctx **ctx, d2d_group *grp, int rank, int p_row, int p_col, int device)
{
   D2D_TRY
   D2D_REQUIRE(grp != nullptr, "null group");
   std::unique_ptr<d2d_ctx> h(new d2d_ctx());
   ctx_common_init(h.get(), grp->nranks, rank, p_row, p_col, device);
   if (grp->nranks > 1) h->c.tr.reset(make_local_transport(grp->g, rank));
   *ctx = h.release();
   D2D_CATCH
}
int d2d_ctx_destroy(d2d_ctx *ctx)
{
   D2D_TRY
   delete ctx;
   D2D_CATCH
}
int d2d_ctx_sync(d2d_ctx *ctx)
{
   D2D_TRY
   ctx->c.sync_all();
   D2D_CATCH
}
int d2d_ctx_set_blocking(d2d_ctx *ctx, int blocking)
{
   ctx->c.blocking = blocking != 0;
   return 0;
}
int d2d_ctx_set_even(d2d_ctx *ctx, int even)
{
   ctx->c.even = even != 0;
   return 0;
}
void *d2d_ctx_stream(d2d_ctx *ctx) { return (void *)ctx->c.stream; }
int d2d_ctx_info(const d2d_ctx *ctx, int *nranks, int *rank, int dims[2], int coord[2], int *transport)
{
   const Ctx &c = ctx->c;
   if (nranks) *nranks = c.nranks;
   if (rank) *rank = c.rank;
   if (dims) { dims[0] = c.p_row; dims[1] = c.p_col; }
   if (coord) { coord[0] = c.c1; coord[1] = c.c2; }
   if (transport) *transport = c.tr ? c.tr->kind() : D2D_TRANSPORT_NONE;
   return 0;
}
int64_t d2d_ctx_launch_count(const d2d_ctx *ctx) { return ctx->c.launches; }
int d2d_ctx_profile(d2d_ctx *ctx, int enable)
{
   D2D_TRY
   if (!enable) ctx->c.prof_flush();
   ctx->c.profiling = enable != 0;
   D2D_CATCH
}
int d2d_ctx_profile_count(d2d_ctx *ctx)
{
   try { ctx->c.prof_flush(); } catch (...) { return -1; }
   return (int)ctx->c.prof.size();
}
int d2d_ctx_profile_get(d2d_ctx *ctx, int i, char label[64], double *total_ms, int64_t *calls, double *bytes)
{
   D2D_TRY
   ctx->c.prof_flush();
   D2D_REQUIRE(i >= 0 && i < (int)ctx->c.prof.size(), "profile index out of range");
   const ProfEntry &e = ctx->c.prof[i];
   snprintf(label, 64, "%s", e.label.c_str());
   if (total_ms) *total_ms = e.total_ms;
   if (calls) *calls = e.calls;
   if (bytes) *bytes = e.bytes;
   D2D_CATCH
}
int d2d_ctx_profile_reset(d2d_ctx *ctx)
{
   D2D_TRY
   ctx->c.prof_flush();
   ctx->c.prof.clear();
   D2D_CATCH
}
int d2d_best_2d_grid(int nproc, int *p_row, int *p_col)
{
   D2D_TRY
   D2D_REQUIRE(nproc >= 1, "nproc must be positive");
   best_2d_grid(nproc, p_row, p_col);
   D2D_CATCH
}

int d2d_decomp_create(d2d_ctx *ctx, int nx, int ny, int nz, d2d_decomp **decomp)
{
   D2D_TRY
   std::unique_ptr<d2d_decomp> h(new d2d_decomp());
   h->ctx = &ctx->c;
   decomp_init(h->d, nx, ny, nz, ctx->c.p_row, ctx->c.p_col, ctx->c.rank);
   *decomp = h.release();
   D2D_CATCH
}
int d2d_decomp_create_for_rank(int nx, int ny, int nz, int p_row, int p_col, int rank, d2d_decomp **decomp)
{
   D2D_TRY
   D2D_REQUIRE(rank >= 0 && rank < p_row * p_col, "rank outside the process grid");
   std::unique_ptr<d2d_decomp> h(new d2d_decomp());
   h->ctx = nullptr;
   decomp_init(h->d, nx, ny, nz, p_row, p_col, rank);
   *decomp = h.release();
   D2D_CATCH
}
int d2d_decomp_destroy(d2d_decomp *decomp)
{
   delete decomp;
   return 0;
}
int d2d_decomp_query(const d2d_decomp *h, int xst[3], int xen[3], int xsz[3], int yst[3], int yen[3], int ysz[3], int zst[3],
                     int zen[3], int zsz[3])
{
   const Decomp &d = h->d;
   for (int i = 0; i < 3; i++) {
      if (xst) xst[i] = d.xst[i] + 1;
      if (xen) xen[i] = d.xen[i] + 1;
      if (xsz) xsz[i] = d.xsz[i];
      if (yst) yst[i] = d.yst[i] + 1;
      if (yen) yen[i] = d.yen[i] + 1;
      if (ysz) ysz[i] = d.ysz[i];
      if (zst) zst[i] = d.zst[i] + 1;
      if (zen) zen[i] = d.zen[i] + 1;
      if (zsz) zsz[i] = d.zsz[i];
   }
   return 0;
}
int d2d_decomp_dist(const d2d_decomp *h, int *x1dist, int *y1dist, int *y2dist, int *z2dist)
{
   const Decomp &d = h->d;
   for (int i = 0; i < d.p_row; i++) {
      if (x1dist) x1dist[i] = d.x1dist[i];
      if (y1dist) y1dist[i] = d.y1dist[i];
   }
   for (int i = 0; i < d.p_col; i++) {
      if (y2dist) y2dist[i] = d.y2dist[i];
      if (z2dist) z2dist[i] = d.z2dist[i];
   }
   return 0;
}
int d2d_decomp_counts(const d2d_decomp *h, int64_t *x1cnts, int64_t *y1cnts, int64_t *y2cnts, int64_t *z2cnts, int64_t *x1disp,
                      int64_t *y1disp, int64_t *y2disp, int64_t *z2disp)
{
   const Decomp &d = h->d;
   for (int i = 0; i < d.p_row; i++) {
      if (x1cnts) x1cnts[i] = d.x1cnts[i];
      if (y1cnts) y1cnts[i] = d.y1cnts[i];
      if (x1disp) x1disp[i] = d.x1disp[i];
      if (y1disp) y1disp[i] = d.y1disp[i];
   }
   for (int i = 0; i < d.p_col; i++) {
      if (y2cnts) y2cnts[i] = d.y2cnts[i];
      if (z2cnts) z2cnts[i] = d.z2cnts[i];
      if (y2disp) y2disp[i] = d.y2disp[i];
      if (z2disp) z2disp[i] = d.z2disp[i];
   }
   return 0;
}

int d2d_decomp_even(const d2d_decomp *h, int64_t *x1count, int64_t *y1count, int64_t *y2count, int64_t *z2count, int *even)
{
   const Decomp &d = h->d;
   if (x1count) *x1count = d.x1count;
   if (y1count) *y1count = d.y1count;
   if (y2count) *y2count = d.y2count;
   if (z2count) *z2count = d.z2count;
   if (even) *even = d.even;
   return 0;
}

int d2d_transpose(d2d_ctx *ctx, const d2d_decomp *decomp, int direction, int dtype, int is_complex, const void *src, void *dst)
{
   D2D_TRY
   D2D_REQUIRE(dtype == D2D_F32 || dtype == D2D_F64, "dtype must be D2D_F32 or D2D_F64");
   transpose(&ctx->c, decomp->d, direction, elem_size(dtype, is_complex), src, dst);
   ctx->c.finish_call();
   D2D_CATCH
}
int d2d_transpose_x_to_y(d2d_ctx *c, const d2d_decomp *d, int dtype, int is_complex, const void *src, void *dst)
{
   return d2d_transpose(c, d, D2D_X_TO_Y, dtype, is_complex, src, dst);
}
int d2d_transpose_y_to_z(d2d_ctx *c, const d2d_decomp *d, int dtype, int is_complex, const void *src, void *dst)
{
   return d2d_transpose(c, d, D2D_Y_TO_Z, dtype, is_complex, src, dst);
}
int d2d_transpose_z_to_y(d2d_ctx *c, const d2d_decomp *d, int dtype, int is_complex, const void *src, void *dst)
{
   return d2d_transpose(c, d, D2D_Z_TO_Y, dtype, is_complex, src, dst);
}
int d2d_transpose_y_to_x(d2d_ctx *c, const d2d_decomp *d, int dtype, int is_complex, const void *src, void *dst)
{
   return d2d_transpose(c, d, D2D_Y_TO_X, dtype, is_complex, src, dst);
}

int d2d_halo_update(d2d_ctx *ctx, const d2d_decomp *decomp, int pencil, int level, int dtype, int is_complex, const int periodic[3],
                    const void *in, void *out)
{
   D2D_TRY
   D2D_REQUIRE(dtype == D2D_F32 || dtype == D2D_F64, "dtype must be D2D_F32 or D2D_F64");
   halo_update(&ctx->c, decomp->d, pencil, level, elem_size(dtype, is_complex), periodic, in, out);
   ctx->c.finish_call();
   D2D_CATCH
}

int d2d_fft_plan_create(d2d_ctx *ctx, int format, int nx, int ny, int nz, int dtype, int inplace, const int skip[3], d2d_fft_plan **plan)
{
   D2D_TRY
   *plan = reinterpret_cast<d2d_fft_plan *>(plan_create(&ctx->c, format, nx, ny, nz, dtype, inplace, skip));
   D2D_CATCH
}
int d2d_fft_plan_destroy(d2d_fft_plan *plan)
{
   D2D_TRY
   plan_destroy(reinterpret_cast<Plan *>(plan));
   D2D_CATCH
}
int d2d_fft_plan_ph(const d2d_fft_plan *plan, const d2d_decomp **ph)
{
   *ph = plan_ph(reinterpret_cast<const Plan *>(plan));
   return 0;
}
int d2d_fft_plan_sp(const d2d_fft_plan *plan, const d2d_decomp **sp)
{
   *sp = plan_sp(reinterpret_cast<const Plan *>(plan));
   return 0;
}
int d2d_fft_get_size(const d2d_fft_plan *plan, int istart[3], int iend[3], int isize[3])
{
   plan_get_size(reinterpret_cast<const Plan *>(plan), istart, iend, isize);
   return 0;
}
int d2d_fft_3d_c2c(d2d_fft_plan *plan, void *in, void *out, int isign)
{
   D2D_TRY
   Plan *p = reinterpret_cast<Plan *>(plan);
   fft_3d_c2c(p, in, out, isign);
   plan_ctx(p)->finish_call();
   D2D_CATCH
}
int d2d_fft_3d_r2c(d2d_fft_plan *plan, const void *in_r, void *out_c)
{
   D2D_TRY
   Plan *p = reinterpret_cast<Plan *>(plan);
   fft_3d_r2c(p, in_r, out_c);
   plan_ctx(p)->finish_call();
   D2D_CATCH
}
int d2d_fft_3d_c2r(d2d_fft_plan *plan, void *in_c, void *out_r)
{
   D2D_TRY
   Plan *p = reinterpret_cast<Plan *>(plan);
   fft_3d_c2r(p, in_c, out_r);
   plan_ctx(p)->finish_call();
   D2D_CATCH
}
int d2d_fft_3d_r2c_host(d2d_fft_plan *plan, const void *in_r_host, void *out_c_host)
{
   D2D_TRY
   fft_3d_host(reinterpret_cast<Plan *>(plan), 0, in_r_host, out_c_host, D2D_FFT_FORWARD);
   D2D_CATCH
}
int d2d_fft_3d_c2r_host(d2d_fft_plan *plan, const void *in_c_host, void *out_r_host)
{
   D2D_TRY
   fft_3d_host(reinterpret_cast<Plan *>(plan), 1, in_c_host, out_r_host, D2D_FFT_BACKWARD);
   D2D_CATCH
}
int d2d_fft_3d_c2c_host(d2d_fft_plan *plan, const void *in_host, void *out_host, int isign)
{
   D2D_TRY
   fft_3d_host(reinterpret_cast<Plan *>(plan), 2, in_host, out_host, isign);
   D2D_CATCH
}

int d2d_fft_c2c_1m(d2d_ctx *ctx, int dtype, int axis, int n1, int n2, int n3, const void *in, void *out, int isign)
{
   D2D_TRY
   D2D_REQUIRE(isign == D2D_FFT_FORWARD || isign == D2D_FFT_BACKWARD, "isign must be -1 or +1");
   fft_1m(&ctx->c, dtype, MODE_C2C, axis, n1, n2, n3, in, out, isign);
   ctx->c.finish_call();
   D2D_CATCH
}
int d2d_fft_r2c_1m(d2d_ctx *ctx, int dtype, int axis, int n1, int n2, int n3, const void *in_r, void *out_c)
{
   D2D_TRY
   fft_1m(&ctx->c, dtype, MODE_R2C, axis, n1, n2, n3, in_r, out_c, D2D_FFT_FORWARD);
   ctx->c.finish_call();
   D2D_CATCH
}
int d2d_fft_c2r_1m(d2d_ctx *ctx, int dtype, int axis, int n1, int n2, int n3, const void *in_c, void *out_r)
{
   D2D_TRY
   fft_1m(&ctx->c, dtype, MODE_C2R, axis, n1, n2, n3, in_c, out_r, D2D_FFT_BACKWARD);
   ctx->c.finish_call();
   D2D_CATCH
}

int d2d_dev_alloc(void **ptr, int64_t bytes)
{
   D2D_TRY
   D2D_CHECK_CUDA(cudaMalloc(ptr, (size_t)bytes));
   D2D_CATCH
}
int d2d_dev_free(void *ptr)
{
   D2D_TRY
   D2D_CHECK_CUDA(cudaFree(ptr));
   D2D_CATCH
}
int d2d_host_alloc_pinned(void **ptr, int64_t bytes)
{
   D2D_TRY
   D2D_CHECK_CUDA(cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocMapped)); // block_gpu.f90:90
   D2D_CATCH
}
int d2d_host_free(void *ptr)
{
   D2D_TRY
   D2D_CHECK_CUDA(cudaFreeHost(ptr));
   D2D_CATCH
}
int d2d_host_get_device_pointer(void **dev_ptr, void *host_ptr)
{
   D2D_TRY
   D2D_CHECK_CUDA(cudaHostGetDevicePointer(dev_ptr, host_ptr, 0)); // decomp_pool.f90:219
   D2D_CATCH
}
static cudaMemcpyKind kind_of(int kind)
{
   return kind == D2D_MEMCPY_H2D ? cudaMemcpyHostToDevice : kind == D2D_MEMCPY_D2H ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
}
int d2d_memcpy(void *dst, const void *src, int64_t bytes, int kind)
{
   D2D_TRY
   D2D_CHECK_CUDA(cudaMemcpy(dst, src, (size_t)bytes, kind_of(kind)));
   D2D_CATCH
}
int d2d_memcpy_async(d2d_ctx *ctx, void *dst, const void *src, int64_t bytes, int kind)
{
   D2D_TRY
   D2D_CHECK_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, kind_of(kind), ctx->c.stream));
   D2D_CATCH
}

int d2d_fft_kernel_count(void) { return fft_registry_size(); }
int d2d_debug_link_map(const d2d_decomp *decomp, int pencil, int other, int consumer, int padq, int *np, int e0[9], int64_t off[8],
                       int in_self[8], int64_t se[8], int64_t sa[8], int64_t sb[8], int *na, int *nb, int64_t cnt[8], int64_t disp[8])
{
   D2D_TRY
   D2D_REQUIRE((pencil == 1 && (other == 0 || other == 2)) || (other == 1 && (pencil == 0 || pencil == 2)), "not a link");
   char *peers = reinterpret_cast<char *>(uintptr_t(1) << 44), *self = reinterpret_cast<char *>(uintptr_t(1) << 45);
   const PieceMap m = fft_link_map(decomp->d, pencil, other, peers, self, 1, consumer != 0, padq);
   LinkSide L;
   fft_link_side(decomp->d, pencil, other, padq, L);
   for (int p = 0; p < L.np; p++) { cnt[p] = L.cnt[p]; disp[p] = L.disp[p]; }
   *np = m.np;
   for (int p = 0; p <= m.np; p++) e0[p] = m.e0[p];
   for (int p = 0; p < m.np; p++) {
      const uintptr_t a = reinterpret_cast<uintptr_t>(m.ptr[p]);
      in_self[p] = (a >> 45) & 1;
      off[p] = (int64_t)(a & ((uintptr_t(1) << 44) - 1));
      se[p] = m.se[p]; sa[p] = m.sa[p]; sb[p] = m.sb[p];
   }
   fft_stage_batch(decomp->d, pencil, *na, *nb);
   D2D_CATCH
}
int d2d_debug_link_chunk(const d2d_decomp *decomp, int pencil, int other, int padq, int f0, int f1, int *np, int *axis_is_a, int *nf,
                         int64_t off[8], int64_t cnt[8])
{
   D2D_TRY
   D2D_REQUIRE((pencil == 1 && (other == 0 || other == 2)) || (other == 1 && (pencil == 0 || pencil == 2)), "not a link");
   LinkChunk c;
   fft_link_chunk(decomp->d, pencil, other, padq, f0, f1, c);
   *np = c.np; *axis_is_a = c.axis_is_a; *nf = c.nf;
   for (int p = 0; p < c.np; p++) { off[p] = c.off[p]; cnt[p] = c.cnt[p]; }
   D2D_CATCH
}
int d2d_debug_user_map(const d2d_decomp *decomp, int pencil, int64_t *se, int64_t *sa, int64_t *sb, int *n, int *na, int *nb)
{
   D2D_TRY
   const PieceMap m = fft_user_map(decomp->d, pencil, nullptr);
   *se = m.se[0]; *sa = m.sa[0]; *sb = m.sb[0]; *n = m.e0[1];
   fft_stage_batch(decomp->d, pencil, *na, *nb);
   D2D_CATCH
}
int d2d_fft_kernel_describe(int i, char *buf, int buflen)
{
   if (i < 0 || i >= fft_registry_size()) return 1;
   const FftKernelInfo *k = fft_registry_at(i);
   snprintf(buf, buflen, "n=%d %s %s %s pairvec=%d tx=%d ly=%d threads=%d minb=%d smem=%zu radices=%d,%d,%d,%d", k->n,
            k->f64 ? "f64" : "f32", k->v2 ? (k->inl == IN_TILE ? "v2-tma-tile" : "v2-tma-line") : k->kind == KIND_LINE ? "line" : "tile",
            k->mode == MODE_C2C ? "c2c" : k->mode == MODE_R2C ? "r2c" : "c2r", k->pairvec, k->tx, k->ly, k->threads, k->minb,
            k->smem, k->radix[0], k->radix[1], k->radix[2], k->radix[3]);
   return 0;
}

} // extern "C"
