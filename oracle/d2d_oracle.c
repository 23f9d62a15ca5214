/*
 * d2d_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked into or called by the product path).
 *
 * CPU restatement, in plain C, of the algorithm of 2DECOMP&FFT's distributed 3-D FFT hot path with
 * the `generic` (Glassman) backend.  It is the parity checker for the CUDA library and the
 * "port" CPU baseline of bench.py.  Every function cites the reference file:line it follows
 * (paths relative to the reference tree).
 *
 * PARITY PINNING.  The reference is 100 % Fortran + MPI and cannot be compiled in this image (no
 * Fortran compiler, no MPI), so there is no oracle/_ref.  The reference holds no golden files; its
 * tests generate fixtures in code:
 *   - transposes: index-encoded field, exact equality after each transpose
 *     (examples/test2d/test2d.f90:74-199)                    -> pinned (tests/test_oracle_*.py)
 *   - partition: size conservation (examples/init_test/init_test.f90:82-112) -> pinned
 *   - FFT: round-trip error thresholds only (examples/fft_physical_x/fft_c2c_x.f90:64-154 ...)
 *     -> round trip pinned; FORWARD SPECTRA ARE "parity unpinned" by the reference itself.  We pin
 *     them against numpy/pocketfft and the analytic DFT of the reference's ramp field instead.
 *
 * The file is compiled twice (REAL=double -> suffix _f64, REAL=float -> suffix _f32), mirroring the
 * reference's compile-time `mytype` (src/decomp_2d_constants.f90:15-32).
 *
 * "World" model: all p_row*p_col MPI ranks are simulated inside one process.  Rank-local work is an
 * OpenMP loop over ranks (one thread per rank, like one MPI rank per core); MPI_ALLTOALLV becomes
 * memcpy between the rank buffers between two implicit barriers.
 *
 * Index conventions are Fortran's: column-major, element (i,j,k) 0-based at i + n1*(j + n2*k).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef ORC_DOUBLE
typedef double REAL;
#define SFX(name) name##_f64
#else
typedef float REAL;
#define SFX(name) name##_f32
#endif

#define ORC_MAXP 64

/* ------------------------------------------------------------------------------------------------
 * decomp_info  (src/info.f90:11-47).  0-based starts here (the reference is 1-based).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
   int xst[3], xen[3], xsz[3];
   int yst[3], yen[3], ysz[3];
   int zst[3], zen[3], zsz[3];
   int x1dist[ORC_MAXP], y1dist[ORC_MAXP], y2dist[ORC_MAXP], z2dist[ORC_MAXP];
   int64_t x1cnts[ORC_MAXP], y1cnts[ORC_MAXP], y2cnts[ORC_MAXP], z2cnts[ORC_MAXP];
   int64_t x1disp[ORC_MAXP], y1disp[ORC_MAXP], y2disp[ORC_MAXP], z2disp[ORC_MAXP];
} orc_decomp;

#ifdef ORC_DOUBLE /* precision-independent host arithmetic: compiled once */

/* distribute  (src/decomp_2d.f90:1070-1105): base n/p, the LAST n mod p ranks get one extra. */
void orc_distribute(int data1, int proc, int *st, int *en, int *sz)
{
   int size1 = data1 / proc;
   int nu = data1 - size1 * proc;
   int nl = proc - nu;
   st[0] = 1;
   sz[0] = size1;
   en[0] = size1;
   for (int i = 1; i < nl; i++) {
      st[i] = st[i - 1] + size1;
      sz[i] = size1;
      en[i] = en[i - 1] + size1;
   }
   size1 = size1 + 1;
   for (int i = nl; i < proc; i++) {
      st[i] = en[i - 1] + 1;
      sz[i] = size1;
      en[i] = en[i - 1] + size1;
   }
}

/* partition  (src/decomp_2d.f90:1016-1064).  pdim(i): 1 = local, 2 = over dims(1), 3 = over dims(2).
 * Output lstart is converted to 0-based, lend stays inclusive 0-based. */
static void orc_partition(int nx, int ny, int nz, const int pdim[3], const int dims[2], const int coord[2],
                          int lstart[3], int lend[3], int lsize[3])
{
   int st[ORC_MAXP], en[ORC_MAXP], sz[ORC_MAXP];
   for (int i = 0; i < 3; i++) {
      int gsize = (i == 0) ? nx : (i == 1) ? ny : nz;
      if (pdim[i] == 1) {
         lstart[i] = 0;
         lend[i] = gsize - 1;
         lsize[i] = gsize;
      } else {
         int d = pdim[i] - 2; /* 0 -> dims(1)/coord(1), 1 -> dims(2)/coord(2) */
         orc_distribute(gsize, dims[d], st, en, sz);
         lstart[i] = st[coord[d]] - 1;
         lend[i] = en[coord[d]] - 1;
         lsize[i] = sz[coord[d]];
      }
   }
}

/* decomp_info_init  (src/decomp_2d.f90:382-490) = get_dist (:1112-1133) + 3x partition (:413-418)
 * + prepare_buffer (:1138-1183).  rank -> coord follows MPI_CART_CREATE without reorder
 * (src/decomp_2d_init_fin.f90:95-123): coord = (rank / p_col, rank mod p_col). */
void orc_decomp_init(int nx, int ny, int nz, int p_row, int p_col, int rank, orc_decomp *d)
{
   int dims[2] = {p_row, p_col};
   int coord[2] = {rank / p_col, rank % p_col};
   int st[ORC_MAXP], en[ORC_MAXP];
   memset(d, 0, sizeof(*d));
   orc_distribute(nx, dims[0], st, en, d->x1dist);
   orc_distribute(ny, dims[0], st, en, d->y1dist);
   orc_distribute(ny, dims[1], st, en, d->y2dist);
   orc_distribute(nz, dims[1], st, en, d->z2dist);
   const int px[3] = {1, 2, 3}, py[3] = {2, 1, 3}, pz[3] = {2, 3, 1};
   orc_partition(nx, ny, nz, px, dims, coord, d->xst, d->xen, d->xsz);
   orc_partition(nx, ny, nz, py, dims, coord, d->yst, d->yen, d->ysz);
   orc_partition(nx, ny, nz, pz, dims, coord, d->zst, d->zen, d->zsz);
   for (int i = 0; i < dims[0]; i++) {
      d->x1cnts[i] = (int64_t)d->x1dist[i] * d->xsz[1] * d->xsz[2];
      d->y1cnts[i] = (int64_t)d->ysz[0] * d->y1dist[i] * d->ysz[2];
      d->x1disp[i] = (i == 0) ? 0 : d->x1disp[i - 1] + d->x1cnts[i - 1];
      d->y1disp[i] = (i == 0) ? 0 : d->y1disp[i - 1] + d->y1cnts[i - 1];
   }
   for (int i = 0; i < dims[1]; i++) {
      d->y2cnts[i] = (int64_t)d->ysz[0] * d->y2dist[i] * d->ysz[2];
      d->z2cnts[i] = (int64_t)d->zsz[0] * d->zsz[1] * d->z2dist[i];
      d->y2disp[i] = (i == 0) ? 0 : d->y2disp[i - 1] + d->y2cnts[i - 1];
      d->z2disp[i] = (i == 0) ? 0 : d->z2disp[i - 1] + d->z2cnts[i - 1];
   }
}

int orc_sizeof_decomp(void) { return (int)sizeof(orc_decomp); }

/* EVEN builds (src/decomp_2d.f90:1186-1204): the padded all-to-all counts -- "the last blocks along pencils always get
 * assigned more mesh points" -- and decomp%even (:448-454).  out = {x1count, y1count, y2count, z2count}. */
void orc_even_counts(int nx, int ny, int nz, int p_row, int p_col, int rank, int64_t out[4], int *even)
{
   orc_decomp d;
   orc_decomp_init(nx, ny, nz, p_row, p_col, rank, &d);
   out[0] = (int64_t)d.x1dist[p_row - 1] * d.y1dist[p_row - 1] * d.xsz[2];
   out[1] = out[0];
   out[2] = (int64_t)d.y2dist[p_col - 1] * d.z2dist[p_col - 1] * d.zsz[0];
   out[3] = out[2];
   *even = (nx % p_row == 0 && ny % p_row == 0 && ny % p_col == 0 && nz % p_col == 0);
}

/* One OpenMP thread plays one MPI rank of the simulated world.  Launchers such as torchrun export OMP_NUM_THREADS=1,
 * which would silently serialise the ranks: the harness sets the thread count explicitly and reads it back. */
#ifdef _OPENMP
#include <omp.h>
void orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }
int orc_get_max_threads(void) { return omp_get_max_threads(); }
#else
void orc_set_threads(int n) { (void)n; }
int orc_get_max_threads(void) { return 1; }
#endif

/* best_2d_grid  (src/decomp_2d_init_fin.f90:270-300) with findfactor (src/factor.f90).
 * The reference lists the factors of nproc in increasing order (<= sqrt, then their complements)
 * and picks col = factors(nfact/2+1), row = nproc/col. */
void orc_best_2d_grid(int iproc, int *best_p_row, int *best_p_col)
{
   int factors[256], nfact = 0;
   int m = (int)sqrt((double)iproc);
   for (int i = 1; i <= m; i++)
      if (iproc % i == 0) factors[nfact++] = i;
   int nlow = nfact;
   if (factors[nlow - 1] * factors[nlow - 1] != iproc)
      for (int i = nlow + 1; i <= 2 * nlow; i++) factors[nfact++] = iproc / factors[2 * nlow - i];
   else
      for (int i = nlow + 1; i <= 2 * nlow - 1; i++) factors[nfact++] = iproc / factors[2 * nlow - 1 - i];
   *best_p_col = factors[nfact / 2]; /* factors(nfact/2+1), 1-based */
   *best_p_row = iproc / *best_p_col;
}

/* ------------------------------------------------------------------------------------------------
 * pack / unpack, element-size generic (a transpose is a bit-exact copy for every type)
 * ---------------------------------------------------------------------------------------------- */

/* mem_split_xy_*  (src/transpose_x_to_y.f90:266-382), pos formula :314 */
void orc_mem_split_xy(const char *in, int n1, int n2, int n3, char *out, int iproc, const int *dist,
                      const int64_t *disp, int es)
{
   int i1 = 0, i2 = -1;
   for (int m = 0; m < iproc; m++) {
      i1 = i2 + 1;
      i2 = i1 + dist[m] - 1;
      int64_t w = i2 - i1 + 1;
      for (int k = 0; k < n3; k++)
         for (int j = 0; j < n2; j++)
            memcpy(out + es * (disp[m] + j * w + (int64_t)k * n2 * w), in + es * (i1 + (int64_t)n1 * (j + (int64_t)n2 * k)),
                   (size_t)es * w);
   }
}

/* mem_merge_xy_*  (src/transpose_x_to_y.f90:384-500), pos formula :432; also mem_merge_zy
 * (src/transpose_z_to_y.f90:488-605, :536) which has the identical shape with y2dist/y2disp. */
void orc_mem_merge_y(const char *in, int n1, int n2, int n3, char *out, int iproc, const int *dist,
                     const int64_t *disp, int es)
{
   int i1 = 0, i2 = -1;
   for (int m = 0; m < iproc; m++) {
      i1 = i2 + 1;
      i2 = i1 + dist[m] - 1;
      int64_t h = i2 - i1 + 1;
      for (int k = 0; k < n3; k++)
         for (int j = i1; j <= i2; j++)
            memcpy(out + es * ((int64_t)n1 * (j + (int64_t)n2 * k)), in + es * (disp[m] + (j - i1) * (int64_t)n1 + (int64_t)k * h * n1),
                   (size_t)es * n1);
   }
}

/* mem_split_yx_* (src/transpose_y_to_x.f90:266-382, :314) and mem_split_yz_*
 * (src/transpose_y_to_z.f90:370-488, :418): pack a Y-pencil by y-range. */
void orc_mem_split_y(const char *in, int n1, int n2, int n3, char *out, int iproc, const int *dist,
                     const int64_t *disp, int es)
{
   int i1 = 0, i2 = -1;
   for (int m = 0; m < iproc; m++) {
      i1 = i2 + 1;
      i2 = i1 + dist[m] - 1;
      int64_t h = i2 - i1 + 1;
      for (int k = 0; k < n3; k++)
         for (int j = i1; j <= i2; j++)
            memcpy(out + es * (disp[m] + (j - i1) * (int64_t)n1 + (int64_t)k * h * n1), in + es * ((int64_t)n1 * (j + (int64_t)n2 * k)),
                   (size_t)es * n1);
   }
}

/* mem_merge_yx_*  (src/transpose_y_to_x.f90:384-500), pos formula :432 */
void orc_mem_merge_yx(const char *in, int n1, int n2, int n3, char *out, int iproc, const int *dist,
                      const int64_t *disp, int es)
{
   int i1 = 0, i2 = -1;
   for (int m = 0; m < iproc; m++) {
      i1 = i2 + 1;
      i2 = i1 + dist[m] - 1;
      int64_t w = i2 - i1 + 1;
      for (int k = 0; k < n3; k++)
         for (int j = 0; j < n2; j++)
            memcpy(out + es * (i1 + (int64_t)n1 * (j + (int64_t)n2 * k)), in + es * (disp[m] + j * w + (int64_t)k * n2 * w),
                   (size_t)es * w);
   }
}

/* mem_merge_yz_* (src/transpose_y_to_z.f90:490-606, :538) and mem_split_zy_*
 * (src/transpose_z_to_y.f90:370-486, :418): z is the slowest index, the segment of peer m is the
 * contiguous slab k1..k2, so both are plain contiguous copies (the CPU build skips them and
 * sends/receives in place, transpose_y_to_z.f90:153-155, transpose_z_to_y.f90:175-177). */
void orc_mem_copy_z(const char *in, int n1, int n2, int n3, char *out, int es)
{
   memcpy(out, in, (size_t)es * n1 * n2 * n3);
}

/* ------------------------------------------------------------------------------------------------
 * The four transposes on a simulated world.  src[r] / dst[r] are the pencils of rank r.
 * Exchange = MPI_ALLTOALLV(work1, cnts, disp, ..., work2, ...) on DECOMP_2D_COMM_COL (x<->y: ranks
 * sharing coord(2), ordered by coord(1)) or DECOMP_2D_COMM_ROW (y<->z: ranks sharing coord(1),
 * ordered by coord(2))  (src/decomp_2d_init_fin.f90:118-123).
 * dir: 0 x->y (transpose_x_to_y.f90:71-135), 1 y->z (transpose_y_to_z.f90:70-185),
 *      2 z->y (transpose_z_to_y.f90:71-186), 3 y->x (transpose_y_to_x.f90:71-135).
 * dims==1 paths are plain copies (transpose_x_to_y.f90:41-50 etc.).
 * ---------------------------------------------------------------------------------------------- */
void orc_transpose_world(int dir, int nx, int ny, int nz, int p_row, int p_col, char *const *src, char *const *dst, int es)
{
   int np = p_row * p_col;
   orc_decomp *dc = (orc_decomp *)malloc(sizeof(orc_decomp) * np);
   char **wk1 = (char **)calloc(np, sizeof(char *));
   char **wk2 = (char **)calloc(np, sizeof(char *));
   for (int r = 0; r < np; r++) orc_decomp_init(nx, ny, nz, p_row, p_col, r, &dc[r]);
   int over_col = (dir == 0 || dir == 3); /* COL communicator, size p_row */
   int csize = over_col ? p_row : p_col;

   if (csize == 1) {
#pragma omp parallel for schedule(static)
      for (int r = 0; r < np; r++) {
         const int *sz = (dir == 0) ? dc[r].xsz : (dir == 3 || dir == 1) ? dc[r].ysz : dc[r].zsz;
         memcpy(dst[r], src[r], (size_t)es * sz[0] * sz[1] * sz[2]);
      }
      goto done;
   }

   /* pack */
#pragma omp parallel for schedule(static)
   for (int r = 0; r < np; r++) {
      orc_decomp *d = &dc[r];
      int64_t bufsz = (int64_t)d->xsz[0] * d->xsz[1] * d->xsz[2];
      int64_t b2 = (int64_t)d->ysz[0] * d->ysz[1] * d->ysz[2];
      int64_t b3 = (int64_t)d->zsz[0] * d->zsz[1] * d->zsz[2];
      if (b2 > bufsz) bufsz = b2;
      if (b3 > bufsz) bufsz = b3;
      wk1[r] = (char *)malloc((size_t)es * bufsz);
      wk2[r] = (char *)malloc((size_t)es * bufsz);
      switch (dir) {
      case 0: orc_mem_split_xy(src[r], d->xsz[0], d->xsz[1], d->xsz[2], wk1[r], p_row, d->x1dist, d->x1disp, es); break;
      case 1: orc_mem_split_y(src[r], d->ysz[0], d->ysz[1], d->ysz[2], wk1[r], p_col, d->y2dist, d->y2disp, es); break;
      case 2: orc_mem_copy_z(src[r], d->zsz[0], d->zsz[1], d->zsz[2], wk1[r], es); break;
      case 3: orc_mem_split_y(src[r], d->ysz[0], d->ysz[1], d->ysz[2], wk1[r], p_row, d->y1dist, d->y1disp, es); break;
      }
   }
   /* all-to-all-v: segment m of rank r's send buffer goes to segment (r's index) of peer m's recv buffer */
#pragma omp parallel for schedule(static)
   for (int r = 0; r < np; r++) {
      int c1 = r / p_col, c2 = r % p_col;
      int me = over_col ? c1 : c2;
      for (int m = 0; m < csize; m++) {
         int peer = over_col ? (m * p_col + c2) : (c1 * p_col + m);
         /* receive from `peer` its segment destined to `me` */
         const orc_decomp *ds = &dc[peer], *dr = &dc[r];
         const int64_t *scnt, *sdsp, *rcnt, *rdsp;
         switch (dir) {
         case 0: scnt = ds->x1cnts; sdsp = ds->x1disp; rcnt = dr->y1cnts; rdsp = dr->y1disp; break;
         case 1: scnt = ds->y2cnts; sdsp = ds->y2disp; rcnt = dr->z2cnts; rdsp = dr->z2disp; break;
         case 2: scnt = ds->z2cnts; sdsp = ds->z2disp; rcnt = dr->y2cnts; rdsp = dr->y2disp; break;
         default: scnt = ds->y1cnts; sdsp = ds->y1disp; rcnt = dr->x1cnts; rdsp = dr->x1disp; break;
         }
         if (scnt[me] != rcnt[m]) abort(); /* MPI would fail on a size mismatch */
         memcpy(wk2[r] + es * rdsp[m], wk1[peer] + es * sdsp[me], (size_t)es * rcnt[m]);
      }
   }
   /* unpack */
#pragma omp parallel for schedule(static)
   for (int r = 0; r < np; r++) {
      orc_decomp *d = &dc[r];
      switch (dir) {
      case 0: orc_mem_merge_y(wk2[r], d->ysz[0], d->ysz[1], d->ysz[2], dst[r], p_row, d->y1dist, d->y1disp, es); break;
      case 1: orc_mem_copy_z(wk2[r], d->zsz[0], d->zsz[1], d->zsz[2], dst[r], es); break;
      case 2: orc_mem_merge_y(wk2[r], d->ysz[0], d->ysz[1], d->ysz[2], dst[r], p_col, d->y2dist, d->y2disp, es); break;
      case 3: orc_mem_merge_yx(wk2[r], d->xsz[0], d->xsz[1], d->xsz[2], dst[r], p_row, d->x1dist, d->x1disp, es); break;
      }
      free(wk1[r]);
      free(wk2[r]);
   }
done:
   free(wk1);
   free(wk2);
   free(dc);
}

#else
void orc_decomp_init(int nx, int ny, int nz, int p_row, int p_col, int rank, orc_decomp *d);
void orc_transpose_world(int dir, int nx, int ny, int nz, int p_row, int p_col, char *const *src, char *const *dst, int es);
#endif /* ORC_DOUBLE (precision independent part) */

/* ------------------------------------------------------------------------------------------------
 * Glassman FFT  (src/glassman.f90:29-108).  Complex stored interleaved (re, im).
 * ---------------------------------------------------------------------------------------------- */

/* SPCPFT (src/glassman.f90:69-108): UIN viewed as (B,C,A), UOUT as (B,A,C), first index fastest.
 * ANGLE is DOUBLE PRECISION but computed from a `mytype` literal and a `mytype` divide (:86);
 * OMEGA/DELTA/SUM are complex(mytype); OMEGA advances by recurrence (:103). */
static void SFX(spcpft)(int a, int b, int c, const REAL *uin, REAL *uout, int isign)
{
   double angle = (double)((REAL)6.28318530717958 / (REAL)(a * c));
   REAL om_r = 1, om_i = 0;
   REAL de_r = (REAL)cos(angle);
   REAL de_i = (isign == 1) ? (REAL)sin(angle) : (REAL)(-sin(angle));
   for (int ic = 0; ic < c; ic++) {
      for (int ia = 0; ia < a; ia++) {
         for (int ib = 0; ib < b; ib++) {
            const REAL *p = uin + 2 * ((size_t)ib + (size_t)b * ((c - 1) + (size_t)c * ia));
            REAL s_r = p[0], s_i = p[1];
            for (int jc = c - 2; jc >= 0; jc--) {
               const REAL *q = uin + 2 * ((size_t)ib + (size_t)b * (jc + (size_t)c * ia));
               REAL t_r = om_r * s_r - om_i * s_i;
               REAL t_i = om_r * s_i + om_i * s_r;
               s_r = q[0] + t_r;
               s_i = q[1] + t_i;
            }
            REAL *o = uout + 2 * ((size_t)ib + (size_t)b * (ia + (size_t)a * ic));
            o[0] = s_r;
            o[1] = s_i;
         }
         REAL n_r = de_r * om_r - de_i * om_i;
         REAL n_i = de_r * om_i + de_i * om_r;
         om_r = n_r;
         om_i = n_i;
      }
   }
}

/* SPCFFT (src/glassman.f90:29-67): smallest remaining factor first, ping-pong U <-> WORK. */
void SFX(orc_spcfft)(REAL *u, int n, int isign, REAL *work)
{
   int a = 1, b = n, c = 1, inu = 1;
   while (b > 1) {
      a = c * a;
      c = 2;
      while (b % c != 0) c++;
      b = b / c;
      if (inu)
         SFX(spcpft)(a, b, c, u, work, isign);
      else
         SFX(spcpft)(a, b, c, work, u, isign);
      inu = !inu;
   }
   if (!inu) memcpy(u, work, sizeof(REAL) * 2 * (size_t)n);
}

/* ------------------------------------------------------------------------------------------------
 * generic backend 1-D multi-line wrappers  (src/fft_generic.f90:112-384).  Line copy through `buf`.
 * ---------------------------------------------------------------------------------------------- */

/* c2c_1m_x/y/z (src/fft_generic.f90:112-208): axis = 0,1,2 on an (n1,n2,n3) complex array */
static void SFX(c2c_1m)(REAL *a, int n1, int n2, int n3, int axis, int isign)
{
   int n = (axis == 0) ? n1 : (axis == 1) ? n2 : n3;
   size_t stride = (axis == 0) ? 1 : (axis == 1) ? (size_t)n1 : (size_t)n1 * n2;
   /* lines are (j,k) for x, (i,k) for y, (i,j) for z -- loop order as in the reference */
   int na = (axis == 0) ? n2 : n1;                       /* inner batch extent */
   int nb = (axis == 2) ? n2 : n3;                       /* outer batch extent */
   size_t sa = (axis == 0) ? (size_t)n1 : 1;             /* inner batch stride */
   size_t sb = (axis == 2) ? (size_t)n1 : (size_t)n1 * n2; /* outer batch stride */
   REAL *buf = (REAL *)malloc(sizeof(REAL) * 4 * (size_t)n);
   REAL *scratch = buf + 2 * (size_t)n;
   for (int ob = 0; ob < nb; ob++)
      for (int ia = 0; ia < na; ia++) {
         REAL *base = a + 2 * (ia * sa + ob * sb);
         for (int e = 0; e < n; e++) {
            buf[2 * e] = base[2 * e * stride];
            buf[2 * e + 1] = base[2 * e * stride + 1];
         }
         SFX(orc_spcfft)(buf, n, isign, scratch);
         for (int e = 0; e < n; e++) {
            base[2 * e * stride] = buf[2 * e];
            base[2 * e * stride + 1] = buf[2 * e + 1];
         }
      }
   free(buf);
}

/* r2c_1m_x (:211-251) / r2c_1m_z (:254-294): complexify, full c2c(-1), keep bins 0..n/2.
 * input real (s1,s2,s3); output complex with the transform axis cut to d = n/2+1. */
static void SFX(r2c_1m)(const REAL *in, int s1, int s2, int s3, REAL *out, int axis /*0 or 2*/)
{
   int n = (axis == 0) ? s1 : s3;
   int d = n / 2 + 1;
   int o1 = (axis == 0) ? d : s1, o2 = s2;
   REAL *buf = (REAL *)malloc(sizeof(REAL) * 4 * (size_t)n);
   REAL *scratch = buf + 2 * (size_t)n;
   if (axis == 0) {
      for (int k = 0; k < s3; k++)
         for (int j = 0; j < s2; j++) {
            for (int i = 0; i < n; i++) {
               buf[2 * i] = in[i + (size_t)s1 * (j + (size_t)s2 * k)];
               buf[2 * i + 1] = 0;
            }
            SFX(orc_spcfft)(buf, n, -1, scratch);
            memcpy(out + 2 * ((size_t)o1 * (j + (size_t)o2 * k)), buf, sizeof(REAL) * 2 * d);
         }
   } else {
      for (int j = 0; j < s2; j++)
         for (int i = 0; i < s1; i++) {
            for (int k = 0; k < n; k++) {
               buf[2 * k] = in[i + (size_t)s1 * (j + (size_t)s2 * k)];
               buf[2 * k + 1] = 0;
            }
            SFX(orc_spcfft)(buf, n, -1, scratch);
            for (int k = 0; k < d; k++) {
               REAL *o = out + 2 * (i + (size_t)o1 * (j + (size_t)o2 * k));
               o[0] = buf[2 * k];
               o[1] = buf[2 * k + 1];
            }
         }
   }
   free(buf);
}

/* c2r_1m_x (:297-340) / c2r_1m_z (:343-384): buf(1..n/2+1)=in, buf(i)=conjg(buf(n+2-i)) for
 * i=n/2+2..n, c2c(+1), keep the real part.  output real (d1,d2,d3). */
static void SFX(c2r_1m)(const REAL *in, REAL *out, int d1, int d2, int d3, int axis /*0 or 2*/)
{
   int n = (axis == 0) ? d1 : d3;
   int h = n / 2 + 1;
   int i1 = (axis == 0) ? h : d1, i2 = d2;
   REAL *buf = (REAL *)malloc(sizeof(REAL) * 4 * (size_t)n);
   REAL *scratch = buf + 2 * (size_t)n;
   int nb1 = (axis == 0) ? d3 : d2, nb0 = (axis == 0) ? d2 : d1;
   for (int o = 0; o < nb1; o++)
      for (int q = 0; q < nb0; q++) {
         for (int e = 0; e < h; e++) {
            size_t idx = (axis == 0) ? (e + (size_t)i1 * (q + (size_t)i2 * o)) : (q + (size_t)i1 * (o + (size_t)i2 * e));
            buf[2 * e] = in[2 * idx];
            buf[2 * e + 1] = in[2 * idx + 1];
         }
         for (int e = h; e < n; e++) { /* 1-based i = e+1: buf(i) = conjg(buf(n+2-i)) -> 0-based n-e */
            buf[2 * e] = buf[2 * (n - e)];
            buf[2 * e + 1] = -buf[2 * (n - e) + 1];
         }
         SFX(orc_spcfft)(buf, n, 1, scratch);
         for (int e = 0; e < n; e++) {
            size_t idx = (axis == 0) ? (e + (size_t)d1 * (q + (size_t)d2 * o)) : (q + (size_t)d1 * (o + (size_t)d2 * e));
            out[idx] = buf[2 * e];
         }
      }
   free(buf);
}

/* exported single-rank helpers (used by line-sampled parity checks at the full bench sizes) */
void SFX(orc_c2c_1m)(REAL *a, int n1, int n2, int n3, int axis, int isign) { SFX(c2c_1m)(a, n1, n2, n3, axis, isign); }
void SFX(orc_r2c_1m)(const REAL *in, int s1, int s2, int s3, REAL *out, int axis) { SFX(r2c_1m)(in, s1, s2, s3, out, axis); }
void SFX(orc_c2r_1m)(const REAL *in, REAL *out, int d1, int d2, int d3, int axis) { SFX(c2r_1m)(in, out, d1, d2, d3, axis); }

/* ------------------------------------------------------------------------------------------------
 * 3-D drivers on the simulated world  (src/fft_common_3d.f90:9-116, 121-189, 194-296).
 * format: 1 = PHYSICAL_IN_X, 3 = PHYSICAL_IN_Z (src/decomp_2d_constants.f90:92-93)
 * `sp` = decomp of (nx/2+1,ny,nz) for X format, (nx,ny,nz/2+1) for Z (src/fft_common.f90:210-216).
 * skip[3]: opt_skip_XYZ_c2c (src/fft_common.f90:185-193; c2c_1m_* return early, fft_generic.f90:124).
 * The input arrays are never modified (not-inplace semantics: the reference copies `in` to wk1).
 * ---------------------------------------------------------------------------------------------- */
static size_t SFX(pencil_elems)(const int sz[3]) { return (size_t)sz[0] * sz[1] * sz[2]; }

static void SFX(stage_c2c)(REAL **a, const orc_decomp *dc, int np, int pencil /*0 x,1 y,2 z*/, int isign, const int *skip)
{
   if (skip && skip[pencil]) return;
#pragma omp parallel for schedule(static)
   for (int r = 0; r < np; r++) {
      const int *sz = (pencil == 0) ? dc[r].xsz : (pencil == 1) ? dc[r].ysz : dc[r].zsz;
      SFX(c2c_1m)(a[r], sz[0], sz[1], sz[2], pencil, isign);
   }
}

static REAL **SFX(alloc_world)(const orc_decomp *dc, int np, int pencil, int cplx)
{
   REAL **p = (REAL **)malloc(sizeof(REAL *) * np);
   for (int r = 0; r < np; r++) {
      const int *sz = (pencil == 0) ? dc[r].xsz : (pencil == 1) ? dc[r].ysz : dc[r].zsz;
      p[r] = (REAL *)malloc(sizeof(REAL) * (cplx ? 2 : 1) * (SFX(pencil_elems)(sz) + 1));
   }
   return p;
}
static void SFX(free_world)(REAL **p, int np)
{
   for (int r = 0; r < np; r++) free(p[r]);
   free(p);
}

/* fft_3d_c2c (src/fft_common_3d.f90:9-116).  in[r]: X-pencil (X-fwd / Z-bwd ordering x->y->z) or
 * Z-pencil (the other ordering); out[r]: the opposite pencil. */
void SFX(orc_fft_3d_c2c)(int nx, int ny, int nz, int p_row, int p_col, int format, int isign, const int *skip,
                         REAL *const *in, REAL *const *out)
{
   int np = p_row * p_col;
   const int es = 2 * (int)sizeof(REAL);
   orc_decomp *ph = (orc_decomp *)malloc(sizeof(orc_decomp) * np);
   for (int r = 0; r < np; r++) orc_decomp_init(nx, ny, nz, p_row, p_col, r, &ph[r]);
   int xyz = (format == 1 && isign == -1) || (format == 3 && isign == 1);
   REAL **wk1 = SFX(alloc_world)(ph, np, xyz ? 0 : 2, 1);
   REAL **wk2 = SFX(alloc_world)(ph, np, 1, 1);
   for (int r = 0; r < np; r++)
      memcpy(wk1[r], in[r], (size_t)es * SFX(pencil_elems)(xyz ? ph[r].xsz : ph[r].zsz));
   if (xyz) {
      SFX(stage_c2c)(wk1, ph, np, 0, isign, skip);
      orc_transpose_world(0, nx, ny, nz, p_row, p_col, (char *const *)wk1, (char *const *)wk2, es);
      SFX(stage_c2c)(wk2, ph, np, 1, isign, skip);
      orc_transpose_world(1, nx, ny, nz, p_row, p_col, (char *const *)wk2, (char *const *)out, es);
      SFX(stage_c2c)((REAL **)out, ph, np, 2, isign, skip);
   } else {
      SFX(stage_c2c)(wk1, ph, np, 2, isign, skip);
      orc_transpose_world(2, nx, ny, nz, p_row, p_col, (char *const *)wk1, (char *const *)wk2, es);
      SFX(stage_c2c)(wk2, ph, np, 1, isign, skip);
      orc_transpose_world(3, nx, ny, nz, p_row, p_col, (char *const *)wk2, (char *const *)out, es);
      SFX(stage_c2c)((REAL **)out, ph, np, 0, isign, skip);
   }
   SFX(free_world)(wk1, np);
   SFX(free_world)(wk2, np);
   free(ph);
}

/* fft_3d_r2c (src/fft_common_3d.f90:121-189).  in_r[r]: real X-pencil of ph (format 1) or real
 * Z-pencil of ph (format 3); out_c[r]: complex Z-pencil of sp (format 1) / X-pencil of sp (format 3). */
void SFX(orc_fft_3d_r2c)(int nx, int ny, int nz, int p_row, int p_col, int format, const int *skip,
                         REAL *const *in_r, REAL *const *out_c)
{
   int np = p_row * p_col;
   const int es = 2 * (int)sizeof(REAL);
   int sx = (format == 1) ? nx / 2 + 1 : nx, sy = ny, sz_ = (format == 1) ? nz : nz / 2 + 1;
   orc_decomp *ph = (orc_decomp *)malloc(sizeof(orc_decomp) * np);
   orc_decomp *sp = (orc_decomp *)malloc(sizeof(orc_decomp) * np);
   for (int r = 0; r < np; r++) {
      orc_decomp_init(nx, ny, nz, p_row, p_col, r, &ph[r]);
      orc_decomp_init(sx, sy, sz_, p_row, p_col, r, &sp[r]);
   }
   REAL **wk13 = SFX(alloc_world)(sp, np, (format == 1) ? 0 : 2, 1);
   REAL **wk2 = SFX(alloc_world)(sp, np, 1, 1);
   if (format == 1) {
#pragma omp parallel for schedule(static)
      for (int r = 0; r < np; r++) SFX(r2c_1m)(in_r[r], ph[r].xsz[0], ph[r].xsz[1], ph[r].xsz[2], wk13[r], 0);
      orc_transpose_world(0, sx, sy, sz_, p_row, p_col, (char *const *)wk13, (char *const *)wk2, es);
      SFX(stage_c2c)(wk2, sp, np, 1, -1, skip);
      orc_transpose_world(1, sx, sy, sz_, p_row, p_col, (char *const *)wk2, (char *const *)out_c, es);
      SFX(stage_c2c)((REAL **)out_c, sp, np, 2, -1, skip);
   } else {
#pragma omp parallel for schedule(static)
      for (int r = 0; r < np; r++) SFX(r2c_1m)(in_r[r], ph[r].zsz[0], ph[r].zsz[1], ph[r].zsz[2], wk13[r], 2);
      orc_transpose_world(2, sx, sy, sz_, p_row, p_col, (char *const *)wk13, (char *const *)wk2, es);
      SFX(stage_c2c)(wk2, sp, np, 1, -1, skip);
      orc_transpose_world(3, sx, sy, sz_, p_row, p_col, (char *const *)wk2, (char *const *)out_c, es);
      SFX(stage_c2c)((REAL **)out_c, sp, np, 0, -1, skip);
   }
   SFX(free_world)(wk13, np);
   SFX(free_world)(wk2, np);
   free(ph);
   free(sp);
}

/* fft_3d_c2r (src/fft_common_3d.f90:194-296).  in_c[r]: complex Z-pencil of sp (format 1) /
 * X-pencil of sp (format 3); out_r[r]: real X-pencil of ph (format 1) / Z-pencil of ph (format 3). */
void SFX(orc_fft_3d_c2r)(int nx, int ny, int nz, int p_row, int p_col, int format, const int *skip,
                         REAL *const *in_c, REAL *const *out_r)
{
   int np = p_row * p_col;
   const int es = 2 * (int)sizeof(REAL);
   int sx = (format == 1) ? nx / 2 + 1 : nx, sy = ny, sz_ = (format == 1) ? nz : nz / 2 + 1;
   orc_decomp *ph = (orc_decomp *)malloc(sizeof(orc_decomp) * np);
   orc_decomp *sp = (orc_decomp *)malloc(sizeof(orc_decomp) * np);
   for (int r = 0; r < np; r++) {
      orc_decomp_init(nx, ny, nz, p_row, p_col, r, &ph[r]);
      orc_decomp_init(sx, sy, sz_, p_row, p_col, r, &sp[r]);
   }
   REAL **wk1 = SFX(alloc_world)(sp, np, (format == 1) ? 2 : 0, 1);
   REAL **wk2 = SFX(alloc_world)(sp, np, 1, 1);
   REAL **wk13 = SFX(alloc_world)(sp, np, (format == 1) ? 0 : 2, 1);
   for (int r = 0; r < np; r++)
      memcpy(wk1[r], in_c[r], (size_t)es * SFX(pencil_elems)((format == 1) ? sp[r].zsz : sp[r].xsz));
   if (format == 1) {
      SFX(stage_c2c)(wk1, sp, np, 2, 1, skip);
      orc_transpose_world(2, sx, sy, sz_, p_row, p_col, (char *const *)wk1, (char *const *)wk2, es);
      SFX(stage_c2c)(wk2, sp, np, 1, 1, skip);
      orc_transpose_world(3, sx, sy, sz_, p_row, p_col, (char *const *)wk2, (char *const *)wk13, es);
#pragma omp parallel for schedule(static)
      for (int r = 0; r < np; r++) SFX(c2r_1m)(wk13[r], out_r[r], ph[r].xsz[0], ph[r].xsz[1], ph[r].xsz[2], 0);
   } else {
      SFX(stage_c2c)(wk1, sp, np, 0, 1, skip);
      orc_transpose_world(0, sx, sy, sz_, p_row, p_col, (char *const *)wk1, (char *const *)wk2, es);
      SFX(stage_c2c)(wk2, sp, np, 1, 1, skip);
      orc_transpose_world(1, sx, sy, sz_, p_row, p_col, (char *const *)wk2, (char *const *)wk13, es);
#pragma omp parallel for schedule(static)
      for (int r = 0; r < np; r++) SFX(c2r_1m)(wk13[r], out_r[r], ph[r].zsz[0], ph[r].zsz[1], ph[r].zsz[2], 2);
   }
   SFX(free_world)(wk1, np);
   SFX(free_world)(wk2, np);
   SFX(free_world)(wk13, np);
   free(ph);
   free(sp);
}
