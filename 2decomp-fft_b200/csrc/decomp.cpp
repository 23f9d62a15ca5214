// decomp.cpp -- 2-D pencil decomposition arithmetic and the all-to-all layouts derived from it.
// Follows decomp_info_init (src/decomp_2d.f90:382-490): get_dist (:1112-1133) -> partition x3
// (:413-418, :1016-1064) -> prepare_buffer (:1138-1183); rank -> coord as MPI_CART_CREATE without
// reorder (src/decomp_2d_init_fin.f90:95-123).
#include <algorithm>
#include <cmath>
#include <cstring>

#include "common.h"

namespace d2d {

// distribute (src/decomp_2d.f90:1070-1105): n/p each, the last n%p ranks get one more
static void distribute(int n, int p, int *sz, int *off)
{
   const int base = n / p, nu = n - base * p, nl = p - nu;
   off[0] = 0;
   for (int i = 0; i < p; i++) {
      sz[i] = base + (i >= nl ? 1 : 0);
      off[i + 1] = off[i] + sz[i];
   }
}

void decomp_init(Decomp &d, int nx, int ny, int nz, int p_row, int p_col, int rank)
{
   D2D_REQUIRE(p_row >= 1 && p_col >= 1 && p_row <= kMaxP && p_col <= kMaxP, "process grid side must be in 1..8");
   D2D_REQUIRE(nx >= 1 && ny >= 1 && nz >= 1, "grid sizes must be positive");
   // src/decomp_2d_init_fin.f90:43-45 / decomp_2d.f90:399-405: every rank must own at least one point
   D2D_REQUIRE(std::min(nx, ny) >= p_row && std::min(ny, nz) >= p_col,
               "Make sure that min(nx,ny) >= p_row and min(ny,nz) >= p_col");
   std::memset(&d, 0, sizeof(d));
   d.nx = nx; d.ny = ny; d.nz = nz; d.p_row = p_row; d.p_col = p_col; d.rank = rank;
   d.c1 = rank / p_col; d.c2 = rank % p_col;
   distribute(nx, p_row, d.x1dist, d.x1off);
   distribute(ny, p_row, d.y1dist, d.y1off);
   distribute(ny, p_col, d.y2dist, d.y2off);
   distribute(nz, p_col, d.z2dist, d.z2off);
   // X-pencil (nx, ny/p_row, nz/p_col); Y-pencil (nx/p_row, ny, nz/p_col); Z-pencil (nx/p_row, ny/p_col, nz)
   d.xst[0] = 0;              d.xsz[0] = nx;
   d.xst[1] = d.y1off[d.c1];  d.xsz[1] = d.y1dist[d.c1];
   d.xst[2] = d.z2off[d.c2];  d.xsz[2] = d.z2dist[d.c2];
   d.yst[0] = d.x1off[d.c1];  d.ysz[0] = d.x1dist[d.c1];
   d.yst[1] = 0;              d.ysz[1] = ny;
   d.yst[2] = d.z2off[d.c2];  d.ysz[2] = d.z2dist[d.c2];
   d.zst[0] = d.x1off[d.c1];  d.zsz[0] = d.x1dist[d.c1];
   d.zst[1] = d.y2off[d.c2];  d.zsz[1] = d.y2dist[d.c2];
   d.zst[2] = 0;              d.zsz[2] = nz;
   for (int i = 0; i < 3; i++) {
      d.xen[i] = d.xst[i] + d.xsz[i] - 1;
      d.yen[i] = d.yst[i] + d.ysz[i] - 1;
      d.zen[i] = d.zst[i] + d.zsz[i] - 1;
   }
   for (int i = 0; i < p_row; i++) {
      d.x1cnts[i] = (int64_t)d.x1dist[i] * d.xsz[1] * d.xsz[2];
      d.y1cnts[i] = (int64_t)d.ysz[0] * d.y1dist[i] * d.ysz[2];
      d.x1disp[i] = i ? d.x1disp[i - 1] + d.x1cnts[i - 1] : 0;
      d.y1disp[i] = i ? d.y1disp[i - 1] + d.y1cnts[i - 1] : 0;
   }
   for (int i = 0; i < p_col; i++) {
      d.y2cnts[i] = (int64_t)d.ysz[0] * d.y2dist[i] * d.ysz[2];
      d.z2cnts[i] = (int64_t)d.zsz[0] * d.zsz[1] * d.z2dist[i];
      d.y2disp[i] = i ? d.y2disp[i - 1] + d.y2cnts[i - 1] : 0;
      d.z2disp[i] = i ? d.z2disp[i - 1] + d.z2cnts[i - 1] : 0;
   }
   // EVEN (src/decomp_2d.f90:1197-1203): pad every message to the size of the largest one -- the last blocks are the largest
   d.x1count = (int64_t)d.x1dist[p_row - 1] * d.y1dist[p_row - 1] * d.xsz[2];
   d.y1count = d.x1count;
   d.y2count = (int64_t)d.y2dist[p_col - 1] * d.z2dist[p_col - 1] * d.zsz[0];
   d.z2count = d.y2count;
   for (int i = 0; i < p_row; i++) { d.e_cnts_row[i] = d.x1count; d.e_disp_row[i] = (int64_t)i * d.x1count; }
   for (int i = 0; i < p_col; i++) { d.e_cnts_col[i] = d.y2count; d.e_disp_col[i] = (int64_t)i * d.y2count; }
   d.even = (nx % p_row == 0 && ny % p_row == 0 && ny % p_col == 0 && nz % p_col == 0) ? 1 : 0; // :443-454
}

// best_2d_grid (src/decomp_2d_init_fin.f90:270-300) on top of findfactor (src/factor.f90:17-54)
void best_2d_grid(int nproc, int *p_row, int *p_col)
{
   std::vector<int> f;
   const int m = (int)std::sqrt((double)nproc);
   for (int i = 1; i <= m; i++)
      if (nproc % i == 0) f.push_back(i);
   const int nlow = (int)f.size();
   const bool square = f[nlow - 1] * f[nlow - 1] == nproc;
   for (int i = nlow - 1 - (square ? 1 : 0); i >= 0; i--) f.push_back(nproc / f[i]);
   *p_col = f[f.size() / 2];
   *p_row = nproc / *p_col;
}

// ---- all-to-all layouts (pos formulas of mem_split_* / mem_merge_*) ------------
namespace {
struct Side {
   const int *dist, *off;
   const int64_t *cnts, *disp;
   int np;
};
// the dist/disp table used by pencil `pencil` when talking to pencil `other`
Side side_of(const Decomp &d, int pencil, int other, bool even = false)
{
   if (even) { // padded equal counts, segment m at m * count (mem_split_* / mem_merge_* under EVEN: pos = m * count + 1)
      if (pencil == 0) return {d.x1dist, d.x1off, d.e_cnts_row, d.e_disp_row, d.p_row};
      if (pencil == 2) return {d.z2dist, d.z2off, d.e_cnts_col, d.e_disp_col, d.p_col};
      if (other == 0) return {d.y1dist, d.y1off, d.e_cnts_row, d.e_disp_row, d.p_row};
      return {d.y2dist, d.y2off, d.e_cnts_col, d.e_disp_col, d.p_col};
   }
   if (pencil == 0) return {d.x1dist, d.x1off, d.x1cnts, d.x1disp, d.p_row};
   if (pencil == 2) return {d.z2dist, d.z2off, d.z2cnts, d.z2disp, d.p_col};
   if (other == 0) return {d.y1dist, d.y1off, d.y1cnts, d.y1disp, d.p_row};
   return {d.y2dist, d.y2off, d.y2cnts, d.y2disp, d.p_col};
}
// pieces along the pencil's own axis, blocks at base + disp[m] (element size es)
PieceMap pieces(const Decomp &d, int pencil, const Side &s, char *base, int es)
{
   PieceMap m{};
   m.np = s.np;
   const int *sz = pencil == 0 ? d.xsz : pencil == 1 ? d.ysz : d.zsz;
   for (int p = 0; p < s.np; p++) {
      m.e0[p] = s.off[p];
      m.ptr[p] = base + (size_t)es * s.disp[p];
      if (pencil == 0) { // X lines (j,k) flattened: ii + w (j + n2 k)   (transpose_x_to_y.f90:314, transpose_y_to_x.f90:432)
         m.se[p] = 1; m.sa[p] = s.dist[p]; m.sb[p] = 0;
      } else if (pencil == 1) { // Y lines (i,k): i + n1 (jj + h k)  (transpose_x_to_y.f90:432, y_to_x:314, y_to_z:418, z_to_y:536)
         m.se[p] = sz[0]; m.sa[p] = 1; m.sb[p] = (long long)sz[0] * s.dist[p];
      } else { // Z lines (i,j) flattened: i + n1 (j + n2 kk)  (transpose_y_to_z.f90:538, z_to_y:418)
         m.se[p] = (long long)sz[0] * sz[1]; m.sa[p] = 1; m.sb[p] = 0;
      }
   }
   m.e0[s.np] = s.off[s.np];
   return m;
}
} // namespace

int comm_size(const Decomp &d, int from, int to) { return (from == 0 || to == 0) ? d.p_row : d.p_col; }

PieceMap natural_map(const Decomp &d, int pencil, void *ptr)
{
   PieceMap m{};
   m.np = 1;
   const int *sz = pencil == 0 ? d.xsz : pencil == 1 ? d.ysz : d.zsz;
   m.e0[0] = 0;
   m.e0[1] = sz[pencil];
   m.ptr[0] = ptr;
   if (pencil == 0) { m.se[0] = 1; m.sa[0] = sz[0]; m.sb[0] = 0; }
   else if (pencil == 1) { m.se[0] = sz[0]; m.sa[0] = 1; m.sb[0] = (long long)sz[0] * sz[1]; }
   else { m.se[0] = (long long)sz[0] * sz[1]; m.sa[0] = 1; m.sb[0] = 0; }
   return m;
}

PieceMap send_map(const Decomp &d, int from, int to, void *sendbuf, int es, bool even)
{
   return pieces(d, from, side_of(d, from, to, even), (char *)sendbuf, es);
}
PieceMap recv_map(const Decomp &d, int from, int to, void *recvbuf, void *sendbuf, int es, bool even)
{
   PieceMap m = pieces(d, to, side_of(d, to, from, even), (char *)recvbuf, es);
   // the block this rank "sends to itself" is never moved: read it where the producer wrote it
   const Side s = side_of(d, from, to, even);
   const int me = (from == 0 || to == 0) ? d.c1 : d.c2;
   m.ptr[me] = (char *)sendbuf + (size_t)es * s.disp[me];
   return m;
}

int64_t send_total(const Decomp &d, int from, int /*to*/) { return d.pencil_elems(from); }
int64_t recv_total(const Decomp &d, int /*from*/, int to) { return d.pencil_elems(to); }

// ---- private wire layouts of the fused 3-D transforms ---------------------------------------------
// Only pencil contents are contractual: inside decomp_2d_fft_3d the blocks that
// travel between ranks keep the reference's per-peer COUNTS and DISPLACEMENTS (so the exchange is the
// same all-to-all-v) but are ordered INSIDE for the memory system:
//     Z <-> Y link: block stored (z, y, x)   -- z fastest
//     Y <-> X link: block stored (y, x, z)   -- y fastest
// so that every stage either streams whole lines (its transform axis is the unit-stride axis) or
// reads TX-wide tiles whose rows are only a few KB apart.  On B200 a tile whose rows are >= 1 MB
// apart runs at ~1/2 the bandwidth (one 2 MB page per row; measured), which is what
// the reference layouts (x fastest everywhere) force on the z stage.
// Batch axes of a stage (a = the axis TX adjacent lines of a tile run along, b = the other one):
//     X stage: a = y, b = z      Y stage: a = z, b = x      Z stage: a = x, b = y
void fft_stage_batch(const Decomp &d, int pencil, int &na, int &nb)
{
   if (pencil == 0) { na = d.xsz[1]; nb = d.xsz[2]; }
   else if (pencil == 1) { na = d.ysz[2]; nb = d.ysz[0]; }
   else { na = d.zsz[0]; nb = d.zsz[1]; }
}

// the user's dense pencil (x fastest) seen with the stage's (a,b) convention
PieceMap fft_user_map(const Decomp &d, int pencil, void *ptr)
{
   PieceMap m{};
   m.np = 1;
   const int *sz = pencil == 0 ? d.xsz : pencil == 1 ? d.ysz : d.zsz;
   const long long n1 = sz[0], n12 = (long long)sz[0] * sz[1];
   m.e0[0] = 0;
   m.e0[1] = sz[pencil];
   m.ptr[0] = ptr;
   if (pencil == 0) { m.se[0] = 1; m.sa[0] = n1; m.sb[0] = n12; }       // a = y, b = z
   else if (pencil == 1) { m.se[0] = n1; m.sa[0] = n12; m.sb[0] = 1; } // a = z, b = x
   else { m.se[0] = n12; m.sa[0] = 1; m.sb[0] = n1; }                   // a = x, b = y
   return m;
}

// One side of a link, seen from this rank: the stage on `pencil` talking to the stage on `other`.
// Blocks are padded along their unit-stride axis to a multiple of `padq` elements (128 B) so that
// every row of every block starts on a 128-byte boundary even for ragged extents like nz/2+1 = 513
// (measured: a 16-byte-misaligned 513-element row costs 2.2x the L2 sectors and read-modify-write
// of partial sectors on the store side).  Counts / displacements are therefore this library's own
// (the reference's x1cnts... stay untouched for the bare transpose API).
void fft_link_side(const Decomp &d, int pencil, int other, int padq, LinkSide &L)
{
   const Side s = side_of(d, pencil, other);
   auto pad = [&](long long v) { return (v + padq - 1) / padq * padq; };
   L = LinkSide{};
   L.np = s.np;
   L.me = (pencil == 0 || other == 0) ? d.c1 : d.c2;
   int64_t disp = 0;
   for (int p = 0; p < s.np; p++) {
      L.e0[p] = s.off[p];
      const long long ext = s.dist[p]; // extent of this piece along the stage's own axis
      if (pencil == 2) {               // Z stage, Z<->Y link, block (z: pad(ext), y: zsz1, x: zsz0)
         const long long wp = pad(ext);
         L.se[p] = 1; L.sa[p] = wp * d.zsz[1]; L.sb[p] = wp;
         L.cnt[p] = wp * d.zsz[1] * d.zsz[0];
      } else if (pencil == 1 && other == 2) { // Y stage, Z<->Y link, block (z: pad(ysz2), y: ext, x: ysz0)
         const long long wp = pad(d.ysz[2]);
         L.se[p] = wp; L.sa[p] = 1; L.sb[p] = wp * ext;
         L.cnt[p] = wp * ext * d.ysz[0];
      } else if (pencil == 1) {               // Y stage, Y<->X link, block (y: pad(ext), x: ysz0, z: ysz2)
         const long long hp = pad(ext);
         L.se[p] = 1; L.sa[p] = hp * d.ysz[0]; L.sb[p] = hp;
         L.cnt[p] = hp * d.ysz[0] * d.ysz[2];
      } else {                                // X stage, Y<->X link, block (y: pad(xsz1), x: ext, z: xsz2)
         const long long hp = pad(d.xsz[1]);
         L.se[p] = hp; L.sa[p] = 1; L.sb[p] = hp * ext;
         L.cnt[p] = hp * ext * d.xsz[2];
      }
      L.disp[p] = disp;
      disp += L.cnt[p];
   }
   L.e0[s.np] = s.off[s.np];
   L.total = disp;
}

// Chunk of a link along its FREE axis -- the axis neither stage of the link transforms: x for Z<->Y, z for Y<->X.  It is the
// slowest axis of every block (see the layouts above), so lines f0 <= f < f1 of the free axis occupy one contiguous
// sub-range of every block: element offset off[m] = disp[m] + f0 * plane[m], length cnt[m] = (f1 - f0) * plane[m].  The
// free axis is batch axis `a` of the Z stage (Z<->Y) and of the Y stage (Y<->X), batch axis `b` of the other stage of the
// link; its extent nf is the same on both sides of the link and on every rank of the communicator.  This is the geometry
// of the chunk-wise overlap of an exchange with its neighbouring stages: a producer restricted to
// [f0, f1) writes exactly these sub-ranges, the exchange moves them, a consumer restricted to [f0, f1) reads nothing else.
void fft_link_chunk(const Decomp &d, int pencil, int other, int padq, int f0, int f1, LinkChunk &C)
{
   LinkSide L;
   fft_link_side(d, pencil, other, padq, L);
   int na = 0, nb = 0;
   fft_stage_batch(d, pencil, na, nb);
   C = LinkChunk{};
   C.np = L.np;
   C.me = L.me;
   // Z<->Y: free axis x = a of the Z stage, b of the Y stage;  Y<->X: free axis z = a of the Y stage, b of the X stage
   C.axis_is_a = (pencil == 2 || (pencil == 1 && other == 0)) ? 1 : 0;
   C.nf = C.axis_is_a ? na : nb;
   D2D_REQUIRE(0 <= f0 && f0 <= f1 && f1 <= C.nf, "chunk outside the free axis");
   for (int p = 0; p < L.np; p++) {
      const int64_t plane = C.nf ? L.cnt[p] / C.nf : 0; // elements of block p per line of the free axis
      D2D_REQUIRE(plane * C.nf == L.cnt[p], "block size is not a multiple of the free-axis extent");
      D2D_REQUIRE((C.axis_is_a ? L.sa[p] : L.sb[p]) == plane || L.cnt[p] == 0, "free axis is not the slowest axis of the block");
      C.off[p] = L.disp[p] + (int64_t)f0 * plane;
      C.cnt[p] = (int64_t)(f1 - f0) * plane;
   }
}

// largest buffer (in elements) any stage of a 3-D transform on this decomp writes or receives
int64_t fft_work_elems(const Decomp &d, int padq)
{
   int64_t m = d.max_pencil();
   static const int sides[4][2] = {{2, 1}, {1, 2}, {1, 0}, {0, 1}};
   for (auto &sd : sides) {
      LinkSide L;
      fft_link_side(d, sd[0], sd[1], padq, L);
      m = std::max(m, L.total);
   }
   return m;
}

// map of the stage on `pencil` for the link towards `other`.  Producer: every block (its own
// included) is written into `self_buf`.  Consumer: the peers' blocks are read from `peers_buf`
// (the receive buffer); the block this rank kept for itself is read where its producer stage left
// it in `self_buf` (it never travels).
PieceMap fft_link_map(const Decomp &d, int pencil, int other, void *peers_buf, void *self_buf, int es, bool consumer, int padq)
{
   LinkSide L, P;
   fft_link_side(d, pencil, other, padq, L);
   fft_link_side(d, other, pencil, padq, P);
   PieceMap m{};
   m.np = L.np;
   for (int p = 0; p < L.np; p++) {
      m.e0[p] = L.e0[p];
      if (!consumer) m.ptr[p] = (char *)self_buf + (size_t)es * L.disp[p];
      else if (p == L.me) m.ptr[p] = (char *)self_buf + (size_t)es * P.disp[L.me];
      else m.ptr[p] = (char *)peers_buf + (size_t)es * L.disp[p];
      m.se[p] = L.se[p]; m.sa[p] = L.sa[p]; m.sb[p] = L.sb[p];
   }
   m.e0[L.np] = L.e0[L.np];
   return m;
}

// the all-to-all(v) of a fused 3-D transform: producer side on pencil `from`, consumer on `to`
void fft_exchange(Ctx *ctx, const Decomp &d, int from, int to, const void *sendbuf, void *recvbuf, int es, int padq)
{
   LinkSide S, R;
   fft_link_side(d, from, to, padq, S);
   fft_link_side(d, to, from, padq, R);
   if (S.np == 1) return;
   D2D_REQUIRE(ctx->tr != nullptr, "context has no transport but the process grid has more than one rank");
   const bool col = (from == 0 || to == 0);
   std::vector<PeerXfer> xf;
   double bytes = 0;
   for (int k = 1; k < S.np; k++) {
      const int m = (S.me + k) % S.np; // stagger the peers
      PeerXfer x;
      x.peer = col ? ctx->peer_rank_col(m) : ctx->peer_rank_row(m);
      x.sendptr = (const char *)sendbuf + (size_t)es * S.disp[m];
      x.sendbytes = (size_t)es * S.cnt[m];
      x.recvptr = (char *)recvbuf + (size_t)es * R.disp[m];
      x.recvbytes = (size_t)es * R.cnt[m];
      bytes += (double)x.sendbytes;
      xf.push_back(x);
   }
   static const char *names[3][3] = {{"", "a2a_x_y", ""}, {"a2a_y_x", "", "a2a_y_z"}, {"", "a2a_z_y", ""}};
   ProfScope ps(ctx, names[from][to], bytes);
   ctx->tr->exchange(xf, ctx->stream);
}

// all-to-all(v) with the peers of the row / column communicator (self excluded).
// Replaces decomp_2d_nccl_alltoall_{col,row}_* (src/decomp_2d_nccl.f90:214-473) / MPI_ALLTOALLV.
void exchange(Ctx *ctx, const Decomp &d, int from, int to, const void *sendbuf, void *recvbuf, int es, int send_w, int recv_w, bool even)
{
   const bool col = (from == 0 || to == 0);
   const int np = col ? d.p_row : d.p_col;
   const int me = col ? d.c1 : d.c2;
   if (np == 1) return;
   D2D_REQUIRE(ctx->tr != nullptr, "context has no transport but the process grid has more than one rank");
   const Side s = side_of(d, from, to, even), r = side_of(d, to, from, even);
   const bool p2p = p2p_active(ctx);
   std::vector<PeerXfer> xf;
   std::vector<size_t> dst_off;
   double bytes = 0;
   for (int k = 1; k < np; k++) {
      const int m = (me + k) % np; // stagger the peers
      PeerXfer x;
      x.peer = col ? ctx->peer_rank_col(m) : ctx->peer_rank_row(m);
      x.sendptr = (const char *)sendbuf + (size_t)es * s.disp[m];
      x.sendbytes = (size_t)es * s.cnts[m];
      x.recvptr = (char *)recvbuf + (size_t)es * r.disp[m];
      x.recvbytes = (size_t)es * r.cnts[m];
      bytes += (double)x.sendbytes;
      xf.push_back(x);
      if (p2p) { // where the destination rank expects the block from this rank
         Decomp dm;
         decomp_init(dm, d.nx, d.ny, d.nz, d.p_row, d.p_col, x.peer);
         const Side rm = side_of(dm, to, from, even);
         D2D_REQUIRE(rm.cnts[me] == s.cnts[m], "exchange: block sizes of the two sides disagree");
         dst_off.push_back((size_t)es * rm.disp[me]);
      }
   }
   static const char *names[3][3] = {{"", "a2a_x_y", ""}, {"a2a_y_x", "", "a2a_y_z"}, {"", "a2a_z_y", ""}};
   ProfScope ps(ctx, names[from][to], bytes);
   if (p2p) {
      D2D_REQUIRE(recv_w >= 0 && recvbuf == ctx->work[recv_w], "peer-memory exchange: the receive side must be a work buffer");
      p2p_exchange(ctx, xf, dst_off, recv_w, send_w);
   } else {
      ctx->tr->exchange(xf, ctx->stream);
   }
}

size_t uniform_pencil_bytes(const Ctx *ctx, const Decomp &d, int es, bool even)
{
   int64_t m = 0;
   for (int r = 0; r < ctx->nranks; r++) {
      Decomp a;
      decomp_init(a, d.nx, d.ny, d.nz, ctx->p_row, ctx->p_col, r);
      m = std::max(m, a.max_pencil());
      if (even) m = std::max(m, std::max(a.x1count * a.p_row, a.y2count * a.p_col)); // src/decomp_2d.f90:443-446
   }
   return (size_t)es * (size_t)m;
}

} // namespace d2d
