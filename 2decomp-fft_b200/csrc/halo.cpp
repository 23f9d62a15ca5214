// halo.cpp -- halo-cell exchange between neighbouring pencils on the device.
// Replaces update_halo_{real,complex} (src/halo.f90:101-198 with src/halo_common.f90: allocate `out` with `level` ghost
// layers around the two decomposed axes of the pencil, copy `in` into its interior) and halo_exchange_{real,complex}
// (src/halo.f90:311-399 with src/halo_exchange_{x,y,z}_body.f90: two successive exchanges, first along the axis that the
// first process-grid dimension splits, then along the axis of the second one; every message spans the FULL extent of the
// other axes including their ghost layers, so the second exchange carries the corners).  Neighbours are those of
// init_neighbour (src/halo.f90:55-99: MPI_CART_SHIFT on the pencil's Cartesian communicator, periodic per axis as given to
// decomp_2d_init); a missing neighbour (MPI_PROC_NULL) leaves its ghost layers untouched.
//
// The reference describes the strips with MPI_TYPE_VECTOR and lets MPI pack them; here a strip is packed by one box-copy
// kernel into the work buffer, the (at most two) messages of an exchange go through the context's transport -- copy-engine
// pushes into the neighbour's CUDA-IPC mapped buffer, NCCL send/recv, or device copies between rank-threads -- and one
// box-copy kernel per strip writes the ghost layers.
#include "common.h"

namespace d2d {

void launch_box_copy(Ctx *ctx, void *dst, long long d1, long long d2, const void *src, long long s1, long long s2, int e1, int e2, int e3, int es);

namespace {
struct Strip {
   long long off; // element offset of the strip's first element inside the haloed array
   int e[3];      // extents
};
} // namespace

void halo_update(Ctx *ctx, const Decomp &d, int pencil, int level, int es, const int periodic[3], const void *in, void *out)
{
   D2D_REQUIRE(pencil >= 0 && pencil < 3, "Invalid data passed to update_halo"); // src/halo.f90:299-306
   D2D_REQUIRE(level >= 0, "halo level must not be negative");
   D2D_CHECK_CUDA(cudaSetDevice(ctx->device));
   ProfScope ps(ctx, "halo_update");
   const int *sz = pencil == 0 ? d.xsz : pencil == 1 ? d.ysz : d.zsz;
   // the two decomposed axes of the pencil, in exchange order: the one split by dims(1) (coord c1), then dims(2) (coord c2)
   const int ax1 = pencil == 0 ? 1 : 0, ax2 = pencil == 2 ? 1 : 2;
   int h[3] = {0, 0, 0};
   h[ax1] = level;
   h[ax2] = level;
   const long long n[3] = {sz[0] + 2LL * h[0], sz[1] + 2LL * h[1], sz[2] + 2LL * h[2]};
   char *o = (char *)out;
   // interior: out(h + i) = in(i)   (halo_common.f90:75-84)
   launch_box_copy(ctx, o + (size_t)es * (h[0] + n[0] * (h[1] + n[1] * h[2])), n[0], n[1], in, sz[0], sz[1], sz[0], sz[1], sz[2], es);
   if (level == 0) return;
   // two strips of the larger of the two exchanges, for the largest pencil of ANY rank: every rank grows (and republishes) its
   // work buffers at the same call
   size_t need = 0;
   for (int r = 0; r < ctx->nranks; r++) {
      Decomp a;
      decomp_init(a, d.nx, d.ny, d.nz, ctx->p_row, ctx->p_col, r);
      const int *s = pencil == 0 ? a.xsz : pencil == 1 ? a.ysz : a.zsz;
      const size_t m[3] = {(size_t)s[0] + 2 * h[0], (size_t)s[1] + 2 * h[1], (size_t)s[2] + 2 * h[2]};
      for (int ax : {ax1, ax2}) need = std::max(need, 2 * (size_t)es * (size_t)level * (m[0] * m[1] * m[2] / m[ax]));
   }

   for (int step = 0; step < 2; step++) {
      const int ax = step == 0 ? ax1 : ax2;
      const int np = step == 0 ? ctx->p_row : ctx->p_col, me = step == 0 ? ctx->c1 : ctx->c2;
      const bool per = periodic && periodic[ax] != 0;
      // neighbours along this process-grid dimension (MPI_CART_SHIFT): -1 = MPI_PROC_NULL
      const int im = me > 0 ? me - 1 : (per ? np - 1 : -1), ip = me < np - 1 ? me + 1 : (per ? 0 : -1);
      auto rank_of = [&](int idx) { return step == 0 ? idx * ctx->p_col + ctx->c2 : ctx->c1 * ctx->p_col + idx; };
      D2D_REQUIRE(sz[ax] >= level, "halo level larger than the local extent");
      // strips of `level` layers along ax, full extent (ghost layers included) along the other axes
      auto strip = [&](long long first) {
         Strip s;
         long long st[3] = {0, 0, 0};
         st[ax] = first;
         s.off = st[0] + n[0] * (st[1] + n[1] * st[2]);
         for (int a = 0; a < 3; a++) s.e[a] = (int)n[a];
         s.e[ax] = level;
         return s;
      };
      const Strip send_m = strip(level), send_p = strip(n[ax] - 2 * level), recv_m = strip(0), recv_p = strip(n[ax] - level);
      const size_t sbytes = (size_t)es * send_m.e[0] * send_m.e[1] * send_m.e[2];
      auto ghost_from = [&](const Strip &dst, const char *src_packed) { // packed strip -> ghost layers
         launch_box_copy(ctx, o + (size_t)es * dst.off, n[0], n[1], src_packed, dst.e[0], dst.e[1], dst.e[0], dst.e[1], dst.e[2], es);
      };
      auto pack = [&](const Strip &src, char *dst_packed) {
         launch_box_copy(ctx, dst_packed, src.e[0], src.e[1], o + (size_t)es * src.off, n[0], n[1], src.e[0], src.e[1], src.e[2], es);
      };
      if (im < 0 && ip < 0) continue;
      // work buffers: [to-minus strip][to-plus strip] in work[0], [from-minus][from-plus] in work[1]
      ctx->ensure_buffers(2, need, false);
      char *w0 = (char *)ctx->work[0], *w1 = (char *)ctx->work[1];
      if (np == 1) { // periodic with a single rank along this dimension: the neighbours are this rank itself
         ctx->wait_buffer_idle(0, ctx->stream);
         pack(send_p, w0);
         pack(send_m, w0 + sbytes);
         ghost_from(recv_m, w0);          // what arrives from "minus" is the to-plus strip of that neighbour
         ghost_from(recv_p, w0 + sbytes);
         continue;
      }
      ctx->wait_buffer_idle(0, ctx->stream);
      if (im >= 0) pack(send_m, w0);
      if (ip >= 0) pack(send_p, w0 + sbytes);
      // messages: one per DISTINCT peer (with two ranks and periodicity both neighbours are the same rank: one message
      // carries both strips).  A peer stores what it gets from this rank: my to-minus strip in its from-plus slot (offset
      // sbytes), my to-plus strip in its from-minus slot (offset 0).
      std::vector<PeerXfer> xf;
      std::vector<size_t> dst_off;
      const bool p2p = p2p_active(ctx);
      if (im >= 0 && ip >= 0 && im == ip) {
         // same peer on both sides: send [to-minus][to-plus]; it arrives as the peer's [from-plus][from-minus], so the peer
         // -- and, symmetrically, this rank -- receives the pair swapped: slot 0 = from-plus, slot 1 = from-minus
         PeerXfer x;
         x.peer = rank_of(im);
         x.sendptr = w0; x.sendbytes = 2 * sbytes;
         x.recvptr = w1; x.recvbytes = 2 * sbytes;
         xf.push_back(x);
         dst_off.push_back(0);
      } else {
         if (im >= 0) {
            PeerXfer x;
            x.peer = rank_of(im);
            x.sendptr = w0; x.sendbytes = sbytes;
            x.recvptr = w1; x.recvbytes = sbytes; // from-minus slot
            xf.push_back(x);
            dst_off.push_back(sbytes); // lands in the peer's from-plus slot
         }
         if (ip >= 0) {
            PeerXfer x;
            x.peer = rank_of(ip);
            x.sendptr = w0 + sbytes; x.sendbytes = sbytes;
            x.recvptr = w1 + sbytes; x.recvbytes = sbytes; // from-plus slot
            xf.push_back(x);
            dst_off.push_back(0); // lands in the peer's from-minus slot
         }
      }
      if (p2p) p2p_exchange(ctx, xf, dst_off, 1, 0);
      else ctx->tr->exchange(xf, ctx->stream);
      if (im >= 0 && ip >= 0 && im == ip) {
         ghost_from(recv_p, w1);          // the peer's to-minus strip
         ghost_from(recv_m, w1 + sbytes); // the peer's to-plus strip
      } else {
         if (im >= 0) ghost_from(recv_m, w1);
         if (ip >= 0) ghost_from(recv_p, w1 + sbytes);
      }
   }
}

} // namespace d2d
