// transpose.cpp -- the four stand-alone pencil transposes (the bare transpose_* API).
// Replaces transpose_{x_to_y,y_to_x,y_to_z,z_to_y}_{real,complex}_long and their workers
// (src/transpose_x_to_y.f90:25-135, transpose_y_to_x.f90:25-135, transpose_y_to_z.f90:25-185,
// transpose_z_to_y.f90:25-186).  Element-size generic: a transpose is a bit-exact copy.
//
// Differences from the reference's GPU path (same results, fewer sweeps):
//  - pack / unpack are one kernel launch each instead of P cudaMemcpy2D calls;
//  - the block a rank sends to itself never moves through the exchange;
//  - y->z receives straight into dst and z->y sends straight from src (the reference stages both
//    through work buffers with full-array cudaMemcpy, transpose_y_to_z.f90:176-181,
//    transpose_z_to_y.f90:112-118);
//  - the exchange is stream-ordered (no host synchronisation inside).
#include "common.h"

namespace d2d {

static void pencil_space(const Decomp &d, int pencil, CopyArgs &c)
{
   const int *sz = pencil == 0 ? d.xsz : pencil == 1 ? d.ysz : d.zsz;
   if (pencil == 0) { c.ne = sz[0]; c.na = sz[1] * sz[2]; c.nb = 1; c.fast_is_a = 0; }
   else if (pencil == 1) { c.ne = sz[1]; c.na = sz[0]; c.nb = sz[2]; c.fast_is_a = 1; }
   else { c.ne = sz[2]; c.na = sz[0] * sz[1]; c.nb = 1; c.fast_is_a = 1; }
}

void transpose(Ctx *ctx, const Decomp &d, int direction, int es, const void *src, void *dst)
{
   static const int kFrom[4] = {0, 1, 2, 1}, kTo[4] = {1, 2, 1, 0};
   static const char *kName[4] = {"transp_x_y", "transp_y_z", "transp_z_y", "transp_y_x"};
   D2D_REQUIRE(direction >= 0 && direction < 4, "invalid transpose direction");
   const int from = kFrom[direction], to = kTo[direction];
   D2D_CHECK_CUDA(cudaSetDevice(ctx->device));
   ProfScope ps(ctx, kName[direction], 2.0 * es * (double)d.pencil_elems(from));
   const int np = comm_size(d, from, to);
   if (np == 1) { // dims==1: dst = src (transpose_x_to_y.f90:41-50 etc.)
      D2D_CHECK_CUDA(cudaMemcpyAsync(dst, src, (size_t)es * d.pencil_elems(from), cudaMemcpyDeviceToDevice, ctx->stream));
      return;
   }
   const int me = (from == 0 || to == 0) ? d.c1 : d.c2;
   void *nsrc = const_cast<void *>(src);
   CopyArgs c{};
   // pack / receive buffers: the context's work buffers 0 and 1, sized for the largest pencil of ANY rank so that every rank
   // grows (and republishes to its peers) at the same call -- a rank-local size would leave the peers with stale mappings
   const bool even = ctx->even;
   ctx->ensure_buffers(2, uniform_pencil_bytes(ctx, d, es, even), false);
   void *w1 = ctx->work[0], *w2 = ctx->work[1];
   const bool p2p = p2p_active(ctx);
   if (even) {
      // EVEN builds of the reference: every message padded to one count per communicator (MPI_ALLTOALL instead of
      // MPI_ALLTOALLV, src/transpose_x_to_y.f90:99-110), segment m of both buffers at m * count; y<->z stage through the work
      // buffers like x<->y (the Z-pencil is no longer the receive / send buffer itself).  Same pencils as the default layout.
      ctx->wait_buffer_idle(0, ctx->stream);
      c.in = natural_map(d, from, nsrc);
      c.out = send_map(d, from, to, w1, es, true);
      pencil_space(d, from, c);
      launch_copy(ctx, c, es, "pack");
      exchange(ctx, d, from, to, w1, w2, es, 0, 1, true);
      c.in = recv_map(d, from, to, w2, w1, es, true);
      c.out = natural_map(d, to, dst);
      pencil_space(d, to, c);
      launch_copy(ctx, c, es, "unpack");
      return;
   }
   switch (direction) {
   case D2D_X_TO_Y:
   case D2D_Y_TO_X: {
      ctx->wait_buffer_idle(0, ctx->stream);
      c.in = natural_map(d, from, nsrc);
      c.out = send_map(d, from, to, w1, es);
      pencil_space(d, from, c);
      launch_copy(ctx, c, es, "pack");
      exchange(ctx, d, from, to, w1, w2, es, 0, 1);
      c.in = recv_map(d, from, to, w2, w1, es);
      c.out = natural_map(d, to, dst);
      pencil_space(d, to, c);
      launch_copy(ctx, c, es, "unpack");
      break;
   }
   case D2D_Y_TO_Z: { // the receive buffer IS the Z pencil (transpose_y_to_z.f90:538)
      ctx->wait_buffer_idle(0, ctx->stream);
      c.in = natural_map(d, 1, nsrc);
      c.out = send_map(d, 1, 2, w1, es);
      c.out.ptr[me] = (char *)dst + (size_t)es * d.z2disp[me]; // own block goes straight home
      pencil_space(d, 1, c);
      launch_copy(ctx, c, es, "pack");
      if (!p2p) {
         exchange(ctx, d, 1, 2, w1, dst, es, 0, -1);
      } else { // peers push into the mapped work buffer; the blocks (contiguous z-slabs) then move home
         exchange(ctx, d, 1, 2, w1, w2, es, 0, 1);
         for (int m = 0; m < d.p_col; m++)
            if (m != me && d.z2cnts[m])
               D2D_CHECK_CUDA(cudaMemcpyAsync((char *)dst + (size_t)es * d.z2disp[m], (char *)w2 + (size_t)es * d.z2disp[m],
                                              (size_t)es * d.z2cnts[m], cudaMemcpyDeviceToDevice, ctx->stream));
      }
      break;
   }
   case D2D_Z_TO_Y: { // the send buffer IS the Z pencil (transpose_z_to_y.f90:418)
      exchange(ctx, d, 2, 1, src, w2, es, -1, 1);
      c.in = recv_map(d, 2, 1, w2, nsrc, es);
      c.out = natural_map(d, 1, dst);
      pencil_space(d, 1, c);
      launch_copy(ctx, c, es, "unpack");
      break;
   }
   }
}

} // namespace d2d
