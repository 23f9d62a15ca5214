// p2p.cpp -- peer-memory plumbing for the fused FFT + exchange path (one process per GPU, NVLink/NVSwitch).
//
// Inside decomp_2d_fft_3d the all-to-all of a transpose (src/decomp_2d_nccl.f90:214-473 in the reference: grouped
// ncclSend/ncclRecv after a pack pass) disappears as a separate step: the store map of the producing FFT kernel points
// piece m straight INTO RANK m's receive buffer (CUDA-IPC mapped), so the transfer happens tile by tile while the
// kernel computes.  This file provides what that needs:
//   * exchange of cudaIpcMemHandle_t of the work buffers and of a small flag array (all-gather over the transport),
//   * stream-ordered flags (cuStreamWriteValue32 / cuStreamWaitValue32 on peer-mapped memory):
//       ready_from[r] >= e : rank r's stream reached the producer of exchange e  -> its receive buffer may be written
//       done_from[r]  >= e : rank r's producer kernel of exchange e completed    -> its block has landed here
// No host synchronisation, no SMs spent on communication.
#include <cuda.h>

#include <cstring>

#include "common.h"

namespace d2d {

namespace {
typedef CUresult (*WriteValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
typedef CUresult (*WaitValueFn)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);

void *driver_fn(const char *name)
{
   void *p = nullptr;
   cudaDriverEntryPointQueryResult q;
   if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
   return p;
}

struct Published {
   cudaIpcMemHandle_t work[kWorkBuffers];
   unsigned long long bytes[kWorkBuffers];
   cudaIpcMemHandle_t flags;
   int device;
   int ok;
};
} // namespace

struct P2P {
   int nranks = 0, rank = 0;
   bool ok = false;
   void *peer_work[kWorkBuffers][kMaxRanks] = {};
   size_t peer_bytes[kWorkBuffers][kMaxRanks] = {};
   void *opened[kWorkBuffers][kMaxRanks] = {};
   uint32_t *flags = nullptr; // local: [0..nranks) ready_from, [nranks..2 nranks) done_from, [2 nranks..3 nranks) arrived_from (chunk sequence numbers), [3 nranks..4 nranks) pushed_from (counters)
   uint32_t *peer_flags[kMaxRanks] = {};
   void *opened_flags[kMaxRanks] = {};
   bool flags_published = false;
   uint32_t epoch = 0, seq = 0;
   uint32_t pushes_expected[kMaxRanks] = {};
   WriteValueFn write_value = nullptr;
   WaitValueFn wait_value = nullptr;
};

static void close_work(P2P *p)
{
   for (int w = 0; w < kWorkBuffers; w++)
      for (int r = 0; r < p->nranks; r++) {
         if (p->opened[w][r]) cudaIpcCloseMemHandle(p->opened[w][r]);
         p->opened[w][r] = nullptr;
         p->peer_work[w][r] = nullptr;
         p->peer_bytes[w][r] = 0;
      }
}

void p2p_destroy(P2P *p)
{
   if (!p) return;
   close_work(p);
   for (int r = 0; r < p->nranks; r++)
      if (p->opened_flags[r]) cudaIpcCloseMemHandle(p->opened_flags[r]);
   if (p->flags) cudaFree(p->flags);
   delete p;
}

// Collective over all ranks of the context: publish the current work buffers (and, once, the flag array).
// Must be called by every rank in the same order (plan creation is collective, like communicator creation).
void p2p_publish(Ctx *ctx)
{
   static const bool enabled = getenv("D2D_P2P") ? atoi(getenv("D2D_P2P")) != 0 : true;
   if (ctx->nranks <= 1 || !ctx->tr || ctx->tr->kind() == D2D_TRANSPORT_LOCAL) return;
   if (!enabled && ctx->tr->kind() == D2D_TRANSPORT_NCCL) return; // D2D_P2P=0: the NCCL exchange path
   D2D_REQUIRE(ctx->nranks <= kMaxRanks, "too many ranks for the peer-memory path");
   if (!ctx->p2p) {
      ctx->p2p = new P2P();
      ctx->p2p->nranks = ctx->nranks;
      ctx->p2p->rank = ctx->rank;
      ctx->p2p->write_value = (WriteValueFn)driver_fn("cuStreamWriteValue32");
      ctx->p2p->wait_value = (WaitValueFn)driver_fn("cuStreamWaitValue32");
      D2D_CHECK_CUDA(cudaMalloc((void **)&ctx->p2p->flags, 4 * kMaxRanks * sizeof(uint32_t)));
      D2D_CHECK_CUDA(cudaMemset(ctx->p2p->flags, 0, 4 * kMaxRanks * sizeof(uint32_t)));
   }
   P2P *p = ctx->p2p;
   D2D_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
   for (int k = 0; k < 2 * kMaxP; k++)
      if (ctx->copy_stream[k]) D2D_CHECK_CUDA(cudaStreamSynchronize(ctx->copy_stream[k]));
   if (ctx->push_stream) D2D_CHECK_CUDA(cudaStreamSynchronize(ctx->push_stream));
   close_work(p);
   Published mine;
   memset(&mine, 0, sizeof(mine));
   mine.ok = (p->write_value && p->wait_value) ? 1 : 0;
   mine.device = ctx->device;
   for (int w = 0; w < kWorkBuffers; w++) {
      mine.bytes[w] = ctx->work_bytes[w];
      if (ctx->work[w] && cudaIpcGetMemHandle(&mine.work[w], ctx->work[w]) != cudaSuccess) mine.ok = 0;
   }
   if (cudaIpcGetMemHandle(&mine.flags, p->flags) != cudaSuccess) mine.ok = 0;
   cudaGetLastError();
   std::vector<Published> all(p->nranks);
   ctx->tr->allgather(&mine, all.data(), sizeof(Published), ctx->stream);
   int ok = 1;
   for (int r = 0; r < p->nranks; r++) ok = ok && all[r].ok;
   if (ok) {
      for (int r = 0; r < p->nranks && ok; r++) {
         if (r == p->rank) {
            for (int w = 0; w < kWorkBuffers; w++) { p->peer_work[w][r] = ctx->work[w]; p->peer_bytes[w][r] = ctx->work_bytes[w]; }
            p->peer_flags[r] = p->flags;
            continue;
         }
         for (int w = 0; w < kWorkBuffers && ok; w++) {
            if (!all[r].bytes[w]) continue;
            void *q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[r].work[w], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; break; }
            p->opened[w][r] = q;
            p->peer_work[w][r] = q;
            p->peer_bytes[w][r] = (size_t)all[r].bytes[w];
         }
         if (ok && !p->flags_published) {
            void *q = nullptr;
            if (cudaIpcOpenMemHandle(&q, all[r].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) ok = 0;
            else { p->opened_flags[r] = q; p->peer_flags[r] = (uint32_t *)q; }
         }
      }
      cudaGetLastError();
   }
   // every rank must take the same path: agree on the outcome
   std::vector<int> oks(p->nranks);
   ctx->tr->allgather(&ok, oks.data(), sizeof(int), ctx->stream);
   for (int r = 0; r < p->nranks; r++) ok = ok && oks[r];
   if (ok) p->flags_published = true;
   else close_work(p);
   p->ok = ok != 0;
   D2D_REQUIRE(p->ok || ctx->tr->kind() != D2D_TRANSPORT_BOOT,
               "bootstrap transport: CUDA IPC / stream memory operations are unavailable between the ranks' devices, and there is no other data plane");
}

void p2p_unpublish(Ctx *ctx)
{
   P2P *p = ctx->p2p;
   if (!p || ctx->nranks <= 1 || !ctx->tr) return;
   D2D_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
   for (int k = 0; k < 2 * kMaxP; k++)
      if (ctx->copy_stream[k]) D2D_CHECK_CUDA(cudaStreamSynchronize(ctx->copy_stream[k]));
   if (ctx->push_stream) D2D_CHECK_CUDA(cudaStreamSynchronize(ctx->push_stream));
   close_work(p);
   p->ok = false;
   int one = 1;
   std::vector<int> all(p->nranks);
   ctx->tr->allgather(&one, all.data(), sizeof(int), ctx->stream); // host-level barrier: everybody has closed
}

bool p2p_active(const Ctx *ctx) { return ctx->p2p && ctx->p2p->ok; }
void p2p_invalidate(Ctx *ctx)
{
   if (ctx->p2p) ctx->p2p->ok = false;
}
void *p2p_peer_work(const Ctx *ctx, int w, int rank) { return ctx->p2p->peer_work[w][rank]; }
size_t p2p_peer_bytes(const Ctx *ctx, int w, int rank) { return ctx->p2p->peer_bytes[w][rank]; }
uint32_t p2p_next_epoch(Ctx *ctx) { return ++ctx->p2p->epoch; }
uint32_t p2p_reserve_seq(Ctx *ctx, uint32_t n)
{
   const uint32_t base = ctx->p2p->seq;
   ctx->p2p->seq += n;
   return base;
}

uint32_t p2p_expect_push(Ctx *ctx, int peer)
{
   P2P *p = ctx->p2p;
   p->pushes_expected[peer] += (uint32_t)push_ctas();
   return p->pushes_expected[peer];
}
uint32_t *p2p_push_flag(Ctx *ctx, int peer) { return ctx->p2p->peer_flags[peer] + 3 * ctx->p2p->nranks + ctx->p2p->rank; }

// tell `peer` that stream `st` of this rank (default: the context's stream) reached the point `which` with `value`
void p2p_signal(Ctx *ctx, int peer, int which, uint32_t value, cudaStream_t st)
{
   P2P *p = ctx->p2p;
   uint32_t *addr = p->peer_flags[peer] + which * p->nranks + p->rank;
   CUresult r = p->write_value((CUstream)(st ? st : ctx->stream), (CUdeviceptr)(uintptr_t)addr, value, 0);
   D2D_REQUIRE(r == CUDA_SUCCESS, "cuStreamWriteValue32 failed");
}
// make stream `st` of this rank wait until `peer` signalled `which` with at least `value`
void p2p_wait(Ctx *ctx, int peer, int which, uint32_t value, cudaStream_t st)
{
   P2P *p = ctx->p2p;
   uint32_t *addr = p->flags + which * p->nranks + peer;
   CUresult r = p->wait_value((CUstream)(st ? st : ctx->stream), (CUdeviceptr)(uintptr_t)addr, value, CU_STREAM_WAIT_VALUE_GEQ);
   D2D_REQUIRE(r == CUDA_SUCCESS, "cuStreamWaitValue32 failed");
}

// All-to-all(v) of the bare transposes over peer memory: copy-engine pushes into the peers' work[recv_w], one copy stream
// per peer, ordered by the flags: ready (the destination is done with the previous contents of its receive buffer) and
// arrived (the block has landed).  Replaces decomp_2d_nccl_alltoall_* (src/decomp_2d_nccl.f90:214-473) for these calls.
void p2p_exchange(Ctx *ctx, const std::vector<PeerXfer> &xf, const std::vector<size_t> &dst_off, int recv_w, int send_w)
{
   D2D_REQUIRE(p2p_active(ctx), "peer-memory exchange is not active");
   const uint32_t epoch = p2p_next_epoch(ctx);
   const uint32_t seq = p2p_reserve_seq(ctx, 1) + 1;
   for (const auto &x : xf) p2p_signal(ctx, x.peer, 0, epoch); // my receive buffer is free (stream order)
   cudaEvent_t packed = ctx->new_sync_event();
   D2D_CHECK_CUDA(cudaEventRecord(packed, ctx->stream));
   for (size_t i = 0; i < xf.size(); i++) {
      const PeerXfer &x = xf[i];
      cudaStream_t cs = ctx->copy_stream_for((int)i);
      D2D_CHECK_CUDA(cudaStreamWaitEvent(cs, packed, 0));
      p2p_wait(ctx, x.peer, 0, epoch, cs);
      char *dst = (char *)p2p_peer_work(ctx, recv_w, x.peer);
      D2D_REQUIRE(dst != nullptr && dst_off[i] + x.sendbytes <= p2p_peer_bytes(ctx, recv_w, x.peer), "peer-memory exchange: destination buffer too small");
      if (x.sendbytes) D2D_CHECK_CUDA(cudaMemcpyAsync(dst + dst_off[i], x.sendptr, x.sendbytes, cudaMemcpyDefault, cs));
      p2p_signal(ctx, x.peer, 2, seq, cs);
   }
   if (send_w >= 0) ctx->mark_buffer_busy(send_w, (int)xf.size());
   else { // the copies read a user array: the context's stream must not run ahead of them
      for (size_t i = 0; i < xf.size(); i++) {
         cudaEvent_t e = ctx->new_sync_event();
         D2D_CHECK_CUDA(cudaEventRecord(e, ctx->copy_stream_for((int)i)));
         D2D_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, e, 0));
      }
   }
   for (const auto &x : xf) p2p_wait(ctx, x.peer, 2, seq);
}

} // namespace d2d
