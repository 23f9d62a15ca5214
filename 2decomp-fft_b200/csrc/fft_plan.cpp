// fft_plan.cpp -- FFT plans and the 3-D drivers.
// Replaces the engine / plan management of the cuFFT backend (src/fft_cufft.f90:37-61 engine type,
// :263-431 init_fft_engine with its 12 cufft plans, :436-483 finalize/use) and the GPU 3-D drivers
// fft_3d_c2c / fft_3d_r2c / fft_3d_c2r (src/fft_cufft.f90:676-790, 795-934, 939-1170; CPU twins in
// src/fft_common_3d.f90).  Stage order and the dims==1 special cases follow those drivers; what
// differs is the data path: each 1-D stage reads through a receive-layout piece map and writes
// through a send-layout piece map, so the transposes shrink to the bare exchange.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>

#include "common.h"
#include "fft_registry.h"

namespace d2d {

bool fft_v2_try_launch(Ctx *ctx, const FftArgs &g, int f64, int mode, cudaError_t *err);
cudaError_t fft_any_launch(Ctx *ctx, const FftArgs &g, int f64, int mode); // fft_any.cu: lengths without a compiled kernel

// Gates between the host-array entry points and the chains: the user's input arrives over PCIe in pieces and the
// output leaves in pieces.  A user X-pencil is contiguous in z-slabs, and z is a batch axis of the x stage, so a first /
// last stage on the X-pencil runs slab by slab as the slabs arrive / leaves slab by slab as they are produced.
struct IoGate {
   int nio = 8;                                    // slabs of an x stage that would otherwise run whole
   std::function<void(size_t, size_t)> need_in;   // the compute stream must not read bytes [lo, hi) of `in` before they arrived
   std::function<void(size_t, size_t)> have_out;  // bytes [lo, hi) of `out` are final once the work enqueued so far has run
};

struct Plan {
   Ctx *ctx;
   int format, nx, ny, nz, f64, inplace;
   int skip[3];
   d2d_decomp ph, sp;
   // device staging for the *_host entry points: one (in, out) pair per kind of transform (0 r2c, 1 c2r, 2 c2c) so that
   // consecutive calls overlap (the download of one call with the upload of the next)
   struct HostIo {
      void *in = nullptr, *out = nullptr;
      size_t in_bytes = 0, out_bytes = 0;
      cudaEvent_t in_done = nullptr, out_done = nullptr; // last compute that read `in` / last download that read `out`
   } io[3];
   const IoGate *gate = nullptr; // set around a chain by fft_3d_host
};

static const int *pencil_size(const Decomp &d, int p) { return p == 0 ? d.xsz : p == 1 ? d.ysz : d.zsz; }

// Picks the kernel of one batched 1-D transform described by g (maps, batch extents, n, flags) and launches it on the
// context's stream: a compiled plan (TMA-staged when the alignment rules hold, cp.async otherwise) or the any-length kernel.
// Also the inner transform of the Bluestein path (fft_any.cu).
cudaError_t fft_dispatch(Ctx *ctx, FftArgs &g, int f64, int mode, int kind, int pairvec, bool wide_real)
{
   const int n = g.n;
   const FftKernelInfo *k = nullptr;
   if (wide_real) k = fft_find(n, f64, KIND_TILE_WIDE, mode, pairvec);
   // inputs that are contiguous along the transform axis land through the line-major kernels
   const bool line_like = kind == KIND_TILE && ((mode == MODE_C2C && g.in.se[0] == 1) || (mode == MODE_R2C && g.rse == 1));
   if (!k && line_like) k = fft_find(n, f64, kind, mode, pairvec, 1);
   if (!k) k = fft_find(n, f64, kind, mode, pairvec);
   if (k) g.tw = twiddles_for(ctx->device, n, f64);
   cudaError_t e = cudaSuccess;
   if (!k) e = fft_any_launch(ctx, g, f64, mode); // no compiled plan for this length: mixed-radix shared-memory kernel
   else if (!fft_v2_try_launch(ctx, g, f64, mode, &e)) e = k->launch(g, ctx->stream);
   if (e == cudaSuccess) ctx->launches++;
   return e;
}

// One batched 1-D stage along `pencil` of the complex-side decomp dc (real side: dr, R2C/C2R only).
// This is c2c_1m_{x,y,z} / r2c_1m_{x,z} / c2r_1m_{x,z} (src/fft_cufft.f90:489-671) fused with the
// neighbouring mem_split_* / mem_merge_* through the maps.  `chain`: the stage is part of a 3-D
// transform and uses the (a,b) batch convention of the private wire layouts (decomp.cpp); otherwise
// it works on one dense local array.
// `rg` restricts a chain stage to lines f0 <= f < f1 of one batch axis (0 = a, 1 = b; -1 = the whole batch): the chunks of the
// overlapped chain (run_chain_overlap).  Along `a` a real transform needs an even f0 (real lines are transformed in pairs).
struct StageRange {
   int axis = -1, f0 = 0, f1 = 0;
};

static void run_stage(Ctx *ctx, int f64, int mode, int pencil, const Decomp &dc, const Decomp *dr, const PieceMap &in,
                      const PieceMap &out, void *rptr, int backward, int passthrough, bool chain, StageRange rg = StageRange())
{
   const int *cs = pencil_size(dc, pencil);
   FftArgs g{};
   g.in = in;
   g.out = out;
   g.backward = backward;
   g.passthrough = passthrough;
#ifdef D2D_DEBUG_KNOBS // experiments only (make EXTRA=-DD2D_DEBUG_KNOBS): kernels that drop their loads / stores return garbage
   static const int dbg = getenv("D2D_DEBUG_SKIP") ? atoi(getenv("D2D_DEBUG_SKIP")) : 0;
   g.debug = dbg;
#else
   g.debug = 0;
#endif
   g.sm_limit = ctx->fft_grid_limit;
   int kind;
   if (chain) {
      fft_stage_batch(dc, pencil, g.na, g.nb);
      kind = KIND_TILE;
   } else if (pencil == 0) { g.na = cs[1] * cs[2]; g.nb = 1; kind = KIND_LINE; }
   else if (pencil == 1) { g.na = cs[0]; g.nb = cs[2]; kind = KIND_TILE; }
   else { g.na = cs[0] * cs[1]; g.nb = 1; kind = KIND_TILE; }
   const int rs = f64 ? 8 : 4;
   if (rg.axis >= 0) {
      D2D_REQUIRE(chain && rg.f0 >= 0 && rg.f0 <= rg.f1 && rg.f1 <= (rg.axis == 0 ? g.na : g.nb), "invalid stage range");
      D2D_REQUIRE(mode == MODE_C2C || rg.axis == 1 || rg.f0 % 2 == 0, "a real stage starts its chunks on even lines");
      auto advance = [&](PieceMap &m) {
         for (int q = 0; q < m.np; q++)
            m.ptr[q] = (char *)m.ptr[q] + (long long)(2 * rs) * rg.f0 * (rg.axis == 0 ? m.sa[q] : m.sb[q]);
      };
      advance(g.in);
      advance(g.out);
      (rg.axis == 0 ? g.na : g.nb) = rg.f1 - rg.f0;
   }
   const long long lines = (long long)g.na * g.nb;
   int n = cs[pencil];
   int pairvec = 0;
   double bytes = 2.0 * (double)lines * n * 2 * rs;
   if (mode != MODE_C2C) {
      D2D_REQUIRE(dr != nullptr && (pencil == 0 || pencil == 2), "real transforms run along x or z only");
      const int *rsz = pencil_size(*dr, pencil);
      n = rsz[pencil];
      D2D_REQUIRE(cs[pencil] == n / 2 + 1, "complex extent must be n/2+1 along the real-transform axis");
      g.na_real = g.na;
      g.na = (g.na_real + 1) / 2;
      const long long r1 = rsz[0], r12 = (long long)rsz[0] * rsz[1];
      if (pencil == 0) { // lines along x; chain: a = y, b = z; dense: a = (y,z) flattened
         g.rse = 1; g.rsa = r1; g.rsb = chain ? r12 : 0;
      } else { // lines along z; chain: a = x, b = y; dense: a = (x,y) flattened
         g.rse = r12; g.rsa = 1; g.rsb = chain ? r1 : 0;
      }
      if (rg.axis >= 0) rptr = (char *)rptr + (long long)rs * rg.f0 * (rg.axis == 0 ? g.rsa : g.rsb);
      g.rptr = rptr;
      if (pencil != 0) {
         // pairs (a, a+1) are adjacent reals: one vector access when every pair is 2*sizeof(T) aligned
         const bool even_rows = chain ? (r1 % 2 == 0) : (r12 % 2 == 0);
         pairvec = (even_rows && ((uintptr_t)rptr % (2 * rs)) == 0) ? 1 : 0;
      }
      bytes = (double)lines * n * rs + (double)lines * (n / 2 + 1) * 2 * rs;
   }
   g.n = n;
   if (lines == 0) return;
   static const bool wide = getenv("D2D_WIDE_REAL_TILES") && atoi(getenv("D2D_WIDE_REAL_TILES")) != 0;
   static const char *axes = "xyz";
   char label[32];
   snprintf(label, sizeof(label), "fft_%s_%c%s", mode == MODE_C2C ? "c2c" : mode == MODE_R2C ? "r2c" : "c2r", axes[pencil],
            (mode == MODE_C2C && chain) ? (backward ? "_bwd" : "_fwd") : "");
   ProfScope ps(ctx, label, bytes);
   const cudaError_t e = fft_dispatch(ctx, g, f64, mode, kind, pairvec, wide && chain && mode != MODE_C2C && pencil == 2 && pairvec);
   if (e != cudaSuccess) throw Error(1000 + (int)e, std::string("FFT kernel launch failed: ") + cudaGetErrorString(e));
}

struct StageDef {
   int pencil, mode;
};

// Stage s of a chain, with the host-array gates: an x stage that reads the user's input (s == 0) or writes the user's
// output (s == 2) runs slab by slab along z (batch axis b), waiting for / announcing the byte range of each slab.
static void run_chain_stage(Plan &p, int s, int mode, int pencil, const Decomp &dc, const Decomp *dr, const PieceMap &in, const PieceMap &out,
                            void *rptr, int backward, int passthrough, StageRange rg = StageRange())
{
   Ctx *ctx = p.ctx;
   const IoGate *g = p.gate;
   const bool gin = g && g->need_in && s == 0 && pencil == 0;
   const bool gout = g && g->have_out && s == 2 && pencil == 0;
   if (!gin && !gout) return run_stage(ctx, p.f64, mode, pencil, dc, dr, in, out, rptr, backward, passthrough, true, rg);
   D2D_REQUIRE(rg.axis != 0, "an x stage is chunked along z only");
   const int rs = p.f64 ? 8 : 4;
   const Decomp &di = (mode == MODE_R2C) ? *dr : dc, &dq = (mode == MODE_C2R) ? *dr : dc;
   const size_t plane_in = (size_t)(mode == MODE_R2C ? rs : 2 * rs) * di.xsz[0] * di.xsz[1];
   const size_t plane_out = (size_t)(mode == MODE_C2R ? rs : 2 * rs) * dq.xsz[0] * dq.xsz[1];
   int na = 0, nb = 0;
   fft_stage_batch(dc, pencil, na, nb);
   std::vector<int> cut;
   if (rg.axis == 1) { cut.push_back(rg.f0); cut.push_back(rg.f1); }
   else {
      const int k = std::max(1, std::min(g->nio, nb));
      for (int c = 0; c <= k; c++) cut.push_back((int)((long long)nb * c / k));
   }
   for (size_t c = 0; c + 1 < cut.size(); c++) {
      if (cut[c] == cut[c + 1]) continue;
      if (gin) g->need_in(plane_in * cut[c], plane_in * cut[c + 1]);
      StageRange r;
      r.axis = 1; r.f0 = cut[c]; r.f1 = cut[c + 1];
      run_stage(ctx, p.f64, mode, pencil, dc, dr, in, out, rptr, backward, passthrough, true, r);
      if (gout) g->have_out(plane_out * cut[c], plane_out * cut[c + 1]);
   }
}

// Generic 3-stage chain.  `in`/`out` are the user arrays (dense pencils); exactly one of them is real
// for r2c/c2r.  Between two stages the data lives in a work buffer in the private wire layout of
// that link (decomp.cpp): the producer writes its send blocks (the block it keeps for itself
// included), the exchange moves the other blocks into a second buffer, the consumer gathers from
// both.  With a 1-rank communicator the exchange vanishes and the consumer reads the producer's
// buffer.  Three work buffers rotate; a single-rank complex-output transform borrows `out` as the
// first one.
static void run_chain_p2p(Plan &p, const Decomp &dc, const Decomp *dr, const StageDef st[3], void *in, void *out, int backward);
static size_t uniform_work_bytes(const Plan &p, bool c2c);

// ---- chunk-wise pipelining of the exchanges with their neighbouring stages ---------------------------------------------
// The default of every multi-rank chain.  Every link with more than one rank is cut into K chunks along its free axis
// (fft_link_chunk: contiguous sub-ranges of every block).  The producer runs chunk by chunk on the context's stream at HBM
// speed into a local send buffer (its own block included: that one never moves); as soon as chunk c is written its
// sub-ranges travel while the producer works on chunk c + 1, and the consumer of the link follows chunk by chunk as they
// land -- unless it is itself the chunked producer of the next link (the middle stage of a p_row > 1, p_col > 1 grid: its
// input chunks run along x, its output chunks along z), in which case it starts after the last chunk of its input.
//
// Two data planes:
//  * peer memory (default with one process per GPU): the COPY ENGINES push every sub-range straight into the destination
//    rank's receive buffer over NVLink (CUDA-IPC mapped, p2p.cpp), one copy stream per peer, so the SMs only ever run
//    HBM-bound FFT kernels and the link runs at the DMA rate while they compute.  Ranks are ordered by stream-ordered
//    flags: ready (epoch e: the destination is done with the previous contents of its receive buffer; waited for by the
//    copy stream, never by the compute stream) and arrived (chunk sequence number: the sub-range has landed; waited for by
//    the consumer).  No host synchronisation, no SMs spent on communication, no NCCL.
//  * a transport (NCCL grouped send/recv with D2D_P2P=0, or the in-process transport of the thread-per-rank groups): the
//    same pipeline with the chunk exchanges on a communication stream, ordered by events.
// Replaces the pack -> all-to-all -> host sync -> unpack sequence of the reference (src/decomp_2d_nccl.f90:214-252 and the
// cudaStreamSynchronize at :249, :324, :395, :470).  Link s uses work buffer 2 s as send and 2 s + 1 as receive buffer.
static int pipe_chunks()
{
   const char *v = getenv("D2D_CHUNKS"); // read per call: tests switch it at run time
   if (!v) v = getenv("D2D_OVERLAP");    // round-1 name
   const int k = v ? atoi(v) : 4;
   return k > 64 ? 64 : k < 1 ? 1 : k;
}

static void run_chain_pipe(Plan &p, const Decomp &dc, const Decomp *dr, const StageDef st[3], void *in, void *out, int backward, int K)
{
   Ctx *ctx = p.ctx;
   const int es = p.f64 ? 16 : 8;
   const int padq = 128 / es;
   const size_t wbytes = uniform_work_bytes(p, dr == nullptr);
   const bool ce = p2p_active(ctx);
   // peer-memory plane: the chunks are pushed by the copy engines (default: one cudaMemcpyAsync per peer and chunk, 780 GB/s
   // for large copies) or by the exchange kernel (D2D_PUSH=sm, push_kernels.cu: one launch per chunk for all peers; SM-issued
   // writes reach 715 GB/s on NVLink and one driving thread per CTA sustains ~45 GB/s, so it needs 16+ SMs)
   static const bool push_sm = getenv("D2D_PUSH") && std::string(getenv("D2D_PUSH")) == "sm";
   const bool sm = ce && push_sm;
   if (!ce && !ctx->comm_stream) D2D_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking));
   struct LimitGuard { // the FFT kernels leave push_ctas() SMs to the exchange kernels for the duration of the chain
      Ctx *c;
      ~LimitGuard() { c->fft_grid_limit = 0; }
   } limit_guard{ctx};
   if (sm) {
      if (!ctx->push_stream) {
         int lo = 0, hi = 0;
         D2D_CHECK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
         D2D_CHECK_CUDA(cudaStreamCreateWithPriority(&ctx->push_stream, cudaStreamNonBlocking, hi));
         D2D_CHECK_CUDA(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, ctx->device));
      }
      ctx->fft_grid_limit = std::max(1, ctx->sm_count - push_ctas());
   }
   const bool real_link[2] = {comm_size(dc, st[0].pencil, st[1].pencil) > 1, comm_size(dc, st[1].pencil, st[2].pencil) > 1};
   // chunk boundaries of a link: K pieces of the free axis, cut on multiples of 16 lines (tiles and real pairs stay whole)
   // Always K chunks (empty ones allowed): every rank of the grid then issues the same number of exchanges whatever its
   // own extents are, which the in-process transport (one barrier over all ranks per exchange) and the chunk sequence
   // numbers of the peer-memory path rely on.
   // D2D_CHUNK_EDGE < 1 makes the first and the last chunk shorter than the others (the exchange cannot start before the first
   // chunk is written, the consumer's last chunk cannot start before the last one has landed); measured on 2 B200s at 1024^3 it
   // changes nothing (10.9-11.1 ms per pair for 0.3 / 0.5 / 1.0): the chain is bound by HBM traffic there, not by those edges.
   static const double edge = getenv("D2D_CHUNK_EDGE") ? atof(getenv("D2D_CHUNK_EDGE")) : 1.0;
   auto bounds = [&](int nf) {
      std::vector<int> b(K + 1, 0);
      const double total = K <= 2 ? (double)K : (K - 2) + 2 * edge;
      double acc = 0;
      for (int c = 1; c < K; c++) {
         acc += (K <= 2 || (c > 1)) ? 1.0 : edge; // weight of chunk c - 1
         b[c] = std::max(b[c - 1], (int)((double)nf * acc / total) / 16 * 16);
      }
      b[K] = nf;
      return b;
   };
   for (int w = 0; w < kCtxBuffers; w++) D2D_REQUIRE(ctx->work_bytes[w] >= wbytes, "pipelined chain: work buffers were not reserved");
   // "ready" for every link of the chain right away: the receive buffers 1 and 3 were last read by the previous chain, which
   // precedes this point in stream order, so the peers never have to wait for this rank to REACH a link before they push
   uint32_t link_epoch[2] = {0, 0};
   if (ce)
      for (int s = 0; s < 2; s++) {
         if (!real_link[s]) continue;
         link_epoch[s] = p2p_next_epoch(ctx);
         const bool col = (st[s].pencil == 0 || st[s + 1].pencil == 0);
         const int np = col ? ctx->p_row : ctx->p_col, me = col ? ctx->c1 : ctx->c2;
         for (int k = 1; k < np; k++) {
            const int m = (me + k) % np;
            p2p_signal(ctx, col ? ctx->peer_rank_col(m) : ctx->peer_rank_row(m), 0, link_epoch[s]);
         }
      }
   PieceMap cur = fft_user_map(dc, st[0].pencil, in);
   // the link feeding the current stage
   struct Feed {
      bool chunked = false;
      std::vector<int> bounds;
      std::vector<cudaEvent_t> arrived; // transport plane: the exchange of chunk c has completed
      uint32_t seq_base = 0;            // copy engines: chunk c has landed when arrived_from[peer] >= seq_base + c + 1
      std::vector<std::vector<uint32_t>> pushed; // exchange kernel: ... when pushed_from[peers[i]] >= pushed[c][i]
      std::vector<int> peers;           // global ranks of the peers of the link
   } feed;
   auto wait_chunk = [&](int c) {
      ProfScope ps(ctx, "pipe_wait");
      if (sm) for (size_t i = 0; i < feed.peers.size(); i++) p2p_wait(ctx, feed.peers[i], 3, feed.pushed[c][i]);
      else if (ce) for (int r : feed.peers) p2p_wait(ctx, r, 2, feed.seq_base + (uint32_t)c + 1);
      else D2D_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, feed.arrived[c], 0));
   };
   for (int s = 0; s < 3; s++) {
      const int pen = st[s].pencil, mode = st[s].mode;
      const bool last = (s == 2);
      const int passthrough = (mode == MODE_C2C && p.skip[pen]) ? 1 : 0;
      PieceMap om{};
      void *rptr = nullptr, *sendbuf = nullptr;
      const int send_w = 2 * s, recv_w = 2 * s + 1;
      if (mode == MODE_R2C) rptr = in;
      if (last) {
         if (mode == MODE_C2R) rptr = out;
         else om = fft_user_map(dc, pen, out);
      } else {
         ctx->wait_buffer_idle(send_w, ctx->stream); // earlier pushes out of this buffer
         sendbuf = ctx->work[send_w];
         om = fft_link_map(dc, pen, st[s + 1].pencil, sendbuf, sendbuf, es, false, padq);
      }
      const bool produce_chunked = !last && real_link[s];
      if (produce_chunked) {
         const int nxt = st[s + 1].pencil;
         void *recvbuf = ctx->work[recv_w];
         if (feed.chunked) { // the whole input must be here before the first chunk
            wait_chunk((int)feed.bounds.size() - 2);
            feed = Feed();
         }
         LinkChunk probe;
         fft_link_chunk(dc, pen, nxt, padq, 0, 0, probe);
         const std::vector<int> b = bounds(probe.nf);
         const bool col = (pen == 0 || nxt == 0);
         const int np = probe.np, me = probe.me;
         Feed nf;
         nf.chunked = true;
         nf.bounds = b;
         std::vector<Decomp> dpeer(np);
         for (int k = 1; k < np; k++) {
            const int m = (me + k) % np;
            const int prank = col ? ctx->peer_rank_col(m) : ctx->peer_rank_row(m);
            nf.peers.push_back(prank);
            if (ce) decomp_init(dpeer[m], dc.nx, dc.ny, dc.nz, ctx->p_row, ctx->p_col, prank);
         }
         uint32_t epoch = 0;
         static const char *names[3][3] = {{"", "ce_x_y", ""}, {"ce_y_x", "", "ce_y_z"}, {"", "ce_z_y", ""}};
         std::vector<Ctx::Pending> spans(np);
         static const int lanes_env = getenv("D2D_COPY_LANES") ? atoi(getenv("D2D_COPY_LANES")) : 1;
         const int lanes = std::max(1, std::min(lanes_env, 2));
         std::vector<cudaEvent_t> ready_seen(np), flagged(np);
         if (ce) {
            epoch = link_epoch[s];
            if (sm) {
               nf.pushed.resize(K);
               for (int c = 0; c < K; c++)
                  for (int r : nf.peers) nf.pushed[c].push_back(p2p_expect_push(ctx, r));
            } else {
               nf.seq_base = p2p_reserve_seq(ctx, (uint32_t)K);
            }
         }
         for (int c = 0; c < K; c++) {
            StageRange rg;
            rg.axis = probe.axis_is_a ? 0 : 1; rg.f0 = b[c]; rg.f1 = b[c + 1];
            run_chain_stage(p, s, mode, pen, dc, dr, cur, om, rptr, backward, passthrough, rg);
            cudaEvent_t written = ctx->new_sync_event();
            D2D_CHECK_CUDA(cudaEventRecord(written, ctx->stream));
            LinkChunk S;
            fft_link_chunk(dc, pen, nxt, padq, b[c], b[c + 1], S);
            if (sm) {
               cudaStream_t ps = ctx->push_stream;
               D2D_CHECK_CUDA(cudaStreamWaitEvent(ps, written, 0));
               if (c == 0) {
                  for (int r : nf.peers) p2p_wait(ctx, r, 0, epoch, ps);
                  if (ctx->profiling) ctx->prof_begin(names[pen][nxt], 0, spans[0], ps);
               }
               PushArgs pa{};
               D2D_REQUIRE(np - 1 <= kMaxPushSegs, "too many peers for one exchange kernel");
               for (int k = 1; k < np; k++) {
                  const int m = (me + k) % np;
                  const int prank = nf.peers[k - 1];
                  LinkChunk R; // the consumer side ON THE DESTINATION rank: where it expects this rank's block
                  fft_link_chunk(dpeer[m], nxt, pen, padq, b[c], b[c + 1], R);
                  D2D_REQUIRE(R.cnt[me] == S.cnt[m], "pipelined chain: chunk sizes of the two sides disagree");
                  char *dst = (char *)p2p_peer_work(ctx, recv_w, prank);
                  const size_t off = (size_t)es * R.off[me], nbytes = (size_t)es * S.cnt[m];
                  D2D_REQUIRE(dst != nullptr && off + nbytes <= p2p_peer_bytes(ctx, recv_w, prank), "pipelined chain: destination buffer too small");
                  pa.flag[pa.nflag++] = p2p_push_flag(ctx, prank);
                  if (nbytes) pa.seg[pa.nseg++] = PushSeg{(const char *)sendbuf + (size_t)es * S.off[m], dst + off, nbytes};
                  if (ctx->profiling) ctx->prof[spans[0].idx].bytes += (double)nbytes;
               }
               launch_push(pa, ps);
               ctx->launches++;
               if (ctx->profiling && c == K - 1) ctx->prof_end(spans[0], ps);
            } else if (ce) {
               // copy engines: one cudaMemcpyAsync per peer and chunk.  A peer copy costs 25-40 us of dead time on its
               // stream (launch, completion flush, flag), so consecutive chunks of a peer alternate between `lanes` streams:
               // the data of chunk c + 1 moves while chunk c completes.  The arrival flags are monotonic sequence numbers,
               // so the flag of chunk c + 1 is chained behind the flag of chunk c with an event.
               for (int k = 1; k < np; k++) {
                  const int m = (me + k) % np;
                  const int prank = nf.peers[k - 1];
                  cudaStream_t cs = ctx->copy_stream_for((k - 1) * lanes + c % lanes);
                  LinkChunk R; // the consumer side ON THE DESTINATION rank: where it expects this rank's block
                  fft_link_chunk(dpeer[m], nxt, pen, padq, b[c], b[c + 1], R);
                  D2D_REQUIRE(R.cnt[me] == S.cnt[m], "pipelined chain: chunk sizes of the two sides disagree");
                  D2D_CHECK_CUDA(cudaStreamWaitEvent(cs, written, 0));
                  if (c == 0) {
                     p2p_wait(ctx, prank, 0, epoch, cs);
                     if (ctx->profiling) ctx->prof_begin(names[pen][nxt], 0, spans[k], cs);
                     if (lanes > 1) { // the other lanes may write into the peer once lane 0 has seen its "ready"
                        ready_seen[k] = ctx->new_sync_event();
                        D2D_CHECK_CUDA(cudaEventRecord(ready_seen[k], cs));
                     }
                  } else if (c < lanes) {
                     D2D_CHECK_CUDA(cudaStreamWaitEvent(cs, ready_seen[k], 0));
                  }
                  char *dst = (char *)p2p_peer_work(ctx, recv_w, prank);
                  const size_t off = (size_t)es * R.off[me], nbytes = (size_t)es * S.cnt[m];
                  D2D_REQUIRE(dst != nullptr && off + nbytes <= p2p_peer_bytes(ctx, recv_w, prank), "pipelined chain: destination buffer too small");
                  if (nbytes) D2D_CHECK_CUDA(cudaMemcpyAsync(dst + off, (const char *)sendbuf + (size_t)es * S.off[m], nbytes, cudaMemcpyDefault, cs));
                  if (lanes > 1 && c > 0) D2D_CHECK_CUDA(cudaStreamWaitEvent(cs, flagged[k], 0)); // flag of chunk c - 1 first
                  p2p_signal(ctx, prank, 2, nf.seq_base + (uint32_t)c + 1, cs);
                  if (lanes > 1 && c + 1 < K) {
                     flagged[k] = ctx->new_sync_event();
                     D2D_CHECK_CUDA(cudaEventRecord(flagged[k], cs));
                  }
                  if (ctx->profiling) {
                     ctx->prof[spans[k].idx].bytes += (double)nbytes;
                     if (c == K - 1) ctx->prof_end(spans[k], cs);
                  }
               }
            } else {
               D2D_CHECK_CUDA(cudaStreamWaitEvent(ctx->comm_stream, written, 0));
               LinkChunk R;
               fft_link_chunk(dc, nxt, pen, padq, b[c], b[c + 1], R);
               std::vector<PeerXfer> xf;
               for (int k = 1; k < np; k++) {
                  const int m = (me + k) % np;
                  PeerXfer x;
                  x.peer = nf.peers[k - 1];
                  x.sendptr = (const char *)sendbuf + (size_t)es * S.off[m];
                  x.sendbytes = (size_t)es * S.cnt[m];
                  x.recvptr = (char *)recvbuf + (size_t)es * R.off[m];
                  x.recvbytes = (size_t)es * R.cnt[m];
                  xf.push_back(x);
               }
               ctx->tr->exchange(xf, ctx->comm_stream);
               cudaEvent_t e = ctx->new_sync_event();
               D2D_CHECK_CUDA(cudaEventRecord(e, ctx->comm_stream));
               nf.arrived.push_back(e);
            }
         }
         if (sm) ctx->mark_buffer_busy_on(send_w, ctx->push_stream);
         else if (ce) ctx->mark_buffer_busy(send_w, (np - 1) * lanes);
         else { // the communication stream still reads the send buffer: the next writer of it waits for that
            if (ctx->buf_busy[send_w].empty()) {
               cudaEvent_t e;
               D2D_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
               ctx->buf_busy[send_w].push_back(e);
            }
            D2D_CHECK_CUDA(cudaEventRecord(ctx->buf_busy[send_w][0], ctx->comm_stream));
         }
         feed = nf;
         cur = fft_link_map(dc, nxt, pen, recvbuf, sendbuf, es, true, padq);
         continue;
      }
      if (feed.chunked) { // chunk by chunk as the sub-ranges of the feeding link land (same free axis, seen from this side)
         LinkChunk probe;
         fft_link_chunk(dc, pen, st[s - 1].pencil, padq, 0, 0, probe);
         for (size_t c = 0; c + 1 < feed.bounds.size(); c++) {
            wait_chunk((int)c);
            StageRange rg;
            rg.axis = probe.axis_is_a ? 0 : 1; rg.f0 = feed.bounds[c]; rg.f1 = feed.bounds[c + 1];
            run_chain_stage(p, s, mode, pen, dc, dr, cur, om, rptr, backward, passthrough, rg);
         }
         feed = Feed();
      } else {
         run_chain_stage(p, s, mode, pen, dc, dr, cur, om, rptr, backward, passthrough);
      }
      if (last) break;
      // a link inside one rank: the consumer reads what this stage wrote
      cur = fft_link_map(dc, st[s + 1].pencil, pen, nullptr, sendbuf, es, true, padq);
   }
}

static void run_chain(Plan &p, const Decomp &dc, const Decomp *dr, const StageDef st[3], void *in, void *out, int backward)
{
   Ctx *ctx = p.ctx;
   // Multi-rank chains over peer memory (D2D_EXCHANGE = auto | pipe | fused):
   //   pipe  : chunk-pipelined chain, the chunks pushed by the copy engines (run_chain_pipe)
   //   fused : the producer kernels store straight into the peers' buffers, no chunks (run_chain_p2p)
   // auto picks by measurement (1024^3 fp64 pair on B200s, profiles/r02_*): with one peer per link the copy engines move a
   // chunk at 690-780 GB/s while the SMs run HBM-bound stages (2 GPUs 10.9 vs 12.9 ms fused, 4 GPUs 7.6 vs 8.4); with three
   // peers per link every cudaMemcpyAsync carries ~35 us of dead time and the chunks of the three peers serialise on the
   // engine (500 GB/s), so the fused stores -- 650-670 GB/s of the ~715 GB/s SM-issued writes reach on NVLink -- win
   // (8 GPUs 5.2 vs 6.1 ms).  Without peer memory (D2D_P2P=0, in-process groups) the pipelined chain runs on the transport;
   // D2D_CHUNKS=0 selects the plain stage -> exchange -> stage sequence there.
   static const std::string xmode = getenv("D2D_EXCHANGE") ? getenv("D2D_EXCHANGE") : ((getenv("D2D_FUSED") && atoi(getenv("D2D_FUSED")) != 0) ? "fused" : "auto");
   const bool fused = xmode == "fused" || (xmode == "auto" && std::max(ctx->p_row, ctx->p_col) > 2);
   if (p2p_active(ctx) && fused) return run_chain_p2p(p, dc, dr, st, in, out, backward);
   const char *ck = getenv("D2D_CHUNKS") ? getenv("D2D_CHUNKS") : getenv("D2D_OVERLAP");
   const bool plain = ck && atoi(ck) <= 0;
   if (ctx->nranks > 1 && (p2p_active(ctx) || !plain)) return run_chain_pipe(p, dc, dr, st, in, out, backward, pipe_chunks());
   const int es = p.f64 ? 16 : 8;
   const int padq = 128 / es;
   const size_t wbytes = uniform_work_bytes(p, dr == nullptr);
   // `out` can stand in for the first work buffer only when the padded wire layout fits in it
   const bool borrow_out = (ctx->nranks == 1) && st[2].mode == MODE_C2C && wbytes <= (size_t)es * dc.pencil_elems(st[2].pencil);
   // opt_inplace (src/fft_common.f90:172-176, src/fft_cufft.f90:696-706, 961-971): a complex input may be overwritten -- it
   // then serves as the work buffer between stages 1 and 2, so a single-rank c2c whose padded wire layout fits the pencils
   // needs no work buffer at all (out carries link 0, in carries link 1)
   const bool reuse_in = p.inplace && ctx->nranks == 1 && st[0].mode != MODE_R2C && wbytes <= (size_t)es * dc.pencil_elems(st[0].pencil);
   int live_a = -1, live_b = -1; // work buffers holding the current stage's input
   auto pick = [&]() {
      for (int i = 0; i < 3; i++)
         if (i != live_a && i != live_b) return i;
      return -1;
   };
   PieceMap cur = fft_user_map(dc, st[0].pencil, in); // complex input map (unused for R2C)
   for (int s = 0; s < 3; s++) {
      const int pen = st[s].pencil, mode = st[s].mode;
      const bool last = (s == 2);
      const int passthrough = (mode == MODE_C2C && p.skip[pen]) ? 1 : 0;
      PieceMap om{};
      void *rptr = nullptr, *sendbuf = nullptr;
      int send_w = -1;
      if (mode == MODE_R2C) rptr = in;
      if (last) {
         if (mode == MODE_C2R) rptr = out;
         else om = fft_user_map(dc, pen, out);
      } else {
         if (s == 0 && borrow_out) sendbuf = out;
         else if (s == 1 && reuse_in) sendbuf = in; // opt_inplace: stage 0 has consumed the user's input (stream order)
         else { send_w = pick(); sendbuf = ctx->reserve(send_w, wbytes); }
         om = fft_link_map(dc, pen, st[s + 1].pencil, sendbuf, sendbuf, es, false, padq);
      }
      run_chain_stage(p, s, mode, pen, dc, dr, cur, om, rptr, backward, passthrough);
      if (last) break;
      const int nxt = st[s + 1].pencil;
      // the input buffers of this stage are free once it has run (stream order)
      const int old_a = live_a, old_b = live_b;
      live_a = send_w;
      live_b = -1;
      void *recvbuf = nullptr;
      if (comm_size(dc, pen, nxt) > 1) {
         int recv_w = (old_a >= 0 && old_a != send_w) ? old_a : (old_b >= 0 && old_b != send_w) ? old_b : pick();
         recvbuf = ctx->reserve(recv_w, wbytes);
         live_b = recv_w;
         fft_exchange(ctx, dc, pen, nxt, sendbuf, recvbuf, es, padq);
      }
      cur = fft_link_map(dc, nxt, pen, recvbuf, sendbuf, es, true, padq);
   }
}

// work-buffer size (bytes) that is the same on every rank of the grid (ragged decompositions give the last
// ranks bigger pencils): peers address each other's buffers, so all of them must (re)allocate together
//
// r2c / c2r chains move complex pencils of `sp`, c2c chains complex pencils of `ph` (about twice as large).  The reference
// sizes its work buffers for the largest complex pencil of any decomposition seen (src/decomp_2d.f90:461-485), so an
// r2c-only user pays for c2c buffers; here the c2c size is reserved by the first c2c call instead (2048^3 fp32 r2c on one
// B200: 2 x 34 GB instead of 2 x 69 GB of work space).
static size_t uniform_work_bytes(const Plan &p, bool c2c)
{
   const int es = p.f64 ? 16 : 8;
   const Decomp &g = c2c ? p.ph.d : p.sp.d;
   int64_t m = 0;
   for (int r = 0; r < p.ctx->nranks; r++) {
      Decomp a;
      decomp_init(a, g.nx, g.ny, g.nz, p.ctx->p_row, p.ctx->p_col, r);
      m = std::max(m, fft_work_elems(a, 128 / es));
   }
   return (size_t)es * (size_t)m;
}

// The same chain when the peers' work buffers are mapped (p2p.cpp): every producer writes each block -- its own
// included -- straight into the DESTINATION rank's buffer, at the displacement where that rank's consumer expects
// the block from this rank.  There is no exchange step and no send buffer; two buffers alternate.  Flags order the
// ranks: "ready" before a producer may write into a peer, "done" before a consumer may read what peers wrote.
static void run_chain_p2p(Plan &p, const Decomp &dc, const Decomp *dr, const StageDef st[3], void *in, void *out, int backward)
{
   Ctx *ctx = p.ctx;
   const int es = p.f64 ? 16 : 8;
   const int padq = 128 / es;
   PieceMap cur = fft_user_map(dc, st[0].pencil, in);
   // "ready" for both links right away: buffers 0 and 1 were last read by the previous chain, which precedes this point in
   // stream order, so a producer never waits for a peer to REACH the link (measured at 8 GPUs: 0.16 ms per pair of such waits)
   uint32_t link_epoch[2] = {0, 0};
   for (int s = 0; s < 2; s++) {
      const bool lcol = (st[s].pencil == 0 || st[s + 1].pencil == 0);
      const int lnp = lcol ? ctx->p_row : ctx->p_col, lme = lcol ? ctx->c1 : ctx->c2;
      if (lnp <= 1) continue;
      ProfScope ps(ctx, "p2p_ready");
      link_epoch[s] = p2p_next_epoch(ctx);
      for (int k = 1; k < lnp; k++) {
         const int m = (lme + k) % lnp;
         p2p_signal(ctx, lcol ? ctx->peer_rank_col(m) : ctx->peer_rank_row(m), 0, link_epoch[s]);
      }
   }
   for (int s = 0; s < 3; s++) {
      const int pen = st[s].pencil, mode = st[s].mode;
      const bool last = (s == 2);
      const int passthrough = (mode == MODE_C2C && p.skip[pen]) ? 1 : 0;
      PieceMap om{};
      void *rptr = nullptr;
      if (mode == MODE_R2C) rptr = in;
      int np = 1, me = 0;
      bool col = false;
      uint32_t epoch = 0;
      double remote_bytes = 0;
      const int w = s & 1;
      if (last) {
         if (mode == MODE_C2R) rptr = out;
         else om = fft_user_map(dc, pen, out);
      } else {
         const int nxt = st[s + 1].pencil;
         LinkSide L;
         fft_link_side(dc, pen, nxt, padq, L);
         np = L.np;
         me = L.me;
         col = (pen == 0 || nxt == 0);
         om.np = np;
         for (int m = 0; m < np; m++) {
            const int prank = col ? ctx->peer_rank_col(m) : ctx->peer_rank_row(m);
            Decomp dm;
            decomp_init(dm, dc.nx, dc.ny, dc.nz, ctx->p_row, ctx->p_col, prank);
            LinkSide R;
            fft_link_side(dm, nxt, pen, padq, R); // the consumer side on the destination rank
            D2D_REQUIRE(R.cnt[me] == L.cnt[m], "peer-store: block sizes of the two sides disagree");
            char *base = (char *)(m == me ? ctx->work[w] : p2p_peer_work(ctx, w, prank));
            const size_t cap = m == me ? ctx->work_bytes[w] : p2p_peer_bytes(ctx, w, prank);
            D2D_REQUIRE(base != nullptr && (size_t)es * (size_t)(R.disp[me] + R.cnt[me]) <= cap, "peer-store: destination buffer too small");
            om.e0[m] = L.e0[m];
            om.ptr[m] = base + (size_t)es * R.disp[me];
            om.se[m] = L.se[m]; om.sa[m] = L.sa[m]; om.sb[m] = L.sb[m];
            if (m != me) remote_bytes += (double)es * (double)L.cnt[m];
         }
         om.e0[np] = L.e0[np];
         if (np > 1) {
            ProfScope ps(ctx, "p2p_ready");
            epoch = link_epoch[s];
            for (int k = 1; k < np; k++) {
               const int m = (me + k) % np;
               p2p_wait(ctx, col ? ctx->peer_rank_col(m) : ctx->peer_rank_row(m), 0, epoch);
            }
         }
      }
      {
         // the producer kernel IS the exchange: time it a second time under the link's name with the bytes that
         // leave this GPU, so that the NVLink rate of the fused kernel can be read next to its HBM rate
         static const char *names[3][3] = {{"", "p2p_x_y", ""}, {"p2p_y_x", "", "p2p_y_z"}, {"", "p2p_z_y", ""}};
         ProfScope ps(ctx, (!last && np > 1) ? names[pen][st[s + 1].pencil] : "p2p_none", remote_bytes);
         run_chain_stage(p, s, mode, pen, dc, dr, cur, om, rptr, backward, passthrough);
      }
      if (last) break;
      const int nxt = st[s + 1].pencil;
      if (np > 1) {
         ProfScope ps(ctx, "p2p_done");
         for (int k = 1; k < np; k++) {
            const int m = (me + k) % np;
            p2p_signal(ctx, col ? ctx->peer_rank_col(m) : ctx->peer_rank_row(m), 1, epoch);
         }
         for (int k = 1; k < np; k++) {
            const int m = (me + k) % np;
            p2p_wait(ctx, col ? ctx->peer_rank_col(m) : ctx->peer_rank_row(m), 1, epoch);
         }
      }
      // consumer: every block (the one this rank wrote for itself included) sits in this rank's buffer w
      LinkSide C;
      fft_link_side(dc, nxt, pen, padq, C);
      cur = PieceMap{};
      cur.np = C.np;
      for (int m = 0; m < C.np; m++) {
         cur.e0[m] = C.e0[m];
         cur.ptr[m] = (char *)ctx->work[w] + (size_t)es * C.disp[m];
         cur.se[m] = C.se[m]; cur.sa[m] = C.sa[m]; cur.sb[m] = C.sb[m];
      }
      cur.e0[C.np] = C.e0[C.np];
   }
}

// Grow the context's work buffers to what a chain of this kind needs.  Collective when it grows (every rank computes the
// same size and has the same history of sizes, so all ranks take the same branch): peers map each other's buffers.
// Called at plan creation for the r2c / c2r size and by the first c2c call for the c2c size -- never inside a steady loop.
static void ensure_work(Plan &p, bool c2c, bool at_plan_creation)
{
   Ctx *ctx = p.ctx;
   // one rank: run_chain reserves what it needs when it needs it (nothing at all for an in-place c2c whose wire layout fits
   // the user's arrays); several ranks: the buffers are mapped between the ranks, so they are sized and published here
   if (ctx->nranks == 1) return;
   ctx->ensure_buffers(kCtxBuffers, uniform_work_bytes(p, c2c), at_plan_creation);
}

Plan *plan_create(Ctx *ctx, int format, int nx, int ny, int nz, int dtype, int inplace, const int skip[3])
{
   D2D_REQUIRE(format == D2D_PHYSICAL_IN_X || format == D2D_PHYSICAL_IN_Z, "format must be PHYSICAL_IN_X (1) or PHYSICAL_IN_Z (3)");
   D2D_REQUIRE(dtype == D2D_F32 || dtype == D2D_F64, "dtype must be D2D_F32 or D2D_F64");
   std::unique_ptr<Plan> p(new Plan());
   p->ctx = ctx;
   p->format = format; p->nx = nx; p->ny = ny; p->nz = nz;
   p->f64 = dtype == D2D_F64; p->inplace = inplace;
   for (int i = 0; i < 3; i++) p->skip[i] = skip ? (skip[i] != 0) : 0;
   p->ph.ctx = ctx; p->sp.ctx = ctx;
   decomp_init(p->ph.d, nx, ny, nz, ctx->p_row, ctx->p_col, ctx->rank);
   // sp (src/fft_common.f90:210-216)
   if (format == D2D_PHYSICAL_IN_X) decomp_init(p->sp.d, nx / 2 + 1, ny, nz, ctx->p_row, ctx->p_col, ctx->rank);
   else decomp_init(p->sp.d, nx, ny, nz / 2 + 1, ctx->p_row, ctx->p_col, ctx->rank);
   D2D_CHECK_CUDA(cudaSetDevice(ctx->device));
   ensure_work(*p, false, true);
   return p.release();
}

void plan_destroy(Plan *p)
{
   if (!p) return;
   try { p->ctx->sync_all(); } catch (...) {}
   for (auto &io : p->io) {
      if (io.in) cudaFree(io.in);
      if (io.out) cudaFree(io.out);
      if (io.in_done) cudaEventDestroy(io.in_done);
      if (io.out_done) cudaEventDestroy(io.out_done);
   }
   delete p;
}

const d2d_decomp *plan_ph(const Plan *p) { return &p->ph; }
const d2d_decomp *plan_sp(const Plan *p) { return &p->sp; }

static void plan_bytes(const Plan *p, size_t b[4], int isign);
// in and out of a 3-D transform are different pencils; the stages write `out` (and, in place, `in`) while `in` is still being
// read, so overlapping arrays would be silently corrupted: refuse them
static void require_disjoint(const Plan *p, const void *in, int bi, const void *out, int bo, int isign)
{
   size_t b[4];
   plan_bytes(p, b, isign);
   const char *a = (const char *)in, *o = (const char *)out;
   D2D_REQUIRE(a != nullptr && o != nullptr, "null array");
   D2D_REQUIRE(a + b[bi] <= o || o + b[bo] <= a, "decomp_2d_fft_3d: the input and output arrays overlap");
}

// fft_3d_c2c (src/fft_cufft.f90:676-790): X-forward / Z-backward run x -> y -> z, the others z -> y -> x
void fft_3d_c2c(Plan *p, void *in, void *out, int isign)
{
   D2D_REQUIRE(isign == D2D_FFT_FORWARD || isign == D2D_FFT_BACKWARD, "isign must be -1 or +1");
   require_disjoint(p, in, 2, out, 3, isign);
   D2D_CHECK_CUDA(cudaSetDevice(p->ctx->device));
   ensure_work(*p, true, false); // first c2c call of this context: grows the work buffers to the ph-complex size (collective)
   ProfScope ps(p->ctx, "fft_c2c");
   const bool xyz = (p->format == D2D_PHYSICAL_IN_X && isign == D2D_FFT_FORWARD) ||
                    (p->format == D2D_PHYSICAL_IN_Z && isign == D2D_FFT_BACKWARD);
   const StageDef a[3] = {{0, MODE_C2C}, {1, MODE_C2C}, {2, MODE_C2C}};
   const StageDef b[3] = {{2, MODE_C2C}, {1, MODE_C2C}, {0, MODE_C2C}};
   run_chain(*p, p->ph.d, nullptr, xyz ? a : b, in, out, isign == D2D_FFT_BACKWARD);
}

// fft_3d_r2c (src/fft_cufft.f90:795-934)
void fft_3d_r2c(Plan *p, const void *in_r, void *out_c)
{
   require_disjoint(p, in_r, 0, out_c, 1, D2D_FFT_FORWARD);
   D2D_CHECK_CUDA(cudaSetDevice(p->ctx->device));
   ProfScope ps(p->ctx, "fft_r2c");
   const StageDef x[3] = {{0, MODE_R2C}, {1, MODE_C2C}, {2, MODE_C2C}};
   const StageDef z[3] = {{2, MODE_R2C}, {1, MODE_C2C}, {0, MODE_C2C}};
   run_chain(*p, p->sp.d, &p->ph.d, p->format == D2D_PHYSICAL_IN_X ? x : z, const_cast<void *>(in_r), out_c, 0);
}

// fft_3d_c2r (src/fft_cufft.f90:939-1170)
void fft_3d_c2r(Plan *p, void *in_c, void *out_r)
{
   require_disjoint(p, in_c, 1, out_r, 0, D2D_FFT_BACKWARD);
   D2D_CHECK_CUDA(cudaSetDevice(p->ctx->device));
   ProfScope ps(p->ctx, "fft_c2r");
   const StageDef x[3] = {{2, MODE_C2C}, {1, MODE_C2C}, {0, MODE_C2R}};
   const StageDef z[3] = {{0, MODE_C2C}, {1, MODE_C2C}, {2, MODE_C2R}};
   run_chain(*p, p->sp.d, &p->ph.d, p->format == D2D_PHYSICAL_IN_X ? x : z, in_c, out_r, 1);
}

// sizes (bytes) of the user arrays of a plan: [0] real physical, [1] complex spectral, [2] complex physical (c2c in), [3] complex c2c out
static void plan_bytes(const Plan *p, size_t b[4], int isign)
{
   const int rs = p->f64 ? 8 : 4;
   const bool fx = p->format == D2D_PHYSICAL_IN_X;
   b[0] = (size_t)rs * p->ph.d.pencil_elems(fx ? 0 : 2);
   b[1] = (size_t)2 * rs * p->sp.d.pencil_elems(fx ? 2 : 0);
   const bool xyz = (fx && isign == D2D_FFT_FORWARD) || (!fx && isign == D2D_FFT_BACKWARD);
   b[2] = (size_t)2 * rs * p->ph.d.pencil_elems(xyz ? 0 : 2);
   b[3] = (size_t)2 * rs * p->ph.d.pencil_elems(xyz ? 2 : 0);
}

static void ensure_stage(Plan *p, int which, size_t bi, size_t bo)
{
   Plan::HostIo &io = p->io[which];
   if (bi > io.in_bytes || bo > io.out_bytes) p->ctx->sync_all(); // nothing may still use the old staging arrays
   if (bi > io.in_bytes) {
      if (io.in) D2D_CHECK_CUDA(cudaFree(io.in));
      io.in = nullptr; io.in_bytes = 0;
      D2D_CHECK_CUDA(cudaMalloc(&io.in, bi));
      io.in_bytes = bi;
   }
   if (bo > io.out_bytes) {
      if (io.out) D2D_CHECK_CUDA(cudaFree(io.out));
      io.out = nullptr; io.out_bytes = 0;
      D2D_CHECK_CUDA(cudaMalloc(&io.out, bo));
      io.out_bytes = bo;
   }
   if (!io.in_done) D2D_CHECK_CUDA(cudaEventCreateWithFlags(&io.in_done, cudaEventDisableTiming));
   if (!io.out_done) D2D_CHECK_CUDA(cudaEventCreateWithFlags(&io.out_done, cudaEventDisableTiming));
}

void plan_reserve_host_staging(Plan *p)
{
   size_t b[4];
   plan_bytes(p, b, D2D_FFT_FORWARD);
   D2D_CHECK_CUDA(cudaSetDevice(p->ctx->device));
   ensure_stage(p, 0, b[0], b[1]);
   ensure_stage(p, 1, b[1], b[0]);
}

// decomp_2d_fft_3d on HOST arrays: upload, transform, download -- as three pipelines on three streams.
//   * uploads (H2D) run on the context's upload stream in kHostPieces pieces, downloads (D2H) on its download stream, the
//     transform on the context's stream; PCIe is full duplex, so the download of one call overlaps the upload of the next;
//   * an x stage on the user's X-pencil starts on the z-slabs that have arrived and hands finished z-slabs to the download
//     stream while it works on the next ones (IoGate);
//   * host-memory dependencies between calls are tracked per piece: an upload from a host range that an earlier download
//     is still writing waits only for the pieces it reads, so the spectrum of r2c_host can travel back up for c2r_host
//     while its tail is still coming down.
// With the context in blocking mode (default, like the reference) every call returns when its output is complete on the
// host; with d2d_ctx_set_blocking(ctx, 0) calls return once enqueued and d2d_ctx_sync() completes them.
constexpr int kHostPieces = 16;

void fft_3d_host(Plan *p, int which /*0 r2c, 1 c2r, 2 c2c*/, const void *in_h, void *out_h, int isign)
{
   Ctx *ctx = p->ctx;
   size_t b[4];
   plan_bytes(p, b, isign);
   const size_t bi = which == 0 ? b[0] : which == 1 ? b[1] : b[2];
   const size_t bo = which == 0 ? b[1] : which == 1 ? b[0] : b[3];
   D2D_CHECK_CUDA(cudaSetDevice(ctx->device));
   ensure_stage(p, which, bi, bo);
   Plan::HostIo &io = p->io[which];
   for (int k = 0; k < 2; k++)
      if (!ctx->io_stream[k]) D2D_CHECK_CUDA(cudaStreamCreateWithFlags(&ctx->io_stream[k], cudaStreamNonBlocking));
   cudaStream_t up = ctx->io_stream[0], dn = ctx->io_stream[1], st = ctx->stream;
   Ctx::HostTrack &ht = ctx->host;
   auto overlaps = [](const char *a, size_t na, const char *b2, size_t nb) { return a < b2 + nb && b2 < a + na; };

   // ---- upload ----------------------------------------------------------------------------------------------------
   D2D_CHECK_CUDA(cudaStreamWaitEvent(up, io.in_done, 0)); // the previous transform of this kind has read its staging input
   const char *hin = (const char *)in_h;
   const bool raw = ht.d2h_base && overlaps(hin, bi, ht.d2h_base, ht.d2h_bytes);
   std::vector<size_t> ucut(kHostPieces + 1);
   for (int c = 0; c <= kHostPieces; c++) ucut[c] = c == kHostPieces ? bi : (bi / kHostPieces * c) / 4096 * 4096;
   std::vector<cudaEvent_t> uev(kHostPieces);
   {
      ProfScope ps(ctx, "h2d", (double)bi, up);
      for (int c = 0; c < kHostPieces; c++) {
         const size_t lo = ucut[c], hi = ucut[c + 1];
         if (raw) { // wait for the download pieces that write [hin + lo, hin + hi): they are ordered, so the last one that overlaps
            int lastp = -1;
            for (size_t q = 0; q < ht.d2h_cut.size(); q++) {
               const size_t plo = ht.d2h_cut[q].first, phi = ht.d2h_cut[q].second;
               if (overlaps(hin + lo, hi - lo, ht.d2h_base + plo, phi - plo)) lastp = (int)q;
            }
            if (lastp >= 0) D2D_CHECK_CUDA(cudaStreamWaitEvent(up, ht.d2h_ev[lastp], 0));
         }
         if (hi > lo) D2D_CHECK_CUDA(cudaMemcpyAsync((char *)io.in + lo, hin + lo, hi - lo, cudaMemcpyHostToDevice, up));
         uev[c] = ctx->new_sync_event();
         D2D_CHECK_CUDA(cudaEventRecord(uev[c], up));
      }
   }
   if (!ht.h2d_done) D2D_CHECK_CUDA(cudaEventCreateWithFlags(&ht.h2d_done, cudaEventDisableTiming));
   D2D_CHECK_CUDA(cudaEventRecord(ht.h2d_done, up));
   ht.h2d_base = hin;
   ht.h2d_bytes = bi;

   // ---- download side bookkeeping ----------------------------------------------------------------------------------
   D2D_CHECK_CUDA(cudaStreamWaitEvent(st, io.out_done, 0)); // the previous download out of this staging output has finished
   char *hout = (char *)out_h;
   if (ht.h2d_base && overlaps(hout, bo, ht.h2d_base, ht.h2d_bytes)) D2D_CHECK_CUDA(cudaStreamWaitEvent(dn, ht.h2d_done, 0));
   if (ht.d2h_base && overlaps(hout, bo, ht.d2h_base, ht.d2h_bytes) && !ht.d2h_cut.empty())
      D2D_CHECK_CUDA(cudaStreamWaitEvent(dn, ht.d2h_ev[ht.d2h_cut.size() - 1], 0)); // write after write: keep the order
   ht.d2h_base = hout;
   ht.d2h_bytes = bo;
   ht.d2h_cut.clear();
   Ctx::Pending dspan{};
   bool dspan_open = false;
   auto download = [&](size_t lo, size_t hi) {
      hi = std::min(hi, bo);
      if (lo >= hi) return;
      cudaEvent_t done = ctx->new_sync_event();
      D2D_CHECK_CUDA(cudaEventRecord(done, st));
      D2D_CHECK_CUDA(cudaStreamWaitEvent(dn, done, 0));
      if (ctx->profiling && !dspan_open) { ctx->prof_begin("d2h", (double)bo, dspan, dn); dspan_open = true; }
      D2D_CHECK_CUDA(cudaMemcpyAsync(hout + lo, (const char *)io.out + lo, hi - lo, cudaMemcpyDeviceToHost, dn));
      const size_t q = ht.d2h_cut.size();
      D2D_REQUIRE(q < Ctx::HostTrack::kMaxPieces, "too many download pieces");
      if (!ht.d2h_ev[q]) D2D_CHECK_CUDA(cudaEventCreateWithFlags(&ht.d2h_ev[q], cudaEventDisableTiming));
      D2D_CHECK_CUDA(cudaEventRecord(ht.d2h_ev[q], dn));
      ht.d2h_cut.push_back({lo, hi});
   };
   IoGate gate;
   gate.nio = kHostPieces;
   gate.need_in = [&](size_t lo, size_t hi) {
      hi = std::min(hi, bi);
      int lastc = -1;
      for (int c = 0; c < kHostPieces; c++)
         if (ucut[c] < hi && lo < ucut[c + 1]) lastc = c;
      if (lastc >= 0) D2D_CHECK_CUDA(cudaStreamWaitEvent(st, uev[lastc], 0)); // pieces are uploaded in order
   };
   gate.have_out = download;

   // which pencils the user arrays are: the gates act slab-wise on X-pencils only
   const bool fx = p->format == D2D_PHYSICAL_IN_X;
   int pin, pout;
   if (which == 0) { pin = fx ? 0 : 2; pout = fx ? 2 : 0; }
   else if (which == 1) { pin = fx ? 2 : 0; pout = fx ? 0 : 2; }
   else {
      const bool xyz = (fx && isign == D2D_FFT_FORWARD) || (!fx && isign == D2D_FFT_BACKWARD);
      pin = xyz ? 0 : 2; pout = xyz ? 2 : 0;
   }
   if (pin != 0) gate.need_in(0, bi); // the whole input first
   p->gate = &gate;
   const int keep_inplace = p->inplace;
   p->inplace = 1; // the staging copy of the input may be clobbered
   try {
      if (which == 0) fft_3d_r2c(p, io.in, io.out);
      else if (which == 1) fft_3d_c2r(p, io.in, io.out);
      else fft_3d_c2c(p, io.in, io.out, isign);
   } catch (...) {
      p->gate = nullptr;
      p->inplace = keep_inplace;
      throw;
   }
   p->gate = nullptr;
   p->inplace = keep_inplace;
   D2D_CHECK_CUDA(cudaEventRecord(io.in_done, st));
   if (pout != 0 || ht.d2h_cut.empty()) download(0, bo); // not announced slab by slab: the whole output now
   if (dspan_open) ctx->prof_end(dspan, dn);
   D2D_CHECK_CUDA(cudaEventRecord(io.out_done, dn));
   ctx->finish_call(); // blocking contexts: host results are complete on return
}

void plan_get_size(const Plan *p, int istart[3], int iend[3], int isize[3])
{
   // decomp_2d_fft_get_size (src/fft_common.f90:311-327): Z-pencil of sp for PHYSICAL_IN_X, X-pencil for PHYSICAL_IN_Z
   const Decomp &s = p->sp.d;
   const bool fx = p->format == D2D_PHYSICAL_IN_X;
   for (int i = 0; i < 3; i++) {
      istart[i] = (fx ? s.zst[i] : s.xst[i]) + 1;
      iend[i] = (fx ? s.zen[i] : s.xen[i]) + 1;
      isize[i] = fx ? s.zsz[i] : s.xsz[i];
   }
}
Ctx *plan_ctx(Plan *p) { return p->ctx; }

// ---- bare batched 1-D transforms on one local array (c2c_1m_* / r2c_1m_* / c2r_1m_*) -------------
void fft_1m(Ctx *ctx, int dtype, int mode, int axis, int n1, int n2, int n3, const void *in, void *out, int isign)
{
   D2D_REQUIRE(axis >= 0 && axis < 3, "axis must be 0, 1 or 2");
   D2D_REQUIRE(dtype == D2D_F32 || dtype == D2D_F64, "dtype must be D2D_F32 or D2D_F64");
   D2D_CHECK_CUDA(cudaSetDevice(ctx->device));
   const int f64 = dtype == D2D_F64;
   Decomp dr, dc;
   decomp_init(dr, n1, n2, n3, 1, 1, 0);
   if (mode == MODE_C2C) {
      run_stage(ctx, f64, mode, axis, dr, nullptr, natural_map(dr, axis, const_cast<void *>(in)), natural_map(dr, axis, out), nullptr,
                isign == D2D_FFT_BACKWARD, 0, false);
      return;
   }
   D2D_REQUIRE(axis == 0 || axis == 2, "real transforms run along x or z only");
   const int c1 = axis == 0 ? n1 / 2 + 1 : n1, c3 = axis == 2 ? n3 / 2 + 1 : n3;
   decomp_init(dc, c1, n2, c3, 1, 1, 0);
   if (mode == MODE_R2C)
      run_stage(ctx, f64, mode, axis, dc, &dr, PieceMap{}, natural_map(dc, axis, out), const_cast<void *>(in), 0, 0, false);
   else
      run_stage(ctx, f64, mode, axis, dc, &dr, natural_map(dc, axis, const_cast<void *>(in)), PieceMap{}, out, 1, 0, false);
}

} // namespace d2d
