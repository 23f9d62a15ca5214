// common.h -- internal types of libd2dfft_b200 (context, decomposition, transport, errors).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/d2d_b200.h"
#include "fft_kernel.cuh"

namespace d2d {

struct Error : std::runtime_error {
   int code;
   Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &m);
const std::string &get_last_error();

#define D2D_STR2(x) #x
#define D2D_STR(x) D2D_STR2(x)
#define D2D_CHECK_CUDA(call)                                                                                           \
   do {                                                                                                                \
      cudaError_t e__ = (call);                                                                                        \
      if (e__ != cudaSuccess)                                                                                          \
         throw ::d2d::Error(1000 + (int)e__, std::string(__FILE__ ":" D2D_STR(__LINE__) " " #call ": ") +              \
                                                 cudaGetErrorString(e__));                                             \
   } while (0)
#define D2D_REQUIRE(cond, msg)                                                                                         \
   do {                                                                                                                \
      if (!(cond)) throw ::d2d::Error(2, std::string(__FILE__ ":" D2D_STR(__LINE__) " ") + (msg));                     \
   } while (0)

constexpr int kMaxP = kMaxPieces; // max ranks along one side of the process grid
constexpr int kMaxRanks = kMaxPieces * kMaxPieces;
constexpr int kCtxBuffers = 4;   // link s of a chain: send buffer 2 s, receive buffer 2 s + 1 (pipelined chains); rotated otherwise
constexpr int kWorkBuffers = kCtxBuffers; // all of them are mapped between the ranks (p2p.cpp)

// decomp_info (src/info.f90:19-47).  0-based starts internally; the C ABI converts to 1-based.
struct Decomp {
   int nx, ny, nz, p_row, p_col, rank, c1, c2;
   int xst[3], xen[3], xsz[3], yst[3], yen[3], ysz[3], zst[3], zen[3], zsz[3];
   int x1dist[kMaxP], y1dist[kMaxP], y2dist[kMaxP], z2dist[kMaxP];
   int x1off[kMaxP + 1], y1off[kMaxP + 1], y2off[kMaxP + 1], z2off[kMaxP + 1]; // exclusive prefix sums of the dists
   int64_t x1cnts[kMaxP], y1cnts[kMaxP], y2cnts[kMaxP], z2cnts[kMaxP];
   int64_t x1disp[kMaxP], y1disp[kMaxP], y2disp[kMaxP], z2disp[kMaxP];
   // EVEN builds of the reference (padded MPI_ALLTOALL, src/decomp_2d.f90:1186-1204): one count per communicator, segment m
   // of a buffer at m * count; `even` = the data is evenly distributed (:443-454)
   int64_t x1count, y1count, y2count, z2count;
   int64_t e_cnts_row[kMaxP], e_disp_row[kMaxP], e_cnts_col[kMaxP], e_disp_col[kMaxP]; // x<->y (p_row peers) / y<->z (p_col peers)
   int even;
   int64_t pencil_elems(int p) const
   {
      const int *s = p == 0 ? xsz : p == 1 ? ysz : zsz;
      return (int64_t)s[0] * s[1] * s[2];
   }
   int64_t max_pencil() const { return std::max(pencil_elems(0), std::max(pencil_elems(1), pencil_elems(2))); }
};
void decomp_init(Decomp &d, int nx, int ny, int nz, int p_row, int p_col, int rank);
void best_2d_grid(int nproc, int *p_row, int *p_col);

// ---- exchange transports ------------------------------------------------------------------------
struct PeerXfer {
   int peer;           // global rank
   const void *sendptr;
   size_t sendbytes;
   void *recvptr;
   size_t recvbytes;
};

struct Transport {
   virtual ~Transport() {}
   virtual int kind() const = 0;
   // all-to-all(v) among the listed peers (self excluded by the caller); stream-ordered where the
   // transport allows it.
   virtual void exchange(const std::vector<PeerXfer> &xf, cudaStream_t st) = 0;
   virtual void barrier(cudaStream_t st) = 0;
   // all-gather of `bytes` host bytes per rank (blocking; bootstrap use only)
   virtual void allgather(const void *send_host, void *recv_host, size_t bytes, cudaStream_t st) = 0;
};

Transport *make_nccl_transport(const unsigned char id[128], int nranks, int rank);
// bootstrap-only transport: the caller supplies the all-gather (MPI_Allgather in the Fortran shim, torch.distributed in the
// Python mirror); the data plane is the peer-memory path of p2p.cpp (copy engines / peer stores over NVLink), no NCCL
Transport *make_boot_transport(d2d_allgather_fn fn, void *user, int nranks, int rank);
void nccl_unique_id(unsigned char id[128]);
struct Group;
Group *group_create(int nranks);
void group_destroy(Group *);
void group_abort(Group *); // wake every rank waiting in the group's barrier with an error (a rank failed)
Transport *make_local_transport(Group *g, int rank);

struct ProfEntry {
   std::string label;
   double total_ms = 0;
   int64_t calls = 0;
   double bytes = 0;
};

struct Ctx {
   int nranks = 1, rank = 0, p_row = 1, p_col = 1, c1 = 0, c2 = 0, device = 0;
   cudaStream_t stream = nullptr;
   std::unique_ptr<Transport> tr;
   bool blocking = true;
   bool even = false; // bare transposes use the padded equal-count layout of the reference's EVEN builds
   int64_t launches = 0;
   // grow-only work buffers (the reference's work1/work2 high-water mark, src/decomp_2d.f90:461-485)
   void *work[kCtxBuffers] = {nullptr, nullptr, nullptr, nullptr};
   size_t work_bytes[kCtxBuffers] = {0, 0, 0, 0};
   void *scratch = nullptr; // global-memory scratch of the two-kernel transforms of very long lines (fft_any.cu)
   size_t scratch_bytes = 0;
   void *scratch2 = nullptr; // zero-padded lines of the Bluestein transforms (fft_any.cu); their M-point transforms may use `scratch`
   size_t scratch2_bytes = 0;
   struct P2P *p2p = nullptr; // peer-mapped work buffers + flags (p2p.cpp); null when unused
   // chunk-wise pipelined chains (fft_plan.cpp run_chain_pipe): the exchanges run on their own streams -- one per peer
   // of a communicator for the copy-engine pushes of the peer-memory path, comm_stream for the transports' exchanges
   cudaStream_t comm_stream = nullptr;
   cudaStream_t copy_stream[2 * kMaxP] = {};
   cudaStream_t io_stream[2] = {}; // host-array entry points: [0] uploads (H2D), [1] downloads (D2H)
   std::vector<cudaEvent_t> sync_events; // untimed events ordering the streams, reused between calls
   size_t next_sync_event = 0;
   std::vector<cudaEvent_t> buf_busy[kCtxBuffers]; // copies still reading work[w] (recorded on the copy streams)
   cudaStream_t push_stream = nullptr; // the exchange kernels of the pipelined chains (highest priority)
   int fft_grid_limit = 0;             // SMs the FFT kernels may occupy while exchange kernels share the device (0 = all)
   cudaStream_t copy_stream_for(int k);
   cudaEvent_t new_sync_event();
   void wait_buffer_idle(int w, cudaStream_t st); // st waits for the copies that read work[w]
   void mark_buffer_busy(int w, int nstreams);
   void mark_buffer_busy_on(int w, cudaStream_t st); // ... the current tail of `st`
   int sm_count = 0;    // record the current tail of copy streams 0..nstreams-1 against work[w]
   // profiling
   bool profiling = false;
   std::vector<ProfEntry> prof;
   struct Pending {
      int idx;
      cudaEvent_t a, b;
   };
   std::vector<Pending> pending;
   std::vector<cudaEvent_t> event_pool;

   void *reserve(int which, size_t bytes);
   // Collective over the ranks of the context: grow work[0..nbuf) to at least `bytes` (a size every rank computes
   // identically) and (re)publish them to the peers when anything changed -- peers hold mappings of these buffers, so they
   // are never resized by one rank alone.
   void ensure_buffers(int nbuf, size_t bytes, bool force_publish);
   bool published = false;
   void prof_begin(const char *label, double bytes, Pending &p, cudaStream_t st = nullptr);
   void prof_end(Pending &p, cudaStream_t st = nullptr);
   void prof_flush();
   int peer_rank_col(int m) const { return m * p_col + c2; } // COL communicator: same coord(2), index = coord(1)
   int peer_rank_row(int m) const { return c1 * p_col + m; } // ROW communicator: same coord(1), index = coord(2)
   // host-memory dependencies between the host-array entry points (fft_plan.cpp fft_3d_host)
   struct HostTrack {
      static constexpr size_t kMaxPieces = 72;
      const char *d2h_base = nullptr; // host range the most recent download writes, in pieces [first, second) with one event each
      size_t d2h_bytes = 0;
      std::vector<std::pair<size_t, size_t>> d2h_cut;
      cudaEvent_t d2h_ev[kMaxPieces] = {};
      const char *h2d_base = nullptr; // host range the most recent upload reads
      size_t h2d_bytes = 0;
      cudaEvent_t h2d_done = nullptr;
   } host;
   void sync_all(); // every stream of the context
   void finish_call()
   {
      if (blocking) sync_all();
   }
   ~Ctx();
};

struct ProfScope {
   Ctx *c;
   Ctx::Pending p{};
   bool on;
   cudaStream_t st;
   ProfScope(Ctx *ctx, const char *label, double bytes = 0, cudaStream_t stream = nullptr) : c(ctx), on(ctx->profiling), st(stream)
   {
      if (on) c->prof_begin(label, bytes, p, st);
   }
   ~ProfScope()
   {
      if (on) c->prof_end(p, st);
   }
};

// ---- copy (pack / unpack) kernels ----------------------------------------------------------------
struct CopyArgs {
   PieceMap in, out;  // element units of `es` bytes
   int ne, na, nb;    // extents of the index space (e: pieces axis, a, b)
   int fast_is_a;     // 1: `a` is the unit-stride axis of both sides, 0: `e` is
};
void launch_copy(Ctx *ctx, const CopyArgs &c, int es, const char *label = nullptr);

// ---- exchange kernel of the pipelined chains (push_kernels.cu) -------------------------------------------------------
constexpr int kMaxPushSegs = 8;
struct PushSeg {
   const char *src; // this rank's send buffer
   char *dst;       // the peer's receive buffer (CUDA-IPC mapped)
   size_t bytes;
};
struct PushArgs {
   int nseg, nflag;
   PushSeg seg[kMaxPushSegs];
   uint32_t *flag[kMaxPushSegs]; // arrival counters in the peers' memory: every CTA adds 1 when its stores are complete
};
int push_ctas(); // CTAs of a push launch (D2D_PUSH_CTAS): what one arrival is worth on the peers' counters
void launch_push(const PushArgs &a, cudaStream_t st);

// ---- twiddles -----------------------------------------------------------------------------------
const void *twiddles_for(int device, int n, int f64, int compact = 0);
void twiddles_release_all();

inline int elem_size(int dtype, int is_complex) { return (dtype == D2D_F64 ? 8 : 4) * (is_complex ? 2 : 1); }

// piece-map builders.  `es`-agnostic: element units.
PieceMap natural_map(const Decomp &d, int pencil, void *ptr);
// send-side map of the transpose leaving pencil `from` towards pencil `to`; the self block stays in `sendbuf`
PieceMap send_map(const Decomp &d, int from, int to, void *sendbuf, int es, bool even = false);
// recv-side map for pencil `to` fed from pencil `from`: peers' blocks in recvbuf, own block in sendbuf
PieceMap recv_map(const Decomp &d, int from, int to, void *recvbuf, void *sendbuf, int es, bool even = false);
// the exchange itself (peers other than self); element size es
// send_w / recv_w: indices of the context's work buffers holding sendbuf / recvbuf (-1: a user array); the peer-memory
// path needs recvbuf == work[recv_w]
void exchange(Ctx *ctx, const Decomp &d, int from, int to, const void *sendbuf, void *recvbuf, int es, int send_w, int recv_w, bool even = false);
size_t uniform_pencil_bytes(const Ctx *ctx, const Decomp &d, int es, bool even = false); // largest pencil (EVEN: padded buffer) of any rank, in bytes
int64_t send_total(const Decomp &d, int from, int to);
int64_t recv_total(const Decomp &d, int from, int to);
int comm_size(const Decomp &d, int from, int to);
// private wire layouts of the fused 3-D transforms (decomp.cpp)
void fft_stage_batch(const Decomp &d, int pencil, int &na, int &nb);
PieceMap fft_user_map(const Decomp &d, int pencil, void *ptr);
struct LinkSide {
   int np, me;
   int e0[kMaxP + 1];
   int64_t cnt[kMaxP], disp[kMaxP], total;
   long long se[kMaxP], sa[kMaxP], sb[kMaxP];
};
void fft_link_side(const Decomp &d, int pencil, int other, int padq, LinkSide &L);
struct LinkChunk {
   int np, me, axis_is_a, nf; // axis_is_a: the free axis is batch axis a (else b) of the stage on `pencil`; nf: its extent
   int64_t off[kMaxP], cnt[kMaxP];
};
void fft_link_chunk(const Decomp &d, int pencil, int other, int padq, int f0, int f1, LinkChunk &C);
int64_t fft_work_elems(const Decomp &d, int padq);
PieceMap fft_link_map(const Decomp &d, int pencil, int other, void *peers_buf, void *self_buf, int es, bool consumer, int padq);
void fft_exchange(Ctx *ctx, const Decomp &d, int from, int to, const void *sendbuf, void *recvbuf, int es, int padq);

void transpose(Ctx *ctx, const Decomp &d, int direction, int es, const void *src, void *dst);

// peer-memory path of the fused transforms (p2p.cpp)
void p2p_publish(Ctx *ctx);
void p2p_unpublish(Ctx *ctx); // collective: close the mappings of the peers' work buffers (before they are reallocated)
void p2p_destroy(struct P2P *p);
bool p2p_active(const Ctx *ctx);
void p2p_invalidate(Ctx *ctx);
void *p2p_peer_work(const Ctx *ctx, int w, int rank);
size_t p2p_peer_bytes(const Ctx *ctx, int w, int rank);
uint32_t p2p_next_epoch(Ctx *ctx);
uint32_t p2p_reserve_seq(Ctx *ctx, uint32_t n); // n consecutive chunk sequence numbers; returns the number before the first
// per-peer arrival counters of the kernel-driven exchange (flags `which` = 3): the sender's kernels add push_ctas() per chunk,
// the receiver waits for push_ctas() x (chunks expected from that peer so far)
uint32_t p2p_expect_push(Ctx *ctx, int peer); // one more chunk expected from `peer`; returns the counter value to wait for
uint32_t *p2p_push_flag(Ctx *ctx, int peer);  // address of this rank's arrival counter in `peer`'s memory
// flags `which`: 0 ready (epochs), 1 done (epochs of the fused chain), 2 arrived (chunk sequence numbers of the pipelined chain)
void p2p_signal(Ctx *ctx, int peer, int which, uint32_t value, cudaStream_t st = nullptr);
void p2p_wait(Ctx *ctx, int peer, int which, uint32_t value, cudaStream_t st = nullptr);
// all-to-all(v) over peer memory for the bare transposes: every PeerXfer's recvptr must lie in work[recv_w] (the peers push
// into the same offsets of their peers' work[recv_w]); dst_off[i] = byte offset in the DESTINATION rank's work[recv_w]
void p2p_exchange(Ctx *ctx, const std::vector<PeerXfer> &xf, const std::vector<size_t> &dst_off, int recv_w, int send_w);

} // namespace d2d

struct d2d_ctx {
   d2d::Ctx c;
};
struct d2d_decomp {
   d2d::Decomp d;
   d2d::Ctx *ctx;
};
