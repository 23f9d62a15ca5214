// ctx.cpp -- rank context (stream, work buffers, per-stage device timers), twiddle tables, the FFT
// kernel registry and the in-process (thread-per-rank) transport.
#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <tuple>
#include <cstring>

#include "common.h"
#include "fft_registry.h"

namespace d2d {

static thread_local std::string g_last_error;
void set_last_error(const std::string &m) { g_last_error = m; }
const std::string &get_last_error() { return g_last_error; }

// ---- registry ----------------------------------------------------------------------------------
static std::vector<FftKernelInfo> &registry()
{
   static std::vector<FftKernelInfo> r;
   return r;
}
void fft_register(const FftKernelInfo &k) { registry().push_back(k); }
const FftKernelInfo *fft_find(int n, int f64, int kind, int mode, int pairvec, int line_in)
{
   for (const auto &k : registry())
      if (!k.v2 && k.n == n && k.f64 == f64 && k.kind == kind && k.mode == mode && k.pairvec == pairvec && k.line_in == line_in) return &k;
   return nullptr;
}
const FftKernelInfo *fft_find_v2(int n, int f64, int mode, int inl, int row_bytes, int merged)
{
   // row_bytes: bytes of TX adjacent complex lines (64: two blocks per SM; 128: one L2 line per row)
   for (const auto &k : registry())
      if (k.v2 && k.n == n && k.f64 == f64 && k.mode == mode && k.inl == inl && k.tx * (f64 ? 16 : 8) == row_bytes && k.merged == merged) return &k;
   return nullptr;
}
int fft_registry_size() { return (int)registry().size(); }
const FftKernelInfo *fft_registry_at(int i) { return &registry()[i]; }

// ---- twiddles: per-pass tables [r-1][q] = exp(-2 pi i r q / (Ns R)), q in [0,Ns) -----------------
namespace {
struct TwKey {
   int device, n, f64;
   bool operator<(const TwKey &o) const { return std::tie(device, n, f64) < std::tie(o.device, o.n, o.f64); }
};
std::mutex g_tw_mutex;
std::map<TwKey, void *> g_tw;
} // namespace

const void *twiddles_for(int device, int n, int f64, int compact)
{
   // compact = 1: the v2 layout (fft_kernel_v2.cuh PlanInfo2): radix-2/4 passes keep only the r = 1 row
   std::lock_guard<std::mutex> lk(g_tw_mutex);
   TwKey key{device, n, f64 + 2 * compact};
   auto it = g_tw.find(key);
   if (it != g_tw.end()) return it->second;
   const FftKernelInfo *k = nullptr;
   for (int i = 0; i < fft_registry_size(); i++)
      if (fft_registry_at(i)->n == n && fft_registry_at(i)->f64 == f64 && fft_registry_at(i)->v2 == compact) { k = fft_registry_at(i); break; }
   D2D_REQUIRE(k != nullptr, "no FFT kernel compiled for this length");
   std::vector<double> hd;
   long long ns = 1;
   for (int p = 0; p < k->npass; p++) {
      const int R = k->radix[p];
      if (p >= 1) {
         const int rmax = (compact && (R == 2 || R == 4)) ? 1 : R - 1;
         for (int r = 1; r <= rmax; r++)
            for (long long q = 0; q < ns; q++) {
               const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)(r * q) / (long double)(ns * R);
               hd.push_back((double)cosl(ang));
               hd.push_back((double)sinl(ang));
            }
      }
      ns *= R;
   }
   D2D_REQUIRE((int)(hd.size() / 2) == k->tw_total, "twiddle table size mismatch");
   if (hd.empty()) hd.resize(2, 0.0);
   void *dptr = nullptr;
   const size_t bytes = (hd.size() / 2) * (f64 ? 16 : 8);
   D2D_CHECK_CUDA(cudaMalloc(&dptr, bytes));
   if (f64) {
      D2D_CHECK_CUDA(cudaMemcpy(dptr, hd.data(), bytes, cudaMemcpyHostToDevice));
   } else {
      std::vector<float> hf(hd.begin(), hd.end());
      D2D_CHECK_CUDA(cudaMemcpy(dptr, hf.data(), bytes, cudaMemcpyHostToDevice));
   }
   g_tw[key] = dptr;
   return dptr;
}

void twiddles_release_all()
{
   std::lock_guard<std::mutex> lk(g_tw_mutex);
   for (auto &kv : g_tw) cudaFree(kv.second);
   g_tw.clear();
}

// ---- context -------------------------------------------------------------------------------------
void *Ctx::reserve(int which, size_t bytes)
{
   if (bytes > work_bytes[which]) {
      D2D_CHECK_CUDA(cudaStreamSynchronize(stream));
      p2p_invalidate(this); // peers hold mappings of the old buffer; the next plan creation publishes again
      if (work[which]) D2D_CHECK_CUDA(cudaFree(work[which]));
      work[which] = nullptr;
      work_bytes[which] = 0;
      D2D_CHECK_CUDA(cudaMalloc(&work[which], bytes));
      work_bytes[which] = bytes;
   }
   return work[which];
}

void Ctx::sync_all()
{
   D2D_CHECK_CUDA(cudaSetDevice(device));
   D2D_CHECK_CUDA(cudaStreamSynchronize(stream));
   if (comm_stream) D2D_CHECK_CUDA(cudaStreamSynchronize(comm_stream));
   for (int k = 0; k < 2 * kMaxP; k++)
      if (copy_stream[k]) D2D_CHECK_CUDA(cudaStreamSynchronize(copy_stream[k]));
   for (cudaStream_t s : io_stream)
      if (s) D2D_CHECK_CUDA(cudaStreamSynchronize(s));
   if (push_stream) D2D_CHECK_CUDA(cudaStreamSynchronize(push_stream));
}

cudaStream_t Ctx::copy_stream_for(int k)
{
   D2D_REQUIRE(k >= 0 && k < 2 * kMaxP, "copy stream index out of range");
   if (!copy_stream[k]) D2D_CHECK_CUDA(cudaStreamCreateWithFlags(&copy_stream[k], cudaStreamNonBlocking));
   return copy_stream[k];
}

cudaEvent_t Ctx::new_sync_event()
{
   // a ring: an event is re-recorded only long after the waits that referred to its previous record were enqueued (a
   // wait captures the record that is current when it is enqueued, so re-recording never disturbs it)
   if (sync_events.size() < 512) {
      cudaEvent_t e;
      D2D_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      sync_events.push_back(e);
      return e;
   }
   return sync_events[next_sync_event++ % sync_events.size()];
}

void Ctx::wait_buffer_idle(int w, cudaStream_t st)
{
   for (cudaEvent_t e : buf_busy[w]) D2D_CHECK_CUDA(cudaStreamWaitEvent(st, e, 0));
}

void Ctx::mark_buffer_busy(int w, int nstreams)
{
   while ((int)buf_busy[w].size() < nstreams) {
      cudaEvent_t e;
      D2D_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      buf_busy[w].push_back(e);
   }
   for (int k = 0; k < nstreams; k++) D2D_CHECK_CUDA(cudaEventRecord(buf_busy[w][k], copy_stream_for(k)));
}

void Ctx::ensure_buffers(int nbuf, size_t bytes, bool force_publish)
{
   bool grow = false;
   for (int i = 0; i < nbuf; i++) grow = grow || work_bytes[i] < bytes;
   if (!grow && !force_publish && published) return;
   if (grow) {
      // nobody may still be copying out of / into the old buffers
      for (int k = 0; k < 2 * kMaxP; k++)
         if (copy_stream[k]) D2D_CHECK_CUDA(cudaStreamSynchronize(copy_stream[k]));
      p2p_unpublish(this); // collective: every rank closes its mappings of the peers' buffers before anybody frees one
      for (int i = 0; i < nbuf; i++) reserve(i, bytes);
   }
   p2p_publish(this); // collective (no-op for a single rank / in-process groups)
   published = true;
}

void Ctx::mark_buffer_busy_on(int w, cudaStream_t st)
{
   if (buf_busy[w].empty()) {
      cudaEvent_t e;
      D2D_CHECK_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      buf_busy[w].push_back(e);
   }
   D2D_CHECK_CUDA(cudaEventRecord(buf_busy[w][0], st));
}

void Ctx::prof_begin(const char *label, double bytes, Pending &p, cudaStream_t st)
{
   int idx = -1;
   for (size_t i = 0; i < prof.size(); i++)
      if (prof[i].label == label) { idx = (int)i; break; }
   if (idx < 0) {
      prof.push_back(ProfEntry{label, 0, 0, 0});
      idx = (int)prof.size() - 1;
   }
   prof[idx].bytes += bytes;
   auto get_event = [&]() {
      cudaEvent_t e;
      if (!event_pool.empty()) { e = event_pool.back(); event_pool.pop_back(); }
      else D2D_CHECK_CUDA(cudaEventCreate(&e));
      return e;
   };
   p.idx = idx;
   p.a = get_event();
   p.b = get_event();
   D2D_CHECK_CUDA(cudaEventRecord(p.a, st ? st : stream));
}
void Ctx::prof_end(Pending &p, cudaStream_t st)
{
   cudaEventRecord(p.b, st ? st : stream);
   pending.push_back(p);
}
void Ctx::prof_flush()
{
   if (pending.empty()) return;
   D2D_CHECK_CUDA(cudaStreamSynchronize(stream));
   if (comm_stream) D2D_CHECK_CUDA(cudaStreamSynchronize(comm_stream));
   if (push_stream) D2D_CHECK_CUDA(cudaStreamSynchronize(push_stream));
   for (int k = 0; k < 2 * kMaxP; k++)
      if (copy_stream[k]) D2D_CHECK_CUDA(cudaStreamSynchronize(copy_stream[k]));
   for (auto &p : pending) {
      float ms = 0;
      D2D_CHECK_CUDA(cudaEventElapsedTime(&ms, p.a, p.b));
      prof[p.idx].total_ms += ms;
      prof[p.idx].calls += 1;
      event_pool.push_back(p.a);
      event_pool.push_back(p.b);
   }
   pending.clear();
}

Ctx::~Ctx()
{
   cudaSetDevice(device);
   if (stream) cudaStreamSynchronize(stream);
   if (p2p) p2p_destroy(p2p);
   p2p = nullptr;
   tr.reset();
   if (comm_stream) { cudaStreamSynchronize(comm_stream); cudaStreamDestroy(comm_stream); }
   for (int k = 0; k < 2 * kMaxP; k++)
      if (copy_stream[k]) { cudaStreamSynchronize(copy_stream[k]); cudaStreamDestroy(copy_stream[k]); }
   for (cudaStream_t s : io_stream)
      if (s) { cudaStreamSynchronize(s); cudaStreamDestroy(s); }
   if (push_stream) { cudaStreamSynchronize(push_stream); cudaStreamDestroy(push_stream); }
   for (auto e : host.d2h_ev)
      if (e) cudaEventDestroy(e);
   if (host.h2d_done) cudaEventDestroy(host.h2d_done);
   for (auto e : sync_events) cudaEventDestroy(e);
   for (auto &v : buf_busy)
      for (auto e : v) cudaEventDestroy(e);
   for (auto &p : pending) { cudaEventDestroy(p.a); cudaEventDestroy(p.b); }
   for (auto e : event_pool) cudaEventDestroy(e);
   for (int i = 0; i < kCtxBuffers; i++)
      if (work[i]) cudaFree(work[i]);
   if (scratch) cudaFree(scratch);
   if (scratch2) cudaFree(scratch2);
   if (stream) cudaStreamDestroy(stream);
}

// ---- in-process transport: one host thread per rank, exchange = device-to-device copies ----------
struct Group {
   int n;
   std::mutex m;
   std::condition_variable cv;
   int waiting = 0;
   long long generation = 0;
   bool aborted = false; // a rank failed: every present and future barrier throws instead of waiting for it
   std::vector<std::vector<PeerXfer>> posted;
   explicit Group(int n_) : n(n_), posted(n_) {}
   void barrier()
   {
      std::unique_lock<std::mutex> lk(m);
      const long long gen = generation;
      if (aborted) throw Error(3, "rank group aborted: another rank of this group failed");
      if (++waiting == n) {
         waiting = 0;
         generation++;
         cv.notify_all();
      } else {
         cv.wait(lk, [&] { return generation != gen || aborted; });
         if (generation == gen) throw Error(3, "rank group aborted: another rank of this group failed");
      }
   }
   void abort()
   {
      std::lock_guard<std::mutex> lk(m);
      aborted = true;
      cv.notify_all();
   }
};
Group *group_create(int nranks) { return new Group(nranks); }
void group_destroy(Group *g) { delete g; }
void group_abort(Group *g) { g->abort(); }

namespace {
struct LocalTransport : Transport {
   Group *g;
   int rank;
   LocalTransport(Group *g_, int r) : g(g_), rank(r) {}
   int kind() const override { return D2D_TRANSPORT_LOCAL; }
   void exchange(const std::vector<PeerXfer> &xf, cudaStream_t st) override
   {
      D2D_CHECK_CUDA(cudaStreamSynchronize(st)); // my send segments are complete
      g->posted[rank] = xf;
      g->barrier();
      for (const auto &x : xf) {
         const PeerXfer *src = nullptr;
         for (const auto &y : g->posted[x.peer])
            if (y.peer == rank) src = &y;
         D2D_REQUIRE(src != nullptr && src->sendbytes == x.recvbytes, "local exchange: send/recv size mismatch");
         if (x.recvbytes) D2D_CHECK_CUDA(cudaMemcpyAsync(x.recvptr, src->sendptr, x.recvbytes, cudaMemcpyDefault, st));
      }
      D2D_CHECK_CUDA(cudaStreamSynchronize(st));
      g->barrier(); // peers may now overwrite their send buffers
   }
   void barrier(cudaStream_t st) override
   {
      D2D_CHECK_CUDA(cudaStreamSynchronize(st));
      g->barrier();
   }
   void allgather(const void *, void *, size_t, cudaStream_t) override { D2D_REQUIRE(false, "allgather: not available on the in-process transport"); }
};
} // namespace
Transport *make_local_transport(Group *g, int rank) { return new LocalTransport(g, rank); }

// ---- bootstrap-only transport: the caller's all-gather (MPI / torch.distributed), no data plane of its own ----------
namespace {
struct BootTransport : Transport {
   d2d_allgather_fn fn;
   void *user;
   int nranks, rank;
   BootTransport(d2d_allgather_fn f, void *u, int n, int r) : fn(f), user(u), nranks(n), rank(r) {}
   int kind() const override { return D2D_TRANSPORT_BOOT; }
   void exchange(const std::vector<PeerXfer> &, cudaStream_t) override
   {
      D2D_REQUIRE(false, "bootstrap transport: the exchange needs the peer-memory path (CUDA IPC between the ranks' devices), which is not active");
   }
   void allgather(const void *send_host, void *recv_host, size_t bytes, cudaStream_t) override
   {
      D2D_REQUIRE(fn(user, send_host, recv_host, (int64_t)bytes) == 0, "bootstrap all-gather callback failed");
   }
   void barrier(cudaStream_t st) override
   {
      D2D_CHECK_CUDA(cudaStreamSynchronize(st));
      std::vector<char> r(nranks);
      char s = 0;
      allgather(&s, r.data(), 1, st);
   }
};
} // namespace
Transport *make_boot_transport(d2d_allgather_fn fn, void *user, int nranks, int rank) { return new BootTransport(fn, user, nranks, rank); }

} // namespace d2d
