// fft_any.cu -- host side of the arbitrary-length FFT kernel (fft_any.cuh): factorisation, geometry,
// the table of n-th roots of unity, launch.  Used by run_stage (fft_plan.cpp) for every transform length
// without a compiled power-of-two kernel.
#include <cmath>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "common.h"
#include "fft_any.cuh"
#include "fft_registry.h"

namespace d2d {

cudaError_t fft_dispatch(Ctx *ctx, FftArgs &g, int f64, int mode, int kind, int pairvec, bool wide_real); // fft_plan.cpp

namespace {
struct RootKey {
   int device, n, f64;
   bool operator<(const RootKey &o) const { return std::tie(device, n, f64) < std::tie(o.device, o.n, o.f64); }
};
std::mutex g_root_mutex;
std::map<RootKey, void *> g_roots;

constexpr size_t kAnyLineBudget = 200 * 1024; // bounds the longest line of the shared-memory kernel (fft_any_max_n)
constexpr size_t kAnySmemBudget = 210 * 1024; // what the geometry may use
constexpr size_t kAnySmemMax = 220 * 1024;    // opt-in limit of sm_100a (227 KB) minus the kernel's static shared memory (6 KB of line tables)

// W[k] = exp(-2 pi i k / n): octant-reduced so that every entry is accurate to the last bit or so
const void *roots_for(int device, int n, int f64)
{
   std::lock_guard<std::mutex> lk(g_root_mutex);
   RootKey key{device, n, f64};
   auto it = g_roots.find(key);
   if (it != g_roots.end()) return it->second;
   std::vector<double> hd(2 * (size_t)n);
   const long double pi = 3.14159265358979323846264338327950288L;
   for (int k = 0; k < n; k++) {
      // angle = 2 pi k / n, reduced to [0, pi/4] through the symmetries of the circle (exact integer arithmetic on 8k/n)
      long long num = 8LL * k; // angle = (pi/4) * num / n
      const int oct = (int)(num / n);
      long long rem = num - (long long)oct * n; // in [0, n)
      long double c, s;
      if (oct & 1) { // odd octant: measure from the end of the octant
         const long double t = (pi / 4) * (long double)(n - rem) / (long double)n;
         c = sinl(t); s = cosl(t);
      } else {
         const long double t = (pi / 4) * (long double)rem / (long double)n;
         c = cosl(t); s = sinl(t);
      }
      // (c, s) = (cos, sin) of the angle folded into the first quadrant's two octants; unfold by quadrant
      long double cc, ss;
      switch (oct >> 1) {
      case 0: cc = c; ss = s; break;
      case 1: cc = -s; ss = c; break;
      case 2: cc = -c; ss = -s; break;
      default: cc = s; ss = -c; break;
      }
      hd[2 * k] = (double)cc;
      hd[2 * k + 1] = (double)(-ss);
   }
   void *dptr = nullptr;
   const size_t bytes = (size_t)n * (f64 ? 16 : 8);
   D2D_CHECK_CUDA(cudaMalloc(&dptr, bytes));
   if (f64) {
      D2D_CHECK_CUDA(cudaMemcpy(dptr, hd.data(), bytes, cudaMemcpyHostToDevice));
   } else {
      std::vector<float> hf(hd.begin(), hd.end());
      D2D_CHECK_CUDA(cudaMemcpy(dptr, hf.data(), bytes, cudaMemcpyHostToDevice));
   }
   g_roots[key] = dptr;
   return dptr;
}

template <typename T, int MODE, bool BIG> cudaError_t launch_any_build(const FftArgsAny &ga, size_t smem, cudaStream_t st)
{
   auto kern = fft_any_kernel<T, MODE, BIG>;
   constexpr int threads = BIG ? kAnyThreadsBig : kAnyThreads;
   // per-device caches (function attributes belong to a device)
   static bool smem_set_of[kMaxDevices] = {false};
   static int sms_of[kMaxDevices] = {0};
   int dev = 0;
   if (cudaError_t e0 = cudaGetDevice(&dev); e0 != cudaSuccess) return e0;
   if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
   if (smem + 8192 > 48 * 1024 && !smem_set_of[dev]) { // the kernel also has 6 KB of static shared memory (line tables)
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAnySmemMax);
      if (e != cudaSuccess) return e;
      smem_set_of[dev] = true;
   }
   int &sms = sms_of[dev];
   if (!sms) {
      cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      if (e != cudaSuccess) return e;
   }
   int per_sm = 0;
   cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem);
   if (e != cudaSuccess) return e;
   if (per_sm < 1) return cudaErrorLaunchOutOfResources;
   const long long total = (long long)ga.a.na * ga.a.nb;
   const long long groups = (total + (1 << ga.lines_log2) - 1) >> ga.lines_log2;
   if (groups <= 0) return cudaSuccess;
   const long long resident = (long long)sms * per_sm;
   const unsigned blocks = (unsigned)(groups < resident ? groups : resident);
   kern<<<blocks, threads, smem, st>>>(ga);
   return cudaGetLastError();
}
template <typename T, int MODE> cudaError_t launch_any(const FftArgsAny &ga, size_t smem, cudaStream_t st)
{
   return ga.big ? launch_any_build<T, MODE, true>(ga, smem, st) : launch_any_build<T, MODE, false>(ga, smem, st);
}
} // namespace

// radices of n in pass order (fft_any.cuh any_factorize) and the build of the kernel that runs them: the big build (radices
// up to 31, 128 threads) when it needs fewer passes or turns a run-time radix (a prime in 17..31) into a register one.
// D2D_ANY_BIG: 0 never, 1 as described (default), 2 whenever a radix above 16 helps or not
int fft_any_factorize(int n, int *radix, int maxp, int *big)
{
   static const int mode = getenv("D2D_ANY_BIG") ? atoi(getenv("D2D_ANY_BIG")) : 1;
   int rs[kMaxAnyPass], rb[kMaxAnyPass];
   bool nb = false;
   const int ns = any_factorize(n, rs, kMaxAnyPass, false, nullptr);
   const int nbg = any_factorize(n, rb, kMaxAnyPass, true, &nb);
   bool use_big = false;
   if (mode >= 1 && nb && nbg <= kMaxAnyPass) {
      bool runtime_small = false; // a prime in 17..31 that the small build would run from shared memory
      for (int i = 0; i < ns && i < kMaxAnyPass; i++) runtime_small = runtime_small || (rs[i] > kAnyMaxFixedSmall && rs[i] <= kAnyMaxFixed);
      use_big = mode >= 2 || nbg < ns || runtime_small;
   }
   const int np = use_big ? nbg : ns;
   for (int i = 0; i < np && i < maxp; i++) radix[i] = use_big ? rb[i] : rs[i];
   if (big) *big = use_big ? 1 : 0;
   return np;
}

void fft_any_release_blue(); // below (Bluestein tables)
void fft_any_release_all()
{
   {
      std::lock_guard<std::mutex> lk(g_root_mutex);
      for (auto &kv : g_roots) cudaFree(kv.second);
      g_roots.clear();
   }
   fft_any_release_blue();
}

// largest transform length the shared-memory kernel takes (one line per block, two buffers + the root table)
int fft_any_max_n(int f64) { return (int)(kAnyLineBudget / (3 * (size_t)(f64 ? 16 : 8))) - 1; }

// geometry of one launch: lines per block (a power of two), shared-memory pitch, which axis the lanes follow on each side
static size_t configure_any(FftArgsAny &ga, int f64, int mode, bool split_step)
{
   const FftArgs &g = ga.a;
   const int n = g.n;
   const size_t ces = f64 ? 16 : 8;
   ga.npass = fft_any_factorize(n, ga.radix, kMaxAnyPass, &ga.big);
   D2D_REQUIRE(ga.npass <= kMaxAnyPass, "too many factors");
   const int max_lines = ga.big ? kAnyThreadsBig : kAnyThreads; // at most one line per thread
   ga.pitch = n | 1;
   if (!split_step) { // which axis is unit-stride on each side
      if (mode == MODE_R2C) ga.in_fast_a = (g.rsa == 1 && g.rse != 1);
      else ga.in_fast_a = (g.in.sa[0] == 1 && g.in.se[0] != 1);
      if (mode == MODE_C2R) ga.out_fast_a = (g.rsa == 1 && g.rse != 1);
      else ga.out_fast_a = (g.out.sa[0] == 1 && g.out.se[0] != 1);
   }
   // The input of the next group of lines is prefetched with cp.async into a third line buffer while the current group runs
   // its passes (plain c2c stages and r2c; D2D_ANY_ASYNC=0 turns it off): ncu showed the two-buffer kernel waiting for its
   // global loads 33-52 % of the time (profiles/r02_k_ncu_any510.txt).
   static const int async_enabled = getenv("D2D_ANY_ASYNC") ? atoi(getenv("D2D_ANY_ASYNC")) : 1;
   const bool can_async = async_enabled && !split_step && (mode == MODE_C2C || mode == MODE_R2C);
   // lines per block (a power of two): 128-byte rows when lines are strided, about 2048 elements of work per block,
   // within the shared-memory budget; narrower (never below the row width) when that lets two blocks share an SM
   auto pow2_ceil = [](long long v) { int l = 0; while ((1LL << l) < v) l++; return l; };
   static const int row_bytes = getenv("D2D_ANY_ROW_BYTES") ? atoi(getenv("D2D_ANY_ROW_BYTES")) : 128; // experiments: 64 halves the tile
   static const int min_row_bytes = getenv("D2D_ANY_MIN_ROW_BYTES") ? atoi(getenv("D2D_ANY_MIN_ROW_BYTES")) : 64; // experiments
   const int want_rows_log2 = pow2_ceil((long long)(std::max(16, row_bytes) / (int)ces));
   const int min_rows_log2 = pow2_ceil((long long)(std::max<int>(min_row_bytes, (int)ces) / (int)ces));
   auto pick = [&](int nbuf, int &ll) {
      auto smem_of = [&](int lines) { return ces * ((size_t)n + (size_t)nbuf * (size_t)lines * ga.pitch); };
      ll = std::max(want_rows_log2, pow2_ceil((2048 + n - 1) / n));
      ll = std::min(ll, pow2_ceil(max_lines));
      ll = std::min(ll, pow2_ceil((long long)g.na * g.nb));
      while (ll > 0 && smem_of(1 << ll) > kAnySmemBudget) ll--;
      // two blocks per SM (one loads / stores while the other computes) are worth more than 128-byte rows: 510^3 pair 14.4 ->
      // 11.4 ms with 64-byte rows (profiles/r01_l_any_and_transposes.txt)
      while (ll > min_rows_log2 && smem_of(1 << ll) > 108 * 1024) ll--;
      return smem_of(1 << ll);
   };
   int ll2 = 0, ll3 = 0;
   const size_t s2 = pick(2, ll2);
   size_t s3 = 0;
   // three buffers when the line still fits; a block that then has the SM to itself keeps its rows as wide as two buffers allow
   bool use3 = false;
   if (can_async) {
      s3 = pick(3, ll3);
      use3 = s3 <= kAnySmemBudget && ll3 >= std::min(ll2, min_rows_log2);
   }
   ga.nbuf = use3 ? 3 : 2;
   ga.async_in = use3 ? 1 : 0;
   ga.lines_log2 = use3 ? ll3 : ll2;
   return use3 ? s3 : s2;
}

// Lengths beyond the shared-memory kernel (n > fft_any_max_n): n = n1 n2 with both factors within its reach, as TWO launches of
// the same kernel through a scratch array in global memory (the classic four-step split; the reference's generic backend takes
// any length, src/glassman.f90:29-67):
//   step A  for every i2 < n2: the n1-point transform over i1 of x[i1 n2 + i2], times exp(-2 pi i i2 k1 / n)  -> scratch[k1 n2 + i2]
//   step B  for every k1 < n1: the n2-point transform over i2 of scratch[k1 n2 + i2]                          -> X[k1 + n1 k2]
// Real transforms run as complex ones with the conversions of the reference's generic backend at the two ends (zero imaginary
// part in, bins 0 .. n/2 out, src/fft_generic.f90:236-244; Hermitian completion in, real part out, :320-337).  A prime length
// above the limit still has no kernel.
static cudaError_t fft_any_launch_split(Ctx *ctx, const FftArgs &g, int f64, int mode)
{
   const int n = g.n, lim = fft_any_max_n(f64);
   int n1 = 0;
   for (int d = 2; (long long)d * d <= n; d++)
      if (n % d == 0 && n / d <= lim) { n1 = d; if (d <= lim) break; }
   D2D_REQUIRE(n1 >= 2 && n1 <= lim && n / n1 <= lim, "transform length " + std::to_string(n) + " is not supported (a factor above " +
                                                         std::to_string(lim) + " has no kernel)");
   const int n2 = n / n1;
   const size_t ces = f64 ? 16 : 8;
   const long long lines_a = (mode == MODE_C2C) ? g.na : g.na_real; // real transforms: one complex line per real line
   const long long lines = lines_a * g.nb;
   D2D_REQUIRE(lines_a * std::max(n1, n2) < (1LL << 31), "too many lines for the two-kernel transform");
   const size_t need = (size_t)lines * (size_t)n * ces;
   if (need > ctx->scratch_bytes) {
      D2D_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
      if (ctx->scratch) D2D_CHECK_CUDA(cudaFree(ctx->scratch));
      ctx->scratch = nullptr;
      ctx->scratch_bytes = 0;
      D2D_CHECK_CUDA(cudaMalloc(&ctx->scratch, need));
      ctx->scratch_bytes = need;
   }
   PieceMap sm{};
   sm.np = 1;
   sm.e0[0] = 0; sm.e0[1] = n;
   sm.ptr[0] = ctx->scratch;
   sm.se[0] = 1; sm.sa[0] = n; sm.sb[0] = (long long)n * lines_a;
   const void *big = roots_for(ctx->device, n, f64);
   cudaError_t e = cudaSuccess;
   for (int step = 0; step < 2 && e == cudaSuccess; step++) {
      FftArgsAny ga{};
      ga.a = g;
      ga.a.n = step == 0 ? n1 : n2;
      ga.a.tw = roots_for(ctx->device, ga.a.n, f64);
      ga.split = step == 0 ? n2 : n1;
      ga.a.na = (int)(lines_a * ga.split);
      ga.a.na_real = 0;
      ga.n_full = n;
      ga.big_tw = big;
      if (step == 0) {
         ga.a.out = sm;
         ga.in_mul = n2; ga.in_add = 1; ga.out_mul = n2; ga.out_add = 1;
         ga.tw_n = g.passthrough ? 0 : n;
         ga.real_in = mode == MODE_R2C;
         ga.herm_in = mode == MODE_C2R;
         ga.in_fast_a = 1; ga.out_fast_a = 1; // consecutive sub-indices are consecutive elements on both sides
      } else {
         ga.a.in = sm;
         ga.in_mul = 1; ga.in_add = n2;
         ga.out_mul = g.passthrough ? 1 : n1; ga.out_add = g.passthrough ? n2 : 1;
         ga.half_out = mode == MODE_R2C;
         ga.real_out = mode == MODE_C2R;
         ga.in_fast_a = 0; ga.out_fast_a = g.passthrough ? 0 : 1;
      }
      const size_t smem = configure_any(ga, f64, MODE_C2C, true);
      e = f64 ? launch_any<double, MODE_C2C>(ga, smem, ctx->stream) : launch_any<float, MODE_C2C>(ga, smem, ctx->stream);
   }
   return e;
}

// ---- Bluestein: lengths with a prime factor beyond the shared-memory kernel --------------------------------------------
// The reference's generic backend takes ANY length (src/glassman.f90:29-67 falls back to an O(n * factor) loop); here a
// length that neither fits the shared-memory kernel nor splits into two factors that do becomes a circular convolution of
// power-of-two length M >= 2n - 1 (chirp-z):  with c[j] = exp(-pi i j^2 / n),
//    X[k] = c[k] * sum_j (x[j] c[j]) conj(c[k - j])  =  c[k] * IFFT_M( FFT_M(x c, zero-padded) . FFT_M(b) )[k],
//    b[j] = b[M - j] = conj(c[j]) for j < n, 0 elsewhere.
// Five launches per chunk of lines through a second scratch array: chirp-in, FFT_M, spectrum product, inverse FFT_M, chirp-out;
// the M-point transforms run on the compiled power-of-two kernels (or their two-kernel split).  Real transforms use the
// conversions of the reference's generic backend at the two ends, like the two-kernel split above.
namespace {

template <typename T, int MODE>
__global__ void __launch_bounds__(256) blue_in_kernel(const __grid_constant__ FftArgs g, const typename Vec2<T>::type *__restrict__ chirp,
                                                      typename Vec2<T>::type *__restrict__ scr, int logM, int lines_a, long long l0, long long nl)
{
   using T2 = typename Vec2<T>::type;
   const int n = g.n;
   const long long total = nl << logM;
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
      const long long l = idx >> logM;
      const int j = (int)(idx - (l << logM));
      T2 x = T2{0, 0};
      if (j < n) {
         const long long id = l0 + l, b = id / lines_a, a = id - b * lines_a;
         if constexpr (MODE == MODE_R2C) {
            x.x = reinterpret_cast<const T *>(g.rptr)[(long long)j * g.rse + a * g.rsa + b * g.rsb];
         } else if constexpr (MODE == MODE_C2R) { // Hermitian completion (src/fft_generic.f90:320-337)
            if (2 * j > n) {
               x = load_piece<T2>(g.in, n - j, a, b);
               x.y = -x.y;
            } else {
               x = load_piece<T2>(g.in, j, a, b);
               if (j == 0 || 2 * j == n) x.y = 0;
            }
         } else {
            x = load_piece<T2>(g.in, j, a, b);
         }
         if (g.backward) x.y = -x.y;
         if (!g.passthrough) x = cmul(x, chirp[j]);
      }
      scr[idx] = x;
   }
}

template <typename T>
__global__ void __launch_bounds__(256) blue_mul_kernel(typename Vec2<T>::type *__restrict__ scr, const typename Vec2<T>::type *__restrict__ bspec, int logM,
                                                       long long total)
{
   const long long mask = (1LL << logM) - 1;
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x)
      scr[idx] = cmul(scr[idx], bspec[idx & mask]);
}

template <typename T, int MODE>
__global__ void __launch_bounds__(256) blue_out_kernel(const __grid_constant__ FftArgs g, const typename Vec2<T>::type *__restrict__ chirp,
                                                       const typename Vec2<T>::type *__restrict__ scr, int logM, int lines_a, long long l0, long long nl)
{
   using T2 = typename Vec2<T>::type;
   const int n = g.n, nout = (MODE == MODE_R2C) ? n / 2 + 1 : n;
   const long long total = nl * nout;
   for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
      const long long l = idx / nout;
      const int k = (int)(idx - l * nout);
      const long long id = l0 + l, b = id / lines_a, a = id - b * lines_a;
      T2 y = scr[(l << logM) + k];
      if (!g.passthrough) y = cmul(y, chirp[k]);
      if (g.backward) y.y = -y.y;
      if constexpr (MODE == MODE_C2R) reinterpret_cast<T *>(g.rptr)[(long long)k * g.rse + a * g.rsa + b * g.rsb] = y.x;
      else store_piece<T2>(g.out, k, a, b, y);
   }
}

struct BlueTab {
   void *chirp = nullptr, *bspec = nullptr;
   int logM = 0;
};
std::map<RootKey, BlueTab> g_blue;

// chirp c[j] = exp(-pi i j^2 / n) (j^2 reduced mod 2n in integers) and B = FFT_M(b) / M, both in extended precision on the host
const BlueTab &blue_tables(int device, int n, int f64)
{
   std::lock_guard<std::mutex> lk(g_root_mutex);
   RootKey key{device, n, f64};
   auto it = g_blue.find(key);
   if (it != g_blue.end()) return it->second;
   int logM = 1;
   while ((1LL << logM) < 2LL * n - 1) logM++;
   const size_t M = (size_t)1 << logM;
   const long double pi = 3.14159265358979323846264338327950288L;
   std::vector<long double> cr(n), ci(n);
   for (int j = 0; j < n; j++) {
      const long long r = ((long long)j * j) % (2LL * n);
      const long double ang = pi * (long double)r / (long double)n;
      cr[j] = cosl(ang);
      ci[j] = -sinl(ang);
   }
   // b in bit-reversed order, then an in-place radix-2 decimation-in-time transform with a table of M-th roots
   std::vector<long double> br(M, 0.0L), bi(M, 0.0L), wr(M / 2), wi(M / 2);
   auto rev = [&](size_t i) { size_t r = 0; for (int t = 0; t < logM; t++) r |= ((i >> t) & 1) << (logM - 1 - t); return r; };
   for (int j = 0; j < n; j++) {
      br[rev((size_t)j)] = cr[j]; bi[rev((size_t)j)] = -ci[j];
      if (j > 0) { br[rev(M - (size_t)j)] = cr[j]; bi[rev(M - (size_t)j)] = -ci[j]; }
   }
   for (size_t k = 0; k < M / 2; k++) {
      const long double ang = -2.0L * pi * (long double)k / (long double)M;
      wr[k] = cosl(ang); wi[k] = sinl(ang);
   }
   for (size_t len = 2; len <= M; len <<= 1) {
      const size_t half = len / 2, step = M / len;
      for (size_t i0 = 0; i0 < M; i0 += len)
         for (size_t k = 0; k < half; k++) {
            const long double ur = br[i0 + k], ui = bi[i0 + k];
            const long double xr = br[i0 + k + half], xi = bi[i0 + k + half];
            const long double tr = xr * wr[k * step] - xi * wi[k * step], ti = xr * wi[k * step] + xi * wr[k * step];
            br[i0 + k] = ur + tr; bi[i0 + k] = ui + ti;
            br[i0 + k + half] = ur - tr; bi[i0 + k + half] = ui - ti;
         }
   }
   std::vector<double> hc(2 * (size_t)n), hb(2 * M);
   for (int j = 0; j < n; j++) { hc[2 * j] = (double)cr[j]; hc[2 * j + 1] = (double)ci[j]; }
   for (size_t k = 0; k < M; k++) { hb[2 * k] = (double)(br[k] / (long double)M); hb[2 * k + 1] = (double)(bi[k] / (long double)M); }
   auto upload = [&](const std::vector<double> &h) {
      void *d = nullptr;
      const size_t bytes = h.size() / 2 * (f64 ? 16 : 8);
      D2D_CHECK_CUDA(cudaMalloc(&d, bytes));
      if (f64) D2D_CHECK_CUDA(cudaMemcpy(d, h.data(), bytes, cudaMemcpyHostToDevice));
      else {
         std::vector<float> hf(h.begin(), h.end());
         D2D_CHECK_CUDA(cudaMemcpy(d, hf.data(), bytes, cudaMemcpyHostToDevice));
      }
      return d;
   };
   BlueTab t;
   t.chirp = upload(hc);
   t.bspec = upload(hb);
   t.logM = logM;
   return g_blue[key] = t;
}

template <typename T, int MODE>
cudaError_t blue_run(Ctx *ctx, const FftArgs &g, const BlueTab &t, int lines_a, long long lines, int sms)
{
   using T2 = typename Vec2<T>::type;
   const int f64 = sizeof(T) == 8;
   const size_t M = (size_t)1 << t.logM;
   const long long per_chunk = std::max<long long>(1, std::min<long long>(lines, (long long)(ctx->scratch2_bytes / (M * sizeof(T2)))));
   T2 *scr = reinterpret_cast<T2 *>(ctx->scratch2);
   const T2 *chirp = reinterpret_cast<const T2 *>(t.chirp), *bspec = reinterpret_cast<const T2 *>(t.bspec);
   auto grid_for = [&](long long items) { return (unsigned)std::max<long long>(1, std::min<long long>((items + 255) / 256, (long long)sms * 8)); };
   for (long long l0 = 0; l0 < lines; l0 += per_chunk) {
      const long long nl = std::min(per_chunk, lines - l0);
      const long long total = nl << t.logM;
      blue_in_kernel<T, MODE><<<grid_for(total), 256, 0, ctx->stream>>>(g, chirp, scr, t.logM, lines_a, l0, nl);
      if (!g.passthrough) {
         FftArgs gi{};
         gi.in.np = gi.out.np = 1;
         gi.in.e0[0] = gi.out.e0[0] = 0;
         gi.in.e0[1] = gi.out.e0[1] = (int)M;
         gi.in.ptr[0] = gi.out.ptr[0] = scr;
         gi.in.se[0] = gi.out.se[0] = 1;
         gi.in.sa[0] = gi.out.sa[0] = (long long)M;
         gi.in.sb[0] = gi.out.sb[0] = 0;
         gi.na = (int)nl;
         gi.nb = 1;
         gi.n = (int)M;
         gi.sm_limit = g.sm_limit;
         cudaError_t e = fft_dispatch(ctx, gi, f64, MODE_C2C, KIND_LINE, 0, false);
         if (e != cudaSuccess) return e;
         blue_mul_kernel<T><<<grid_for(total), 256, 0, ctx->stream>>>(scr, bspec, t.logM, total);
         gi.backward = 1;
         e = fft_dispatch(ctx, gi, f64, MODE_C2C, KIND_LINE, 0, false);
         if (e != cudaSuccess) return e;
      }
      const long long nout = (MODE == MODE_R2C) ? g.n / 2 + 1 : g.n;
      blue_out_kernel<T, MODE><<<grid_for(nl * nout), 256, 0, ctx->stream>>>(g, chirp, scr, t.logM, lines_a, l0, nl);
   }
   return cudaGetLastError();
}

cudaError_t fft_any_launch_bluestein(Ctx *ctx, const FftArgs &g, int f64, int mode)
{
   const int n = g.n;
   D2D_REQUIRE(n < (1 << 24), "transform length " + std::to_string(n) + " is not supported");
   const BlueTab &t = blue_tables(ctx->device, n, f64);
   const size_t M = (size_t)1 << t.logM, ces = f64 ? 16 : 8;
   const int lines_a = (mode == MODE_C2C) ? g.na : g.na_real; // real transforms: one complex line per real line
   const long long lines = (long long)lines_a * g.nb;
   if (lines <= 0) return cudaSuccess;
   // scratch: whole lines of M points, at most ~1 GiB per chunk of lines (grow-only)
   const size_t want = std::max<size_t>(M * ces, std::min<size_t>((size_t)lines * M * ces, (size_t)1 << 30));
   if (want > ctx->scratch2_bytes) {
      D2D_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
      if (ctx->scratch2) D2D_CHECK_CUDA(cudaFree(ctx->scratch2));
      ctx->scratch2 = nullptr;
      ctx->scratch2_bytes = 0;
      D2D_CHECK_CUDA(cudaMalloc(&ctx->scratch2, want));
      ctx->scratch2_bytes = want;
   }
   int sms = 0;
   D2D_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
   if (mode == MODE_C2C) return f64 ? blue_run<double, MODE_C2C>(ctx, g, t, lines_a, lines, sms) : blue_run<float, MODE_C2C>(ctx, g, t, lines_a, lines, sms);
   if (mode == MODE_R2C) return f64 ? blue_run<double, MODE_R2C>(ctx, g, t, lines_a, lines, sms) : blue_run<float, MODE_R2C>(ctx, g, t, lines_a, lines, sms);
   return f64 ? blue_run<double, MODE_C2R>(ctx, g, t, lines_a, lines, sms) : blue_run<float, MODE_C2R>(ctx, g, t, lines_a, lines, sms);
}

// does n = n1 n2 with both factors within the shared-memory kernel's reach?
bool any_splits(int n, int lim)
{
   for (int d = 2; (long long)d * d <= n; d++)
      if (n % d == 0 && n / d <= lim) return true; // d <= sqrt(n) <= n / d <= lim
   return false;
}
} // namespace

void fft_any_release_blue()
{
   std::lock_guard<std::mutex> lk(g_root_mutex);
   for (auto &kv : g_blue) {
      cudaFree(kv.second.chirp);
      cudaFree(kv.second.bspec);
   }
   g_blue.clear();
}

cudaError_t fft_any_launch(Ctx *ctx, const FftArgs &g, int f64, int mode)
{
   const int n = g.n;
   D2D_REQUIRE(n >= 1, "transform length must be positive");
   // a large prime factor p costs the shared-memory kernel n * p operations per line (its direct p-point DFT); beyond a few
   // hundred the five sweeps of the chirp-z path are cheaper (D2D_BLUESTEIN_MIN_PRIME moves the threshold, 0 = never)
   static const int blue_min = getenv("D2D_BLUESTEIN_MIN_PRIME") ? atoi(getenv("D2D_BLUESTEIN_MIN_PRIME")) : 600;
   if (blue_min > 0 && n > blue_min && !g.passthrough) {
      int m = n, big = 1;
      for (int f = 2; (long long)f * f <= m; f++)
         while (m % f == 0) { big = f; m /= f; }
      if (m > 1) big = m;
      if (big > blue_min) return fft_any_launch_bluestein(ctx, g, f64, mode);
   }
   if (n > fft_any_max_n(f64)) return any_splits(n, fft_any_max_n(f64)) ? fft_any_launch_split(ctx, g, f64, mode) : fft_any_launch_bluestein(ctx, g, f64, mode);
   FftArgsAny ga{};
   ga.a = g;
   ga.a.tw = roots_for(ctx->device, n, f64);
   const size_t smem = configure_any(ga, f64, mode, false);
   if (mode == MODE_C2C) return f64 ? launch_any<double, MODE_C2C>(ga, smem, ctx->stream) : launch_any<float, MODE_C2C>(ga, smem, ctx->stream);
   if (mode == MODE_R2C) return f64 ? launch_any<double, MODE_R2C>(ga, smem, ctx->stream) : launch_any<float, MODE_R2C>(ga, smem, ctx->stream);
   return f64 ? launch_any<double, MODE_C2R>(ga, smem, ctx->stream) : launch_any<float, MODE_C2R>(ga, smem, ctx->stream);
}

} // namespace d2d
