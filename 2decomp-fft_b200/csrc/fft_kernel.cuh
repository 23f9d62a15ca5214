// fft_kernel.cuh -- batched 1-D FFT kernels for sm_100a (B200), the compute core of the library.
//
// Replaces the cuFFT plans/executions of the reference's GPU backend
// (src/fft_cufft.f90:73-258 plan builders, :489-671 c2c_1m_{x,y,z}, r2c_1m_{x,z}, c2r_1m_{x,z})
// AND the pack/unpack passes of the transposes (mem_split_* / mem_merge_*, src/transpose_*.f90):
// every kernel reads its lines through a piecewise-strided map and writes them through another, so
// the producing FFT writes straight into the per-destination send segments and the consuming FFT
// gathers straight from the receive segments.  No stand-alone pack/unpack sweep touches HBM.
//
// Algorithm: Stockham autosort, radix 16/8/4/2 butterflies held in registers (E elements per
// thread), inter-pass exchange through padded shared memory, per-pass twiddle tables laid out
// [r][q] (bank-conflict-free, L1-resident, read with ld.global.nc).  fp64 and fp32.
// One thread block processes TX*LY lines: TX adjacent lines of the fastest batch axis (lanes run
// along that axis first, so a strided pencil is read in TX*sizeof(complex) contiguous runs) times
// LY tiles.  TX=1 is the contiguous (X-pencil) variant.
//
// Real transforms use the two-for-one trick: two real lines (a, a+1 of the fastest batch axis) are
// transformed as one complex line and separated in the epilogue (r2c) / combined in the prologue
// (c2r).  c2r ignores Im(bin 0) and Im(bin n/2), which is exactly the effect of the reference's
// "c2c then take the real part" (src/fft_generic.f90:320-337).
//
// Direction: the butterflies implement the forward transform exp(-2 pi i jk/n)
// (DECOMP_2D_FFT_FORWARD = -1, src/decomp_2d_constants.f90:86); backward = conj . forward . conj,
// applied for free at load/store.  No normalisation in either direction, like the reference.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace d2d {

constexpr int kMaxPieces = 8;

enum FftMode { MODE_C2C = 0, MODE_R2C = 1, MODE_C2R = 2 };

// Piecewise-strided view of the complex lines of one stage.  The transform axis e in [0,n) is cut
// into `np` pieces [e0[m], e0[m+1]); element e of line (a,b) lives at
//    ptr[m] + (e - e0[m])*se[m] + a*sa[m] + b*sb[m]        (in complex elements).
// np == 1 is a plain strided pencil; np > 1 is a send/recv buffer in the reference's all-to-all
// layout -- or the peers' buffers themselves.
struct PieceMap {
   int np;
   int e0[kMaxPieces + 1];
   void *ptr[kMaxPieces];
   long long se[kMaxPieces], sa[kMaxPieces], sb[kMaxPieces];
};

struct FftArgs {
   PieceMap in, out; // complex side(s).  R2C uses only `out`, C2R only `in`.
   // real side (R2C input / C2R output): natural pencil, element e of real line ar at
   //    rptr + e*rse + ar*rsa + b*rsb   (in real elements); complex line pair index a <-> ar = 2a, 2a+1
   void *rptr;
   long long rse, rsa, rsb;
   int na, nb;         // batch extents: na = fastest batch axis (complex lines; pairs count for real modes = ceil(na_real/2))
   int na_real;        // number of real lines along a (R2C/C2R)
   int n;              // transform length (checked against the template)
   int backward;       // 1: conj-in / conj-out (isign=+1), C2C only
   int passthrough;    // 1: opt_skip_XYZ_c2c -- move data through the maps without transforming
   int debug;          // experiments only (D2D_DEBUG_SKIP): bit 0 = drop the global stores, bit 1 = drop the global loads
   const void *tw;     // twiddle tables of this n / dtype
   int sm_limit;       // host side only: SMs the launch may occupy (0 = all): exchange kernels share the device
};

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

// arithmetic shared with host-side checks of the pass arithmetic (tools/micro/test_any_host.cu)
#define D2D_HD __host__ __device__ __forceinline__

template <typename T2> D2D_HD T2 cadd(T2 a, T2 b) { return T2{a.x + b.x, a.y + b.y}; }
template <typename T2> D2D_HD T2 csub(T2 a, T2 b) { return T2{a.x - b.x, a.y - b.y}; }
// fp32: one packed FADD2 per complex add / subtract on sm_100a (the fp32 kernels are bound by instruction issue, not by HBM:
// same instruction count as fp64 for half the bytes).  -DD2D_NO_F32X2 builds the scalar version for A/B runs.
#if defined(__CUDA_ARCH__) && !defined(D2D_NO_F32X2)
template <> __device__ __forceinline__ float2 cadd<float2>(float2 a, float2 b)
{
   float2 r;
   asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
       : "=f"(r.x), "=f"(r.y)
       : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
   return r;
}
template <> __device__ __forceinline__ float2 csub<float2>(float2 a, float2 b)
{
   float2 r;
   asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
       : "=f"(r.x), "=f"(r.y)
       : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
   return r;
}
#endif
template <typename T2> D2D_HD T2 cmul(T2 a, T2 b) { return T2{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
// multiply by -i : (x,y) -> (y,-x)
template <typename T2> D2D_HD T2 mul_mi(T2 a) { return T2{a.y, -a.x}; }

template <typename T> struct Consts;
template <> struct Consts<double> {
   static constexpr double rsqrt2 = 0.70710678118654752440;
   static constexpr double c8 = 0.92387953251128675613; // cos(pi/8)
   static constexpr double s8 = 0.38268343236508977173; // sin(pi/8)
};
template <> struct Consts<float> {
   static constexpr float rsqrt2 = 0.70710678118654752440f;
   static constexpr float c8 = 0.92387953251128675613f;
   static constexpr float s8 = 0.38268343236508977173f;
};

// ---- forward butterflies, in place on v[B], v[B+S], ... ; output r ends up at v[B + out_idx(r)*S]
template <typename T, int R> struct Bfly;

template <typename T> struct Bfly<T, 1> {
   using T2 = typename Vec2<T>::type;
   template <int B, int S> static D2D_HD void run(T2 *) {}
   static constexpr int out_idx(int r) { return r; }
};

template <typename T> struct Bfly<T, 2> {
   using T2 = typename Vec2<T>::type;
   template <int B, int S> static D2D_HD void run(T2 *v)
   {
      T2 a = v[B], b = v[B + S];
      v[B] = cadd(a, b);
      v[B + S] = csub(a, b);
   }
   static constexpr int out_idx(int r) { return r; }
};

template <typename T> struct Bfly<T, 4> {
   using T2 = typename Vec2<T>::type;
   template <int B, int S> static D2D_HD void run(T2 *v)
   {
      T2 a0 = cadd(v[B], v[B + 2 * S]), a1 = csub(v[B], v[B + 2 * S]);
      T2 a2 = cadd(v[B + S], v[B + 3 * S]), a3 = mul_mi(csub(v[B + S], v[B + 3 * S]));
      v[B] = cadd(a0, a2);
      v[B + S] = cadd(a1, a3);
      v[B + 2 * S] = csub(a0, a2);
      v[B + 3 * S] = csub(a1, a3);
   }
   static constexpr int out_idx(int r) { return r; }
};

template <typename T> struct Bfly<T, 8> {
   using T2 = typename Vec2<T>::type;
   // DIF split 2 x 4: X[2m] = DFT4(x_k + x_{k+4})[m], X[2m+1] = DFT4((x_k - x_{k+4}) W8^k)[m]
   template <int B, int S> static D2D_HD void run(T2 *v)
   {
      const T h = Consts<T>::rsqrt2;
#pragma unroll
      for (int k = 0; k < 4; k++) {
         T2 a = v[B + k * S], b = v[B + (k + 4) * S];
         v[B + k * S] = cadd(a, b);
         v[B + (k + 4) * S] = csub(a, b);
      }
      { // W8^1 = (1 - i)/sqrt2 ; W8^2 = -i ; W8^3 = (-1 - i)/sqrt2
         T2 t = v[B + 5 * S];
         v[B + 5 * S] = T2{(t.x + t.y) * h, (t.y - t.x) * h};
         v[B + 6 * S] = mul_mi(v[B + 6 * S]);
         t = v[B + 7 * S];
         v[B + 7 * S] = T2{(t.y - t.x) * h, -(t.x + t.y) * h};
      }
      Bfly<T, 4>::template run<B, S>(v);
      Bfly<T, 4>::template run<B + 4 * S, S>(v);
   }
   static constexpr int out_idx(int r) { return (r & 1) * 4 + (r >> 1); }
};

template <typename T> struct Bfly<T, 16> {
   using T2 = typename Vec2<T>::type;
   // 4 x 4: i = k + 4l, r = m + 4n : X[m+4n] = sum_k W4^{kn} W16^{km} [ sum_l x_{k+4l} W4^{lm} ]
   template <int B, int S> static D2D_HD void run(T2 *v)
   {
      const T h = Consts<T>::rsqrt2, c = Consts<T>::c8, s = Consts<T>::s8;
      Bfly<T, 4>::template run<B, 4 * S>(v); // u_k[m] at v[B + (k + 4m) S]
      Bfly<T, 4>::template run<B + S, 4 * S>(v);
      Bfly<T, 4>::template run<B + 2 * S, 4 * S>(v);
      Bfly<T, 4>::template run<B + 3 * S, 4 * S>(v);
      // twiddles W16^{km}, W16^p = (cos(p pi/8), -sin(p pi/8))
      // m = 1: k=1 W^1, k=2 W^2, k=3 W^3
      v[B + 5 * S] = cmul(v[B + 5 * S], T2{c, -s});
      {
         T2 t = v[B + 6 * S];
         v[B + 6 * S] = T2{(t.x + t.y) * h, (t.y - t.x) * h};
      }
      v[B + 7 * S] = cmul(v[B + 7 * S], T2{s, -c});
      // m = 2: k=1 W^2, k=2 W^4 = -i, k=3 W^6 = (-1 - i)/sqrt2
      {
         T2 t = v[B + 9 * S];
         v[B + 9 * S] = T2{(t.x + t.y) * h, (t.y - t.x) * h};
         v[B + 10 * S] = mul_mi(v[B + 10 * S]);
         t = v[B + 11 * S];
         v[B + 11 * S] = T2{(t.y - t.x) * h, -(t.x + t.y) * h};
      }
      // m = 3: k=1 W^3 = (s, -c), k=2 W^6, k=3 W^9 = (-c, s)
      v[B + 13 * S] = cmul(v[B + 13 * S], T2{s, -c});
      {
         T2 t = v[B + 14 * S];
         v[B + 14 * S] = T2{(t.y - t.x) * h, -(t.x + t.y) * h};
      }
      v[B + 15 * S] = cmul(v[B + 15 * S], T2{-c, s});
      Bfly<T, 4>::template run<B, S>(v); // X[m+4n] at v[B + (4m+n) S]
      Bfly<T, 4>::template run<B + 4 * S, S>(v);
      Bfly<T, 4>::template run<B + 8 * S, S>(v);
      Bfly<T, 4>::template run<B + 12 * S, S>(v);
   }
   static constexpr int out_idx(int r) { return 4 * (r & 3) + (r >> 2); }
};

template <typename T> struct Bfly<T, 32> {
   using T2 = typename Vec2<T>::type;
   // DIF split 2 x 16: X[2m] = DFT16(x_k + x_{k+16})[m], X[2m+1] = DFT16((x_k - x_{k+16}) W32^k)[m]
   template <int B, int S> static D2D_HD void run(T2 *v)
   {
      // W32^k = (cos(k pi/16), -sin(k pi/16)), k = 1..15
      constexpr T c1 = (T)0.98078528040323044913, s1 = (T)0.19509032201612826785; // pi/16
      constexpr T c2 = Consts<T>::c8, s2 = Consts<T>::s8;                           // 2 pi/16
      constexpr T c3 = (T)0.83146961230254523708, s3 = (T)0.55557023301960222474; // 3 pi/16
      constexpr T h = Consts<T>::rsqrt2;
#pragma unroll
      for (int k = 0; k < 16; k++) {
         T2 a = v[B + k * S], b = v[B + (k + 16) * S];
         v[B + k * S] = cadd(a, b);
         v[B + (k + 16) * S] = csub(a, b);
      }
      v[B + 17 * S] = cmul(v[B + 17 * S], T2{c1, -s1});
      v[B + 18 * S] = cmul(v[B + 18 * S], T2{c2, -s2});
      v[B + 19 * S] = cmul(v[B + 19 * S], T2{c3, -s3});
      { T2 t = v[B + 20 * S]; v[B + 20 * S] = T2{(t.x + t.y) * h, (t.y - t.x) * h}; }
      v[B + 21 * S] = cmul(v[B + 21 * S], T2{s3, -c3});
      v[B + 22 * S] = cmul(v[B + 22 * S], T2{s2, -c2});
      v[B + 23 * S] = cmul(v[B + 23 * S], T2{s1, -c1});
      v[B + 24 * S] = mul_mi(v[B + 24 * S]);
      v[B + 25 * S] = cmul(v[B + 25 * S], T2{-s1, -c1});
      v[B + 26 * S] = cmul(v[B + 26 * S], T2{-s2, -c2});
      v[B + 27 * S] = cmul(v[B + 27 * S], T2{-s3, -c3});
      { T2 t = v[B + 28 * S]; v[B + 28 * S] = T2{(t.y - t.x) * h, -(t.x + t.y) * h}; }
      v[B + 29 * S] = cmul(v[B + 29 * S], T2{-c3, -s3});
      v[B + 30 * S] = cmul(v[B + 30 * S], T2{-c2, -s2});
      v[B + 31 * S] = cmul(v[B + 31 * S], T2{-c1, -s1});
      Bfly<T, 16>::template run<B, S>(v);
      Bfly<T, 16>::template run<B + 16 * S, S>(v);
   }
   static constexpr int out_idx(int r) { return (r & 1) * 16 + Bfly<T, 16>::out_idx(r >> 1); }
};

} // namespace d2d
#include "fft_bfly_mixed.cuh" // radices 3, 5, 6, 10, 12, 20, 24
namespace d2d {

// ---- radix plans: E elements per thread, T = n/E threads per line; every radix divides E --------
// (the name dates from the time when only powers of two had compiled kernels)
template <int N> struct Pow2Plan;
#define D2D_PLAN(N_, E_, A, B, C, D)                                                                                   \
   template <> struct Pow2Plan<N_> {                                                                                   \
      static constexpr int N = N_, E = E_, T = N_ / E_;                                                                \
      static constexpr int R0 = A, R1 = B, R2 = C, R3 = D;                                                             \
   };
D2D_PLAN(2, 2, 2, 1, 1, 1)
D2D_PLAN(4, 4, 4, 1, 1, 1)
D2D_PLAN(8, 8, 8, 1, 1, 1)
D2D_PLAN(16, 16, 16, 1, 1, 1)
D2D_PLAN(32, 8, 8, 4, 1, 1)
D2D_PLAN(64, 8, 8, 8, 1, 1)
D2D_PLAN(128, 16, 16, 8, 1, 1)
D2D_PLAN(256, 16, 16, 16, 1, 1)
// 512 and 1024: two passes = ONE shared-memory exchange per transform (32 elements per thread).  Measured on B200,
// 1024^3 fp64: c2c stages 2.9-3.0 ms (87-91 % of the measured HBM copy rate) against 3.1-3.6 ms with the three-pass
// plans 8.8.8 / 16.16.4 (make VARIANT=_r16 EXTRA=-DD2D_PLAN_R16 builds those for A/B runs): the kernels were bound by
// the shared-memory pipe (l1tex 63-74 % busy), and the second exchange was 40 % of its traffic.
#ifdef D2D_PLAN_R16
D2D_PLAN(512, 8, 8, 8, 8, 1)
#else
D2D_PLAN(512, 32, 32, 16, 1, 1)
#endif
#ifdef D2D_PLAN_R16
D2D_PLAN(1024, 16, 16, 16, 4, 1)
#else
D2D_PLAN(1024, 32, 32, 32, 1, 1)
#endif
D2D_PLAN(2048, 16, 16, 16, 8, 1)
D2D_PLAN(4096, 16, 16, 16, 16, 1)
D2D_PLAN(8192, 16, 16, 16, 16, 2)
D2D_PLAN(16384, 16, 16, 16, 16, 4)
// 3 * 2^k: 24 elements per thread (12 below 96), first pass radix 24 = 3 x 8, then radix 8 / 4 passes (radix-4 passes keep
// one twiddle per butterfly in the TMA kernels: the kernels are bound by the shared-memory pipe, not by the FP64 pipe)
D2D_PLAN(6, 6, 6, 1, 1, 1)
D2D_PLAN(12, 12, 12, 1, 1, 1)
D2D_PLAN(24, 24, 24, 1, 1, 1)
D2D_PLAN(48, 12, 12, 4, 1, 1)
D2D_PLAN(96, 24, 24, 4, 1, 1)
D2D_PLAN(192, 24, 24, 8, 1, 1)
D2D_PLAN(384, 24, 24, 4, 4, 1)
D2D_PLAN(768, 24, 24, 8, 4, 1)
D2D_PLAN(1536, 24, 24, 8, 8, 1)
D2D_PLAN(3072, 24, 24, 8, 4, 4)
// 5 * 2^k: 20 elements per thread (radices 20 = 5 x 4, 4, 2)
D2D_PLAN(10, 10, 10, 1, 1, 1)
D2D_PLAN(20, 20, 20, 1, 1, 1)
D2D_PLAN(40, 20, 20, 2, 1, 1)
D2D_PLAN(80, 20, 20, 4, 1, 1)
D2D_PLAN(160, 20, 20, 4, 2, 1)
D2D_PLAN(320, 20, 20, 4, 4, 1)
D2D_PLAN(640, 20, 20, 4, 4, 2)
D2D_PLAN(1280, 20, 20, 4, 4, 4)
#undef D2D_PLAN
// (A one-exchange fp32 plan 2048 = 64 . 32 with 64 elements per thread was measured at 2048^3 fp32 on one B200: 109.8 ms
// per pair against 109.0 ms with 16 . 16 . 8 -- three variants spill -- so single precision keeps the plans above.)

template <class P> struct PlanInfo {
   static constexpr int radix(int p) { return p == 0 ? P::R0 : p == 1 ? P::R1 : p == 2 ? P::R2 : P::R3; }
   static constexpr int npass = (P::R1 == 1) ? 1 : (P::R2 == 1) ? 2 : (P::R3 == 1) ? 3 : 4;
   static constexpr int ns(int p) { return p == 0 ? 1 : ns(p - 1) * radix(p - 1); } // product of earlier radices
   // offset (in complex elements) of the twiddle table of pass p (p >= 1): sum_{q=1}^{p-1} (R_q - 1) * Ns_q
   static constexpr int tw_off(int p) { return p <= 1 ? 0 : tw_off(p - 1) + (radix(p - 1) - 1) * ns(p - 1); }
   static constexpr int tw_total = tw_off(npass);
};

// shared-memory index of line position p: one padding element every PADK elements
template <int PADK> D2D_HD constexpr int padix(int p) { return PADK > 0 ? p + p / (PADK > 0 ? PADK : 1) : p; }

template <typename T2> __device__ __forceinline__ T2 ldg_nc(const T2 *p);
template <> __device__ __forceinline__ double2 ldg_nc<double2>(const double2 *p) { return __ldg(p); }
template <> __device__ __forceinline__ float2 ldg_nc<float2>(const float2 *p) { return __ldg(p); }

__device__ __forceinline__ long long piece_addr(const PieceMap &m, int e, long long a, long long b, int &pc)
{
   int p = 0;
   if (m.np > 1) {
#pragma unroll 1
      while (p + 1 < m.np && e >= m.e0[p + 1]) p++;
   }
   pc = p;
   return (long long)(e - m.e0[p]) * m.se[p] + a * m.sa[p] + b * m.sb[p];
}

template <typename T2> __device__ __forceinline__ T2 load_piece(const PieceMap &m, int e, long long a, long long b)
{
   int pc;
   long long off = piece_addr(m, e, a, b, pc);
   return reinterpret_cast<const T2 *>(m.ptr[pc])[off];
}
template <typename T2> __device__ __forceinline__ void store_piece(const PieceMap &m, int e, long long a, long long b, T2 val)
{
   int pc;
   long long off = piece_addr(m, e, a, b, pc);
   reinterpret_cast<T2 *>(m.ptr[pc])[off] = val;
}

// One Stockham pass PASS on the registers of one thread (all loops are compile-time).
template <typename T, class P, int PASS, int SP, int PADK, bool TWS> struct PassOp {
   using T2 = typename Vec2<T>::type;
   using PI = PlanInfo<P>;
   static constexpr int E = P::E, TPL = P::T, N = P::N;
   static constexpr int R = PI::radix(PASS), NS = PI::ns(PASS), NB = E / R; // NB butterflies per thread

   static D2D_HD void twiddle(T2 *v, int j, const T2 *__restrict__ tw)
   {
      if (PASS == 0) return;
#pragma unroll
      for (int u = 0; u < NB; u++) {
         const int q = (j + TPL * u) % NS;
#pragma unroll
         for (int r = 1; r < R; r++) {
            T2 w;
            if constexpr (TWS) w = tw[PI::tw_off(PASS) + (r - 1) * NS + q]; // table staged in shared memory
            else w = ldg_nc(tw + PI::tw_off(PASS) + (r - 1) * NS + q);
            v[u + r * NB] = cmul(v[u + r * NB], w);
         }
      }
   }
   static D2D_HD void butterflies(T2 *v)
   {
      run_b<0>(v);
   }
   template <int U> static D2D_HD void run_b(T2 *v)
   {
      if constexpr (U < NB) {
         Bfly<T, R>::template run<U, NB>(v);
         run_b<U + 1>(v);
      }
   }
   // scatter the outputs to shared memory at their Stockham positions
   static D2D_HD void scatter(const T2 *v, int j, T2 *lsm)
   {
#pragma unroll
      for (int u = 0; u < NB; u++) {
         const int jj = j + TPL * u;
         const int base = (jj / NS) * (NS * R) + (jj % NS);
#pragma unroll
         for (int r = 0; r < R; r++) lsm[padix<PADK>(base + r * NS) * SP] = v[u + Bfly<T, R>::out_idx(r) * NB];
      }
   }
   // last pass: put output slot s = u + r*NB into w[s] (compile-time permutation)
   static D2D_HD void unpermute(const T2 *v, T2 *w)
   {
#pragma unroll
      for (int u = 0; u < NB; u++)
#pragma unroll
         for (int r = 0; r < R; r++) w[u + r * NB] = v[u + Bfly<T, R>::out_idx(r) * NB];
   }
};

struct NoHook {
   __device__ __forceinline__ void operator()() const {}
};

// All passes.  SYNC0: a barrier is needed before the first scatter too (the exchange buffer was used
// as the landing zone of this tile's input).  `hook` runs right after the LAST read of the exchange
// buffer: the buffer is free from there on (the pipelined kernel starts the next tile's loads there).
template <typename T, class P, int PASS, int SP, int PADK, bool TWS, bool SYNC0, class Hook> struct RunPasses {
   using T2 = typename Vec2<T>::type;
   using PI = PlanInfo<P>;
   static __device__ __forceinline__ void run(T2 *v, int j, T2 *lsm, const T2 *__restrict__ tw, Hook &hook)
   {
      using Op = PassOp<T, P, PASS, SP, PADK, TWS>;
      Op::twiddle(v, j, tw);
      Op::butterflies(v);
      if constexpr (PASS + 1 < PI::npass) {
         if (PASS > 0 || SYNC0) __syncthreads(); // WAR: everyone has read the previous contents
         Op::scatter(v, j, lsm);
         __syncthreads();
#pragma unroll
         for (int s = 0; s < P::E; s++) v[s] = lsm[padix<PADK>(j + P::T * s) * SP];
         if constexpr (PASS + 2 == PI::npass) hook();
         RunPasses<T, P, PASS + 1, SP, PADK, TWS, SYNC0, Hook>::run(v, j, lsm, tw, hook);
      } else {
         T2 w[P::E];
         Op::unpermute(v, w);
#pragma unroll
         for (int s = 0; s < P::E; s++) v[s] = w[s];
      }
   }
};

template <typename T, class P, int TX, int LY, int PADK, bool LM = false> struct KernelGeom {
   using T2 = typename Vec2<T>::type;
   static constexpr int threads = TX * LY * P::T;
   // Two shared-memory layouts, chosen so that what is contiguous in global memory lands contiguously
   // (LDGSTS moves 32-byte sectors; splitting one across distant shared addresses halves its rate,
   // measured): LM = false, interleaved [position][tx] for tile-like inputs (TX adjacent lines are
   // contiguous in memory); LM = true, line-major [tx][position] for line-like inputs (the transform
   // axis is contiguous): line l of the block at l * line_sm, position p at padix(p).
   // line_sm == bankq/TX (mod bankq = 128 B / sizeof(T2)) makes a quarter-warp (fp64) / half-warp (fp32)
   // made of TX lines x bankq/TX consecutive positions hit distinct banks, for the tile-ordered lanes (tx fastest)
   // as well as for the line-ordered lanes used to land line-like inputs.
   static constexpr int bankq = 128 / (int)sizeof(T2);
   static constexpr int line_raw = padix<PADK>(P::N - 1) + 1;
   static constexpr int lane_step = (bankq / TX) > 1 ? (bankq / TX) : 1; // line_sm == lane_step (mod bankq)
   static constexpr int line_sm = (TX > 1 && LM) ? ((line_raw + bankq - 1 - lane_step) / bankq * bankq + lane_step) : line_raw;
   static constexpr size_t tile_bytes = (size_t)line_sm * TX * LY * sizeof(T2);
   static constexpr bool needs_smem = (PlanInfo<P>::npass > 1);
   // twiddle tables live in shared memory when two blocks per SM still fit (L1 misses on the
   // ld.global.nc path cost ~1.7x the data in L2 reads, measured)
   static constexpr size_t tw_bytes = (size_t)PlanInfo<P>::tw_total * sizeof(T2);
   static constexpr size_t kSmemPerSM = 227 * 1024;
   static constexpr bool tw_in_smem = needs_smem && (tile_bytes + tw_bytes <= kSmemPerSM) &&
                                      (kSmemPerSM / (tile_bytes + tw_bytes) == kSmemPerSM / tile_bytes || kSmemPerSM / (tile_bytes + tw_bytes) >= 2);
   static constexpr size_t smem_bytes = tile_bytes + (tw_in_smem ? tw_bytes : 0);
};

// ---- cp.async (LDGSTS) helpers: global -> shared without a register round trip -------------------
template <int BYTES> __device__ __forceinline__ void cp_async(void *smem, const void *gmem)
{
   const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
   if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem) : "memory");
   else if constexpr (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem) : "memory");
   else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;\n" ::: "memory"); }

// conditional conjugation without touching the FP64 pipe: flip the sign bit of the imaginary part
__device__ __forceinline__ double flip_sign(double y, unsigned mask_hi)
{
   return __hiloint2double(__double2hiint(y) ^ (int)mask_hi, __double2loint(y));
}
__device__ __forceinline__ float flip_sign(float y, unsigned mask_hi) { return __int_as_float(__float_as_int(y) ^ (int)mask_hi); }

// The kernel.  MODE: C2C / R2C / C2R.  PAIRVEC: real pairs (2a,2a+1) are adjacent and 2*sizeof(T)
// aligned in memory (rsa == 1, even row pitch): load/store them as one vector.
//
// Persistent blocks (grid = resident blocks) walk over the tile groups.  When the transform needs
// shared memory (more than one pass) and the input is read through plain element loads (C2C, R2C),
// the kernel is software-pipelined: the input of tile i+1 is brought in with cp.async into the very
// buffer tile i used for its exchanges, as soon as tile i has read it for the last time, and lands
// while tile i runs its last pass and its stores.
template <typename T, class P, int TX, int LY, int PADK, int MODE, bool PAIRVEC, int MINB, bool LM>
__global__ void __launch_bounds__(TX *LY *P::T, MINB) fft_kernel(const __grid_constant__ FftArgs g)
{
   using T2 = typename Vec2<T>::type;
   using G = KernelGeom<T, P, TX, LY, PADK, LM>;
   constexpr int SP = LM ? 1 : TX; // shared-memory stride between consecutive positions of a line
   constexpr int N = P::N, E = P::E, TPL = P::T;
   constexpr bool PIPE = G::needs_smem && MODE != MODE_C2R;
   extern __shared__ __align__(16) unsigned char smem_raw[];
   T2 *sm = reinterpret_cast<T2 *>(smem_raw);

   const int tid = threadIdx.x;
   const int tx = tid % TX;
   const int j = (tid / TX) % TPL;
   const int ly = tid / (TX * TPL);
   const int tiles_a = (g.na + TX - 1) / TX;
   const long long ntiles = (long long)tiles_a * g.nb;
   const long long ngroups = (ntiles + LY - 1) / LY;
   T2 *lsm = LM ? sm + (size_t)(ly * TX + tx) * G::line_sm : sm + (size_t)ly * (G::line_sm * TX) + tx;
   constexpr bool TWS = G::tw_in_smem;
   const T2 *__restrict__ tw = reinterpret_cast<const T2 *>(g.tw);
   if constexpr (TWS) {
      T2 *tws = sm + G::tile_bytes / sizeof(T2);
      for (int i = tid; i < PlanInfo<P>::tw_total; i += TX * LY * TPL) tws[i] = ldg_nc(tw + i);
      tw = tws;
      __syncthreads();
   }
   const unsigned conj_mask = g.backward ? 0x80000000u : 0u;

   long long a = 0, b = 0;
   bool valid = false, v1 = false;
   // Staging identity.  cp.async only merges ADJACENT lanes into one L2 request (measured: a lane
   // order that hops between lines costs 3.6x the requests and 2.6x the sectors), so the thread
   // that brings an element in is chosen for coalescing, not for the butterflies: lanes run along
   // the transform axis when the input is line-like (unit stride along e), along the tile axis
   // otherwise.  When the two identities differ a barrier separates landing and reading back.
   constexpr bool line_in = PIPE && TX > 1 && LM; // the host picks LM kernels for line-like inputs
   const int stx = line_in ? (tid / TPL) % TX : tx;
   const int sj = line_in ? tid % TPL : j;
   T2 *slsm = LM ? sm + (size_t)(ly * TX + stx) * G::line_sm : sm + (size_t)ly * (G::line_sm * TX) + stx;
   long long sa_ = 0;
   bool svalid = false, sv1 = false;
   auto locate = [&](long long grp) {
      const long long tile = grp * LY + ly;
      const bool in_range = tile < ntiles;
      b = tile / tiles_a;
      const long long a0 = (tile - b * tiles_a) * TX;
      a = a0 + tx;
      sa_ = a0 + stx;
      valid = in_range && (a < g.na);
      v1 = valid && (2 * a + 1 < g.na_real);
      svalid = in_range && (sa_ < g.na);
      sv1 = svalid && (2 * sa_ + 1 < g.na_real);
   };

   // start the asynchronous loads of the current tile into its shared-memory slots
   auto stage = [&]() {
      if constexpr (PIPE) {
         if (svalid && !(g.debug & 2)) {
            if constexpr (MODE == MODE_C2C) {
               if (g.in.np == 1) {
                  const T2 *p = reinterpret_cast<const T2 *>(g.in.ptr[0]) + (long long)sj * g.in.se[0] + sa_ * g.in.sa[0] + b * g.in.sb[0];
                  const long long step = (long long)TPL * g.in.se[0];
#pragma unroll
                  for (int s = 0; s < E; s++) cp_async<sizeof(T2)>(&slsm[padix<PADK>(sj + TPL * s) * SP], p + s * step);
               } else {
#pragma unroll
                  for (int s = 0; s < E; s++) {
                     int pc;
                     const long long off = piece_addr(g.in, sj + TPL * s, sa_, b, pc);
                     cp_async<sizeof(T2)>(&slsm[padix<PADK>(sj + TPL * s) * SP], reinterpret_cast<const T2 *>(g.in.ptr[pc]) + off);
                  }
               }
            } else { // R2C: two real lines -> (re, im) of one complex line
               const T *p = reinterpret_cast<const T *>(g.rptr) + (long long)sj * g.rse + (2 * sa_) * g.rsa + b * g.rsb;
               const long long step = (long long)TPL * g.rse;
#pragma unroll
               for (int s = 0; s < E; s++) {
                  T2 *dst = &slsm[padix<PADK>(sj + TPL * s) * SP];
                  if constexpr (PAIRVEC) cp_async<sizeof(T2)>(dst, p + s * step);
                  else {
                     cp_async<sizeof(T)>(&dst->x, p + s * step);
                     if (sv1) cp_async<sizeof(T)>(&dst->y, p + s * step + g.rsa);
                  }
               }
            }
         }
         cp_async_commit();
      }
   };

   long long grp = blockIdx.x;
   if (grp < ngroups) {
      locate(grp);
      stage();
   }
   for (; grp < ngroups; grp += gridDim.x) {
      T2 v[E];
      const long long a_cur = a, b_cur = b;
      const bool valid_cur = valid, v1_cur = v1;

      // ------------------------------------------------------------------ load
      if constexpr (PIPE) {
         cp_async_wait_all();
         if (line_in) __syncthreads(); // staged by other threads (same identity otherwise: no barrier needed)
#pragma unroll
         for (int s = 0; s < E; s++) {
            T2 x = lsm[padix<PADK>(j + TPL * s) * SP];
            if constexpr (MODE == MODE_C2C) x.y = flip_sign(x.y, conj_mask);
            else if (!v1_cur) x.y = 0;
            v[s] = x;
         }
      } else if constexpr (MODE == MODE_C2C) {
#pragma unroll
         for (int s = 0; s < E; s++) {
            T2 x = T2{0, 0};
            if (valid_cur) x = load_piece<T2>(g.in, j + TPL * s, a_cur, b_cur);
            x.y = flip_sign(x.y, conj_mask);
            v[s] = x;
         }
      } else if constexpr (MODE == MODE_R2C) {
         const T *__restrict__ rp = reinterpret_cast<const T *>(g.rptr);
#pragma unroll
         for (int s = 0; s < E; s++) {
            const long long off = (long long)(j + TPL * s) * g.rse + (2 * a_cur) * g.rsa + b_cur * g.rsb;
            T2 x = T2{0, 0};
            if constexpr (PAIRVEC) {
               if (valid_cur) x = *reinterpret_cast<const T2 *>(rp + off);
            } else {
               if (valid_cur) x.x = rp[off];
               if (v1_cur) x.y = rp[off + g.rsa];
            }
            v[s] = x;
         }
      } else { // C2R: build conj(Z), Z[k] = A[k] + i B[k], Z[n-k] = conj(A[k]) + i conj(B[k])
#pragma unroll
         for (int s = 0; s <= E / 2; s++) {
            const int k = j + TPL * s;
            if (k <= N / 2 && (s < E / 2 || j == 0)) {
               T2 A = T2{0, 0}, B = T2{0, 0};
               if (valid_cur) A = load_piece<T2>(g.in, k, 2 * a_cur, b_cur);
               if (v1_cur) B = load_piece<T2>(g.in, k, 2 * a_cur + 1, b_cur);
               if (k == 0 || 2 * k == N) { A.y = 0; B.y = 0; }
               if constexpr (G::needs_smem) {
                  lsm[padix<PADK>(k) * SP] = T2{A.x - B.y, -(A.y + B.x)};
                  if (k > 0 && 2 * k < N) lsm[padix<PADK>(N - k) * SP] = T2{A.x + B.y, A.y - B.x};
               } else { // single-thread line (T == 1, j == 0, k == s): slots are the positions
                  v[s] = T2{A.x - B.y, -(A.y + B.x)};
                  if (s > 0 && 2 * s < N) v[(N - s) % E] = T2{A.x + B.y, A.y - B.x};
               }
            }
         }
         if constexpr (G::needs_smem) {
            __syncthreads();
#pragma unroll
            for (int s = 0; s < E; s++) v[s] = lsm[padix<PADK>(j + TPL * s) * SP];
            __syncthreads();
         }
      }

      // next tile of this block
      const long long nxt = grp + gridDim.x;
      auto prefetch_next = [&]() {
         if constexpr (PIPE) {
            __syncthreads(); // the whole block is done with the exchange buffer
            if (nxt < ngroups) {
               locate(nxt);
               stage();
            }
         }
      };

      // ------------------------------------------------------------------ transform
      if (!g.passthrough) {
         if constexpr (PIPE && MODE == MODE_C2C) {
            RunPasses<T, P, 0, SP, PADK, TWS, true, decltype(prefetch_next)>::run(v, j, lsm, tw, prefetch_next);
         } else {
            NoHook nh;
            RunPasses<T, P, 0, SP, PADK, TWS, PIPE, NoHook>::run(v, j, lsm, tw, nh);
         }
      } else if constexpr (PIPE && MODE == MODE_C2C) {
         prefetch_next();
      }

      // ------------------------------------------------------------------ store
      if constexpr (MODE == MODE_C2C) {
         if (valid_cur && !(g.debug & 1)) {
            if (g.out.np == 1) {
               T2 *p = reinterpret_cast<T2 *>(g.out.ptr[0]) + (long long)j * g.out.se[0] + a_cur * g.out.sa[0] + b_cur * g.out.sb[0];
               const long long step = (long long)TPL * g.out.se[0];
#pragma unroll
               for (int s = 0; s < E; s++) {
                  T2 x = v[s];
                  x.y = flip_sign(x.y, conj_mask);
                  p[s * step] = x;
               }
            } else {
#pragma unroll
               for (int s = 0; s < E; s++) {
                  T2 x = v[s];
                  x.y = flip_sign(x.y, conj_mask);
                  store_piece<T2>(g.out, j + TPL * s, a_cur, b_cur, x);
               }
            }
         }
      } else if constexpr (MODE == MODE_C2R) {
         T *__restrict__ rp = reinterpret_cast<T *>(g.rptr);
#pragma unroll
         for (int s = 0; s < E; s++) {
            const long long off = (long long)(j + TPL * s) * g.rse + (2 * a_cur) * g.rsa + b_cur * g.rsb;
            if constexpr (PAIRVEC) {
               if (valid_cur) *reinterpret_cast<T2 *>(rp + off) = T2{v[s].x, -v[s].y};
            } else {
               if (valid_cur) rp[off] = v[s].x;
               if (v1_cur) rp[off + g.rsa] = -v[s].y;
            }
         }
         if constexpr (G::needs_smem) __syncthreads(); // next iteration's prologue overwrites the buffer
      } else { // R2C: separate the two spectra.  A[k] = (Z[k] + conj Z[n-k])/2, B[k] = (Z[k] - conj Z[n-k])/(2i)
         if constexpr (G::needs_smem) {
            __syncthreads();
#pragma unroll
            for (int s = 0; s < E; s++) lsm[padix<PADK>(j + TPL * s) * SP] = v[s];
            __syncthreads();
         }
         T2 zn[E / 2 + 1];
#pragma unroll
         for (int s = 0; s <= E / 2; s++) {
            const int k = j + TPL * s;
            zn[s] = T2{0, 0};
            if (k <= N / 2 && (s < E / 2 || j == 0)) {
               if constexpr (G::needs_smem) zn[s] = lsm[padix<PADK>((N - k) % N) * SP];
               else zn[s] = v[(N - s) % N]; // T == 1: k == s
            }
         }
         prefetch_next(); // (no-op unless pipelined) the exchange buffer is free again
#pragma unroll
         for (int s = 0; s <= E / 2; s++) {
            const int k = j + TPL * s;
            if (k <= N / 2 && (s < E / 2 || j == 0)) {
               const T2 zk = v[s]; // slot s holds Z[j + T s] = Z[k]
               const T hf = (T)0.5;
               T2 A = T2{(zk.x + zn[s].x) * hf, (zk.y - zn[s].y) * hf};
               T2 B = T2{(zk.y + zn[s].y) * hf, (zn[s].x - zk.x) * hf};
               if (valid_cur) store_piece<T2>(g.out, k, 2 * a_cur, b_cur, A);
               if (v1_cur) store_piece<T2>(g.out, k, 2 * a_cur + 1, b_cur, B);
            }
         }
      }
      if constexpr (!PIPE) {
         if (nxt < ngroups) locate(nxt);
      }
   }
}

} // namespace d2d
