// fft_any.cuh -- batched 1-D FFT kernel for ARBITRARY transform lengths (sm_100a).
//
// The power-of-two lengths have compiled register-resident kernels (fft_kernel.cuh / fft_kernel_v2.cuh);
// every other length -- the reference examples default to 17 x 13 x 11 grids scaled by small factors
// (examples/fft_physical_x/fft_c2c_x.f90:18,39-42) and fft_multiple_grids uses ny+2 / nz+16 -- runs here.
// Same contract as the other kernels: c2c_1m_{x,y,z} / r2c_1m_{x,z} / c2r_1m_{x,z} of
// src/fft_cufft.f90:489-671 (CPU twin: src/fft_generic.f90:112-384 over src/glassman.f90) fused with
// mem_split_* / mem_merge_* of src/transpose_*.f90 through the piece maps.
//
// Algorithm: mixed-radix Stockham autosort between two shared-memory buffers, one pass per RADIX, where the factors of n
// are fused into as few register-resident radices as possible (any_factorize below): every pass is a round trip through
// shared memory, which is what bounds this kernel.
//   * radices 16, 8, 4, 2 and 6, 10, 12, 15, 20, 24, 30: the register butterflies of the compiled kernels
//     (fft_kernel.cuh, fft_bfly_mixed.cuh), e.g. 510 = 17 . 30 is two passes, 1000 = 25 . 20 . 2 three;
//   * odd radices 3..31: register DFT using the conjugate symmetry of the coefficients (inputs r and R-r are
//     combined first, outputs k and R-k come out of the same two real-weighted sums: half the multiplies);
//   * any larger prime R: direct R-point DFT from shared memory with the same pairing, one pair of outputs
//     (k, R-k) per work item -- O(n * sum of factors) work like the reference's Glassman routine.  These passes run
//     FIRST (no twiddles when nothing precedes them); a second large prime gets its twiddles applied in a separate
//     in-place sweep.  (Lengths whose largest prime factor exceeds a few hundred take the chirp-z path of fft_any.cu.)
//   * odd radices run before even ones: the scatter of a first pass with an even radix R lands R elements apart (bank
//     conflicts of degree gcd(R, 8)), an odd one is conflict-free and leaves an odd sub-transform length for the rest.
// Two builds of the kernel: radices <= 16 in 256 threads and 128 registers; radices up to 31 ("big") in 128 threads and
// 255 registers -- two resident blocks per SM either way (one loads / stores while the other computes).
// All twiddles and DFT coefficients are n-th roots of unity, so ONE table W[k] = exp(-2 pi i k / n)
// (computed in extended precision on the host) staged in shared memory serves every pass.
// A block owns 2^lines_log2 adjacent lines; lanes run along whichever axis is unit-stride in global memory
// on the load side and on the store side (independently), so strided pencils are read and written in
// lines * sizeof(complex) contiguous runs.  No integer division in the inner loops (power-of-two line
// counts, warp-per-line walks, float-reciprocal quotients for the pass indices).
//
// Real transforms: two-for-one like the other kernels (two real lines = one complex line), valid for odd n
// as well: bins 0..floor(n/2) are produced / consumed, Im(bin 0) and (even n) Im(bin n/2) are ignored by c2r.
#pragma once
#include "fft_kernel.cuh"

namespace d2d {

constexpr int kMaxAnyPass = 32;
constexpr int kAnyThreads = 256;     // "small" build; the "big" build runs kAnyThreadsBig
constexpr int kAnyThreadsBig = 128;
constexpr int kAnyMaxLines = 256;    // lines per block (power of two, at most one thread per line)
constexpr int kAnyMaxFixedSmall = 16; // largest register radix of the small build
constexpr int kAnyMaxFixed = 31;      // ... of the big build

// radices with a register butterfly: X(R) for every R of the small build / the additional ones of the big build
#define D2D_ANY_RADICES_SMALL(X) X(2) X(3) X(4) X(5) X(6) X(7) X(8) X(9) X(10) X(11) X(12) X(13) X(15) X(16)
#define D2D_ANY_RADICES_BIG(X) X(17) X(19) X(20) X(23) X(24) X(25) X(27) X(29) X(30) X(31)

inline bool any_radix_is_fixed(int R, bool big)
{
#define D2D_X(r) if (R == r) return true;
   D2D_ANY_RADICES_SMALL(D2D_X)
   if (big) { D2D_ANY_RADICES_BIG(D2D_X) }
#undef D2D_X
   return false;
}

// Pass radices of an n-point transform, in pass order.  Primes above the largest register radix first (run-time radix; the
// first of them needs no twiddles), then the odd register radices in decreasing order, then the even ones in decreasing
// order.  The factors 2, 3, 5 are fused into the radices 2 3 4 5 6 8 9 10 12 15 16 (and 20 24 25 27 30, big build) by an
// exhaustive search for the fewest passes; among those the plan with the largest odd radix, then the largest smallest
// radix (balanced passes) wins.  Returns the number of passes (radix[] holds at most maxp of them); *needs_big: a radix
// above kAnyMaxFixedSmall occurs.
struct AnyFuse {
   int best[kMaxAnyPass], nbest = 0, cur[kMaxAnyPass];
   long long best_cost = -1;
   bool big = false;
   void search(int c2, int c3, int c5, int depth, int last)
   {
      if (c2 == 0 && c3 == 0 && c5 == 0) {
         int odd = 0, mn = 1 << 30;
         for (int i = 0; i < depth; i++) {
            if (cur[i] % 2 && cur[i] > odd) odd = cur[i];
            if (cur[i] < mn) mn = cur[i];
         }
         const long long cost = (long long)depth * 10000 - odd * 100 - (depth ? mn : 0);
         if (best_cost < 0 || cost < best_cost) {
            best_cost = cost;
            nbest = depth;
            for (int i = 0; i < depth; i++) best[i] = cur[i];
         }
         return;
      }
      if (depth >= kMaxAnyPass || (best_cost >= 0 && (long long)(depth + 1) * 10000 - 3100 > best_cost)) return;
      static const int cand[16][4] = {{30, 1, 1, 1}, {27, 0, 3, 0}, {25, 0, 0, 2}, {24, 3, 1, 0}, {20, 2, 0, 1}, {16, 4, 0, 0}, {15, 0, 1, 1}, {12, 2, 1, 0},
                                      {10, 1, 0, 1}, {9, 0, 2, 0},  {8, 3, 0, 0},  {6, 1, 1, 0},  {5, 0, 0, 1},  {4, 2, 0, 0},  {3, 0, 1, 0},  {2, 1, 0, 0}};
      for (int i = 0; i < 16; i++) {
         const int r = cand[i][0];
         if (r > last || (!big && r > kAnyMaxFixedSmall)) continue; // non-increasing radices: every multiset once
         if (cand[i][1] > c2 || cand[i][2] > c3 || cand[i][3] > c5) continue;
         cur[depth] = r;
         search(c2 - cand[i][1], c3 - cand[i][2], c5 - cand[i][3], depth + 1, r);
      }
   }
};

inline int any_factorize(int n, int *radix, int maxp, bool allow_big, bool *needs_big)
{
   int np = 0;
   auto push = [&](int r) {
      if (np < maxp) radix[np] = r;
      np++;
   };
   int m = n, c2 = 0, c3 = 0, c5 = 0;
   while (m % 2 == 0) { m /= 2; c2++; }
   while (m % 3 == 0) { m /= 3; c3++; }
   while (m % 5 == 0) { m /= 5; c5++; }
   int fixed[64], nfixed = 0; // register radices, sorted below
   const int top = allow_big ? kAnyMaxFixed : 13;
   for (int f = 7; (long long)f * f <= m; f += 2)
      while (m % f == 0) {
         if (f <= top) fixed[nfixed++] = f;
         else push(f);
         m /= f;
      }
   if (m > 1) {
      if (m <= top) fixed[nfixed++] = m;
      else push(m);
   }
   AnyFuse fz;
   fz.big = allow_big;
   fz.search(c2, c3, c5, 0, 1 << 30);
   for (int i = 0; i < fz.nbest && nfixed < 64; i++) fixed[nfixed++] = fz.best[i];
   // odd radices first, each group in decreasing order
   for (int i = 1; i < nfixed; i++)
      for (int j = i; j > 0; j--) {
         const int a = fixed[j], b = fixed[j - 1];
         const bool before = (a % 2 != b % 2) ? (a % 2 == 1) : (a > b);
         if (!before) break;
         fixed[j] = b;
         fixed[j - 1] = a;
      }
   for (int i = 0; i < nfixed; i++) push(fixed[i]);
   if (needs_big) {
      *needs_big = false;
      for (int i = 0; i < np && i < maxp; i++)
         if (radix[i] > kAnyMaxFixedSmall && radix[i] <= kAnyMaxFixed) *needs_big = true;
   }
   return np;
}

struct FftArgsAny {
   FftArgs a;                 // a.tw = W[k] = exp(-2 pi i k / n), k in [0, n)
   int npass;
   int radix[kMaxAnyPass];
   int big;                   // 1: the build with radices up to 31 in 128 threads runs it
   int nbuf;                  // line buffers in shared memory: 2, or 3 when the next group's input is prefetched (async_in)
   int async_in;              // 1: the input of group i+1 lands with cp.async in the third buffer while group i runs its passes
   int lines_log2;            // complex lines per block = 1 << lines_log2
   int pitch;                 // shared-memory pitch of a line, in complex elements (odd)
   int in_fast_a, out_fast_a; // 1: adjacent lines (axis a) are contiguous in global memory on that side
   // ---- one step of a two-kernel transform of a length that does not fit in shared memory (n_full = n1 n2, fft_any.cu) ----
   // split > 1 (C2C kernels only): the batch axis a is virtual, a' = a * split + sub; the element index on either side is
   // e * mul + sub * add; a result element k is multiplied by exp(-2 pi i sub k / tw_n) when tw_n > 0 (table big_tw).
   int split, in_mul, in_add, out_mul, out_add, tw_n, n_full;
   const void *big_tw;
   // conversions at the two ends of a real transform done as a complex one (src/fft_generic.f90:236-244, 320-337):
   int real_in;  // read the REAL array (rptr / rse / rsa / rsb), imaginary part 0
   int herm_in;  // the input holds bins 0 .. n_full/2: element e > n_full/2 is conj(bin n_full - e)
   int half_out; // store bins 0 .. n_full/2 only
   int real_out; // store the real part into the REAL array
};

// floor(x / d) for 0 <= x < 2^22, d >= 1, rd = 1.0f / d: (x + 0.5) / d is at least 0.5 / d away from an integer, far
// more than the rounding error of the float product
D2D_HD int any_div(int x, int d, float rd)
{
#if defined(__CUDA_ARCH__)
   return __float2int_rd(((float)x + 0.5f) * rd);
#else
   (void)rd;
   return x / d;
#endif
}

// Stockham pass (R, Ns) of an n-point transform, n = M R:
//   butterfly jj in [0, M): inputs src[jj + r M], r in [0, R), twiddled by exp(-2 pi i q r / (Ns R)), q = jj % Ns;
//   outputs dst[(jj / Ns) Ns R + q + k Ns], k in [0, R).
// Register butterfly for a compile-time radix.
template <typename T, int R> D2D_HD void any_bfly_fixed(const typename Vec2<T>::type *src, typename Vec2<T>::type *dst,
                                                          const typename Vec2<T>::type *W, int n, int Ns, float rNs, int jj)
{
   using T2 = typename Vec2<T>::type;
   const int M = n / R;
   const int q = jj - any_div(jj, Ns, rNs) * Ns;
   const int base = (jj - q) * R + q;
   const int s = n / (Ns * R); // exp(-2 pi i q r / (Ns R)) = W[q r s], and q r s < n
   T2 x[R];
   x[0] = src[jj];
   if (Ns == 1) {
#pragma unroll
      for (int r = 1; r < R; r++) x[r] = src[jj + r * M];
   } else {
      const int qs = q * s;
#pragma unroll
      for (int r = 1; r < R; r++) x[r] = cmul(src[jj + r * M], W[qs * r]);
   }
   if constexpr (R == 2 || R == 4 || R == 8 || R == 16 || R == 6 || R == 10 || R == 12 || R == 15 || R == 20 || R == 24 || R == 30) {
      Bfly<T, R>::template run<0, 1>(x);
#pragma unroll
      for (int k = 0; k < R; k++) dst[base + k * Ns] = x[Bfly<T, R>::out_idx(k)];
   } else {
      // odd R: with sr = x[r] + x[R-r], dr = x[r] - x[R-r], w^m = (c_m, -s_m):
      //   X[k], X[R-k] = (x0 + sum_r c_{rk} sr)  -/+  i (sum_r s_{rk} dr),  r = 1..(R-1)/2
      static_assert(R % 2 == 1, "odd radix expected");
      constexpr int H = (R - 1) / 2;
      T wc[H + 1], ws[H + 1]; // c_m, s_m for m = 0..H (c_{R-m} = c_m, s_{R-m} = -s_m)
      wc[0] = (T)1; // m = r k mod R is 0 only for composite R (9, 25, 27)
      ws[0] = (T)0;
#pragma unroll
      for (int m = 1; m <= H; m++) {
         const T2 w = W[m * M];
         wc[m] = w.x;
         ws[m] = -w.y;
      }
      T2 sum = x[0];
#pragma unroll
      for (int r = 1; r <= H; r++) {
         const T2 a = x[r], b = x[R - r];
         x[r] = cadd(a, b);
         x[R - r] = csub(a, b);
         sum = cadd(sum, x[r]);
      }
      dst[base] = sum;
#pragma unroll
      for (int k = 1; k <= H; k++) {
         T2 p = x[0], qq = T2{0, 0};
#pragma unroll
         for (int r = 1; r <= H; r++) {
            const int m = (r * k) % R;
            const T c = m <= H ? wc[m] : wc[R - m];
            const T sn = m <= H ? ws[m] : -ws[R - m];
            p.x += c * x[r].x;
            p.y += c * x[r].y;
            qq.x += sn * x[R - r].x;
            qq.y += sn * x[R - r].y;
         }
         dst[base + k * Ns] = T2{p.x + qq.y, p.y - qq.x};       // P - i Q
         dst[base + (R - k) * Ns] = T2{p.x - qq.y, p.y + qq.x}; // P + i Q
      }
   }
}

// Large prime radix R (run time): work item (butterfly jj, kp in [0, (R-1)/2]) produces X[kp] and X[R-kp] of that
// butterfly from inputs that already carry their twiddles.
template <typename T> D2D_HD void any_pair_runtime(const typename Vec2<T>::type *src, typename Vec2<T>::type *dst,
                                                    const typename Vec2<T>::type *W, int n, int R, int Ns, float rNs, int jj, int kp)
{
   using T2 = typename Vec2<T>::type;
   const int M = n / R, H = (R - 1) / 2;
   const int q = jj - any_div(jj, Ns, rNs) * Ns;
   const int base = (jj - q) * R + q;
   T2 p = src[jj], qq = T2{0, 0};
   const int step = kp * M; // W[r kp M mod n]: coefficient of input r for output kp
   int idx = 0;
   for (int r = 1; r <= H; r++) {
      idx += step;
      if (idx >= n) idx -= n;
      const T2 a = src[jj + r * M], b = src[jj + (R - r) * M];
      const T2 w = W[idx]; // (c, -s)
      p.x += w.x * (a.x + b.x);
      p.y += w.x * (a.y + b.y);
      qq.x -= w.y * (a.x - b.x);
      qq.y -= w.y * (a.y - b.y);
   }
   if (kp == 0) {
      dst[base] = p; // all coefficients are 1: p = sum of the inputs, qq = 0
   } else {
      dst[base + kp * Ns] = T2{p.x + qq.y, p.y - qq.x};
      dst[base + (R - kp) * Ns] = T2{p.x - qq.y, p.y + qq.x};
   }
}

#if defined(__CUDACC__)

template <typename T, int R>
__device__ __forceinline__ void any_pass_fixed(const typename Vec2<T>::type *src, typename Vec2<T>::type *dst, const typename Vec2<T>::type *W,
                                               int n, int Ns, float rNs, int pt, int TL)
{
   const int M = n / R;
   for (int jj = pt; jj < M; jj += TL) any_bfly_fixed<T, R>(src, dst, W, n, Ns, rNs, jj);
}

template <typename T, int MODE, bool BIG>
__global__ void __launch_bounds__(BIG ? kAnyThreadsBig : kAnyThreads, 2) fft_any_kernel(const __grid_constant__ FftArgsAny ga)
{
   constexpr int kThreads = BIG ? kAnyThreadsBig : kAnyThreads;
   constexpr int kLogThreads = BIG ? 7 : 8;
   static_assert(kThreads == (1 << kLogThreads), "thread indexing assumes a power of two");
   using T2 = typename Vec2<T>::type;
   const FftArgs &g = ga.a;
   const int n = g.n, LL = ga.lines_log2, L = 1 << LL, pitch = ga.pitch;
   const int nh = n / 2 + 1;
   extern __shared__ __align__(16) unsigned char any_smem[];
   T2 *W = reinterpret_cast<T2 *>(any_smem);
   T2 *const bufs[3] = {W + n, W + n + (size_t)L * pitch, W + n + 2 * (size_t)L * pitch}; // [2] exists when ga.nbuf == 3
   // (a, b) of the block's lines; a = -1: beyond the batch; sub: split steps, the sub-index of the virtual batch index.
   // Two sets: the prefetch of group i+1 fills the other one while group i still needs its own for the stores.
   __shared__ int line_a_[2][kAnyMaxLines], line_b_[2][kAnyMaxLines], line_sub_[2][kAnyMaxLines];
   const int split = (MODE == MODE_C2C && ga.split > 1) ? ga.split : 1;
   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   constexpr int NW = kThreads / 32;
   // pass identity: TL threads per line
   const int TL = kThreads >> LL;
   const int pl = tid >> (kLogThreads - LL), pt = tid & (TL - 1);
   {
      const T2 *__restrict__ wg = reinterpret_cast<const T2 *>(g.tw);
      for (int i = tid; i < n; i += kThreads) W[i] = ldg_nc(wg + i);
   }
   const long long total = (long long)g.na * g.nb; // complex lines (pairs of real lines for r2c / c2r)
   const long long ngroups = (total + L - 1) >> LL;
   const bool bw = g.backward != 0;
   const bool async = (MODE != MODE_C2R) && ga.async_in != 0 && ga.nbuf == 3 && split == 1;

   auto fill_table = [&](long long grp, int par) {
      if (tid < L) {
         const long long id = (grp << LL) + tid;
         int a = -1, b = 0, sub = 0;
         if (id < total) {
            b = (int)(id / g.na);
            a = (int)(id - (long long)b * g.na);
            if (split > 1) {
               sub = a % split;
               a /= split;
            }
         }
         line_a_[par][tid] = a;
         line_b_[par][tid] = b;
         line_sub_[par][tid] = sub;
      }
   };
   // fast_a: lanes run over the lines first (l = i mod L); otherwise one warp walks along a line
   auto for_each = [&](int len, bool fast_a, auto &&fn) {
      if (fast_a) {
         const int items = len << LL;
         for (int i = tid; i < items; i += kThreads) fn(i & (L - 1), i >> LL);
      } else {
         for (int l = warp; l < L; l += NW)
            for (int e = lane; e < len; e += 32) fn(l, e);
      }
   };
   // cp.async (LDGSTS) of one group's input straight into a line buffer: C2C elements as they are (the conjugation of a
   // backward transform happens in shared memory afterwards), R2C the two real lines of a pair into the two halves
   auto issue_async = [&](T2 *dstbuf, int par) {
      const int *la = line_a_[par], *lb = line_b_[par];
      if constexpr (MODE == MODE_C2C) {
         for_each(n, ga.in_fast_a != 0, [&](int l, int e) {
            const int a = la[l];
            T2 *d = &dstbuf[l * pitch + e];
            if (a >= 0) {
               int pc;
               const long long off = piece_addr(g.in, e, a, lb[l], pc);
               cp_async<sizeof(T2)>(d, reinterpret_cast<const T2 *>(g.in.ptr[pc]) + off);
            } else {
               *d = T2{0, 0};
            }
         });
      } else if constexpr (MODE == MODE_R2C) {
         const T *__restrict__ rp = reinterpret_cast<const T *>(g.rptr);
         for_each(n, ga.in_fast_a != 0, [&](int l, int e) {
            const int a = la[l];
            T2 *d = &dstbuf[l * pitch + e];
            if (a >= 0) {
               const long long off = (long long)e * g.rse + (2LL * a) * g.rsa + (long long)lb[l] * g.rsb;
               cp_async<sizeof(T)>(&d->x, rp + off);
               if (2 * a + 1 < g.na_real) cp_async<sizeof(T)>(&d->y, rp + off + g.rsa);
               else d->y = 0;
            } else {
               *d = T2{0, 0};
            }
         });
      }
      cp_async_commit();
   };

   long long grp = blockIdx.x;
   int cur = 0, par = 0; // bufs[cur] holds / receives the input of the current group; par: its set of line tables
   if (async && grp < ngroups) {
      fill_table(grp, 0);
      __syncthreads();
      issue_async(bufs[0], 0);
   }
   for (; grp < ngroups; grp += gridDim.x, par ^= 1) {
      T2 *const buf0 = bufs[cur], *const buf1 = bufs[cur + 1 < ga.nbuf ? cur + 1 : cur + 1 - ga.nbuf];
      T2 *const my0 = buf0 + pl * pitch, *const my1 = buf1 + pl * pitch;
      const int *line_a = line_a_[par], *line_b = line_b_[par], *line_sub = line_sub_[par];
      if (async) {
         cp_async_wait_all();
         __syncthreads(); // this group's input has landed (and W); the previous group's stores have read their buffer
         const long long nxt = grp + gridDim.x;
         if (nxt < ngroups) {
            fill_table(nxt, par ^ 1);
            __syncthreads();
            issue_async(bufs[(cur + 2) % 3], par ^ 1);
         }
         // (a backward c2c needs no conjugation here: its input lands as it is and the stores read the result of the
         // forward passes in reversed order, X_b[k] = X_f[(n - k) mod n])
      } else {
      __syncthreads(); // W is staged; the previous group's stores have read the buffers and the line table
      fill_table(grp, par);
      __syncthreads();
      // ------------------------------------------------------------------ load -> buf0
      // Global loads are batched: a thread issues U independent loads before it touches shared memory.  One load per loop
      // trip (the first version of this kernel) left ~8 KB in flight per SM and bounded the whole kernel by memory latency.
      auto for_each_ld = [&](int len, bool fast_a, auto &&ld, auto &&st) {
         constexpr int U = ((MODE == MODE_C2R) ? 4 : 8) * (BIG ? 2 : 1) / (sizeof(T) == 8 ? 2 : 1); // ~32 KB in flight per SM
         using Item = decltype(ld(0, 0));
         if (fast_a) {
            const int items = len << LL;
            for (int i0 = tid; i0 < items; i0 += U * kThreads) {
               Item v[U];
#pragma unroll
               for (int u = 0; u < U; u++) {
                  const int i = i0 + u * kThreads;
                  if (i < items) v[u] = ld(i & (L - 1), i >> LL);
               }
#pragma unroll
               for (int u = 0; u < U; u++) {
                  const int i = i0 + u * kThreads;
                  if (i < items) st(i & (L - 1), i >> LL, v[u]);
               }
            }
         } else {
            for (int l = warp; l < L; l += NW)
               for (int e0 = lane; e0 < len; e0 += 32 * U) {
                  Item v[U];
#pragma unroll
                  for (int u = 0; u < U; u++)
                     if (e0 + 32 * u < len) v[u] = ld(l, e0 + 32 * u);
#pragma unroll
                  for (int u = 0; u < U; u++)
                     if (e0 + 32 * u < len) st(l, e0 + 32 * u, v[u]);
               }
         }
      };
      struct Pair2 {
         T2 A, B;
      };
      if constexpr (MODE == MODE_C2C) {
         auto put = [&](int l, int e, T2 x) { buf0[l * pitch + e] = x; };
         if (split > 1) { // a step of a two-kernel transform
            for_each_ld(n, ga.in_fast_a != 0, [&](int l, int e) {
               const int a = line_a[l];
               T2 x = T2{0, 0};
               if (a >= 0) {
                  const int ea = e * ga.in_mul + line_sub[l] * ga.in_add;
                  if (ga.real_in) {
                     x.x = reinterpret_cast<const T *>(g.rptr)[(long long)ea * g.rse + (long long)a * g.rsa + (long long)line_b[l] * g.rsb];
                  } else if (ga.herm_in && 2 * ea > ga.n_full) {
                     x = load_piece<T2>(g.in, ga.n_full - ea, a, line_b[l]);
                     x.y = -x.y;
                  } else {
                     x = load_piece<T2>(g.in, ea, a, line_b[l]);
                     if (ga.herm_in && (ea == 0 || 2 * ea == ga.n_full)) x.y = 0; // like the other c2r kernels (and cuFFT)
                  }
                  if (bw) x.y = -x.y;
               }
               return x;
            }, put);
         } else {
            for_each_ld(n, ga.in_fast_a != 0, [&](int l, int e) {
               const int a = line_a[l];
               T2 x = T2{0, 0};
               if (a >= 0) {
                  x = load_piece<T2>(g.in, e, a, line_b[l]);
                  if (bw) x.y = -x.y;
               }
               return x;
            }, put);
         }
      } else if constexpr (MODE == MODE_R2C) {
         const T *__restrict__ rp = reinterpret_cast<const T *>(g.rptr);
         for_each_ld(n, ga.in_fast_a != 0, [&](int l, int e) {
            const int a = line_a[l];
            T2 x = T2{0, 0};
            if (a >= 0) {
               const long long off = (long long)e * g.rse + (2LL * a) * g.rsa + (long long)line_b[l] * g.rsb;
               x.x = rp[off];
               if (2 * a + 1 < g.na_real) x.y = rp[off + g.rsa];
            }
            return x;
         }, [&](int l, int e, T2 x) { buf0[l * pitch + e] = x; });
      } else { // C2R: Z[k] = A[k] + i B[k], Z[n-k] = conj(A[k]) + i conj(B[k]); the forward passes get conj(Z)
         for_each_ld(nh, ga.in_fast_a != 0, [&](int l, int k) {
            const int a = line_a[l];
            Pair2 v{T2{0, 0}, T2{0, 0}};
            if (a >= 0) {
               v.A = load_piece<T2>(g.in, k, 2LL * a, line_b[l]);
               if (2 * a + 1 < g.na_real) v.B = load_piece<T2>(g.in, k, 2LL * a + 1, line_b[l]);
            }
            return v;
         }, [&](int l, int k, Pair2 v) {
            T2 A = v.A, B = v.B;
            const bool selfconj = (k == 0) || (2 * k == n);
            if (selfconj) { A.y = 0; B.y = 0; }
            buf0[l * pitch + k] = T2{A.x - B.y, -(A.y + B.x)};
            if (!selfconj) buf0[l * pitch + n - k] = T2{A.x + B.y, A.y - B.x};
         });
      }
      __syncthreads();
      } // synchronous load

      // ------------------------------------------------------------------ passes (ping-pong), line pl, thread pt of TL
      T2 *src = my0, *dst = my1;
      if (!g.passthrough) {
         int Ns = 1;
         for (int p = 0; p < ga.npass; p++) {
            const int R = ga.radix[p];
            const float rNs = 1.0f / (float)Ns;
            bool done = true;
            switch (R) {
#define D2D_X(r) case r: any_pass_fixed<T, r>(src, dst, W, n, Ns, rNs, pt, TL); break;
               D2D_ANY_RADICES_SMALL(D2D_X)
            default: done = false;
            }
            if constexpr (BIG) {
               if (!done) {
                  done = true;
                  switch (R) {
                     D2D_ANY_RADICES_BIG(D2D_X)
                  default: done = false;
                  }
               }
            }
#undef D2D_X
            if (!done) {
               const int M = n / R, H = (R - 1) / 2;
               if (Ns > 1) { // twiddles in place: src[jj + r M] *= W[q r s]
                  const int s = n / (Ns * R);
                  for (int i = pt; i < n; i += TL) {
                     const int r = any_div(i, M, 1.0f / (float)M), jj = i - r * M;
                     const int q = jj - any_div(jj, Ns, rNs) * Ns;
                     if (r > 0 && q > 0) src[i] = cmul(src[i], W[q * s * r]);
                  }
                  __syncthreads();
               }
               const int items = M * (H + 1);
               const float rM = 1.0f / (float)M;
               for (int i = pt; i < items; i += TL) {
                  const int kp = any_div(i, M, rM), jj = i - kp * M;
                  any_pair_runtime<T>(src, dst, W, n, R, Ns, rNs, jj, kp);
               }
            }
            __syncthreads();
            T2 *t = src; src = dst; dst = t;
            Ns *= R;
         }
      }
      T2 *const res = src - pl * pitch; // block view of the buffer holding the result

      // ------------------------------------------------------------------ store
      if (async) cur = (cur + 2) % 3; // the prefetched buffer becomes the input of the next group
      if (g.debug & 1) continue;
      if constexpr (MODE == MODE_C2C) {
         if (split > 1) {
            const T2 *__restrict__ bt = reinterpret_cast<const T2 *>(ga.big_tw);
            for_each(n, ga.out_fast_a != 0, [&](int l, int e) {
               const int a = line_a[l];
               if (a >= 0) {
                  T2 x = res[l * pitch + e];
                  const int sub = line_sub[l];
                  if (ga.tw_n > 0) x = cmul(x, ldg_nc(bt + (long long)sub * e)); // sub * e < tw_n
                  if (bw) x.y = -x.y;
                  const int eo = e * ga.out_mul + sub * ga.out_add;
                  if (ga.real_out) {
                     reinterpret_cast<T *>(g.rptr)[(long long)eo * g.rse + (long long)a * g.rsa + (long long)line_b[l] * g.rsb] = x.x;
                  } else if (!ga.half_out || 2 * eo <= ga.n_full) {
                     store_piece<T2>(g.out, eo, a, line_b[l], x);
                  }
               }
            });
         } else {
            const bool rev = bw && async && !g.passthrough; // prefetched input was not conjugated: index reversal instead of conj . forward . conj
            for_each(n, ga.out_fast_a != 0, [&](int l, int e) {
               const int a = line_a[l];
               if (a >= 0) {
                  T2 x = res[l * pitch + ((rev && e) ? n - e : e)];
                  if (bw && !async) x.y = -x.y; // the synchronous loads conjugated the input
                  store_piece<T2>(g.out, e, a, line_b[l], x);
               }
            });
         }
      } else if constexpr (MODE == MODE_C2R) {
         T *__restrict__ rp = reinterpret_cast<T *>(g.rptr);
         for_each(n, ga.out_fast_a != 0, [&](int l, int e) {
            const int a = line_a[l];
            if (a >= 0) {
               const long long off = (long long)e * g.rse + (2LL * a) * g.rsa + (long long)line_b[l] * g.rsb;
               const T2 x = res[l * pitch + e];
               rp[off] = x.x;
               if (2 * a + 1 < g.na_real) rp[off + g.rsa] = -x.y;
            }
         });
      } else { // R2C: A[k] = (Z[k] + conj Z[n-k]) / 2, B[k] = (Z[k] - conj Z[n-k]) / (2i), k in [0, n/2]
         for_each(nh, ga.out_fast_a != 0, [&](int l, int k) {
            const int a = line_a[l];
            if (a >= 0) {
               const T2 zk = res[l * pitch + k];
               const T2 zn = res[l * pitch + (k == 0 ? 0 : n - k)];
               const T hf = (T)0.5;
               store_piece<T2>(g.out, k, 2LL * a, line_b[l], T2{(zk.x + zn.x) * hf, (zk.y - zn.y) * hf});
               if (2 * a + 1 < g.na_real) store_piece<T2>(g.out, k, 2LL * a + 1, line_b[l], T2{(zk.y + zn.y) * hf, (zn.x - zk.x) * hf});
            }
         });
      }
   }
}

#endif // __CUDACC__

} // namespace d2d
