// fft_any.cuh -- batched 1-D FFT kernel for ARBITRARY transform lengths (sm_100a).
//
// The power-of-two lengths have compiled register-resident kernels (fft_kernel.cuh / fft_kernel_v2.cuh);
// every other length -- the reference examples default to 17 x 13 x 11 grids scaled by small factors
// (examples/fft_physical_x/fft_c2c_x.f90:18,39-42) and fft_multiple_grids uses ny+2 / nz+16 -- runs here.
// Same contract as the other kernels: c2c_1m_{x,y,z} / r2c_1m_{x,z} / c2r_1m_{x,z} of
// src/fft_cufft.f90:489-671 (CPU twin: src/fft_generic.f90:112-384 over src/glassman.f90) fused with
// mem_split_* / mem_merge_* of src/transpose_*.f90 through the piece maps.
//
// Algorithm: mixed-radix Stockham autosort between two shared-memory buffers, one pass per factor of n.
// Factors 2, 3, 4, 5, 7 have register butterflies; any other prime factor R is a direct R-point DFT
// (one output element per work item, R terms) -- the same O(n * sum of factors) work as the reference's
// Glassman routine.  All twiddles and DFT coefficients are n-th roots of unity, so ONE table
// W[k] = exp(-2 pi i k / n) (computed in extended precision on the host) staged in shared memory serves
// every pass.  A block owns `lines` adjacent lines; lanes run along whichever axis is unit-stride in
// global memory on the load side and on the store side (independently), so strided pencils are read and
// written in lines * sizeof(complex) contiguous runs.
//
// Real transforms: two-for-one like the other kernels (two real lines = one complex line), valid for odd n
// as well: bins 0..floor(n/2) are produced / consumed, Im(bin 0) and (even n) Im(bin n/2) are ignored by c2r.
#pragma once
#include "fft_kernel.cuh"

#if defined(__CUDACC__)
#define D2D_HD __host__ __device__ __forceinline__
#else
#define D2D_HD inline
#endif

namespace d2d {

constexpr int kMaxAnyPass = 32;
constexpr int kAnyThreads = 256;

struct FftArgsAny {
   FftArgs a;                 // a.tw = W[k] = exp(-2 pi i k / n), k in [0, n)
   int npass;
   int radix[kMaxAnyPass];
   int lines;                 // complex lines per block
   int pitch;                 // shared-memory pitch of a line, in complex elements (odd)
   int in_fast_a, out_fast_a; // 1: adjacent lines (axis a) are contiguous in global memory on that side
};

template <typename T2> D2D_HD T2 any_cadd(T2 a, T2 b) { return T2{a.x + b.x, a.y + b.y}; }
template <typename T2> D2D_HD T2 any_csub(T2 a, T2 b) { return T2{a.x - b.x, a.y - b.y}; }
template <typename T2> D2D_HD T2 any_cmul(T2 a, T2 b) { return T2{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
template <typename T2> D2D_HD T2 any_mul_mi(T2 a) { return T2{a.y, -a.x}; }

// Stockham pass (R, Ns) of an n-point transform, n = M R:
//   butterfly jj in [0, M): inputs src[jj + r M], r in [0, R), twiddled by exp(-2 pi i q r / (Ns R)), q = jj % Ns;
//   outputs dst[(jj / Ns) Ns R + q + k Ns], k in [0, R).
// Register butterfly for a compile-time radix.
template <typename T2, int R> D2D_HD void any_bfly_fixed(const T2 *src, T2 *dst, const T2 *W, int n, int Ns, int jj)
{
   const int M = n / R;
   const int q = jj % Ns;
   const int base = (jj - q) * R + q;
   const int s = n / (Ns * R); // exp(-2 pi i q r / (Ns R)) = W[q r s], and q r s < n
   T2 x[R];
   x[0] = src[jj];
#pragma unroll
   for (int r = 1; r < R; r++) x[r] = any_cmul(src[jj + r * M], W[q * s * r]);
   if constexpr (R == 2) {
      dst[base] = any_cadd(x[0], x[1]);
      dst[base + Ns] = any_csub(x[0], x[1]);
   } else if constexpr (R == 4) {
      const T2 a0 = any_cadd(x[0], x[2]), a1 = any_csub(x[0], x[2]);
      const T2 a2 = any_cadd(x[1], x[3]), a3 = any_mul_mi(any_csub(x[1], x[3]));
      dst[base] = any_cadd(a0, a2);
      dst[base + Ns] = any_cadd(a1, a3);
      dst[base + 2 * Ns] = any_csub(a0, a2);
      dst[base + 3 * Ns] = any_csub(a1, a3);
   } else {
      // direct R-point DFT in registers: X[k] = sum_r x[r] W[(r k mod R) M]
#pragma unroll
      for (int k = 0; k < R; k++) {
         T2 acc = x[0];
#pragma unroll
         for (int r = 1; r < R; r++) acc = any_cadd(acc, any_cmul(x[r], W[((r * k) % R) * M]));
         dst[base + k * Ns] = acc;
      }
   }
}

// One OUTPUT element o in [0, n) of pass (R, Ns) for a run-time radix R (any prime):
//   o = g Ns R + k Ns + q  ->  sum_r src[g Ns + q + r M] W[r (q s + k M) mod n]
template <typename T2> D2D_HD T2 any_out_runtime(const T2 *src, const T2 *W, int n, int R, int Ns, int o)
{
   const int M = n / R;
   const int q = o % Ns;
   const int k = (o / Ns) % R;
   const int gq = (o / (Ns * R)) * Ns + q; // butterfly index jj
   const int s = n / (Ns * R);
   const int step = q * s + k * M; // < n
   T2 acc = src[gq];
   int idx = 0;
   for (int r = 1; r < R; r++) {
      idx += step;
      if (idx >= n) idx -= n;
      acc = any_cadd(acc, any_cmul(src[gq + r * M], W[idx]));
   }
   return acc;
}

#if defined(__CUDACC__)

template <typename T2, int R>
__device__ __forceinline__ void any_pass_fixed(const T2 *src, T2 *dst, const T2 *W, int n, int Ns, int lines, int pitch)
{
   const int M = n / R;
   const int items = lines * M;
   for (int i = threadIdx.x; i < items; i += kAnyThreads) {
      const int l = i / M, jj = i - l * M;
      any_bfly_fixed<T2, R>(src + l * pitch, dst + l * pitch, W, n, Ns, jj);
   }
}

template <typename T, int MODE> __global__ void __launch_bounds__(kAnyThreads) fft_any_kernel(const __grid_constant__ FftArgsAny ga)
{
   using T2 = typename Vec2<T>::type;
   const FftArgs &g = ga.a;
   const int n = g.n, L = ga.lines, pitch = ga.pitch;
   const int nh = n / 2 + 1;
   extern __shared__ __align__(16) unsigned char any_smem[];
   T2 *W = reinterpret_cast<T2 *>(any_smem);
   T2 *buf0 = W + n;
   T2 *buf1 = buf0 + (size_t)L * pitch;
   {
      const T2 *__restrict__ wg = reinterpret_cast<const T2 *>(g.tw);
      for (int i = threadIdx.x; i < n; i += kAnyThreads) W[i] = ldg_nc(wg + i);
   }
   const long long total = (long long)g.na * g.nb; // complex lines (pairs of real lines for r2c / c2r)
   const long long ngroups = (total + L - 1) / L;
   const bool bw = g.backward != 0;

   for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
      __syncthreads(); // W is staged; the previous group's stores have read the buffers
      const long long id0 = grp * L;
      // ------------------------------------------------------------------ load -> buf0
      if constexpr (MODE == MODE_C2C) {
         const int items = L * n;
         for (int i = threadIdx.x; i < items; i += kAnyThreads) {
            int l, e;
            if (ga.in_fast_a) { e = i / L; l = i - e * L; }
            else { l = i / n; e = i - l * n; }
            const long long id = id0 + l;
            T2 x = T2{0, 0};
            if (id < total) {
               const long long b = id / g.na, a = id - b * g.na;
               x = load_piece<T2>(g.in, e, a, b);
               if (bw) x.y = -x.y;
            }
            buf0[l * pitch + e] = x;
         }
      } else if constexpr (MODE == MODE_R2C) {
         const T *__restrict__ rp = reinterpret_cast<const T *>(g.rptr);
         const int items = L * n;
         for (int i = threadIdx.x; i < items; i += kAnyThreads) {
            int l, e;
            if (ga.in_fast_a) { e = i / L; l = i - e * L; }
            else { l = i / n; e = i - l * n; }
            const long long id = id0 + l;
            T2 x = T2{0, 0};
            if (id < total) {
               const long long b = id / g.na, a = id - b * g.na;
               const long long off = (long long)e * g.rse + (2 * a) * g.rsa + b * g.rsb;
               x.x = rp[off];
               if (2 * a + 1 < g.na_real) x.y = rp[off + g.rsa];
            }
            buf0[l * pitch + e] = x;
         }
      } else { // C2R: Z[k] = A[k] + i B[k], Z[n-k] = conj(A[k]) + i conj(B[k]); the forward passes get conj(Z)
         const int items = L * nh;
         for (int i = threadIdx.x; i < items; i += kAnyThreads) {
            int l, k;
            if (ga.in_fast_a) { k = i / L; l = i - k * L; }
            else { l = i / nh; k = i - l * nh; }
            const long long id = id0 + l;
            T2 A = T2{0, 0}, B = T2{0, 0};
            if (id < total) {
               const long long b = id / g.na, a = id - b * g.na;
               A = load_piece<T2>(g.in, k, 2 * a, b);
               if (2 * a + 1 < g.na_real) B = load_piece<T2>(g.in, k, 2 * a + 1, b);
            }
            const bool selfconj = (k == 0) || (2 * k == n);
            if (selfconj) { A.y = 0; B.y = 0; }
            buf0[l * pitch + k] = T2{A.x - B.y, -(A.y + B.x)};
            if (!selfconj) buf0[l * pitch + n - k] = T2{A.x + B.y, A.y - B.x};
         }
      }
      __syncthreads();

      // ------------------------------------------------------------------ passes (ping-pong)
      T2 *src = buf0, *dst = buf1;
      if (!g.passthrough) {
         int Ns = 1;
         for (int p = 0; p < ga.npass; p++) {
            const int R = ga.radix[p];
            switch (R) {
            case 2: any_pass_fixed<T2, 2>(src, dst, W, n, Ns, L, pitch); break;
            case 3: any_pass_fixed<T2, 3>(src, dst, W, n, Ns, L, pitch); break;
            case 4: any_pass_fixed<T2, 4>(src, dst, W, n, Ns, L, pitch); break;
            case 5: any_pass_fixed<T2, 5>(src, dst, W, n, Ns, L, pitch); break;
            case 7: any_pass_fixed<T2, 7>(src, dst, W, n, Ns, L, pitch); break;
            default: {
               const int items = L * n;
               for (int i = threadIdx.x; i < items; i += kAnyThreads) {
                  const int l = i / n, o = i - l * n;
                  dst[l * pitch + o] = any_out_runtime<T2>(src + l * pitch, W, n, R, Ns, o);
               }
            }
            }
            __syncthreads();
            T2 *t = src; src = dst; dst = t;
            Ns *= R;
         }
      }

      // ------------------------------------------------------------------ store from src
      if (g.debug & 1) continue;
      if constexpr (MODE == MODE_C2C) {
         const int items = L * n;
         for (int i = threadIdx.x; i < items; i += kAnyThreads) {
            int l, e;
            if (ga.out_fast_a) { e = i / L; l = i - e * L; }
            else { l = i / n; e = i - l * n; }
            const long long id = id0 + l;
            if (id < total) {
               const long long b = id / g.na, a = id - b * g.na;
               T2 x = src[l * pitch + e];
               if (bw) x.y = -x.y;
               store_piece<T2>(g.out, e, a, b, x);
            }
         }
      } else if constexpr (MODE == MODE_C2R) {
         T *__restrict__ rp = reinterpret_cast<T *>(g.rptr);
         const int items = L * n;
         for (int i = threadIdx.x; i < items; i += kAnyThreads) {
            int l, e;
            if (ga.out_fast_a) { e = i / L; l = i - e * L; }
            else { l = i / n; e = i - l * n; }
            const long long id = id0 + l;
            if (id < total) {
               const long long b = id / g.na, a = id - b * g.na;
               const long long off = (long long)e * g.rse + (2 * a) * g.rsa + b * g.rsb;
               const T2 x = src[l * pitch + e];
               rp[off] = x.x;
               if (2 * a + 1 < g.na_real) rp[off + g.rsa] = -x.y;
            }
         }
      } else { // R2C: A[k] = (Z[k] + conj Z[n-k]) / 2, B[k] = (Z[k] - conj Z[n-k]) / (2i), k in [0, n/2]
         const int items = L * nh;
         for (int i = threadIdx.x; i < items; i += kAnyThreads) {
            int l, k;
            if (ga.out_fast_a) { k = i / L; l = i - k * L; }
            else { l = i / nh; k = i - l * nh; }
            const long long id = id0 + l;
            if (id < total) {
               const long long b = id / g.na, a = id - b * g.na;
               const T2 zk = src[l * pitch + k];
               const T2 zn = src[l * pitch + (k == 0 ? 0 : n - k)];
               const T hf = (T)0.5;
               store_piece<T2>(g.out, k, 2 * a, b, T2{(zk.x + zn.x) * hf, (zk.y - zn.y) * hf});
               if (2 * a + 1 < g.na_real) store_piece<T2>(g.out, k, 2 * a + 1, b, T2{(zk.y + zn.y) * hf, (zn.x - zk.x) * hf});
            }
         }
      }
   }
}

#endif // __CUDACC__

} // namespace d2d
