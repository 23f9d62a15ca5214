// Host-side check of the arbitrary-length pass arithmetic of csrc/fft_any.cuh (the same __host__ __device__
// functions the kernel runs, driven by the same factorisation) against a naive long-double DFT.
// Build: nvcc -std=c++17 -O1 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -o test_any_host test_any_host.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../2decomp-fft_b200/csrc/fft_any.cuh"

// (the factorisation any_factorize lives in fft_any.cuh and is the one fft_any.cu calls: `big` selects the radix set of the
// 128-thread build, radices up to 31 and the fused 20 / 24 / 30, or of the 256-thread build, radices up to 16)

template <typename T> double run(int n, bool big)
{
   using namespace d2d;
   using T2 = typename Vec2<T>::type;
   int radix[kMaxAnyPass];
   const int np = any_factorize(n, radix, kMaxAnyPass, big, nullptr);
   long long prod = 1;
   for (int p = 0; p < np; p++) prod *= radix[p];
   if (prod != n) { printf("n=%d: radices multiply to %lld\n", n, prod); return 1e30; }
   std::vector<T2> W(n), a(n), b(n);
   std::vector<double> xr(n), xi(n);
   const long double pi2 = 2 * 3.14159265358979323846264338327950288L;
   for (int k = 0; k < n; k++) W[k] = T2{(T)cosl(pi2 * k / n), (T)-sinl(pi2 * k / n)};
   srand(n);
   for (int k = 0; k < n; k++) {
      a[k] = T2{(T)(rand() / (double)RAND_MAX - 0.5), (T)(rand() / (double)RAND_MAX - 0.5)};
      xr[k] = a[k].x; xi[k] = a[k].y;
   }
   T2 *src = a.data(), *dst = b.data();
   int Ns = 1;
   for (int p = 0; p < np; p++) {
      const int R = radix[p], M = n / R;
      const float rNs = 1.0f / Ns;
      bool fixed = true;
      for (int jj = 0; jj < M && fixed; jj++) {
         switch (R) {
#define D2D_X(r) case r: any_bfly_fixed<T, r>(src, dst, W.data(), n, Ns, rNs, jj); break;
            D2D_ANY_RADICES_SMALL(D2D_X)
         default:
            fixed = false;
         }
         if (!fixed && big) {
            fixed = true;
            switch (R) {
               D2D_ANY_RADICES_BIG(D2D_X)
            default:
               fixed = false;
            }
         }
#undef D2D_X
      }
      if (!fixed) { // the kernel's run-time radix path: twiddle sweep, then one item per (butterfly, output pair)
         const int H = (R - 1) / 2;
         if (Ns > 1) {
            const int s = n / (Ns * R);
            for (int i = 0; i < n; i++) {
               const int r = i / M, jj = i - r * M, q = jj % Ns;
               if (r > 0 && q > 0) src[i] = cmul(src[i], W[q * s * r]);
            }
         }
         for (int i = 0; i < M * (H + 1); i++) any_pair_runtime<T>(src, dst, W.data(), n, R, Ns, rNs, i % M, i / M);
      }
      std::swap(src, dst);
      Ns *= R;
   }
   double err = 0, mx = 0;
   for (int k = 0; k < n; k++) {
      long double re = 0, im = 0;
      for (int j = 0; j < n; j++) {
         const long double ang = -pi2 * (long double)((long long)j * k % n) / n;
         re += xr[j] * cosl(ang) - xi[j] * sinl(ang);
         im += xr[j] * sinl(ang) + xi[j] * cosl(ang);
      }
      err = fmax(err, fmax(fabs((double)(src[k].x - re)), fabs((double)(src[k].y - im))));
      mx = fmax(mx, fmax(fabs((double)re), fabs((double)im)));
   }
   printf("n=%5d %s passes=%d [", n, big ? "big  " : "small", np);
   for (int p = 0; p < np; p++) printf("%d%s", radix[p], p + 1 < np ? "." : "");
   printf("] rel err %.2e\n", err / mx);
   return err / mx;
}

int main()
{
   const int sizes[] = {1, 2, 3, 5, 6, 7, 9, 10, 11, 12, 13, 15, 17, 18, 20, 21, 22, 24, 26, 27, 34, 35, 48, 49, 51, 60, 66, 68, 81, 96, 100, 102, 121, 127, 130, 144, 210, 257, 289, 323, 360, 375, 510, 544, 600, 625, 675, 768, 899, 961, 1000, 1001, 1331, 1536, 2187, 3000, 3125};
   double worst = 0, worst32 = 0;
   for (int big = 0; big < 2; big++) {
      for (int n : sizes) worst = fmax(worst, run<double>(n, big != 0));
      for (int n : sizes) worst32 = fmax(worst32, run<float>(n, big != 0));
   }
   printf("worst fp64 %.2e, fp32 %.2e\n", worst, worst32);
   return (worst < 1e-13 && worst32 < 2e-5) ? 0 : 1;
}
