// Host-side check of the arbitrary-length pass arithmetic of csrc/fft_any.cuh (the same __host__ __device__
// functions the kernel runs) against a naive long-double DFT.  Build: nvcc -std=c++17 -O1 -o test_any_host test_any_host.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../2decomp-fft_b200/csrc/fft_any.cuh"

namespace d2d { int fft_any_factorize(int n, int *radix, int maxp); }
// copy of the factorisation (fft_any.cu) so that this file builds alone
static int factorize(int n, int *radix, int maxp)
{
   int np = 0;
   auto push = [&](int r) { if (np < maxp) radix[np] = r; np++; };
   while (n % 4 == 0) { push(4); n /= 4; }
   if (n % 2 == 0) { push(2); n /= 2; }
   for (int f = 3; (long long)f * f <= n; f += 2)
      while (n % f == 0) { push(f); n /= f; }
   if (n > 1) push(n);
   return np;
}

int main()
{
   using namespace d2d;
   const int sizes[] = {1, 2, 3, 5, 6, 7, 9, 10, 11, 12, 13, 15, 17, 18, 20, 21, 22, 24, 26, 27, 34, 35, 49, 51, 60, 66, 68, 100, 102, 121, 127, 130, 210, 257, 289, 360, 1000, 1001};
   double worst = 0;
   for (int n : sizes) {
      int radix[kMaxAnyPass];
      const int np = factorize(n, radix, kMaxAnyPass);
      std::vector<double2> W(n), a(n), b(n), x(n);
      const long double pi2 = 2 * 3.14159265358979323846264338327950288L;
      for (int k = 0; k < n; k++) W[k] = double2{(double)cosl(pi2 * k / n), (double)-sinl(pi2 * k / n)};
      srand(n);
      for (int k = 0; k < n; k++) x[k] = a[k] = double2{rand() / (double)RAND_MAX - 0.5, rand() / (double)RAND_MAX - 0.5};
      double2 *src = a.data(), *dst = b.data();
      int Ns = 1;
      for (int p = 0; p < np; p++) {
         const int R = radix[p], M = n / R;
         for (int jj = 0; jj < M; jj++) {
            switch (R) {
            case 2: any_bfly_fixed<double2, 2>(src, dst, W.data(), n, Ns, jj); break;
            case 3: any_bfly_fixed<double2, 3>(src, dst, W.data(), n, Ns, jj); break;
            case 4: any_bfly_fixed<double2, 4>(src, dst, W.data(), n, Ns, jj); break;
            case 5: any_bfly_fixed<double2, 5>(src, dst, W.data(), n, Ns, jj); break;
            case 7: any_bfly_fixed<double2, 7>(src, dst, W.data(), n, Ns, jj); break;
            default: break;
            }
         }
         if (R != 2 && R != 3 && R != 4 && R != 5 && R != 7)
            for (int o = 0; o < n; o++) dst[o] = any_out_runtime<double2>(src, W.data(), n, R, Ns, o);
         std::swap(src, dst);
         Ns *= R;
      }
      double err = 0, mx = 0;
      for (int k = 0; k < n; k++) {
         long double re = 0, im = 0;
         for (int j = 0; j < n; j++) {
            const long double ang = -pi2 * (long double)((long long)j * k % n) / n;
            re += x[j].x * cosl(ang) - x[j].y * sinl(ang);
            im += x[j].x * sinl(ang) + x[j].y * cosl(ang);
         }
         err = fmax(err, fmax(fabs((double)(src[k].x - re)), fabs((double)(src[k].y - im))));
         mx = fmax(mx, fmax(fabs((double)re), fabs((double)im)));
      }
      printf("n=%5d passes=%d rel err %.2e\n", n, np, err / mx);
      worst = fmax(worst, err / mx);
   }
   printf("worst %.2e\n", worst);
   return worst < 1e-13 ? 0 : 1;
}
