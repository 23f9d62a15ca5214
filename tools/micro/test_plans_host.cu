// Host-side check of the compiled radix plans of csrc/fft_kernel.cuh: the very PassOp / Bfly code the kernels run
// (twiddle, butterflies, scatter through the padded exchange buffer, gather, unpermute), with the T threads of a line
// emulated one after the other, against a naive long-double DFT.  Covers every plan, powers of two and 3 * 2^k / 5 * 2^k.
// Build: nvcc -std=c++17 -O1 --expt-relaxed-constexpr -gencode arch=compute_100a,code=sm_100a -o test_plans_host test_plans_host.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../2decomp-fft_b200/csrc/fft_kernel.cuh"

using namespace d2d;

template <typename T, class P, int PASS> struct Emu {
   using T2 = typename Vec2<T>::type;
   using PI = PlanInfo<P>;
   static constexpr int PADK = P::R0;
   static void run(std::vector<std::vector<T2>> &v, std::vector<T2> &lsm, const T2 *tw)
   {
      using Op = PassOp<T, P, PASS, 1, PADK, true>;
      for (int j = 0; j < P::T; j++) {
         Op::twiddle(v[j].data(), j, tw);
         Op::butterflies(v[j].data());
      }
      if constexpr (PASS + 1 < PI::npass) {
         for (int j = 0; j < P::T; j++) Op::scatter(v[j].data(), j, lsm.data());
         for (int j = 0; j < P::T; j++)
            for (int s = 0; s < P::E; s++) v[j][s] = lsm[padix<PADK>(j + P::T * s)];
         Emu<T, P, PASS + 1>::run(v, lsm, tw);
      } else {
         for (int j = 0; j < P::T; j++) {
            T2 w[P::E];
            Op::unpermute(v[j].data(), w);
            for (int s = 0; s < P::E; s++) v[j][s] = w[s];
         }
      }
   }
};

template <typename T, int N> double check()
{
   using P = Pow2Plan<N>;
   using PI = PlanInfo<P>;
   using T2 = typename Vec2<T>::type;
   static_assert(P::E * P::T == N, "plan geometry");
   static_assert(PI::radix(0) * PI::radix(1) * PI::radix(2) * PI::radix(3) == N, "radices must multiply to N");
   static_assert(P::E % PI::radix(0) == 0 && P::E % PI::radix(1) == 0 && P::E % PI::radix(2) == 0 && P::E % PI::radix(3) == 0, "every radix divides E");
   // twiddle tables as ctx.cpp builds them (full layout)
   std::vector<T2> tw;
   long long ns = 1;
   for (int p = 0; p < PI::npass; p++) {
      const int R = PI::radix(p);
      if (p >= 1)
         for (int r = 1; r < R; r++)
            for (long long q = 0; q < ns; q++) {
               const long double ang = -2.0L * 3.14159265358979323846264338327950288L * (long double)(r * q) / (long double)(ns * R);
               tw.push_back(T2{(T)cosl(ang), (T)sinl(ang)});
            }
      ns *= R;
   }
   if ((int)tw.size() != PI::tw_total) { printf("N=%d: twiddle table size %zu != %d\n", N, tw.size(), PI::tw_total); return 1e30; }
   tw.push_back(T2{0, 0});
   std::vector<double> xr(N), xi(N);
   srand(N);
   for (int k = 0; k < N; k++) { xr[k] = rand() / (double)RAND_MAX - 0.5; xi[k] = rand() / (double)RAND_MAX - 0.5; }
   std::vector<std::vector<T2>> v(P::T, std::vector<T2>(P::E));
   for (int j = 0; j < P::T; j++)
      for (int s = 0; s < P::E; s++) v[j][s] = T2{(T)xr[j + P::T * s], (T)xi[j + P::T * s]};
   std::vector<T2> lsm(padix<P::R0>(N - 1) + 2);
   Emu<T, P, 0>::run(v, lsm, tw.data());
   // reference
   const long double pi2 = 2 * 3.14159265358979323846264338327950288L;
   std::vector<long double> c(N), s(N);
   for (int k = 0; k < N; k++) { c[k] = cosl(pi2 * k / N); s[k] = sinl(pi2 * k / N); }
   double worst = 0, scale = 0;
   for (int k = 0; k < N; k++) {
      long double sr = 0, si = 0;
      for (int i = 0; i < N; i++) {
         const int m = (int)(((long long)i * k) % N);
         sr += xr[i] * c[m] + xi[i] * s[m];
         si += xi[i] * c[m] - xr[i] * s[m];
      }
      const T2 got = v[k % P::T][k / P::T];
      worst = std::max(worst, std::max(std::fabs((double)(got.x - sr)), std::fabs((double)(got.y - si))));
      scale = std::max(scale, std::max(std::fabs((double)sr), std::fabs((double)si)));
   }
   return worst / scale;
}

#define CHECK(N) { const double e64 = check<double, N>(), e32 = check<float, N>(); w64 = std::max(w64, e64); w32 = std::max(w32, e32); \
                   printf("N = %5d  fp64 %.2e  fp32 %.2e\n", N, e64, e32); }

int main()
{
   double w64 = 0, w32 = 0;
   CHECK(2) CHECK(4) CHECK(8) CHECK(16) CHECK(32) CHECK(64) CHECK(128) CHECK(256) CHECK(512) CHECK(1024) CHECK(2048) CHECK(4096)
   CHECK(6) CHECK(12) CHECK(24) CHECK(48) CHECK(96) CHECK(192) CHECK(384) CHECK(768) CHECK(1536) CHECK(3072)
   CHECK(10) CHECK(20) CHECK(40) CHECK(80) CHECK(160) CHECK(320) CHECK(640) CHECK(1280)
   printf("worst fp64 %.2e fp32 %.2e\n", w64, w32);
   return (w64 < 2e-15 && w32 < 2e-6) ? 0 : 1;
}
