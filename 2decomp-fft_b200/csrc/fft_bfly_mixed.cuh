// fft_bfly_mixed.cuh -- register butterflies for the radices 3, 5 and their products (6, 10, 12, 15, 20, 24, 30).
//
// Included by fft_kernel.cuh right after the power-of-two butterflies.  With them the compiled, register-resident
// kernels (fft_kernel.cuh cp.async kernels, fft_kernel_v2.cuh TMA kernels) also cover the lengths 3 * 2^k and 5 * 2^k
// (Pow2Plan specialisations in fft_kernel.cuh), which the reference reaches through cuFFT's mixed-radix plans
// (src/fft_cufft.f90:73-258) and its generic backend through SPCFFT's factor loop (src/glassman.f90:29-67).
// Same conventions as the other butterflies: forward transform exp(-2 pi i jk / R), in place on v[B], v[B+S], ...;
// output r ends up at v[B + out_idx(r) * S].
#pragma once

namespace d2d {

// cos(2 pi m / 120), m in [0, 30]: every root of unity the radices below need (120 = lcm(24, 20)), exact at the axes
constexpr double root120_cos(int m)
{
   constexpr double q[31] = {
      1.0, 0.998629534754573873784, 0.994521895368273336923, 0.98768834059513772619,
      0.978147600733805637929, 0.96592582628906828675, 0.951056516295153572116, 0.93358042649720174899,
      0.913545457642600895502, 0.89100652418836786236, 0.866025403784438646764, 0.838670567945424029638,
      0.809016994374947424102, 0.77714596145697087998, 0.743144825477394235015, 0.707106781186547524401,
      0.669130606358858213826, 0.629320391049837452706, 0.587785252292473129169, 0.544639035015027082224,
      0.5, 0.45399049973954679156, 0.406736643075800207754, 0.358367949545300273484,
      0.309016994374947424102, 0.258819045102520762349, 0.207911690817759337102, 0.15643446504023086901,
      0.1045284632676534714, 0.0523359562429438327221, 0.0};
   m = ((m % 120) + 120) % 120;
   return m <= 30 ? q[m] : m <= 60 ? -q[60 - m] : m <= 90 ? -q[m - 60] : q[120 - m];
}
constexpr double root120_sin(int m) { return root120_cos(m - 30); }

// x * W_R^e, W_R = exp(-2 pi i / R), e and R compile-time: multiples of a quarter turn cost no multiplication
template <typename T, int R, int e> D2D_HD typename Vec2<T>::type mul_root(typename Vec2<T>::type x)
{
   using T2 = typename Vec2<T>::type;
   static_assert(120 % R == 0, "root table covers the divisors of 120");
   constexpr int m = ((e % R) + R) % R;
   if constexpr (m == 0) return x;
   else if constexpr (4 * m == R) return T2{x.y, -x.x};      // -i
   else if constexpr (2 * m == R) return T2{-x.x, -x.y};     // -1
   else if constexpr (4 * m == 3 * R) return T2{-x.y, x.x};  // +i
   else {
      constexpr T c = (T)root120_cos(m * (120 / R)), s = (T)root120_sin(m * (120 / R)); // W = (c, -s)
      return T2{x.x * c + x.y * s, x.y * c - x.x * s};
   }
}

template <typename T> struct Bfly<T, 3> {
   using T2 = typename Vec2<T>::type;
   template <int B, int S> static D2D_HD void run(T2 *v)
   {
      constexpr T h = (T)0.866025403784438646764; // sin(pi/3)
      const T2 x0 = v[B], s = cadd(v[B + S], v[B + 2 * S]), d = csub(v[B + S], v[B + 2 * S]);
      const T2 t = T2{x0.x - (T)0.5 * s.x, x0.y - (T)0.5 * s.y};
      const T2 u = T2{h * d.y, -h * d.x}; // -i sin(pi/3) (x1 - x2)
      v[B] = cadd(x0, s);
      v[B + S] = cadd(t, u);
      v[B + 2 * S] = csub(t, u);
   }
   static constexpr int out_idx(int r) { return r; }
};

template <typename T> struct Bfly<T, 5> {
   using T2 = typename Vec2<T>::type;
   template <int B, int S> static D2D_HD void run(T2 *v)
   {
      constexpr T c1 = (T)0.309016994374947424102, c2 = (T)-0.809016994374947424102; // cos(2 pi/5), cos(4 pi/5)
      constexpr T s1 = (T)0.951056516295153572116, s2 = (T)0.587785252292473129169;  // sin(2 pi/5), sin(4 pi/5)
      const T2 x0 = v[B];
      const T2 p1 = cadd(v[B + S], v[B + 4 * S]), d1 = csub(v[B + S], v[B + 4 * S]);
      const T2 p2 = cadd(v[B + 2 * S], v[B + 3 * S]), d2 = csub(v[B + 2 * S], v[B + 3 * S]);
      const T2 a1 = T2{x0.x + c1 * p1.x + c2 * p2.x, x0.y + c1 * p1.y + c2 * p2.y};
      const T2 a2 = T2{x0.x + c2 * p1.x + c1 * p2.x, x0.y + c2 * p1.y + c1 * p2.y};
      const T2 b1 = T2{s1 * d1.x + s2 * d2.x, s1 * d1.y + s2 * d2.y};
      const T2 b2 = T2{s2 * d1.x - s1 * d2.x, s2 * d1.y - s1 * d2.y};
      v[B] = cadd(x0, cadd(p1, p2));
      v[B + S] = T2{a1.x + b1.y, a1.y - b1.x};     // a1 - i b1
      v[B + 4 * S] = T2{a1.x - b1.y, a1.y + b1.x}; // a1 + i b1
      v[B + 2 * S] = T2{a2.x + b2.y, a2.y - b2.x};
      v[B + 3 * S] = T2{a2.x - b2.y, a2.y + b2.x};
   }
   static constexpr int out_idx(int r) { return r; }
};

// R = A * Bn by Cooley-Tukey: input i = k + Bn l, output r = m + A n:
//    X[m + A n] = sum_k W_Bn^{kn} ( W_R^{km} sum_l x[k + Bn l] W_A^{lm} )
// step 1: Bn transforms of length A over stride Bn (u_k[m] lands at slot k + Bn oA(m)); step 2: the twiddles W_R^{km};
// step 3: A transforms of length Bn over the contiguous slots Bn oA(m) + k.  Output r sits at slot Bn oA(m) + oB(n).
template <typename T, int A, int Bn> struct BflyCT {
   using T2 = typename Vec2<T>::type;
   static constexpr int R = A * Bn;
   template <int B, int S, int K> static D2D_HD void step1(T2 *v)
   {
      if constexpr (K < Bn) {
         Bfly<T, A>::template run<B + K * S, Bn * S>(v);
         step1<B, S, K + 1>(v);
      }
   }
   template <int B, int S, int K, int M> static D2D_HD void step2(T2 *v)
   {
      if constexpr (K < Bn) {
         if constexpr (M < A) {
            constexpr int pos = B + (K + Bn * Bfly<T, A>::out_idx(M)) * S;
            v[pos] = mul_root<T, R, K * M>(v[pos]);
            step2<B, S, K, M + 1>(v);
         } else {
            step2<B, S, K + 1, 1>(v);
         }
      }
   }
   template <int B, int S, int M> static D2D_HD void step3(T2 *v)
   {
      if constexpr (M < A) {
         Bfly<T, Bn>::template run<B + Bn * Bfly<T, A>::out_idx(M) * S, S>(v);
         step3<B, S, M + 1>(v);
      }
   }
   template <int B, int S> static D2D_HD void run(T2 *v)
   {
      step1<B, S, 0>(v);
      step2<B, S, 1, 1>(v);
      step3<B, S, 0>(v);
   }
   static constexpr int out_idx(int r) { return Bn * Bfly<T, A>::out_idx(r % A) + Bfly<T, Bn>::out_idx(r / A); }
};

template <typename T> struct Bfly<T, 6> : BflyCT<T, 3, 2> {};
template <typename T> struct Bfly<T, 12> : BflyCT<T, 3, 4> {};
template <typename T> struct Bfly<T, 24> : BflyCT<T, 3, 8> {};
template <typename T> struct Bfly<T, 10> : BflyCT<T, 5, 2> {};
template <typename T> struct Bfly<T, 20> : BflyCT<T, 5, 4> {};
template <typename T> struct Bfly<T, 15> : BflyCT<T, 3, 5> {};  // any-length kernel only (fft_any.cuh)
template <typename T> struct Bfly<T, 30> : BflyCT<T, 2, 15> {}; // any-length kernel only

} // namespace d2d
