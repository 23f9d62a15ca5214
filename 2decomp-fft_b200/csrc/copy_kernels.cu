// copy_kernels.cu -- pack / unpack of the stand-alone transposes as ONE strided-copy kernel driven
// by the same piece maps as the FFT kernels.  Replaces the per-destination cudaMemcpy2D calls of
// mem_split_* / mem_merge_* (src/transpose_x_to_y.f90:300-306, 418-424; transpose_y_to_x.f90:300-306,
// 418-424; transpose_y_to_z.f90:404-410, 524-530; transpose_z_to_y.f90:404-410, 522-528) and the
// full-array cudaMemcpy staging copies (transpose_y_to_z.f90:176-181, transpose_z_to_y.f90:112-118).
#include "common.h"

namespace d2d {

namespace {

template <typename V> __global__ void __launch_bounds__(256) copy_kernel(const __grid_constant__ CopyArgs c)
{
   const int nfast = c.fast_is_a ? c.na : c.ne;
   const int nmid = c.fast_is_a ? c.ne : c.na;
   const int f = blockIdx.x * blockDim.x + threadIdx.x;
   if (f >= nfast) return;
   const long long b = blockIdx.z;
   for (int m = blockIdx.y; m < nmid; m += gridDim.y) {
      const int e = c.fast_is_a ? m : f;
      const long long a = c.fast_is_a ? f : m;
      int pi, po;
      const long long oi = piece_addr(c.in, e, a, b, pi);
      const long long oo = piece_addr(c.out, e, a, b, po);
      reinterpret_cast<V *>(c.out.ptr[po])[oo] = reinterpret_cast<const V *>(c.in.ptr[pi])[oi];
   }
}

bool divisible(const PieceMap &m, int vw, int es, bool fast_is_a)
{
   for (int p = 0; p < m.np; p++) {
      if ((uintptr_t)m.ptr[p] % ((size_t)vw * es)) return false;
      if (fast_is_a) {
         if (m.sa[p] != 1 || m.se[p] % vw || m.sb[p] % vw) return false;
      } else {
         if (m.se[p] != 1 || m.sa[p] % vw || m.sb[p] % vw || m.e0[p] % vw) return false;
      }
   }
   if (!fast_is_a && m.e0[m.np] % vw) return false;
   return true;
}

void scale(PieceMap &m, int vw, bool fast_is_a)
{
   for (int p = 0; p < m.np; p++) {
      m.sb[p] /= vw;
      if (fast_is_a) m.se[p] /= vw;
      else { m.sa[p] /= vw; m.e0[p] /= vw; }
   }
   if (!fast_is_a) m.e0[m.np] /= vw;
}

} // namespace

void launch_copy(Ctx *ctx, const CopyArgs &c0, int es, const char *label)
{
   CopyArgs c = c0;
   if (c.ne <= 0 || c.na <= 0 || c.nb <= 0) return;
   ProfScope ps(ctx, label ? label : "copy", 2.0 * es * (double)c.ne * c.na * c.nb); // read + write of the pencil
   // widen to 16-byte (or 8-byte) vectors along the unit-stride axis when every offset allows it
   int vw = 16 / es;
   while (vw > 1) {
      const int nfast = c.fast_is_a ? c.na : c.ne;
      if (nfast % vw == 0 && divisible(c.in, vw, es, c.fast_is_a) && divisible(c.out, vw, es, c.fast_is_a)) break;
      vw /= 2;
   }
   if (vw > 1) {
      scale(c.in, vw, c.fast_is_a);
      scale(c.out, vw, c.fast_is_a);
      if (c.fast_is_a) c.na /= vw;
      else c.ne /= vw;
   }
   const int ves = es * vw;
   const int nfast = c.fast_is_a ? c.na : c.ne;
   const int nmid = c.fast_is_a ? c.ne : c.na;
   D2D_REQUIRE(c.nb <= 65535, "copy: slow extent exceeds the grid limit");
   const int threads = 256;
   dim3 grid((nfast + threads - 1) / threads, (unsigned)std::min(nmid, 16384), (unsigned)c.nb);
   if (ves == 16) copy_kernel<uint4><<<grid, threads, 0, ctx->stream>>>(c);
   else if (ves == 8) copy_kernel<uint2><<<grid, threads, 0, ctx->stream>>>(c);
   else if (ves == 4) copy_kernel<uint32_t><<<grid, threads, 0, ctx->stream>>>(c);
   else D2D_REQUIRE(false, "copy: unsupported element size");
   D2D_CHECK_CUDA(cudaGetLastError());
   ctx->launches++;
}

} // namespace d2d

// ---- box copies of the halo exchange (halo.cpp) ---------------------------------------------------------------------
namespace d2d {
namespace {
template <typename V> __global__ void __launch_bounds__(256) box_copy_kernel(V *dst, long long d1, long long d12, const V *src, long long s1, long long s12,
                                                                              int e1, int e2, int e3)
{
   const int i = blockIdx.x * blockDim.x + threadIdx.x;
   if (i >= e1) return;
   for (int k = blockIdx.z; k < e3; k += gridDim.z)
      for (int j = blockIdx.y; j < e2; j += gridDim.y) dst[i + d1 * j + d12 * k] = src[i + s1 * j + s12 * k];
}
} // namespace

// dst(i, j, k) = src(i, j, k) for a box of (e1, e2, e3) elements of `es` bytes inside two Fortran-ordered arrays with leading
// dimensions (d1, d2) / (s1, s2); dst and src point at the first element of the box
void launch_box_copy(Ctx *ctx, void *dst, long long d1, long long d2, const void *src, long long s1, long long s2, int e1, int e2, int e3, int es)
{
   if (e1 <= 0 || e2 <= 0 || e3 <= 0) return;
   const dim3 grid((e1 + 255) / 256, (unsigned)std::min(e2, 4096), (unsigned)std::min(e3, 4096));
   if (es == 16) box_copy_kernel<uint4><<<grid, 256, 0, ctx->stream>>>((uint4 *)dst, d1, d1 * d2, (const uint4 *)src, s1, s1 * s2, e1, e2, e3);
   else if (es == 8) box_copy_kernel<uint2><<<grid, 256, 0, ctx->stream>>>((uint2 *)dst, d1, d1 * d2, (const uint2 *)src, s1, s1 * s2, e1, e2, e3);
   else if (es == 4) box_copy_kernel<uint32_t><<<grid, 256, 0, ctx->stream>>>((uint32_t *)dst, d1, d1 * d2, (const uint32_t *)src, s1, s1 * s2, e1, e2, e3);
   else D2D_REQUIRE(false, "box copy: unsupported element size");
   D2D_CHECK_CUDA(cudaGetLastError());
   ctx->launches++;
}
} // namespace d2d
