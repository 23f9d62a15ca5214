// fft_inst.cu -- instantiates the FFT kernels for one (D2D_N, D2D_F64) pair; compiled once per
// pair (see Makefile) so that the template-heavy code builds in parallel.
#include "fft_registry.h"

#ifndef D2D_N
#error "compile with -DD2D_N=<transform length> -DD2D_F64=<0|1>"
#endif

namespace d2d {
namespace {

#if D2D_F64
using real_t = double;
#else
using real_t = float;
#endif
using T2 = Vec2<real_t>::type;
using P = Pow2Plan<D2D_N>;

constexpr int cmin(int a, int b) { return a < b ? a : b; }
constexpr int cmax(int a, int b) { return a > b ? a : b; }
constexpr int pow2_floor(int v) { int p = 1; while (2 * p <= v) p *= 2; return p; } // tile extents stay powers of two for 3 * 2^k / 5 * 2^k lines

// ---- geometry (see DESIGN.md "kernel geometry") ------------------------------------------------
constexpr int kTargetThreads = 256;
constexpr int kMaxSmem = 128 * 1024;
constexpr int kLineBytes = P::N * (int)sizeof(T2);
// contiguous-line kernel: TX = 1, LY lines per block
constexpr int LY_LINE = pow2_floor(cmax(1, cmin(kTargetThreads / P::T, kMaxSmem / kLineBytes)));
// strided-tile kernel: TX adjacent lines so that one row of the tile is 64 B
constexpr int TX_WANT = 64 / (int)sizeof(T2);
constexpr int TX_TILE = pow2_floor(cmax(1, cmin(TX_WANT, cmin(1024 / P::T, kMaxSmem / kLineBytes))));
constexpr int LY_TILE = pow2_floor(cmax(1, cmin(kTargetThreads / (TX_TILE * P::T), kMaxSmem / (kLineBytes * TX_TILE))));
// wide tiles (128 B rows) for the stages whose far side is a user array with a huge row pitch
constexpr int TX_WIDE = pow2_floor(cmax(1, cmin(2 * TX_WANT, cmin(1024 / P::T, (200 * 1024) / kLineBytes))));
constexpr int LY_WIDE = pow2_floor(cmax(1, cmin(kTargetThreads / (TX_WIDE * P::T), kMaxSmem / (kLineBytes * TX_WIDE))));
constexpr int PADK = P::R0; // one padding element per first-pass butterfly (bank-conflict model: tools/smem_conflicts.py)

// resident blocks per SM the v1 kernels are compiled for; 32-element plans need the whole register file of one block
// (20 / 24 fp64 elements per thread -- the 5 * 2^k / 3 * 2^k plans -- spill at the 128 registers of two 256-thread blocks)
constexpr int minb_for(int threads) { return P::E >= 32 ? (D2D_F64 ? 1 : 2) : (P::E >= 20 && D2D_F64) ? 1 : cmax(1, (D2D_F64 ? 512 : 768) / threads); }

template <int TX, int LY, int MODE, bool PAIRVEC, bool LM = false> struct Inst {
   using G = KernelGeom<real_t, P, TX, LY, PADK, LM>;
   static constexpr int MINB = minb_for(G::threads);
   static cudaError_t launch(const FftArgs &g, cudaStream_t st)
   {
      auto kern = fft_kernel<real_t, P, TX, LY, PADK, MODE, PAIRVEC, MINB, LM>;
      // per-device caches (function attributes belong to a device: a process may drive several, d2d_ctx_create_in_group)
      static int resident_of[kMaxDevices] = {0}; // blocks that fit on the whole GPU at once (persistent grid); 0 = not set up
      const size_t smem = G::needs_smem ? G::smem_bytes : 0;
      int dev = 0;
      if (cudaError_t e0 = cudaGetDevice(&dev); e0 != cudaSuccess) return e0;
      if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
      static int per_sm_of[kMaxDevices] = {0};
      int &resident = resident_of[dev];
      if (!resident) {
         if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
         }
         int sms = 0, per_sm = 0;
         cudaError_t e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
         if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G::threads, smem);
         if (e != cudaSuccess) return e;
         if (per_sm < 1) return cudaErrorLaunchOutOfResources;
         per_sm_of[dev] = per_sm;
         resident = sms * per_sm;
      }
      const long long tiles = (long long)((g.na + TX - 1) / TX) * g.nb;
      const long long groups = (tiles + LY - 1) / LY;
      if (groups <= 0) return cudaSuccess;
      const long long cap = (g.sm_limit > 0 && g.sm_limit * per_sm_of[dev] < resident) ? g.sm_limit * per_sm_of[dev] : resident;
      const unsigned blocks = (unsigned)(groups < cap ? groups : cap);
      kern<<<blocks, G::threads, smem, st>>>(g);
      return cudaGetLastError();
   }
   static void reg(int kind)
   {
      using PI = PlanInfo<P>;
      FftKernelInfo k{};
      k.n = P::N; k.f64 = D2D_F64; k.kind = kind; k.mode = MODE; k.pairvec = PAIRVEC ? 1 : 0;
      k.tx = TX; k.ly = LY; k.threads = G::threads; k.minb = MINB;
      k.smem = G::needs_smem ? G::smem_bytes : 0;
      k.tw_total = PI::tw_total; k.npass = PI::npass;
      for (int p = 0; p < 4; p++) k.radix[p] = PI::radix(p);
      k.func = (const void *)fft_kernel<real_t, P, TX, LY, PADK, MODE, PAIRVEC, MINB, LM>;
      k.line_in = LM ? 1 : 0;
      k.launch = &launch;
      fft_register(k);
   }
};

// ---- v2: TMA-staged kernels (fft_kernel_v2.cuh) --------------------------------------------------
constexpr int TX2N = 64 / (int)sizeof(T2);                                 // 64-byte tile rows, two blocks per SM
constexpr int TX2W = 128 / (int)sizeof(T2);                                // 128-byte tile rows (one L2 line per row)
constexpr size_t kSmemSM = 227 * 1024;

template <int MODE, int INL, int TX2, bool MRG = false> struct Inst2 {
   // merged landing: always two sub-tiles (fp32: 2 x 8 lines = 512 threads; N = 512: two blocks of 128 threads per SM)
   static constexpr int LY2 = MRG ? 2 : cmax(1, kTargetThreads / (TX2 * P::T));
   static constexpr bool mrg_ok = !MRG || (INL == IN_TILE);
   using G = Geom2<real_t, P, TX2, LY2, PADK, MODE, INL, MRG && mrg_ok>;
   static constexpr bool fits = mrg_ok && PlanInfo<P>::npass >= 2 && P::N >= 256 && G::late_fits && G::threads <= 1024 && G::smem_bytes + 1024 <= kSmemSM &&
                                (!MRG || G::late_all <= G::x_bytes);
   static constexpr int MINB = (G::threads <= 256 && 2 * (G::smem_bytes + 1024) <= kSmemSM) ? 2 : 1;
   static cudaError_t launch(const FftArgs2 &g, const TmapPack &tm, cudaStream_t st)
   {
      auto kern = fft_kernel_v2<real_t, P, TX2, LY2, PADK, MODE, INL, MINB, MRG && mrg_ok>;
      static int resident_of[kMaxDevices] = {0}; // per device, 0 = not set up (see Inst::launch)
      const size_t smem = G::smem_bytes;
      int dev = 0;
      if (cudaError_t e0 = cudaGetDevice(&dev); e0 != cudaSuccess) return e0;
      if (dev < 0 || dev >= kMaxDevices) return cudaErrorInvalidDevice;
      static int per_sm_of[kMaxDevices] = {0};
      int &resident = resident_of[dev];
      if (!resident) {
         cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
         if (e != cudaSuccess) return e;
         int sms = 0, per_sm = 0;
         e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
         if (e == cudaSuccess) e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, G::threads, smem);
         if (e != cudaSuccess) return e;
         if (per_sm < 1) return cudaErrorLaunchOutOfResources;
         per_sm_of[dev] = per_sm;
         resident = sms * per_sm;
      }
      const long long tiles_a = g.tiles_a;
      const long long groups = MRG ? ((tiles_a + LY2 - 1) / LY2) * g.a.nb : (tiles_a * g.a.nb + LY2 - 1) / LY2;
      if (groups <= 0) return cudaSuccess;
      const long long cap = (g.a.sm_limit > 0 && g.a.sm_limit * per_sm_of[dev] < resident) ? g.a.sm_limit * per_sm_of[dev] : resident;
      const unsigned blocks = (unsigned)(groups < cap ? groups : cap);
      kern<<<blocks, G::threads, smem, st>>>(g, tm);
      return cudaGetLastError();
   }
   static void reg()
   {
      if constexpr (fits) {
         using PI = PlanInfo<P>;
         FftKernelInfo k{};
         k.n = P::N; k.f64 = D2D_F64; k.kind = KIND_TILE; k.mode = MODE; k.pairvec = 0;
         k.tx = TX2; k.ly = LY2; k.threads = G::threads; k.minb = MINB;
         k.smem = G::smem_bytes;
         k.tw_total = PlanInfo2<P>::tw_total; k.npass = PI::npass;
         for (int p = 0; p < 4; p++) k.radix[p] = PI::radix(p);
         k.func = (const void *)fft_kernel_v2<real_t, P, TX2, LY2, PADK, MODE, INL, MINB, MRG && mrg_ok>;
         k.merged = MRG ? 1 : 0;
         k.line_in = INL == IN_LINE;
         k.launch = nullptr;
         k.v2 = 1; k.inl = INL;
         k.rows = G::rows; k.rows_early = G::rows_early; k.row_bytes = G::row_bytes;
         k.launch2 = &launch;
         fft_register(k);
      }
   }
};

struct Registrar {
   Registrar()
   {
      Inst2<MODE_C2C, IN_TILE, TX2N>::reg();
      Inst2<MODE_R2C, IN_TILE, TX2N, true>::reg(); // two adjacent sub-tiles landing as one 128-byte-row box
      Inst2<MODE_C2C, IN_TILE, TX2N, true>::reg();
      Inst2<MODE_C2C, IN_LINE, TX2N>::reg();
      Inst2<MODE_R2C, IN_TILE, TX2N>::reg();
      Inst2<MODE_R2C, IN_LINE, TX2N>::reg();
      Inst2<MODE_C2R, IN_TILE, TX2N>::reg();
      Inst2<MODE_C2R, IN_LINE, TX2N>::reg();
      Inst2<MODE_C2C, IN_TILE, TX2W>::reg();
      Inst2<MODE_C2C, IN_LINE, TX2W>::reg();
      Inst2<MODE_R2C, IN_TILE, TX2W>::reg();
      Inst2<MODE_R2C, IN_LINE, TX2W>::reg();
      Inst2<MODE_C2R, IN_TILE, TX2W>::reg();
      Inst2<MODE_C2R, IN_LINE, TX2W>::reg();
      Inst<1, LY_LINE, MODE_C2C, false>::reg(KIND_LINE);
      Inst<1, LY_LINE, MODE_R2C, false>::reg(KIND_LINE);
      Inst<1, LY_LINE, MODE_C2R, false>::reg(KIND_LINE);
      Inst<TX_TILE, LY_TILE, MODE_C2C, false>::reg(KIND_TILE);
      Inst<TX_TILE, LY_TILE, MODE_R2C, true>::reg(KIND_TILE);
      Inst<TX_TILE, LY_TILE, MODE_C2R, true>::reg(KIND_TILE);
      Inst<TX_TILE, LY_TILE, MODE_R2C, false>::reg(KIND_TILE);
      Inst<TX_TILE, LY_TILE, MODE_C2R, false>::reg(KIND_TILE);
      if constexpr (TX_TILE > 1 && PlanInfo<P>::npass > 1) { // line-like inputs: line-major shared memory
         Inst<TX_TILE, LY_TILE, MODE_C2C, false, true>::reg(KIND_TILE);
         Inst<TX_TILE, LY_TILE, MODE_R2C, false, true>::reg(KIND_TILE);
      }
      if constexpr (TX_WIDE > TX_TILE) {
         Inst<TX_WIDE, LY_WIDE, MODE_R2C, true>::reg(KIND_TILE_WIDE);
         Inst<TX_WIDE, LY_WIDE, MODE_C2R, true>::reg(KIND_TILE_WIDE);
      }
   }
} registrar_instance;

} // namespace
} // namespace d2d
