// fft_kernel_v2.cuh -- TMA-staged, double-landing variant of the batched 1-D FFT kernel (sm_100a).
//
// Same arithmetic as fft_kernel.cuh (Stockham passes, register butterflies, padded shared-memory
// exchange buffer X) and the same fused pack/unpack through piece maps on the store side; what
// changes is how a tile reaches the SM:
//
//   * the input of a tile is brought in by the TMA engine, not by LSU instructions:
//       IN_TILE  strided pencils (TX adjacent lines are contiguous in memory): cp.async.bulk.tensor
//                3-D boxes {TX lines x <=256 rows x 1} described by tensor maps built on the host,
//                one per piece of the receive layout; out-of-range lines are zero-filled by the
//                hardware, so ragged batch extents need no special code;
//       IN_LINE  lines contiguous along the transform axis: one cp.async.bulk (1-D) per line piece;
//     completion is tracked by two mbarriers (complete_tx::bytes);
//   * the landing zone of the FIRST half of a tile (buffer L) is separate from the exchange buffer,
//     so the first half of tile i+1 is requested as soon as tile i sits in registers and has the
//     whole tile period to arrive; the SECOND half lands in the exchange buffer itself once tile i
//     has read it for the last time (as in v1).  Two resident blocks per SM still fit
//     (N = 1024 fp64: 32 KB L + 69.6 KB X + 7.8 KB twiddles).
//   * twiddle tables of radix-2/4 passes keep only W^q; W^2q and W^3q are formed by multiplication
//     (saves 8 KB of shared memory at N = 1024, which is what lets L fit).
//   * c2r is pipelined like the other modes: the two Hermitian half-lines of a pair land raw and
//     are combined while being read into registers.
//
// Replaces the same reference code as fft_kernel.cuh: the cuFFT executions of src/fft_cufft.f90:489-671
// and mem_split_* / mem_merge_* of src/transpose_*.f90.
#pragma once
#include <cuda.h>

#include "fft_kernel.cuh"

namespace d2d {

enum InLayout { IN_TILE = 0, IN_LINE = 1 };

// Build-time switches of the two most recent kernel changes (make VARIANT=_safe EXTRA=-DD2D_V2_SAFE builds the library
// without them, for A/B runs): the r2c mirror exchange in the tail of X and the even/odd split of real line pairs
// across the two landing zones.
#ifdef D2D_V2_SAFE
constexpr bool kV2Mirror = false, kV2SplitPairs = false;
#else
constexpr bool kV2Mirror = true, kV2SplitPairs = true;
#endif

constexpr int kMaxTmaps = 16;
constexpr int kMaxLoadOps = 40;

struct LoadOp {
   short map;     // index into TmapPack::m
   short late;    // bit 0: 0 lands in L (first rows), 1 lands in the exchange buffer (last rows); bit 1: tensor dims are (a, b, e)
   int c1;        // row coordinate inside the piece's tensor
   int dst_row;   // landing row relative to the start of its zone
   int bytes;     // box bytes (expect_tx bookkeeping)
};

struct alignas(64) TmapPack {
   CUtensorMap m[kMaxTmaps];
};

struct FftArgs2 {
   FftArgs a;
   int nops;
   int c0_mul;               // dim-0 coordinate of a tile = a0 * c0_mul (scalars per line along dim 0)
   int bytes_early, bytes_late; // per sub-tile
   int line_bytes;           // IN_LINE: bytes landed per input line (pieces rounded up to 16 bytes)
   // C2C only: tile columns shifted per b so that the rows of the TILE side (input of IN_TILE stages, output of IN_LINE
   // stages) start on full-row boundaries in global memory even when the row pitch of the pencil is odd (the 513-wide
   // spectral pencils of PHYSICAL_IN_X): first column of tile ta of row b is ta * W - ((shift0 + b * shift_b) mod W),
   // W = lines per landing row; columns < 0 do not exist.  tiles_a counts the (possibly one more) tiles per b.
   int shift_on, shift0, shift_b, tiles_a;
   // merged landing only: the boxes land with the TMA engine's 128-byte swizzle (16-byte chunk index of a row XOR row mod 8)
   // and the two sub-tiles own the chunks of even / odd index instead of the first / second half of a row: the 8 lanes of a
   // quarter warp (4 chunks of row r, 4 of row r + 1) then read 8 distinct 16-byte bank groups instead of the same 4 twice
   int swz;
   LoadOp ops[kMaxLoadOps];
};

// ---- twiddle layout with compact radix-2/4 tables ---------------------------------------------------
template <class P> struct PlanInfo2 {
   using PI = PlanInfo<P>;
   static constexpr bool compact(int p) { return p >= 1 && (PI::radix(p) == 2 || PI::radix(p) == 4); }
   static constexpr int entries(int p) { return p < 1 ? 0 : (compact(p) ? PI::ns(p) : (PI::radix(p) - 1) * PI::ns(p)); }
   static constexpr int tw_off(int p) { return p <= 1 ? 0 : tw_off(p - 1) + entries(p - 1); }
   static constexpr int tw_total = tw_off(PI::npass);
};

// ---- PTX wrappers -----------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void *bar, int count)
{
   asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(void *bar, unsigned bytes)
{
   asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void *bar, unsigned parity)
{
   asm volatile("{\n"
                ".reg .pred p;\n"
                "WAIT_%=:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra DONE_%=;\n"
                "bra WAIT_%=;\n"
                "DONE_%=:\n"
                "}\n" ::"r"(smem_u32(bar)),
                "r"(parity)
                : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *tm, int c0, int c1, int c2, void *bar)
{
   asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(smem_u32(dst)),
                "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
                : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, unsigned bytes, void *bar)
{
   asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
                "r"(bytes), "r"(smem_u32(bar))
                : "memory");
}

// global stores of the results (cache-streaming / cache-global store hints were measured on the 1024^3 pairs: 18.60 / 18.40 ms
// against 18.53 ms with plain stores -- no effect beyond run-to-run noise, so plain stores stay)
template <typename V> __device__ __forceinline__ void gstore(V *p, V v) { *p = v; }

// two adjacent complex values as ONE store (the spectra of a real line pair that are neighbours in memory): 256-bit STG for
// fp64 (sm_100: st.global.v4.f64), 128-bit for fp32 -- instead of two stores that each fill half of every 32-byte sector
__device__ __forceinline__ void store_pair(double2 *q, double2 a, double2 b)
{
   asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};\n" ::"l"(q), "d"(a.x), "d"(a.y), "d"(b.x), "d"(b.y) : "memory");
}
__device__ __forceinline__ void store_pair(float2 *q, float2 a, float2 b) { *reinterpret_cast<float4 *>(q) = make_float4(a.x, a.y, b.x, b.y); }

// One pass with the compact twiddle layout (otherwise PassOp's).
template <typename T, class P, int PASS, int SP, int PADK> struct PassOp2 : PassOp<T, P, PASS, SP, PADK, true> {
   using Base = PassOp<T, P, PASS, SP, PADK, true>;
   using T2 = typename Vec2<T>::type;
   using PI = PlanInfo<P>;
   using PI2 = PlanInfo2<P>;
   static __device__ __forceinline__ void twiddle(T2 *v, int j, const T2 *__restrict__ tw)
   {
      if (PASS == 0) return;
      constexpr int R = Base::R, NS = Base::NS, NB = Base::NB, TPL = Base::TPL;
#pragma unroll
      for (int u = 0; u < NB; u++) {
         const int q = (j + TPL * u) % NS;
         if constexpr (PI2::compact(PASS)) {
            const T2 w1 = tw[PI2::tw_off(PASS) + q];
            v[u + NB] = cmul(v[u + NB], w1);
            if constexpr (R == 4) {
               const T2 w2 = cmul(w1, w1);
               const T2 w3 = cmul(w2, w1);
               v[u + 2 * NB] = cmul(v[u + 2 * NB], w2);
               v[u + 3 * NB] = cmul(v[u + 3 * NB], w3);
            }
         } else {
#pragma unroll
            for (int r = 1; r < R; r++) v[u + r * NB] = cmul(v[u + r * NB], tw[PI2::tw_off(PASS) + (r - 1) * NS + q]);
         }
      }
   }
};

// All passes.  The caller has put a block barrier between the landing reads and this call, so the
// first scatter needs none; `late` runs right after the LAST read of the exchange buffer.
template <typename T, class P, int PASS, int SP, int PADK, class Late> struct RunPasses2 {
   using T2 = typename Vec2<T>::type;
   using PI = PlanInfo<P>;
   static __device__ __forceinline__ void run(T2 *v, int j, T2 *lsm, const T2 *__restrict__ tw, Late &late)
   {
      using Op = PassOp2<T, P, PASS, SP, PADK>;
      Op::twiddle(v, j, tw);
      Op::butterflies(v);
      if constexpr (PASS + 1 < PI::npass) {
         if (PASS > 0) __syncthreads();
         Op::scatter(v, j, lsm);
         __syncthreads();
#pragma unroll
         for (int s = 0; s < P::E; s++) v[s] = lsm[padix<PADK>(j + P::T * s) * SP];
         if constexpr (PASS + 2 == PI::npass) late();
         RunPasses2<T, P, PASS + 1, SP, PADK, Late>::run(v, j, lsm, tw, late);
      } else {
         T2 w[P::E];
         Op::unpermute(v, w);
#pragma unroll
         for (int s = 0; s < P::E; s++) v[s] = w[s];
      }
   }
};

// ---- geometry ---------------------------------------------------------------------------------------
template <typename T, class P, int TX, int LY, int PADK, int MODE, int INL, bool MRG = false> struct Geom2 {
   using T2 = typename Vec2<T>::type;
   static constexpr int N = P::N, NH = N / 2 + 1;
   static constexpr int threads = TX * LY * P::T;
   static constexpr int bankq = 128 / (int)sizeof(T2);
   // exchange buffer: interleaved [padix(position)][tx] per sub-tile (v1's tile layout)
   static constexpr int x_line = padix<PADK>(N - 1) + 1;
   static constexpr size_t x_sub = ((size_t)x_line * TX * sizeof(T2) + 127) / 128 * 128; // bytes per sub-tile (TMA lands on 128-byte boundaries)
   static constexpr size_t x_bytes = x_sub * LY;
   // landing geometry.  "rows" run along the transform axis for IN_TILE, lines are whole for IN_LINE.
   //   IN_TILE : row = ROWU scalars-of-16-bytes... expressed in bytes below
   //   C2C     : row = TX complex              rows = N
   //   R2C     : row = TX real pairs           rows = N
   //   C2R     : row = 2 TX complex            rows = NH
   static constexpr int rows = (MODE == MODE_C2R) ? NH : N;
   static constexpr int row_bytes = (MODE == MODE_C2R ? 2 : 1) * TX * (int)sizeof(T2);
   static constexpr int rows_early = (MODE == MODE_C2R) ? N / 4 : N / 2; // boxes never straddle this row
   // IN_LINE: lines per sub-tile and their shared-memory pitch (bytes)
   //   C2C : TX complex lines of N;  R2C : 2 TX real lines of N;  C2R : 2 TX complex half-lines of NH
   static constexpr int nlines = (MODE == MODE_C2C) ? TX : 2 * TX;
   static constexpr int lines_early = nlines / 2;
   static constexpr int line_elems = (MODE == MODE_C2R) ? NH : N;                       // elements per line
   static constexpr int line_esize = (MODE == MODE_R2C) ? (int)sizeof(T) : (int)sizeof(T2);
   // pitch chosen so that the TX lines read by a quarter/half warp fall into distinct banks
   static constexpr int line_pitch_elems = (MODE == MODE_C2C)   ? N + (bankq / TX > 0 ? bankq / TX : 1)
                                           : (MODE == MODE_R2C) ? (kV2SplitPairs ? N + ((128 / (int)sizeof(T)) / TX > 2 ? (128 / (int)sizeof(T)) / TX : 2)
                                                                                 : N + 16 / (int)sizeof(T) * (TX >= 4 ? 2 : 4))
                                                                : NH;
   static constexpr size_t line_pitch = ((size_t)line_pitch_elems * line_esize + 15) / 16 * 16;
   static constexpr size_t l_sub = (((INL == IN_TILE) ? (size_t)rows_early * row_bytes : (size_t)lines_early * line_pitch) + 127) / 128 * 128;
   static constexpr size_t late_sub = (INL == IN_TILE) ? (size_t)(rows - rows_early) * row_bytes : (size_t)(nlines - lines_early) * line_pitch;
   static constexpr size_t l_bytes = l_sub * LY;
   static constexpr int late_skew = (INL == IN_LINE && MODE == MODE_C2C) ? (bankq / 2) * (int)sizeof(T2) : 0; // bank skew of the late zone
   static constexpr bool late_fits = late_sub + late_skew <= x_sub; // the late half of a tile lands in the exchange buffer of its sub-tile
   // R2C epilogue: only the upper half of Z (positions n/2..n-1) has to change hands to separate the two spectra.  It is
   // exchanged through the TAIL of the exchange buffer, behind the late landing zone, so that the late half of the next
   // tile can be requested right after the last pass exchange (as in C2C) instead of after the epilogue.
   static constexpr size_t mir_off = (late_skew + late_sub + 15) / 16 * 16;
   static constexpr size_t mir_bytes = (size_t)(padix<PADK>(N / 2 - 1) + 1) * TX * sizeof(T2);
   // MRG (IN_TILE, LY == 2): the LY sub-tiles of a group are ADJACENT along the tile axis and land together -- one TMA
   // box row carries LY * row_bytes (128 bytes for fp64), shared landing zones: the early one is L as a whole, the late one
   // spans the front of the exchange buffers (it must fit in front of the mirror regions, which all move into the last
   // sub-tile's buffer).  The sub-tiles still compute and store on their own (TX-wide lanes, full-width store segments).
   static constexpr bool merged = MRG;
   static_assert(!MRG || (INL == IN_TILE && LY >= 2), "merged landing is for tile inputs with several sub-tiles");
   static constexpr int land_row_bytes = MRG ? LY * row_bytes : row_bytes;
   static constexpr size_t late_all = (size_t)LY * late_sub; // bytes of the merged late zone
   static constexpr size_t mir_off_mrg = (size_t)(LY - 1) * x_sub; // mirrors of all sub-tiles: in the last sub-tile's buffer
   static constexpr bool mirror_fits = kV2Mirror && (MODE == MODE_R2C) && (P::E % 2 == 0) &&
                                       (MRG ? (late_all <= mir_off_mrg && (size_t)LY * mir_bytes <= x_sub) : (mir_off + mir_bytes <= x_sub));
   static_assert(!MRG || late_all <= x_bytes, "merged late zone must fit in the exchange buffers");
   static constexpr size_t tw_bytes = ((size_t)PlanInfo2<P>::tw_total * sizeof(T2) + 15) / 16 * 16;
   static constexpr size_t off_x = l_bytes;
   static constexpr size_t off_tw = off_x + x_bytes;
   static constexpr size_t off_bar = off_tw + tw_bytes;
   static constexpr size_t smem_bytes = off_bar + 16;
};

// Walks the pieces of a map for increasing e: the piece lookup and the (a, b) part of the address are redone only
// when e crosses into the next piece (store side of multi-rank chains, where every element used to pay a search).
template <typename T2> struct PieceCursor {
   const PieceMap &m;
   long long a, b;
   int pc, next_e0;
   T2 *base; // &piece[pc](e = 0, a, b), i.e. ptr - e0 se + a sa + b sb
   long long se;
   __device__ __forceinline__ PieceCursor(const PieceMap &m_, long long a_, long long b_) : m(m_), a(a_), b(b_), pc(-1), next_e0(0), base(nullptr), se(0)
   {
      load(0);
   }
   __device__ __forceinline__ void load(int p)
   {
      pc = p;
      next_e0 = (p + 1 < m.np) ? m.e0[p + 1] : 0x7fffffff;
      se = m.se[p];
      base = reinterpret_cast<T2 *>(m.ptr[p]) + a * m.sa[p] + b * m.sb[p] - (long long)m.e0[p] * se;
   }
   __device__ __forceinline__ T2 *at(int e)
   {
      while (e >= next_e0) load(pc + 1);
      return base + (long long)e * se;
   }
   __device__ __forceinline__ long long sa() const { return m.sa[pc]; }
};

// ---- the kernel -------------------------------------------------------------------------------------
template <typename T, class P, int TX, int LY, int PADK, int MODE, int INL, int MINB, bool MRG = false>
__global__ void __launch_bounds__(TX *LY *P::T, MINB) fft_kernel_v2(const __grid_constant__ FftArgs2 g2, const __grid_constant__ TmapPack tm)
{
   using T2 = typename Vec2<T>::type;
   using G = Geom2<T, P, TX, LY, PADK, MODE, INL, MRG>;
   static_assert(PlanInfo<P>::npass >= 2, "v2 kernels need at least one exchange");
   static_assert(G::late_fits, "the late half of a tile must fit in the exchange buffer of its sub-tile");
   static_assert(MODE == MODE_C2C || INL == IN_TILE || TX % 2 == 0, "real line pairs must land in the same zone");
   constexpr int SP = TX;
   constexpr int N = P::N, E = P::E, TPL = P::T, NH = N / 2 + 1;
   extern __shared__ __align__(1024) unsigned char smem2_raw[];
   unsigned char *smem_raw = smem2_raw;
   const FftArgs &g = g2.a;

   unsigned char *Lbase = smem_raw;
   unsigned char *Xbase = smem_raw + G::off_x;
   T2 *tws = reinterpret_cast<T2 *>(smem_raw + G::off_tw);
   unsigned long long *bar = reinterpret_cast<unsigned long long *>(smem_raw + G::off_bar); // [0] early, [1] late

   const int tid = threadIdx.x;
   const int tx = tid % TX;
   const int j = (tid / TX) % TPL;
   const int ly = tid / (TX * TPL);
   const int tiles_a = g2.tiles_a;    // (na + TX - 1) / TX, one more per b when the columns are shifted
   const int ntiles = tiles_a * g.nb; // the host guarantees < 2^31
   constexpr int SW = (MRG ? LY : 1) * TX; // lines per landing row (power of two)
   const bool shifted = (MODE == MODE_C2C) && g2.shift_on != 0;
   auto col_shift = [&](int b) { return shifted ? ((g2.shift0 + b * g2.shift_b) & (SW - 1)) : 0; };
   // MRG: a group is LY tiles adjacent along a, inside one b (groups_a per b); otherwise LY consecutive tiles
   const int groups_a = (tiles_a + LY - 1) / LY;
   const int ngroups = MRG ? groups_a * g.nb : (ntiles + LY - 1) / LY;
   T2 *lsm = reinterpret_cast<T2 *>(Xbase + (size_t)ly * G::x_sub) + tx;
   // landing zones as this thread reads them: row r of the early zone at Lmine + r * land_row_bytes (MRG: the zones are
   // shared, the sub-tile owns row_bytes at offset ly * row_bytes of every row -- or, swizzled, every other 16-byte chunk)
   // mrg_c: index of this thread's line among the LY * TX lines of the group; mrg_off: its byte offset inside a landing row
   constexpr int EPC = 16 / (int)sizeof(T2); // lines per 16-byte chunk of a landing row (C2C / R2C)
   const bool swz = MRG && g2.swz != 0;
   const int mrg_c = swz ? ((tx / EPC) * LY + ly) * EPC + tx % EPC : ly * TX + tx;
   const int mrg_off = swz ? ((((tx / EPC) * LY + ly) ^ (j & 7)) << 4) + (tx % EPC) * (int)sizeof(T2) : mrg_c * (int)sizeof(T2);
   static_assert(!MRG || (TPL % 8 == 0 && G::rows_early % 8 == 0), "swizzled landing: the row phase must be the thread's j mod 8");
   const unsigned char *Lmine = MRG ? Lbase + mrg_off : Lbase + (size_t)ly * G::l_sub;
   const unsigned char *Xmine = MRG ? Xbase + mrg_off : Xbase + (size_t)ly * G::x_sub + G::late_skew;
   const int tile_idx = MRG ? 0 : tx; // element of a landing row (C2C / R2C tiles): MRG rows are addressed through mrg_off

   {
      const T2 *__restrict__ twg = reinterpret_cast<const T2 *>(g.tw);
      for (int i = tid; i < PlanInfo2<P>::tw_total; i += TX * LY * TPL) tws[i] = ldg_nc(twg + i);
   }
   if (tid == 0) {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      fence_mbar_init();
   }
   __syncthreads();
   const unsigned conj_mask = g.backward ? 0x80000000u : 0u;
   const bool noload = (g.debug & 2) != 0;

   // ---- issue the loads of one half (which = 0 early / 1 late) of tile group `grp`; warp 0 only ------
   auto issue = [&](int grp, int which) {
      if (noload) return;
      void *mb = &bar[which];
      if constexpr (INL == IN_TILE && MRG) {
         if (tid == 0) { // one box row = the rows of all LY sub-tiles; columns beyond the batch extent are zero-filled
            mbar_expect_tx(mb, (unsigned)(which ? g2.bytes_late : g2.bytes_early));
            const int b = grp / groups_a;
            const int a0 = (grp - b * groups_a) * (LY * TX) - col_shift(b);
            unsigned char *zone = which ? Xbase : Lbase;
            for (int i = 0; i < g2.nops; i++) {
               const LoadOp &op = g2.ops[i];
               if ((op.late & 1) != which) continue;
               unsigned char *dst = zone + (size_t)op.dst_row * G::land_row_bytes;
               if (op.late & 2) tma_load_3d(dst, &tm.m[op.map], a0 * g2.c0_mul, b, op.c1, mb);
               else tma_load_3d(dst, &tm.m[op.map], a0 * g2.c0_mul, op.c1, b, mb);
            }
         }
      } else if constexpr (INL == IN_TILE) {
         if (tid == 0) {
            unsigned total = 0;
            for (int l = 0; l < LY; l++)
               if (grp * LY + l < ntiles) total += (unsigned)(which ? g2.bytes_late : g2.bytes_early);
            mbar_expect_tx(mb, total);
            // boxes of the same rows for the LY sub-tiles are requested back to back: adjacent sub-tiles are the two
            // halves of the same 128-byte lines of a strided pencil
            for (int i = 0; i < g2.nops; i++) {
               const LoadOp &op = g2.ops[i];
               if ((op.late & 1) != which) continue;
               for (int l = 0; l < LY; l++) {
                  const int tile = grp * LY + l;
                  if (tile >= ntiles) break;
                  const int b = tile / tiles_a;
                  const int a0 = (tile - b * tiles_a) * TX - col_shift(b);
                  unsigned char *zone = which ? (Xbase + (size_t)l * G::x_sub + G::late_skew) : (Lbase + (size_t)l * G::l_sub);
                  unsigned char *dst = zone + (size_t)op.dst_row * G::row_bytes;
                  if (op.late & 2) tma_load_3d(dst, &tm.m[op.map], a0 * g2.c0_mul, b, op.c1, mb); // tensor dims ordered (a, b, e)
                  else tma_load_3d(dst, &tm.m[op.map], a0 * g2.c0_mul, op.c1, b, mb);
               }
            }
         }
      } else {
         if (tid < 32) {
            // one bulk copy per (line, piece).  C2C: lines [lbase, lbase + nl) of each sub-tile belong to this half.
            // Real modes: the early zone takes the EVEN input lines (first line of every pair), the late zone the ODD
            // ones, so that the TX lines read together by a quarter warp sit at odd multiples of the pitch apart
            // (conflict-free; with pairs stored next to each other they were 2 pitches apart: 2-way conflicts).
            constexpr bool SPLIT = kV2SplitPairs && (MODE != MODE_C2C);
            constexpr int nE = G::lines_early, nL = G::nlines - G::lines_early;
            static_assert(!SPLIT || (nE == TX && nL == TX), "real modes: one zone per member of the line pairs");
            const int nl = which ? nL : nE;
            const int lbase = which ? nE : 0;
            const int np = (MODE == MODE_R2C) ? 1 : g.in.np;
            const int limit = (MODE == MODE_C2C) ? g.na : g.na_real; // input lines along a
            if (tid == 0) {
               unsigned lines = 0;
               for (int l = 0; l < LY; l++) {
                  const int tile = grp * LY + l;
                  if (tile >= ntiles) break;
                  int cnt;
                  if constexpr (SPLIT) cnt = (limit - 2 * ((tile % tiles_a) * TX) + (which ? 0 : 1)) / 2; // lines 2 (a0 + q) + which < limit
                  else if constexpr (MODE != MODE_C2C) cnt = limit - (2 * ((tile % tiles_a) * TX) + lbase);
                  else { // lines [lo, lo + nl) of this half that exist: 0 <= line < limit (lo < 0 when the columns are shifted)
                     const int bb = tile / tiles_a;
                     const int lo = (tile - bb * tiles_a) * TX - col_shift(bb) + lbase;
                     const int hi = lo + nl < limit ? lo + nl : limit;
                     cnt = hi - (lo > 0 ? lo : 0);
                  }
                  lines += (unsigned)(cnt < 0 ? 0 : cnt > nl ? nl : cnt);
               }
               mbar_expect_tx(mb, lines * (unsigned)g2.line_bytes);
            }
            __syncwarp();
            const int nop = LY * nl * np;
            for (int i = tid; i < nop; i += 32) {
               const int m = i % np;
               const int q = (i / np) % nl;
               const int l = i / (np * nl);
               const int tile = grp * LY + l;
               if (tile >= ntiles) continue;
               const int b = tile / tiles_a;
               constexpr int lmul = (MODE == MODE_C2C) ? 1 : 2; // input lines per complex line of the tile
               const int aline = SPLIT ? 2 * ((tile - b * tiles_a) * TX + q) + which : lmul * ((tile - b * tiles_a) * TX) - col_shift(b) + lbase + q;
               if (aline >= limit || aline < 0) continue;
               unsigned char *dst = (which ? (Xbase + (size_t)l * G::x_sub + G::late_skew) : (Lbase + (size_t)l * G::l_sub)) + (size_t)q * G::line_pitch;
               if constexpr (MODE == MODE_R2C) {
                  const T *src = reinterpret_cast<const T *>(g.rptr) + (long long)aline * g.rsa + (long long)b * g.rsb;
                  bulk_load(dst, src, (unsigned)(N * sizeof(T)), mb);
               } else {
                  const int e0 = g.in.e0[m], e1 = g.in.e0[m + 1];
                  const T2 *src = reinterpret_cast<const T2 *>(g.in.ptr[m]) + (long long)aline * g.in.sa[m] + (long long)b * g.in.sb[m];
                  bulk_load(dst + (size_t)e0 * sizeof(T2), src, (unsigned)(((e1 - e0) * sizeof(T2) + 15) & ~(size_t)15), mb);
               }
            }
         }
      }
   };

   int grp = blockIdx.x;
   if (grp < ngroups) {
      issue(grp, 0);
      issue(grp, 1);
   }
   unsigned phase = 0;
   for (; grp < ngroups; grp += gridDim.x, phase ^= 1) {
      T2 v[E];
      int b, a;
      bool in_range;
      if constexpr (MRG) {
         b = grp / groups_a;
         const int ta = (grp - b * groups_a) * LY + (swz ? 0 : ly);
         in_range = ta < tiles_a;
         a = (grp - b * groups_a) * (LY * TX) + mrg_c - col_shift(b);
      } else {
         const int tile = grp * LY + ly;
         in_range = tile < ntiles;
         b = tile / tiles_a;
         a = (tile - b * tiles_a) * TX + tx - col_shift(b);
      }
      const bool valid = in_range && a >= 0 && a < g.na;
      const bool v1 = valid && (2 * a + 1 < g.na_real);
      const int nxt = grp + (int)gridDim.x;

      // ------------------------------------------------------------------ landing -> registers
      if (!noload) {
         mbar_wait(&bar[0], phase);
         mbar_wait(&bar[1], phase);
      }
      if constexpr (MODE == MODE_C2C || MODE == MODE_R2C) {
#pragma unroll
         for (int s = 0; s < E; s++) {
            const int row = j + TPL * s;
            T2 x;
            if constexpr (INL == IN_TILE) {
               const unsigned char *src = (row < G::rows_early) ? Lmine + (size_t)row * G::land_row_bytes : Xmine + (size_t)(row - G::rows_early) * G::land_row_bytes;
               x = reinterpret_cast<const T2 *>(src)[tile_idx];
            } else if constexpr (MODE == MODE_C2C) {
               const unsigned char *line = (tx < G::lines_early) ? Lmine + (size_t)tx * G::line_pitch : Xmine + (size_t)(tx - G::lines_early) * G::line_pitch;
               x = reinterpret_cast<const T2 *>(line)[row];
            } else { // R2C from two real lines: even line in the early zone, odd line in the late zone
               if constexpr (kV2SplitPairs) {
                  x.x = reinterpret_cast<const T *>(Lmine + (size_t)tx * G::line_pitch)[row];
                  x.y = reinterpret_cast<const T *>(Xmine + (size_t)tx * G::line_pitch)[row];
               } else {
                  const int l0 = 2 * tx;
                  const unsigned char *la = (l0 < G::lines_early) ? Lmine + (size_t)l0 * G::line_pitch : Xmine + (size_t)(l0 - G::lines_early) * G::line_pitch;
                  x.x = reinterpret_cast<const T *>(la)[row];
                  x.y = reinterpret_cast<const T *>(la + G::line_pitch)[row];
               }
            }
            if constexpr (MODE == MODE_C2C) x.y = flip_sign(x.y, conj_mask);
            else if (!v1) x.y = 0; // an unpaired last real line: its partner must not leak into its spectrum
            // lines beyond the batch extent need no zeroing: TMA zero-fills them (tiles), or they hold stale finite-or-not
            // values whose transform is never stored (4 selects per element saved: 7.7 % of the c2c instruction stream)
            v[s] = x;
         }
      } else { // C2R: Z[k] = A[k] + i B[k], Z[n-k] = conj(A[k]) + i conj(B[k]); the forward passes get conj(Z)
         const unsigned char *la = nullptr, *lb = nullptr; // half-lines of the pair: A in the early zone, B in the late zone
         if constexpr (INL == IN_LINE) {
            if constexpr (kV2SplitPairs) {
               la = Lmine + (size_t)tx * G::line_pitch;
               lb = Xmine + (size_t)tx * G::line_pitch;
            } else {
               const int l0 = 2 * tx;
               la = (l0 < G::lines_early) ? Lmine + (size_t)l0 * G::line_pitch : Xmine + (size_t)(l0 - G::lines_early) * G::line_pitch;
               lb = la + G::line_pitch;
            }
         }
#pragma unroll
         for (int s = 0; s < E; s++) {
            const int p = j + TPL * s;
            // positions of slot s span [TPL s, TPL s + TPL): entirely below / above n/2 except for one slot
            const bool lower = (TPL * s + TPL - 1 <= N / 2) ? true : (TPL * s > N / 2) ? false : (p <= N / 2);
            const int k = lower ? p : N - p;
            T2 A, B;
            if constexpr (INL == IN_TILE) {
               const unsigned char *src = (k < G::rows_early) ? Lmine + (size_t)k * G::land_row_bytes : Xmine + (size_t)(k - G::rows_early) * G::land_row_bytes;
               A = reinterpret_cast<const T2 *>(src)[2 * tx];
               B = reinterpret_cast<const T2 *>(src)[2 * tx + 1];
            } else {
               A = reinterpret_cast<const T2 *>(la)[k];
               B = reinterpret_cast<const T2 *>(lb)[k];
            }
            if (!v1) B = T2{0, 0};
            if (TPL * s == 0 || TPL * s == N / 2) { // only these slots can hold bin 0 / bin n/2 (thread j = 0)
               if (j == 0) { A.y = 0; B.y = 0; }
            }
            v[s] = lower ? T2{A.x - B.y, -(A.y + B.x)} : T2{A.x + B.y, A.y - B.x};
         }
      }
      __syncthreads(); // everybody holds its tile: L (and the late zone, until the first scatter) are free
      if (nxt < ngroups) issue(nxt, 0);

      auto late = [&]() {
         fence_proxy_async(); // generic-proxy writes to X (scatters) before the async-proxy writes of the next landing
         __syncthreads(); // the whole block is done with the exchange buffer
         if (nxt < ngroups) issue(nxt, 1);
      };

      // ------------------------------------------------------------------ transform
      bool late_done = false;
      // C2R of a tile-like input writes real lines that are contiguous in memory (c2r_x of PHYSICAL_IN_X): straight from the
      // registers a warp would store 8-real runs of 8 different lines (64 / 32 bytes each), so the lines are staged through
      // the exchange buffer and leave as 512-byte runs.  The late half of the next tile is requested after that instead of
      // after the last exchange.
      constexpr int SPITCH = N + 16 / (int)sizeof(T); // reals per staged line (+16 bytes: the lines a half-warp writes fall into distinct banks)
      // fp32 only: 2.99 -> 1.72 ms at 1024^3 (32-byte runs before); fp64 loses (3.6 -> 4.6 ms: its 64-byte runs were tolerable
      // and the late half of the next tile, requested only after the staging, is no longer hidden behind the last pass)
      constexpr bool CAN_SOUT = (MODE == MODE_C2R && INL == IN_TILE && !MRG && sizeof(T) == 4 && (size_t)2 * TX * SPITCH * sizeof(T) <= G::x_sub);
      const bool sout = CAN_SOUT && g.rse == 1;
      if (!g.passthrough) {
         if constexpr (MODE == MODE_R2C && !G::mirror_fits) {
            auto nolate = [&]() {};
            RunPasses2<T, P, 0, SP, PADK, decltype(nolate)>::run(v, j, lsm, tws, nolate);
         } else if (CAN_SOUT && sout) {
            auto nolate = [&]() {};
            RunPasses2<T, P, 0, SP, PADK, decltype(nolate)>::run(v, j, lsm, tws, nolate);
         } else {
            RunPasses2<T, P, 0, SP, PADK, decltype(late)>::run(v, j, lsm, tws, late);
            late_done = true;
         }
      }

      // ------------------------------------------------------------------ store
      if constexpr (MODE == MODE_C2C) {
         if (!late_done) late();
         if (valid && !(g.debug & 1)) {
            if (g.out.np == 1) {
               T2 *p = reinterpret_cast<T2 *>(g.out.ptr[0]) + (long long)j * g.out.se[0] + (long long)a * g.out.sa[0] + (long long)b * g.out.sb[0];
               const long long step = (long long)TPL * g.out.se[0];
#pragma unroll
               for (int s = 0; s < E; s++) {
                  T2 x = v[s];
                  x.y = flip_sign(x.y, conj_mask);
                  gstore(p, x);
                  p += step;
               }
            } else {
               PieceCursor<T2> cur(g.out, a, b);
#pragma unroll
               for (int s = 0; s < E; s++) {
                  T2 x = v[s];
                  x.y = flip_sign(x.y, conj_mask);
                  gstore(cur.at(j + TPL * s), x);
               }
            }
         }
      } else if constexpr (MODE == MODE_C2R) {
         if (CAN_SOUT && sout) {
            __syncthreads(); // the last exchange has been read by everybody: the buffer is free (no late landing requested yet)
            T *stg = reinterpret_cast<T *>(Xbase + (size_t)ly * G::x_sub);
#pragma unroll
            for (int s = 0; s < E; s++) {
               stg[(2 * tx) * SPITCH + j + TPL * s] = v[s].x;
               stg[(2 * tx + 1) * SPITCH + j + TPL * s] = -v[s].y;
            }
            __syncthreads();
            constexpr int NW = TX * LY * TPL / 32, VW = 16 / (int)sizeof(T);
            const int warp = tid / 32, lane = tid % 32;
            if (!(g.debug & 1)) {
               for (int l = warp; l < LY * 2 * TX; l += NW) {
                  const int sl = l / (2 * TX), ll = l % (2 * TX);
                  const int tile = grp * LY + sl;
                  if (tile >= ntiles) continue;
                  const int bb = tile / tiles_a;
                  const long long ar = 2LL * ((tile - bb * tiles_a) * TX + ll / 2) + (ll & 1);
                  if (ar >= g.na_real) continue;
                  T *dst = reinterpret_cast<T *>(g.rptr) + ar * g.rsa + (long long)bb * g.rsb;
                  const T *src = reinterpret_cast<const T *>(Xbase + (size_t)sl * G::x_sub) + ll * SPITCH;
                  if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                     for (int c = lane * VW; c < N; c += 32 * VW) *reinterpret_cast<uint4 *>(dst + c) = *reinterpret_cast<const uint4 *>(src + c);
                  } else {
                     for (int c = lane; c < N; c += 32) dst[c] = src[c];
                  }
               }
            }
            late(); // fence + barrier, then the late half of the next tile may land in the buffer
            continue;
         }
         if (!late_done) late();
         // real output: element e of real line ar at rptr + e rse + ar rsa + b rsb; lines 2a, 2a+1 of this thread
         T *__restrict__ rp = reinterpret_cast<T *>(g.rptr) + (long long)j * g.rse + (long long)(2 * a) * g.rsa + (long long)b * g.rsb;
         const long long step = (long long)TPL * g.rse;
         const bool pairvec = (g.rsa == 1) && ((g.rsb & 1) == 0) && ((g.rse & 1) == 0) && ((reinterpret_cast<uintptr_t>(g.rptr) % (2 * sizeof(T))) == 0);
         if (!(g.debug & 1)) {
            if (pairvec && v1) {
#pragma unroll
               for (int s = 0; s < E; s++) {
                  gstore(reinterpret_cast<T2 *>(rp), T2{v[s].x, -v[s].y});
                  rp += step;
               }
            } else if (valid) {
#pragma unroll
               for (int s = 0; s < E; s++) {
                  rp[0] = v[s].x;
                  if (v1) rp[g.rsa] = -v[s].y;
                  rp += step;
               }
            }
         }
      } else { // R2C: separate the two spectra; partner of bin k is Z[n-k]
         T2 zn[E / 2 + 1];
         if constexpr (G::mirror_fits) {
            // slots s >= E/2 hold positions >= n/2; they go to mirror index (position - n/2) in the tail of X
            if (!late_done) late();
            T2 *mir = reinterpret_cast<T2 *>(MRG ? Xbase + G::mir_off_mrg + (size_t)ly * G::mir_bytes : Xbase + (size_t)ly * G::x_sub + G::mir_off) + tx;
#pragma unroll
            for (int s = E / 2; s < E; s++) mir[padix<PADK>(j + TPL * (s - E / 2)) * SP] = v[s];
            __syncthreads();
#pragma unroll
            for (int s = 0; s <= E / 2; s++) {
               const int k = j + TPL * s;
               zn[s] = v[0]; // k == 0 (thread j = 0, slot 0): its own partner
               if (k > 0 && k <= N / 2 && (s < E / 2 || j == 0)) zn[s] = mir[padix<PADK>(N / 2 - k) * SP]; // Z[n-k] sits at mirror index (n - k) - n/2
            }
         } else {
            __syncthreads();
#pragma unroll
            for (int s = 0; s < E; s++) lsm[padix<PADK>(j + TPL * s) * SP] = v[s];
            __syncthreads();
#pragma unroll
            for (int s = 0; s <= E / 2; s++) {
               const int k = j + TPL * s;
               zn[s] = T2{0, 0};
               if (k <= N / 2 && (s < E / 2 || j == 0)) zn[s] = lsm[padix<PADK>((N - k) % N) * SP];
            }
            late();
         }
         if (!(g.debug & 1) && valid) {
            PieceCursor<T2> cur(g.out, 2 * a, b);
#pragma unroll
            for (int s = 0; s <= E / 2; s++) {
               const int k = j + TPL * s;
               if (k <= N / 2 && (s < E / 2 || j == 0)) {
                  const T2 zk = v[s];
                  const T hf = (T)0.5;
                  T2 A = T2{(zk.x + zn[s].x) * hf, (zk.y - zn[s].y) * hf};
                  T2 B = T2{(zk.y + zn[s].y) * hf, (zn[s].x - zk.x) * hf};
                  T2 *q = cur.at(k);
                  if (v1 && cur.sa() == 1 && (reinterpret_cast<uintptr_t>(q) & (2 * sizeof(T2) - 1)) == 0) store_pair(q, A, B);
                  else {
                     gstore(q, A);
                     if (v1) gstore(q + cur.sa(), B);
                  }
               }
            }
         }
      }
   }
}

} // namespace d2d
