// fft_v2_host.cpp -- host side of the TMA-staged FFT kernels (fft_kernel_v2.cuh): decides whether a
// stage can run on them (alignment rules of the TMA engine), builds the tensor maps and the list of
// box loads of a tile, and launches.  Falls back (returns false) to the v1 kernels otherwise.
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "common.h"
#include "fft_registry.h"

namespace d2d {

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
   static EncodeTiledFn fn = []() -> EncodeTiledFn {
      void *p = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
         return nullptr;
      return (EncodeTiledFn)p;
   }();
   return fn;
}

int env_int(const char *name, int dflt)
{
   const char *v = getenv(name);
   return v ? atoi(v) : dflt;
}

bool aligned16(const void *p) { return ((uintptr_t)p & 15) == 0; }

// One piece of a tile-like input seen as a 3-D tensor (dim0 = batch axis a in scalars, then the
// transform axis e and the batch axis b ordered by increasing stride).
struct PieceDesc {
   void *base;
   long long n0;         // extent of dim 0 in scalars
   long long rows;       // extent along e
   long long stride_e, stride_b; // bytes
   long long nb;
};

struct Builder {
   TmapPack tm;
   int nmaps = 0;
   struct Key { int piece, box_rows; } keys[kMaxTmaps];
   int swap[kMaxTmaps];
   bool swizzle = false; // 128-byte swizzle of the landing rows (merged landing, fft_kernel_v2.cuh FftArgs2::swz)
   int find_or_make(int piece, int box_rows, const PieceDesc &d, int f64, int box0)
   {
      for (int i = 0; i < nmaps; i++)
         if (keys[i].piece == piece && keys[i].box_rows == box_rows) return i;
      if (nmaps == kMaxTmaps) return -1;
      EncodeTiledFn fn = encode_fn();
      if (!fn) return -1;
      const int es = f64 ? 8 : 4;
      // outer dims sorted by stride (the engine wants non-decreasing strides to be safe)
      const long long sb = d.nb > 1 ? d.stride_b : 0;
      const bool sw = d.nb > 1 && sb < d.stride_e; // b before e
      cuuint64_t dims[3], strides[2];
      cuuint32_t box[3], estr[3] = {1, 1, 1};
      dims[0] = (cuuint64_t)d.n0;
      box[0] = (cuuint32_t)box0;
      const long long nb_stride = d.nb > 1 ? sb : (d.stride_e * d.rows + 15) / 16 * 16; // any legal value for an extent-1 axis
      if (!sw) {
         dims[1] = (cuuint64_t)d.rows; strides[0] = (cuuint64_t)d.stride_e; box[1] = (cuuint32_t)box_rows;
         dims[2] = (cuuint64_t)d.nb;   strides[1] = (cuuint64_t)nb_stride;  box[2] = 1;
      } else {
         dims[1] = (cuuint64_t)d.nb;   strides[0] = (cuuint64_t)sb;         box[1] = 1;
         dims[2] = (cuuint64_t)d.rows; strides[1] = (cuuint64_t)d.stride_e; box[2] = (cuuint32_t)box_rows;
      }
      static const int promo = env_int("D2D_TMA_L2PROMO", 1);
      const CUtensorMapL2promotion l2 = promo == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE
                                        : promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                     : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
      CUresult r = fn(&tm.m[nmaps], es == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d.base, dims, strides, box,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, l2,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return -1;
      keys[nmaps] = {piece, box_rows};
      swap[nmaps] = sw ? 1 : 0;
      return nmaps++;
   }
};

bool stride_ok(long long bytes) { return bytes > 0 && (bytes & 15) == 0 && bytes < (1LL << 40); }

} // namespace

// returns true when the stage was launched on a v2 kernel
bool fft_v2_try_launch(Ctx *ctx, const FftArgs &g, int f64, int mode, cudaError_t *err)
{
   static const int enabled = env_int("D2D_V2", 1);
   *err = cudaSuccess;
   if (!enabled || g.passthrough) return false;
   const int es = f64 ? 8 : 4; // scalar size
   const int ces = 2 * es;     // complex size
   // ---- which input layout? ----------------------------------------------------------------------
   int inl = -1;
   if (mode == MODE_R2C) {
      if (g.rsa == 1) inl = IN_TILE;
      else if (g.rse == 1) inl = IN_LINE;
   } else {
      bool tile = true, line = true;
      for (int m = 0; m < g.in.np; m++) {
         tile = tile && g.in.sa[m] == 1;
         line = line && g.in.se[m] == 1;
      }
      inl = tile ? IN_TILE : line ? IN_LINE : -1;
      if (tile && line) inl = (g.na > 1) ? IN_TILE : IN_LINE;
   }
   if (inl < 0) return false;
   // tile-like OUTPUT (TX adjacent lines are contiguous on the store side)?  Scattered 64-byte stores cost the
   // memory system far more than 128-byte ones (tools/micro/membench.cu), so those stages may use wide tiles.
   bool tile_out;
   if (mode == MODE_C2R) tile_out = g.rsa == 1 && g.rse != 1;
   else tile_out = g.out.sa[0] == 1 && g.out.se[0] != 1;
   static const int rb_all = env_int("D2D_V2_ROWBYTES", 64), rb_tout = env_int("D2D_V2_ROWBYTES_TILEOUT", 0);
   static const int rb_r2c = env_int("D2D_V2_ROWBYTES_R2C", 0); // experiments: row bytes of r2c stages whose input is a tile
   // default: 128-byte rows for fp64 tile-out stages (one block of 512 threads per SM); fp32 would need 1024 threads
   // fp32 tile-out: 128-byte rows need 16 lines = 512 threads and one block per SM; measured (1024^3, profiles/r01_e_kernels_ab.txt)
   // that pays only when the rows are far apart (the user's Z-pencil: 2.4 vs 5.1 ms), not inside the work buffers
   const long long out_pitch = (mode == MODE_C2R) ? g.rse * (long long)es : g.out.se[0] * (long long)ces;
   const int tout_default = f64 ? 128 : (out_pitch >= (1LL << 20) ? 128 : 64);
   int want = tile_out ? (rb_tout ? rb_tout : tout_default) : rb_all;
   if (rb_r2c && mode == MODE_R2C && inl == IN_TILE) want = rb_r2c;
   // tile inputs: the variant whose two adjacent sub-tiles land as ONE box of 128-byte rows (r2c by default: its rows
   // are the user's z-lines, 8 MB apart at 1024^3, where 64-byte rows cap the memory system at ~4.2 TB/s, profiles/
   // r01_c_membench_pitch.txt; D2D_V2_MERGE: 0 never, 1 r2c, 2 r2c and c2c)
   static const int merge_mode = env_int("D2D_V2_MERGE", 1);
   const FftKernelInfo *k = nullptr;
   // c2c tile inputs take it only when their rows are >= 1 MB apart (the user's Z-pencil read by c2c_z of PHYSICAL_IN_X):
   // with rows a few KB apart the two-way bank conflict of the shared landing rows costs more than it gives (2.9 -> 3.1 ms)
   const bool far_rows = mode == MODE_C2C && g.in.np == 1 && g.in.se[0] * (long long)ces >= (1LL << 20);
   if (inl == IN_TILE && want == 64 && ((merge_mode >= 1 && (mode == MODE_R2C || far_rows)) || (merge_mode >= 2 && mode == MODE_C2C)))
      k = fft_find_v2(g.n, f64, mode, inl, 64, 1);
   if (!k) k = fft_find_v2(g.n, f64, mode, inl, want);
   if (!k) k = fft_find_v2(g.n, f64, mode, inl, 64);
   if (!k) return false;
   if ((long long)((g.na + 3) / 4) * g.nb >= (1LL << 31)) return false; // the kernels count tiles in 32 bits

   FftArgs2 a2{};
   a2.a = g;
   a2.a.tw = twiddles_for(ctx->device, g.n, f64, 1);
   Builder B;
   memset(&B.tm, 0, sizeof(B.tm));
   const int N = g.n, NH = N / 2 + 1;

   if (inl == IN_LINE) {
      if (mode == MODE_R2C) {
         if (!aligned16(g.rptr) || ((g.rsa * es) & 15) || ((g.rsb * es) & 15) || (((long long)N * es) & 15)) return false;
      } else {
         const int len_total = (mode == MODE_C2R) ? NH : N;
         if (g.in.e0[g.in.np] != len_total) return false;
         for (int m = 0; m < g.in.np; m++) {
            if (!aligned16(g.in.ptr[m]) || ((g.in.sa[m] * ces) & 15) || ((g.in.sb[m] * ces) & 15)) return false;
            if (((long long)g.in.e0[m] * ces) & 15) return false;
            // bulk copies move multiples of 16 bytes: an odd fp32 piece (n/2+1 = 513 bins) is read one element long,
            // which is only legal when the line pitch leaves that room (padded wire layouts do, dense user arrays do not)
            const long long plen = g.in.e0[m + 1] - g.in.e0[m];
            const long long rlen = (plen * ces + 15) / 16 * 16 / ces;
            if (rlen != plen) {
               const long long pitch_a = g.in.sa[m] < 0 ? -g.in.sa[m] : g.in.sa[m], pitch_b = g.in.sb[m] < 0 ? -g.in.sb[m] : g.in.sb[m];
               const long long lines_a = (mode == MODE_C2R) ? g.na_real : g.na;
               if ((lines_a > 1 && pitch_a < rlen) || (g.nb > 1 && pitch_b < rlen) || (lines_a <= 1 && g.nb <= 1)) return false;
               if (m + 1 < g.in.np) return false; // the over-read would land on the next piece's slot in shared memory
            }
            a2.line_bytes += (int)(rlen * ces);
         }
      }
      if (mode == MODE_R2C) a2.line_bytes = N * es;
   } else {
      // ---- tile-like: tensor maps + box list ----------------------------------------------------------
      const int np = (mode == MODE_R2C) ? 1 : g.in.np;
      const int rows_total = (mode == MODE_C2R) ? NH : N;
      const int land_row_bytes = k->merged ? k->ly * k->row_bytes : k->row_bytes;
      const int box0 = land_row_bytes / es; // scalars per landing row
      a2.c0_mul = k->row_bytes / es / k->tx; // scalars per line along dim 0
      int row = 0; // global row (along e) of the tile
      a2.nops = 0;
      // swizzled landing (merged kernels, 128-byte rows): every box must start on a multiple of 8 rows, so that the swizzle
      // phase of a row is its index mod 8 whatever the engine derives it from (box row or shared-memory address bits)
      static const int swz_enabled = env_int("D2D_V2_SWIZZLE", 1);
      if (swz_enabled && k->merged && land_row_bytes == 128 && k->rows_early % 8 == 0) {
         bool ok = true;
         for (int m = 0; m < np && mode != MODE_R2C; m++) {
            const int len = g.in.e0[m + 1] - g.in.e0[m];
            ok = ok && (g.in.e0[m] % 8 == 0) && (len % 8 == 0 || m == np - 1);
         }
         if (mode == MODE_R2C) ok = (N % 8 == 0);
         // tail boxes shorter than 8 rows may only come last: a piece of length 8 q + r (r < 8) is cut 256, ..., 8, then 4 / 2 / 1
         const int last_len = (mode == MODE_R2C) ? N : g.in.e0[np] - g.in.e0[np - 1];
         const int tail = last_len % 8;
         ok = ok && (tail == 0 || tail == 1 || tail == 2 || tail == 4); // at most one box below 8 rows
         B.swizzle = ok;
      }
      for (int m = 0; m < np; m++) {
         PieceDesc d{};
         int len;
         if (mode == MODE_R2C) {
            d.base = g.rptr;
            d.n0 = g.na_real;
            len = N;
            d.stride_e = g.rse * (long long)es;
            d.stride_b = g.rsb * (long long)es;
         } else {
            d.base = g.in.ptr[m];
            d.n0 = (mode == MODE_C2R) ? 2LL * g.na_real : 2LL * g.na;
            len = g.in.e0[m + 1] - g.in.e0[m];
            if (g.in.e0[m] != row) return false;
            d.stride_e = g.in.se[m] * (long long)ces;
            d.stride_b = g.in.sb[m] * (long long)ces;
         }
         d.rows = len;
         d.nb = g.nb;
         if (!aligned16(d.base) || !stride_ok(d.stride_e) || (d.nb > 1 && !stride_ok(d.stride_b))) return false;
         if (d.n0 >= (1LL << 32) || d.rows >= (1LL << 32)) return false;
         // cut [row, row+len) at the early/late boundary, then into power-of-two boxes of <= 256 rows
         int done = 0;
         while (done < len) {
            const int grow = row + done;
            const int late = grow >= k->rows_early;
            int room = late ? (rows_total - grow) : (k->rows_early - grow);
            room = std::min(room, len - done);
            static const int box_rows = std::max(8, std::min(256, env_int("D2D_TMA_BOX_ROWS", 256))); // rows per box (power of two)
            int br = 256;
            while (br > room || br > box_rows) br >>= 1;
            const int mi = B.find_or_make(m, br, d, f64, box0);
            if (mi < 0 || a2.nops == kMaxLoadOps) return false;
            LoadOp &op = a2.ops[a2.nops++];
            op.map = (short)mi;
            op.late = (short)(late | (B.swap[mi] << 1));
            op.c1 = done;
            op.dst_row = late ? grow - k->rows_early : grow;
            op.bytes = br * land_row_bytes;
            (late ? a2.bytes_late : a2.bytes_early) += op.bytes;
            done += br;
         }
         row += len;
      }
      if (row != rows_total) return false;
   }
   a2.swz = B.swizzle ? 1 : 0;
   // ---- column shift (C2C): make the rows of the tile side start on full-row boundaries -------------------------
   a2.tiles_a = (g.na + k->tx - 1) / k->tx;
   static const int shift_enabled = env_int("D2D_V2_SHIFT", 1);
   if (shift_enabled && mode == MODE_C2C) {
      const int W = (k->merged ? k->ly : 1) * k->tx; // lines per row of the tile side
      const void *ptr = nullptr;
      long long se = 0, sb = 0;
      // tile INPUTS: measured no gain (TMA boxes are insensitive to the row phase: c2c_z_bwd 3.97 vs 4.10 ms), so only
      // D2D_V2_SHIFT=2 shifts them; tile OUTPUTS (LSU stores) gain 5.35 -> 3.34 ms (profiles/r01_h_kernels_ab.txt)
      if (inl == IN_TILE && g.in.np == 1 && shift_enabled >= 2 && !B.swizzle) { ptr = g.in.ptr[0]; se = g.in.se[0]; sb = g.in.sb[0]; }
      else if (inl == IN_LINE && tile_out && g.out.np == 1) { ptr = g.out.ptr[0]; se = g.out.se[0]; sb = g.out.sb[0]; }
      if (ptr && ((uintptr_t)ptr % ces) == 0 && (se % W) == 0 && se > 0 && sb >= 0) {
         const int off0 = (int)(((uintptr_t)ptr / ces) % W), offb = (int)(sb % W);
         if (off0 != 0 || (offb != 0 && g.nb > 1)) {
            a2.shift_on = 1;
            a2.shift0 = off0;
            a2.shift_b = offb;
            a2.tiles_a = (g.na + 2 * k->tx - 2) / k->tx; // room for the partial first tile of a row
            if (k->merged) a2.tiles_a = ((g.na + W - 1 + W - 1) / W) * k->ly;
         }
      }
   }
   if ((long long)a2.tiles_a * g.nb >= (1LL << 31)) return false;
   *err = k->launch2(a2, B.tm, ctx->stream);
   return true;
}

} // namespace d2d
