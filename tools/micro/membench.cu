// membench.cu -- raw memory-system capability for the access patterns of the strided FFT stages
// (no compute, no shared memory): each block moves tiles of ROWS rows x W bytes.
//   mode 0: contiguous read  -> contiguous write   (copy)
//   mode 1: tile (strided) read -> contiguous write  ("tile-in / line-out")
//   mode 2: contiguous read  -> tile (strided) write ("line-in / tile-out")
//   mode 3: tile read only      mode 4: tile write only     mode 5: contiguous read only   mode 6: contiguous write only
// usage: membench <W bytes per row> <mode> [blocks_per_sm] [row_stride_bytes]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

template <int U> __global__ void __launch_bounds__(256) k(const uint4 *__restrict__ src, uint4 *__restrict__ dst, int W, int mode, long long ntiles,
                                                           int tiles_a, long long row_stride16, long long plane16, int rows)
{
   const int cpr = W / 16;                 // 16-byte chunks per row
   const int tid = threadIdx.x;
   const int per_tile = rows * cpr;        // chunks per tile
   uint4 acc = make_uint4(0, 0, 0, 0);
   for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
      const long long b = t / tiles_a, a0 = t % tiles_a;
      const long long tbase = b * plane16 + a0 * cpr;        // strided view
      const long long cbase = t * (long long)per_tile;        // contiguous view
      for (int c0 = tid; c0 < per_tile; c0 += 256 * U) {
         uint4 v[U];
#pragma unroll
         for (int u = 0; u < U; u++) {
            const int c = c0 + u * 256;
            if (c < per_tile) {
               const int row = c / cpr, col = c % cpr;
               if (mode == 0 || mode == 2 || mode == 5) v[u] = src[cbase + c];
               else if (mode == 1 || mode == 3) v[u] = src[tbase + row * row_stride16 + col];
               else v[u] = make_uint4(c, 1, 2, 3);
            }
         }
#pragma unroll
         for (int u = 0; u < U; u++) {
            const int c = c0 + u * 256;
            if (c < per_tile) {
               const int row = c / cpr, col = c % cpr;
               if (mode == 0 || mode == 1 || mode == 6) dst[cbase + c] = v[u];
               else if (mode == 2 || mode == 4) dst[tbase + row * row_stride16 + col] = v[u];
               else { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
            }
         }
      }
   }
   if (acc.x == 0x12345678 && acc.y == 0x9abcdef0) dst[0] = acc;
}

int main(int argc, char **argv)
{
   const int W = argc > 1 ? atoi(argv[1]) : 64;
   const int mode = argc > 2 ? atoi(argv[2]) : 0;
   const int bps = argc > 3 ? atoi(argv[3]) : 4;
   const long long row_stride = argc > 4 ? atoll(argv[4]) : 16384;   // bytes between rows of a tile
   const int rows = 1024;
   const long long total = 8LL << 30;                                  // 8 GiB per array
   const int tiles_a = (int)(row_stride / W);
   const long long plane = row_stride * rows;                          // bytes per b-plane
   const long long nplanes = total / plane;
   const long long ntiles = nplanes * tiles_a;
   uint4 *src, *dst;
   CK(cudaMalloc(&src, total));
   CK(cudaMalloc(&dst, total));
   CK(cudaMemset(src, 1, total));
   CK(cudaMemset(dst, 2, total));
   cudaEvent_t e0, e1;
   cudaEventCreate(&e0);
   cudaEventCreate(&e1);
   float best = 1e9;
   for (int it = 0; it < 5; it++) {
      cudaEventRecord(e0);
      k<16><<<148 * bps, 256>>>(src, dst, W, mode, ntiles, tiles_a, row_stride / 16, plane / 16, rows);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0 && ms < best) best = ms;
   }
   const double bytes = (mode <= 2 ? 2.0 : 1.0) * (double)total;
   printf("W=%4d mode=%d bps=%d stride=%lld : %.3f ms  %.0f GB/s\n", W, mode, bps, row_stride, best, bytes / best / 1e6);
   return 0;
}
