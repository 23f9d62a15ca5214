// nvlbench.cu -- what store pattern does NVLink like?  One process, two GPUs with peer access: kernels on GPU 0 write
// (or copy) into a buffer on GPU 1.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o nvlbench nvlbench.cu
//   mode 0  cudaMemcpyPeerAsync (copy engine)                  mode 1  LSU stores, every warp one contiguous run of RUN bytes
//   mode 2  LSU copy (load local, store peer), same pattern    mode 3  TMA: cp.async.bulk global->shared->peer, PIECE bytes
// usage: nvlbench <mode> <run_or_piece_bytes> <blocks> [threads]
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void st_kernel(uint4 *dst, const uint4 *src, size_t n16, int run16, int copy)
{
   // a warp owns runs of run16 uint4; consecutive runs of a warp are `pitch` apart (scattered like the lines of a tile)
   const size_t warp = (size_t)blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32, nwarps = (size_t)gridDim.x * (blockDim.x / 32);
   const int lane = threadIdx.x % 32;
   const size_t nruns = n16 / run16;
   const int per = 32 / run16 > 0 ? 32 / run16 : 1; // runs covered by one warp instruction
   for (size_t r = warp * per; r < nruns; r += nwarps * per) {
      // run index of this lane: lanes [0, run16) -> run r, next run16 lanes -> run r + stride ... (scattered by a large stride)
      const size_t sub = lane / run16, e = lane % run16;
      size_t run = r + sub;
      // scatter: permute run index so that the `per` runs of an instruction are far apart
      run = (run % per) * (nruns / per) + run / per;
      if (run >= nruns) continue;
      for (int k = 0; k < run16; k += 32) {
         const size_t idx = run * run16 + e + k;
         if (e + k < (size_t)run16) dst[idx] = copy ? src[idx] : make_uint4(1, 2, 3, 4);
      }
   }
}

__device__ __forceinline__ unsigned s32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(32) tma_kernel(char *dst, const char *src, size_t bytes, int piece, int stages)
{
   extern __shared__ __align__(128) unsigned char sm[];
   unsigned long long *bar = (unsigned long long *)(sm + (size_t)stages * piece);
   if (threadIdx.x) return;
   for (int s = 0; s < stages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar[s])));
   asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
   const long long total = bytes / piece, mine = total > blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
   const int ahead = stages / 2;
   auto load = [&](long long i) {
      const int s = i % stages;
      const size_t off = ((size_t)blockIdx.x + i * gridDim.x) * piece;
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bar[s])), "r"(piece) : "memory");
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (size_t)s * piece)), "l"(src + off), "r"(piece), "r"(s32(&bar[s])) : "memory");
   };
   for (long long i = 0; i < mine && i < ahead; i++) load(i);
   for (long long i = 0; i < mine; i++) {
      const int s = i % stages;
      const unsigned par = (i / stages) & 1;
      asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(s32(&bar[s])), "r"(par) : "memory");
      const size_t off = ((size_t)blockIdx.x + i * gridDim.x) * piece;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"(s32(sm + (size_t)s * piece)), "r"(piece) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (i + ahead < mine) {
         // slot (i + ahead) % stages held piece i + ahead - stages: allow stages - ahead - 1 newer stores pending
         if (stages - ahead == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
         else if (stages - ahead == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
         else if (stages - ahead == 4) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
         else if (stages - ahead == 6) asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory");
         else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
         load(i + ahead);
      }
   }
   asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv)
{
   const int mode = argc > 1 ? atoi(argv[1]) : 0, run = argc > 2 ? atoi(argv[2]) : 512, blocks = argc > 3 ? atoi(argv[3]) : 296;
   const int threads = argc > 4 ? atoi(argv[4]) : 256;
   const int stages = argc > 5 ? atoi(argv[5]) : 6;
   const size_t bytes = (size_t)1 << 31;
   int n = 0;
   CK(cudaGetDeviceCount(&n));
   if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
   char *src, *dst;
   CK(cudaSetDevice(1));
   CK(cudaMalloc(&dst, bytes));
   CK(cudaSetDevice(0));
   CK(cudaDeviceEnablePeerAccess(1, 0));
   CK(cudaMalloc(&src, bytes));
   CK(cudaMemset(src, 1, bytes));
   cudaEvent_t a, b;
   CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
   float best = 1e9;
   for (int it = 0; it < 4; it++) {
      CK(cudaEventRecord(a));
      if (mode == 0) CK(cudaMemcpyPeerAsync(dst, 1, src, 0, bytes));
      else if (mode == 1 || mode == 2) st_kernel<<<blocks, threads>>>((uint4 *)dst, (const uint4 *)src, bytes / 16, run / 16, mode == 2);
      else {
         const size_t smem = (size_t)stages * run + 64;
         CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
         tma_kernel<<<blocks, 32, smem>>>(dst, src, bytes, run, stages);
      }
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      CK(cudaGetLastError());
      float ms;
      CK(cudaEventElapsedTime(&ms, a, b));
      if (it && ms < best) best = ms;
   }
   printf("mode %d run/piece %6d blocks %4d threads %4d stages %d : %.3f ms  %.1f GB/s\n", mode, run, blocks, threads, stages, best, bytes / best / 1e6);
   return 0;
}
