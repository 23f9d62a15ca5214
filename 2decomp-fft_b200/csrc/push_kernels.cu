// push_kernels.cu -- the exchange of the chunk-pipelined 3-D transforms as ONE kernel per chunk: a handful of CTAs
// stream the sub-ranges of a chunk from this rank's send buffer straight into the receive buffers of ALL peers of the
// communicator over NVLink, with the TMA engine in both directions (cp.async.bulk global -> shared, shared -> peer global),
// and tell each peer that the chunk has landed with a release-add on a counter in the peer's memory.
//
// Replaces, for these chunks, the grouped ncclSend / ncclRecv of decomp_2d_nccl_alltoall_* (src/decomp_2d_nccl.f90:214-473)
// and the host synchronisation that follows it (:249, :324, :395, :470).  Opt-in (D2D_PUSH=sm); the default pushes chunks
// with the copy engines.  Measured on B200 (profiles/r02_c_nvlink_store_patterns_2gpu.txt, r02_b_*): a peer cudaMemcpyAsync
// runs at 780 GB/s but costs 25-40 us of dead time per call; SM-issued writes -- LSU stores in runs of >= 128 bytes or TMA
// bulk stores alike -- top out at ~715 GB/s, and one driving thread per CTA sustains ~45 GB/s (latency of the load /
// store ring), so this kernel needs 16-32 CTAs, each owning an SM next to the 1-block-per-SM FFT kernels: 634 GB/s with 16
// CTAs inside the 2-GPU headline chain (11.5 ms per pair against 10.9 ms with the copy engines).
//
// One elected thread per CTA drives a ring of kStages shared-memory stages: loads run kStages - 1 pieces ahead (mbarrier
// complete_tx), each landed piece leaves as a bulk store (bulk async-group), a stage is reloaded once the store that read it
// has drained (wait_group.read 1).  Shared memory is only a staging FIFO: no LSU instruction touches the data.
#include "common.h"
#include "fft_registry.h"

namespace d2d {

namespace {

constexpr int kPieceBytes = 32 * 1024;
constexpr int kStages = 6;

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes)
{
   asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;\n" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;\n" ::: "memory"); }

__global__ void __launch_bounds__(32, 1) push_kernel(const __grid_constant__ PushArgs g)
{
   extern __shared__ __align__(128) unsigned char psm[];
   unsigned long long *bar = reinterpret_cast<unsigned long long *>(psm + (size_t)kStages * kPieceBytes);
   if (threadIdx.x != 0) return;
   for (int s = 0; s < kStages; s++) mbar_init(&bar[s], 1);
   fence_mbar_init();

   // pieces of all segments, numbered segment after segment; this CTA takes pieces blockIdx.x, + gridDim.x, ...
   long long first[kMaxPushSegs + 1];
   first[0] = 0;
   for (int i = 0; i < g.nseg; i++) first[i + 1] = first[i] + (long long)((g.seg[i].bytes + kPieceBytes - 1) / kPieceBytes);
   const long long total = first[g.nseg];
   const long long mine = total > (long long)blockIdx.x ? (total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
   auto locate = [&](long long i, const char *&src, char *&dst, unsigned &bytes) {
      const long long p = (long long)blockIdx.x + i * gridDim.x;
      int s = 0;
      while (p >= first[s + 1]) s++;
      const size_t off = (size_t)(p - first[s]) * kPieceBytes;
      const size_t left = g.seg[s].bytes - off;
      src = g.seg[s].src + off;
      dst = g.seg[s].dst + off;
      bytes = (unsigned)(left < (size_t)kPieceBytes ? left : (size_t)kPieceBytes);
   };
   auto load = [&](long long i) {
      const char *src; char *dst; unsigned bytes;
      locate(i, src, dst, bytes);
      const int s = (int)(i % kStages);
      mbar_expect_tx(&bar[s], bytes);
      bulk_load(psm + (size_t)s * kPieceBytes, src, bytes, &bar[s]);
   };
   for (long long i = 0; i < mine && i < kStages - 1; i++) load(i);
   for (long long i = 0; i < mine; i++) {
      const int s = (int)(i % kStages);
      mbar_wait(&bar[s], (unsigned)((i / kStages) & 1));
      const char *src; char *dst; unsigned bytes;
      locate(i, src, dst, bytes);
      bulk_store(dst, psm + (size_t)s * kPieceBytes, bytes);
      bulk_commit();
      if (i + kStages - 1 < mine) {
         bulk_wait_read_1(); // the store of piece i - 1 has read its stage, which piece i + kStages - 1 reuses
         load(i + kStages - 1);
      }
   }
   bulk_wait_all(); // every store of this CTA is complete
   asm volatile("fence.proxy.async;\n" ::: "memory"); // the stores went through the async proxy; the counter update does not
   __threadfence_system();
   // release-add on every peer's arrival counter: the peer's stream waits for (CTAs of the launch) x (chunks so far)
   for (int i = 0; i < g.nflag; i++)
      asm volatile("red.release.sys.global.add.u32 [%0], %1;\n" ::"l"(g.flag[i]), "r"(1u) : "memory");
}

} // namespace

int push_ctas()
{
   static const int n = [] {
      const char *v = getenv("D2D_PUSH_CTAS");
      const int k = v ? atoi(v) : 12;
      return k < 1 ? 1 : k > 64 ? 64 : k;
   }();
   return n;
}

void launch_push(const PushArgs &a, cudaStream_t st)
{
   static bool ready[kMaxDevices] = {};
   const size_t smem = (size_t)kStages * kPieceBytes + kStages * sizeof(unsigned long long);
   int dev = 0;
   D2D_CHECK_CUDA(cudaGetDevice(&dev));
   D2D_REQUIRE(dev >= 0 && dev < kMaxDevices, "device index out of range");
   if (!ready[dev]) {
      D2D_CHECK_CUDA(cudaFuncSetAttribute(push_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      ready[dev] = true;
   }
   for (int i = 0; i < a.nseg; i++)
      D2D_REQUIRE(((uintptr_t)a.seg[i].src % 16) == 0 && ((uintptr_t)a.seg[i].dst % 16) == 0 && a.seg[i].bytes % 16 == 0,
                  "push: segments must be 16-byte aligned");
   push_kernel<<<push_ctas(), 32, smem, st>>>(a);
   D2D_CHECK_CUDA(cudaGetLastError());
}

} // namespace d2d
