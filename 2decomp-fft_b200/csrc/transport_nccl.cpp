// transport_nccl.cpp -- all-to-all(v) over NCCL: grouped ncclSend/ncclRecv with the peers of the
// row / column communicator on ONE global communicator, like decomp_2d_nccl_alltoall_*
// (src/decomp_2d_nccl.f90:214-473) -- but stream-ordered (no cudaStreamSynchronize after the group,
// :249/:324/:395/:470), byte-typed (complex is not re-described as 2x real, :256-273) and without
// the send-to-self (:231-245): the self block never leaves the producer's buffer.
#include <nccl.h>

#include "common.h"

namespace d2d {

#define D2D_CHECK_NCCL(call)                                                                                           \
   do {                                                                                                                \
      ncclResult_t r__ = (call);                                                                                       \
      if (r__ != ncclSuccess)                                                                                          \
         throw ::d2d::Error(3000 + (int)r__, std::string(__FILE__ ":" D2D_STR(__LINE__) " " #call ": ") +              \
                                                 ncclGetErrorString(r__));                                             \
   } while (0)

void nccl_unique_id(unsigned char id[128])
{
   static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
   ncclUniqueId u;
   D2D_CHECK_NCCL(ncclGetUniqueId(&u));
   memcpy(id, &u, 128);
}

namespace {
struct NcclTransport : Transport {
   ncclComm_t comm = nullptr;
   int nranks, rank;
   NcclTransport(const unsigned char id[128], int n, int r) : nranks(n), rank(r)
   {
      ncclUniqueId u;
      memcpy(&u, id, 128);
      D2D_CHECK_NCCL(ncclCommInitRank(&comm, n, u, r));
   }
   ~NcclTransport() override
   {
      if (comm) ncclCommDestroy(comm);
   }
   int kind() const override { return D2D_TRANSPORT_NCCL; }
   void exchange(const std::vector<PeerXfer> &xf, cudaStream_t st) override
   {
      D2D_CHECK_NCCL(ncclGroupStart());
      for (const auto &x : xf) {
         if (x.sendbytes) D2D_CHECK_NCCL(ncclSend(x.sendptr, x.sendbytes, ncclInt8, x.peer, comm, st));
         if (x.recvbytes) D2D_CHECK_NCCL(ncclRecv(x.recvptr, x.recvbytes, ncclInt8, x.peer, comm, st));
      }
      D2D_CHECK_NCCL(ncclGroupEnd());
   }
   void allgather(const void *send_host, void *recv_host, size_t bytes, cudaStream_t st) override
   {
      void *dbuf = nullptr;
      D2D_CHECK_CUDA(cudaMalloc(&dbuf, bytes * (size_t)(nranks + 1)));
      char *dsend = (char *)dbuf + bytes * (size_t)nranks;
      D2D_CHECK_CUDA(cudaMemcpyAsync(dsend, send_host, bytes, cudaMemcpyHostToDevice, st));
      D2D_CHECK_NCCL(ncclAllGather(dsend, dbuf, bytes, ncclInt8, comm, st));
      D2D_CHECK_CUDA(cudaMemcpyAsync(recv_host, dbuf, bytes * (size_t)nranks, cudaMemcpyDeviceToHost, st));
      D2D_CHECK_CUDA(cudaStreamSynchronize(st));
      D2D_CHECK_CUDA(cudaFree(dbuf));
   }
   void barrier(cudaStream_t st) override
   {
      // a 1-element all-reduce is the cheapest stream-ordered barrier NCCL offers
      static thread_local void *buf = nullptr;
      if (!buf) D2D_CHECK_CUDA(cudaMalloc(&buf, 8));
      D2D_CHECK_NCCL(ncclAllReduce(buf, buf, 1, ncclInt32, ncclSum, comm, st));
   }
};
} // namespace

Transport *make_nccl_transport(const unsigned char id[128], int nranks, int rank) { return new NcclTransport(id, nranks, rank); }

} // namespace d2d
