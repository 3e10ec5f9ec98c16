// chimera-b200: sum of the ranks' partial backward transforms over NVLink peer memory.
//
// The kr-row sharded field solve leaves, on every rank, a partial
// sum of the E (or B) grid arrays in one flat FP64 buffer.  With those buffers allocated
// as symmetric memory (torch.distributed._symmetric_memory: every rank knows every rank's
// buffer address, and on NVSwitch a multicast address that maps all of them) the sum is
// ONE kernel per field instead of an NCCL all-reduce: rank r owns the r-th 1/world block
// of the buffer, reads that block from every rank, adds, and stores the total back into
// every rank's buffer --
//   * peer mode:      `world` 16-byte loads and `world` 16-byte stores per element pair
//                     through the peers' addresses (P2P over NVLink),
//   * multicast mode: one multimem.ld_reduce (the switch adds the `world` copies) and one
//                     multimem.st (the switch fans the result out) per element.
// The caller brackets the launch with two cross-rank barriers on the same stream (all
// partials written / all totals stored): the symmetric-memory handle's barrier().
// No reference counterpart (the reference is single-device).
//
// STATUS: opt-in (CHB_PEER_EXCHANGE=1 / multimem).  Run on 2 and 8 B200 in round 2 (parity
// 1.5e-14 / 8.4e-14 against the replicated solve); NCCL stays the default, see DESIGN.md 5.
#include <cstdlib>
#include "common.cuh"
#include "../../include/chimera_b200.h"

namespace chb {

constexpr int kMaxPeers = 16;

struct PeerBufs {
  double* p[kMaxPeers];
};

// block [begin, end) of this rank, in doubles; begin/end even (16-byte aligned pairs)
__global__ void __launch_bounds__(512)
peer_allreduce_kernel(const __grid_constant__ PeerBufs bufs, int world, size_t begin, size_t end) {
  const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
  for (size_t i = begin + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < end;
       i += stride) {
    double2 v[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < world) v[r] = __ldcg(reinterpret_cast<const double2*>(bufs.p[r] + i));
    double2 s = v[0];
#pragma unroll
    for (int r = 1; r < kMaxPeers; ++r)
      if (r < world) { s.x += v[r].x; s.y += v[r].y; }      // rank order: same bits everywhere
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < world) *reinterpret_cast<double2*>(bufs.p[r] + i) = s;
  }
}

// (multimem.ld_reduce / .st take no vector form for .f64 on sm_100a -- ptxas rejects .v2.f64 --
// so every thread keeps sixteen independent 8-byte reductions in flight instead)
__global__ void __launch_bounds__(512)
multimem_allreduce_kernel(double* mc, size_t begin, size_t end) {
  constexpr int U = 16;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < end;
       i0 += stride * U) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * stride;
      if (i < end)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f64 %0, [%1];"
                     : "=d"(v[u]) : "l"(mc + i) : "memory");
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = i0 + u * stride;
      if (i < end)
        asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc + i), "d"(v[u])
                     : "memory");
    }
  }
}

// all-gather of row blocks: every rank copies ITS block [begin, end) of the local buffer to
// the same place in every other rank's buffer (16-byte P2P stores) ...
__global__ void __launch_bounds__(512)
peer_push_kernel(const __grid_constant__ PeerBufs bufs, int world, int rank, size_t begin,
                 size_t end) {
  const double* __restrict__ src = bufs.p[rank];
  const size_t stride = (size_t)gridDim.x * blockDim.x * 2;
  for (size_t i = begin + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < end;
       i += stride) {
    const double2 v = *reinterpret_cast<const double2*>(src + i);
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < world && r != rank) *reinterpret_cast<double2*>(bufs.p[r] + i) = v;
  }
}

// ... or stores it once through the multicast address (the switch fans it out)
__global__ void __launch_bounds__(512)
multimem_push_kernel(const double* __restrict__ src, double* mc, size_t begin, size_t end) {
  constexpr int U = 8;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = begin + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < end;
       i0 += stride * U) {
    double v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * stride < end) v[u] = src[i0 + u * stride];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i0 + u * stride < end)
        asm volatile("multimem.st.relaxed.sys.global.f64 [%0], %1;" ::"l"(mc + i0 + u * stride),
                     "d"(v[u]) : "memory");
  }
}

}  // namespace chb

using namespace chb;

// CTAs of the exchange kernels.  They run on their own stream next to the step's compute
// kernels: a grid that fills the GPU (the first version: 4 x 148 CTAs) is no faster -- the
// NVLink ports, not the SMs, bound it -- and starves the compute kernels it is meant to
// overlap (8-GPU timeline, profiles/r2_trace_8gpu_multimem_v2.txt: a 24 us contraction
// took 347 us next to it).  CHB_PEER_CTAS overrides.
static int peer_ctas() {
  static const int n = [] { const char* e = getenv("CHB_PEER_CTAS");
                            const int v = e ? atoi(e) : 32;
                            return v < 1 ? 1 : (v > 8 * kSMs ? 8 * kSMs : v); }();
  return n;
}

extern "C" int chb_peer_allgather_f64(const uint64_t* peer_ptrs_host, int world, int rank,
                                      uint64_t multicast_ptr, size_t begin, size_t count,
                                      void* stream) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !peer_ptrs_host)
    return CHB_ERR_ARG;
  if (count == 0 || world == 1) return CHB_OK;
  const int grid = peer_ctas();
  if (multicast_ptr) {
    multimem_push_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double*>(peer_ptrs_host[rank]),
        reinterpret_cast<double*>(multicast_ptr), begin, begin + count);
    CHB_RETURN_LAST_ERROR();
  }
  if ((begin & 1) || (count & 1)) return CHB_ERR_ARG;
  PeerBufs bufs;
  for (int r = 0; r < kMaxPeers; ++r) {
    bufs.p[r] = r < world ? reinterpret_cast<double*>(peer_ptrs_host[r]) : nullptr;
    if (r < world && (peer_ptrs_host[r] & 15)) return CHB_ERR_ARG;
  }
  peer_push_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(bufs, world, rank, begin, begin + count);
  CHB_RETURN_LAST_ERROR();
}

extern "C" int chb_peer_allreduce_f64(const uint64_t* peer_ptrs_host, int world, int rank,
                                      uint64_t multicast_ptr, size_t n, void* stream) {
  if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world || !peer_ptrs_host)
    return CHB_ERR_ARG;
  if (n == 0 || world == 1) return CHB_OK;
  // equal blocks of an even number of doubles; the last rank takes the remainder
  size_t per = (n / world) & ~(size_t)1;
  const size_t begin = per * rank;
  const size_t end = rank == world - 1 ? n : begin + per;
  if (end <= begin) return CHB_OK;
  const int grid = peer_ctas();
  if (multicast_ptr) {
    multimem_allreduce_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<double*>(multicast_ptr), begin, end);
    CHB_RETURN_LAST_ERROR();
  }
  if ((n & 1) || (end & 1)) return CHB_ERR_ARG;     // pairs: the flat buffers are even-sized
  PeerBufs bufs;
  for (int r = 0; r < kMaxPeers; ++r) {
    bufs.p[r] = r < world ? reinterpret_cast<double*>(peer_ptrs_host[r]) : nullptr;
    if (r < world && (peer_ptrs_host[r] & 15)) return CHB_ERR_ARG;
  }
  peer_allreduce_kernel<<<grid, 512, 0, (cudaStream_t)stream>>>(bufs, world, begin, end);
  CHB_RETURN_LAST_ERROR();
}
