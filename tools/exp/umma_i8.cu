// Experiment (round-2 groundwork, not part of the product): bring-up of the int8 tensor
// path of sm_100a -- tcgen05.mma.kind::i8 with int32 accumulators in TMEM -- as the
// building block of an Ozaki-style (int8 slices, exact int32 accumulation) emulation of
// the FP64 DHT contraction.  Operands are pre-tiled in global memory in the UMMA
// canonical K-major no-swizzle layout (core matrix = 8 rows x 16 bytes, contiguous), so
// a tile is a flat copy into shared memory and no TMA tensor map is needed.
//   step 1: one 128 x 64 tile, K = 64*nkb, checked against the host
//   step 2: throughput of back-to-back MMAs on resident operands
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/exp/umma_i8 tools/exp/umma_i8.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return 1; } } while (0)

constexpr int TM = 128, TN = 64, KB = 64;          // tile rows, tile cols, k bytes per block
constexpr int A_BLK = TM * KB, B_BLK = TN * KB;    // bytes per staged block
constexpr uint32_t A_LBO = (TM / 8) * 128, B_LBO = (TN / 8) * 128, SBO = 128;

// canonical K-major no-swizzle offset of element (row, k) of a [rows x KB] block
__host__ __device__ inline int can_off(int row, int k, int rows) {
  return (k / 16) * (rows / 8) * 128 + (row / 8) * 128 + (row % 8) * 16 + (k % 16);
}

__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;      // descriptor version (Blackwell)
  return d;                    // layout_type = SWIZZLE_NONE (0), base_offset 0
}
// kind::i8, D = s32, A = B = s8, both K-major, M x N
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n"
      :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t mbar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" :: "r"(mbar)
               : "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred P1;\n\tLAB_WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra LAB_DONE_%=;\n\tbra LAB_WAIT_%=;\n\tLAB_DONE_%=:\n\t}\n"
      :: "r"(mbar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, int32_t* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
}

// C[128 x 64] (int32) = A[128 x K] . B[64 x K]^T, operands as nkb canonical blocks each.
// reps > 1: the MMA sequence is repeated on the resident last block (throughput probe).
__global__ void __launch_bounds__(128) umma_tile_kernel(const int8_t* A, const int8_t* B, int32_t* C,
                                                        int nkb, int reps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t mbar_storage;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* sA = smem;
  uint8_t* sB = smem + A_BLK;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(&mbar_storage);

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n"
                 :: "r"((uint32_t)__cvta_generic_to_shared(&tmem_base_smem)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
  }
  if (tid == 0) {
    mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_smem;

  const uint32_t idesc = idesc_i8(TM, TN);
  const uint32_t sA_addr = (uint32_t)__cvta_generic_to_shared(sA);
  const uint32_t sB_addr = (uint32_t)__cvta_generic_to_shared(sB);
  uint32_t phase = 0;
  for (int kb = 0; kb < nkb; ++kb) {
    const uint4* gA = reinterpret_cast<const uint4*>(A + (size_t)kb * A_BLK);
    const uint4* gB = reinterpret_cast<const uint4*>(B + (size_t)kb * B_BLK);
    for (int i = tid; i < A_BLK / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = gA[i];
    for (int i = tid; i < B_BLK / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = gB[i];
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic -> async proxy
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      const int r = (kb == nkb - 1) ? reps : 1;
      for (int rep = 0; rep < r; ++rep) {
#pragma unroll
        for (int j = 0; j < KB / 32; ++j) {       // UMMA_K = 32 bytes = two 16-byte chunks
          const uint64_t ad = smem_desc(sA_addr + j * 2 * A_LBO, A_LBO, SBO);
          const uint64_t bd = smem_desc(sB_addr + j * 2 * B_LBO, B_LBO, SBO);
          umma_i8(tmem_base, ad, bd, idesc, (kb > 0 || j > 0 || rep > 0) ? 1u : 0u);
        }
      }
      umma_commit(mbar);
    }
    mbar_wait(mbar, phase);      // MMAs of this block are done: smem may be overwritten
    phase ^= 1;
  }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31 (= tile rows), 64 columns
  int32_t v[32];
  for (int c0 = 0; c0 < TN; c0 += 32) {
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    if (C) {
#pragma unroll
      for (int c = 0; c < 32; ++c) C[(size_t)(blockIdx.x * TM + tid) * TN + c0 + c] = v[c];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" :: "r"(tmem_base), "r"(64u)
                 : "memory");
}

int main() {
  const int nkb = 8, K = nkb * KB;                 // K = 512
  std::vector<int8_t> a(TM * K), b(TN * K), ac(TM * K), bc(TN * K);
  srand(1);
  for (auto& x : a) x = (int8_t)(rand() % 129 - 64);
  for (auto& x : b) x = (int8_t)(rand() % 129 - 64);
  for (int kb = 0; kb < nkb; ++kb)
    for (int k = 0; k < KB; ++k) {
      for (int r = 0; r < TM; ++r) ac[kb * A_BLK + can_off(r, k, TM)] = a[r * K + kb * KB + k];
      for (int r = 0; r < TN; ++r) bc[kb * B_BLK + can_off(r, k, TN)] = b[r * K + kb * KB + k];
    }
  int8_t *dA, *dB;
  int32_t* dC;
  CK(cudaMalloc(&dA, ac.size()));
  CK(cudaMalloc(&dB, bc.size()));
  CK(cudaMalloc(&dC, (size_t)148 * TM * TN * 4));
  CK(cudaMemcpy(dA, ac.data(), ac.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, bc.data(), bc.size(), cudaMemcpyHostToDevice));
  CK(cudaMemset(dC, 0xff, (size_t)TM * TN * 4));
  const int smem_bytes = A_BLK + B_BLK;
  CK(cudaFuncSetAttribute(umma_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
  umma_tile_kernel<<<1, 128, smem_bytes>>>(dA, dB, dC, nkb, 1);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> c(TM * TN);
  CK(cudaMemcpy(c.data(), dC, c.size() * 4, cudaMemcpyDeviceToHost));
  long bad = 0;
  for (int i = 0; i < TM; ++i)
    for (int j = 0; j < TN; ++j) {
      int32_t ref = 0;
      for (int k = 0; k < K; ++k) ref += (int32_t)a[i * K + k] * (int32_t)b[j * K + k];
      if (ref != c[i * TN + j]) {
        if (bad < 5) printf("mismatch (%d,%d): got %d want %d\n", i, j, c[i * TN + j], ref);
        ++bad;
      }
    }
  printf("step 1: 128x64x%d int8 tile on tcgen05.mma.kind::i8: %ld mismatches of %d\n", K, bad, TM * TN);

  // step 2: issue-rate probe, one CTA per SM, MMAs repeated on resident operands
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int reps : {2000, 20000}) {
    float ms = 0;
    for (int it = 0; it < 2; ++it) {
      cudaEventRecord(e0);
      umma_tile_kernel<<<148, 128, smem_bytes>>>(dA, dB, nullptr, 1, reps);
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      cudaEventElapsedTime(&ms, e0, e1);
    }
    const double ops = 148.0 * reps * (KB / 32) * (2.0 * TM * TN * 32);
    printf("step 2: reps %d: %.3f ms  %.1f TOPS (148 CTAs x 1 issuing thread, M=128 N=64 K=32)\n", reps, ms,
           ops / ms / 1e9);
  }
  printf("err %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
