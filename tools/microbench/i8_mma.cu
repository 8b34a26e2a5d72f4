// tcgen05.mma kind::i8 on B200 (sm_100a): correctness of the hand-built shared-memory / instruction descriptors against a CPU
// int32 reference, and the sustained issue rate of int8 MMAs (M = 128, N = 64 / 128 / 256, K = 32 per instruction) with operands
// in the canonical K-major no-swizzle ("interleaved") shared-memory layout -- the building block of an integer-slice (Ozaki)
// FP64 trailing update.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o i8_mma i8_mma.cu
//
// Operand tile in shared memory (R rows x 128 K-bytes, int8, K-major): core matrix = 8 rows x 16 bytes = 128 contiguous bytes;
//     offset(r, k) = (k / 16) * (R / 8 * 128) + (r / 8) * 128 + (r % 8) * 16 + (k % 16)
// i.e. SBO (stride between 8-row groups) = 128 B, LBO (stride between 16-byte K chunks) = R * 16 B.  One MMA consumes two K
// chunks (K = 32); the next K step advances the descriptor's start address by 2 * LBO.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) {                                                                   \
      printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);                  \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
  // c_format S32 = 2 @ [4,6); a_format signed = 1 @ [7,10); b_format signed = 1 @ [10,13); K-major both; n_dim = N >> 3 @ [17,23);
  // m_dim = M >> 4 @ [24,29)
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint64_t make_sdesc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version 1 (sm_100)
  return d;                // layout_type 0 = no swizzle, base_offset 0
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(done)
        : "r"(s_u32(bar)), "r"(parity)
        : "memory");
    if (!done && ++spins > (1u << 26)) __trap();
  }
}

template <int N>
__global__ void __launch_bounds__(128, 1) i8_mma_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D,
                                                        int reps, int nacc, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;               // 128 x 128 bytes
  uint8_t* sB = smem + 128 * 128;   // N x 128 bytes
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 128 / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(A)[i];
  for (int i = tid; i < N * 128 / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(B)[i];
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor-core (async) proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc = make_idesc_i8(128, N);
  const uint32_t a0 = s_u32(sA), b0 = s_u32(sB);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t dcol = (uint32_t)((r % nacc) * N);  // accumulator r % nacc
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = make_sdesc(a0 + ks * 2 * (128 * 16), 128 * 16, 128);
        const uint64_t bd = make_sdesc(b0 + ks * 2 * (N * 16), N * 16, 128);
        mma_i8(tmem + dcol, ad, bd, idesc, (r >= nacc || ks > 0) ? 1u : 0u);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) {
    t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // accumulator 0 -> global (block 0 only): warp w reads TMEM lanes 32w .. 32w+31 (= rows), 32 columns per load
  if (D && blockIdx.x == 0) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
          "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
            "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
            "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
            "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- cta_group::2: one MMA spans a CTA pair (M = 256: each CTA supplies its own 128 rows of A and HALF of B's N rows, and holds
// the accumulator rows of its own A rows in its own TMEM).  Per SM and MMA that is 4 KB of A + 2 KB of B from shared memory
// instead of 8 KB.  Only the leader (cluster rank 0) issues; the commit is multicast to the barrier of both CTAs.
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mma_i8_2cta(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_2cta(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(s_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

// A: [2 CTAs][128 rows x 128 K] canonical; B: N = 128 rows x 128 K canonical [chunk][128 rows][16 B]; CTA r stages rows 64r .. 64r+63 of B as
// [chunk (8)][64 rows][16 B] (LBO = 1024).  D: [256 x 128] int32.
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1)
    i8_mma_2cta_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D, int reps, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const uint32_t rank = cluster_rank();
  uint8_t* sA = smem;              // 128 x 128 bytes (own rows)
  uint8_t* sB = smem + 128 * 128;  // 64 x 128 bytes (own half of N)
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 128 / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(A + (size_t)rank * 128 * 128)[i];
  for (int i = tid; i < 64 * 128 / 16; i += 128) {  // unit i = (chunk, row in half): source [chunk][128 rows][16]
    const int chunk = i / 64, row = i % 64;
    reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(B)[chunk * 128 + rank * 64 + row];
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();  // both CTAs' operands staged, barriers initialised, TMEM allocated
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc = make_idesc_i8(256, 128);
  const uint32_t a0 = s_u32(sA), b0 = s_u32(sB);
  long long t0 = 0, t1 = 0;
  if (rank == 0 && tid == 0) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t dcol = (uint32_t)((r % 4) * 128);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = make_sdesc(a0 + ks * 2 * (128 * 16), 128 * 16, 128);
        const uint64_t bd = make_sdesc(b0 + ks * 2 * (64 * 16), 64 * 16, 128);
        mma_i8_2cta(tmem + dcol, ad, bd, idesc, (r >= 4 || ks > 0) ? 1u : 0u);
      }
    }
    mma_commit_2cta(&bar);
  }
  mbar_wait(&bar, 0);
  if (rank == 0 && tid == 0) {
    t1 = clock64();
    if (cycles) cycles[blockIdx.x / 2] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (D && blockIdx.x < 2) {
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
          "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
            "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
            "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
            "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[((size_t)rank * 128 + tid) * 128 + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- A operand from TENSOR MEMORY (tcgen05.mma ... [d], [a_tmem], b_desc): the A planes of a stage are copied shared -> tensor memory
// once (tcgen05.cp 128x256b: 128 rows x 32 bytes = one K = 32 operand, 8 TMEM columns) and every MMA then reads only B from shared
// memory -- 4 KB instead of 8 KB per MMA.  Checks the copy + the TS form against the CPU and times the TS issue rate.
__device__ __forceinline__ void cp_128x256b(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__global__ void __launch_bounds__(128, 1) i8_mma_ts_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D,
                                                           int reps, int recopy, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 128 / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(A)[i];
  for (int i = tid; i < 128 * 128 / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(B)[i];
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc = make_idesc_i8(128, 128);
  const uint32_t a0 = s_u32(sA), b0 = s_u32(sB);
  const uint32_t acol = 384;  // A staging area: columns 384 .. 415 (4 K-steps x 8 columns)
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (r == 0 || recopy) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) cp_128x256b(tmem + acol + 8 * ks, make_sdesc(a0 + ks * 2 * (128 * 16), 128 * 16, 128));
      }
      const uint32_t dcol = (uint32_t)((r % 3) * 128);
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        mma_i8_ts(tmem + dcol, tmem + acol + 8 * ks, make_sdesc(b0 + ks * 2 * (128 * 16), 128 * 16, 128), idesc, (r >= 3 || ks > 0) ? 1u : 0u);
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) {
    t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (D && blockIdx.x == 0) {
    for (int c0 = 0; c0 < 128; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
          "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
            "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
            "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
            "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[(size_t)tid * 128 + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---- collector reuse of the A operand: consecutive MMAs with the SAME A (different B, different accumulators) issued as
// .collector::a::fill / ::use / ::lastuse read A from shared memory once per group instead of once per MMA (SASS: A_KEEP / A_REUSE).
// acc g (g = 0..2) = sum_ks A_ks * B_{(ks+g) % 4}' -- checked against the CPU; rate with and without the reuse.
template <int MODE>  // 0: plain, 1: fill, 2: use, 3: lastuse
__device__ __forceinline__ void mma_i8_coll(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  if constexpr (MODE == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::fill [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else if constexpr (MODE == 2)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::use [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else if constexpr (MODE == 3)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8.collector::a::lastuse [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
  else
    mma_i8(tmem_d, adesc, bdesc, idesc, accumulate);
}
template <int REUSE>
__global__ void __launch_bounds__(128, 1) i8_mma_coll_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int32_t* __restrict__ D,
                                                             int reps, long long* __restrict__ cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* sA = smem;
  uint8_t* sB = smem + 128 * 128;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 128 / 16; i += 128) reinterpret_cast<uint4*>(sA)[i] = reinterpret_cast<const uint4*>(A)[i];
  for (int i = tid; i < 128 * 128 / 16; i += 128) reinterpret_cast<uint4*>(sB)[i] = reinterpret_cast<const uint4*>(B)[i];
  if (tid == 0) {
    mbar_init(&bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  constexpr uint32_t idesc = make_idesc_i8(128, 128);
  const uint32_t a0 = s_u32(sA), b0 = s_u32(sB);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        const uint64_t ad = make_sdesc(a0 + ks * 4096, 2048, 128);
        const uint32_t acc = (r > 0 || ks > 0) ? 1u : 0u;
        mma_i8_coll<REUSE ? 1 : 0>(tmem + 0, ad, make_sdesc(b0 + ((ks + 0) & 3) * 4096, 2048, 128), idesc, acc);
        mma_i8_coll<REUSE ? 2 : 0>(tmem + 128, ad, make_sdesc(b0 + ((ks + 1) & 3) * 4096, 2048, 128), idesc, acc);
        mma_i8_coll<REUSE ? 3 : 0>(tmem + 256, ad, make_sdesc(b0 + ((ks + 2) & 3) * 4096, 2048, 128), idesc, acc);
      }
    }
    mma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  if (tid == 0) {
    t1 = clock64();
    if (cycles) cycles[blockIdx.x] = t1 - t0;
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  if (D && blockIdx.x == 0) {
    for (int c0 = 0; c0 < 384; c0 += 32) {
      uint32_t v[32];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, "
          "%20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
            "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
            "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
            "=r"(v[31])
          : "r"(taddr)
          : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      for (int j = 0; j < 32; ++j) D[(size_t)tid * 384 + c0 + j] = (int32_t)v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

static size_t canon(int R, int r, int k) { return (size_t)(k / 16) * (R / 8 * 128) + (size_t)(r / 8) * 128 + (r % 8) * 16 + (k % 16); }

template <int N>
static int run(int sms, double clock_ghz) {
  std::vector<int8_t> hA(128 * 128), hB((size_t)N * 128), cA(128 * 128), cB((size_t)N * 128);
  srand(7 + N);
  for (auto& v : hA) v = (int8_t)(rand() % 129 - 64);
  for (auto& v : hB) v = (int8_t)(rand() % 129 - 64);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < 128; ++k) cA[canon(128, r, k)] = hA[r * 128 + k];
  for (int r = 0; r < N; ++r)
    for (int k = 0; k < 128; ++k) cB[canon(N, r, k)] = hB[r * 128 + k];
  int8_t *dA, *dB;
  int32_t* dD;
  long long* dC;
  CK(cudaMalloc(&dA, cA.size()));
  CK(cudaMalloc(&dB, cB.size()));
  CK(cudaMalloc(&dD, (size_t)128 * N * 4));
  CK(cudaMalloc(&dC, 1024 * sizeof(long long)));
  CK(cudaMemcpy(dA, cA.data(), cA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, cB.data(), cB.size(), cudaMemcpyHostToDevice));
  const size_t smem = 128 * 128 + (size_t)N * 128;
  CK(cudaFuncSetAttribute(i8_mma_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // ---- correctness: one pass (4 MMAs, K = 128) into accumulator 0
  i8_mma_kernel<N><<<1, 128, smem>>>(dA, dB, dD, 1, 1, dC);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hD((size_t)128 * N);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < N; ++j) {
      int32_t s = 0;
      for (int k = 0; k < 128; ++k) s += (int32_t)hA[i * 128 + k] * (int32_t)hB[j * 128 + k];
      if (s != hD[(size_t)i * N + j]) {
        if (bad < 4) printf("  mismatch N=%d (%d,%d): got %d want %d\n", N, i, j, hD[(size_t)i * N + j], s);
        ++bad;
      }
    }
  printf("{\"test\": \"i8_mma_correct\", \"M\": 128, \"N\": %d, \"K\": 128, \"mismatches\": %lld}\n", N, bad);
  // ---- accumulate check: 3 passes into one accumulator = 3x
  i8_mma_kernel<N><<<1, 128, smem>>>(dA, dB, dD, 3, 1, dC);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hD3((size_t)128 * N);
  CK(cudaMemcpy(hD3.data(), dD, hD3.size() * 4, cudaMemcpyDeviceToHost));
  long long bad3 = 0;
  for (size_t i = 0; i < hD.size(); ++i) bad3 += (hD3[i] != 3 * hD[i]);
  printf("{\"test\": \"i8_mma_accumulate\", \"N\": %d, \"mismatches\": %lld}\n", N, bad3);
  // ---- rate: reps passes of 4 MMAs round-robin over nacc accumulators; 1 CTA alone, then one CTA per SM
  const int nacc = 512 / N;
  for (int grid : {1, sms}) {
    const int reps = 4096;
    i8_mma_kernel<N><<<grid, 128, smem>>>(dA, dB, nullptr, 64, nacc, dC);  // warm-up
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    i8_mma_kernel<N><<<grid, 128, smem>>>(dA, dB, nullptr, reps, nacc, dC);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> hc(grid);
    CK(cudaMemcpy(hc.data(), dC, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long cmax = 0;
    for (long long c : hc) cmax = c > cmax ? c : cmax;
    const double macs = (double)reps * 4 * 128.0 * N * 32;
    printf("{\"test\": \"i8_mma_rate\", \"N\": %d, \"ctas\": %d, \"mma_per_cta\": %d, \"cycles_per_mma\": %.2f, \"mac_per_clk_per_sm\": %.1f, "
           "\"smem_bytes_per_clk\": %.1f, \"ms\": %.3f, \"total_tops\": %.1f}\n",
           N, grid, reps * 4, (double)cmax / (reps * 4), macs / (double)cmax, (double)reps * 4 * (128 + N) * 32 / (double)cmax, ms,
           2.0 * macs * grid / (ms * 1e-3) / 1e12);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad == 0 && bad3 == 0;
}

static int run_2cta(int sms) {
  std::vector<int8_t> hA(256 * 128), hB(128 * 128), cA(256 * 128), cB(128 * 128);
  srand(99);
  for (auto& v : hA) v = (int8_t)(rand() % 129 - 64);
  for (auto& v : hB) v = (int8_t)(rand() % 129 - 64);
  for (int h = 0; h < 2; ++h)
    for (int r = 0; r < 128; ++r)
      for (int k = 0; k < 128; ++k) cA[(size_t)h * 128 * 128 + canon(128, r, k)] = hA[(h * 128 + r) * 128 + k];
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < 128; ++k) cB[canon(128, r, k)] = hB[r * 128 + k];
  int8_t *dA, *dB;
  int32_t* dD;
  long long* dC;
  CK(cudaMalloc(&dA, cA.size()));
  CK(cudaMalloc(&dB, cB.size()));
  CK(cudaMalloc(&dD, (size_t)256 * 128 * 4));
  CK(cudaMalloc(&dC, 1024 * sizeof(long long)));
  CK(cudaMemcpy(dA, cA.data(), cA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, cB.data(), cB.size(), cudaMemcpyHostToDevice));
  const size_t smem = 128 * 128 + 64 * 128;
  i8_mma_2cta_kernel<<<2, 128, smem>>>(dA, dB, dD, 1, dC);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hD((size_t)256 * 128);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int i = 0; i < 256; ++i)
    for (int j = 0; j < 128; ++j) {
      int32_t s = 0;
      for (int k = 0; k < 128; ++k) s += (int32_t)hA[i * 128 + k] * (int32_t)hB[j * 128 + k];
      if (s != hD[(size_t)i * 128 + j]) {
        if (bad < 6) printf("  2cta mismatch (%d,%d): got %d want %d\n", i, j, hD[(size_t)i * 128 + j], s);
        ++bad;
      }
    }
  printf("{\"test\": \"i8_mma_2cta_correct\", \"M\": 256, \"N\": 128, \"K\": 128, \"mismatches\": %lld}\n", bad);
  for (int pairs : {1, sms / 2}) {
    const int reps = 4096;
    i8_mma_2cta_kernel<<<2 * pairs, 128, smem>>>(dA, dB, nullptr, 64, dC);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    i8_mma_2cta_kernel<<<2 * pairs, 128, smem>>>(dA, dB, nullptr, reps, dC);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> hc(pairs);
    CK(cudaMemcpy(hc.data(), dC, pairs * sizeof(long long), cudaMemcpyDeviceToHost));
    long long cmax = 0;
    for (long long c : hc) cmax = c > cmax ? c : cmax;
    const double macs = (double)reps * 4 * 256.0 * 128 * 32;
    printf("{\"test\": \"i8_mma_2cta_rate\", \"M\": 256, \"N\": 128, \"cta_pairs\": %d, \"cycles_per_mma\": %.2f, \"mac_per_clk_per_sm\": %.1f, "
           "\"smem_bytes_per_clk_per_sm\": %.1f, \"ms\": %.3f, \"total_tops\": %.1f}\n",
           pairs, (double)cmax / (reps * 4), macs / 2 / (double)cmax, (double)reps * 4 * (128 + 64) * 32 / (double)cmax, ms,
           2.0 * macs * pairs / (ms * 1e-3) / 1e12);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad == 0;
}

static int run_ts(int sms) {
  std::vector<int8_t> hA(128 * 128), hB(128 * 128), cA(128 * 128), cB(128 * 128);
  srand(1234);
  for (auto& v : hA) v = (int8_t)(rand() % 256 - 128);
  for (auto& v : hB) v = (int8_t)(rand() % 256 - 128);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < 128; ++k) {
      cA[canon(128, r, k)] = hA[r * 128 + k];
      cB[canon(128, r, k)] = hB[r * 128 + k];
    }
  int8_t *dA, *dB;
  int32_t* dD;
  long long* dC;
  CK(cudaMalloc(&dA, cA.size()));
  CK(cudaMalloc(&dB, cB.size()));
  CK(cudaMalloc(&dD, (size_t)128 * 128 * 4));
  CK(cudaMalloc(&dC, 1024 * sizeof(long long)));
  CK(cudaMemcpy(dA, cA.data(), cA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, cB.data(), cB.size(), cudaMemcpyHostToDevice));
  const size_t smem = 2 * 128 * 128;
  i8_mma_ts_kernel<<<1, 128, smem>>>(dA, dB, dD, 1, 1, dC);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hD((size_t)128 * 128);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int i = 0; i < 128; ++i)
    for (int j = 0; j < 128; ++j) {
      int32_t s = 0;
      for (int k = 0; k < 128; ++k) s += (int32_t)hA[i * 128 + k] * (int32_t)hB[j * 128 + k];
      if (s != hD[(size_t)i * 128 + j]) {
        if (bad < 6) printf("  TS mismatch (%d,%d): got %d want %d\n", i, j, hD[(size_t)i * 128 + j], s);
        ++bad;
      }
    }
  printf("{\"test\": \"i8_mma_ts_correct\", \"what\": \"tcgen05.cp 128x256b smem -> TMEM, then tcgen05.mma with A from TMEM\", \"mismatches\": %lld}\n", bad);
  for (int recopy : {0, 1})
    for (int grid : {1, sms}) {
      const int reps = 4096;
      i8_mma_ts_kernel<<<grid, 128, smem>>>(dA, dB, nullptr, 64, recopy, dC);
      CK(cudaDeviceSynchronize());
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0));
      CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      i8_mma_ts_kernel<<<grid, 128, smem>>>(dA, dB, nullptr, reps, recopy, dC);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<long long> hc(grid);
      CK(cudaMemcpy(hc.data(), dC, grid * sizeof(long long), cudaMemcpyDeviceToHost));
      long long cmax = 0;
      for (long long c : hc) cmax = c > cmax ? c : cmax;
      printf("{\"test\": \"i8_mma_ts_rate\", \"copy_A_every_pass\": %d, \"ctas\": %d, \"cycles_per_mma\": %.2f, \"mac_per_clk_per_sm\": %.1f, \"ms\": %.3f}\n",
             recopy, grid, (double)cmax / (reps * 4), (double)reps * 4 * 128.0 * 128 * 32 / (double)cmax, ms);
    }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad == 0;
}

template <int REUSE>
static int run_coll(int sms) {
  std::vector<int8_t> hA(128 * 128), hB(128 * 128), cA(128 * 128), cB(128 * 128);
  srand(4321);
  for (auto& v : hA) v = (int8_t)(rand() % 256 - 128);
  for (auto& v : hB) v = (int8_t)(rand() % 256 - 128);
  for (int r = 0; r < 128; ++r)
    for (int k = 0; k < 128; ++k) {
      cA[canon(128, r, k)] = hA[r * 128 + k];
      cB[canon(128, r, k)] = hB[r * 128 + k];
    }
  int8_t *dA, *dB;
  int32_t* dD;
  long long* dC;
  CK(cudaMalloc(&dA, cA.size()));
  CK(cudaMalloc(&dB, cB.size()));
  CK(cudaMalloc(&dD, (size_t)128 * 384 * 4));
  CK(cudaMalloc(&dC, 1024 * sizeof(long long)));
  CK(cudaMemcpy(dA, cA.data(), cA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, cB.data(), cB.size(), cudaMemcpyHostToDevice));
  const size_t smem = 2 * 128 * 128;
  i8_mma_coll_kernel<REUSE><<<1, 128, smem>>>(dA, dB, dD, 1, dC);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hD((size_t)128 * 384);
  CK(cudaMemcpy(hD.data(), dD, hD.size() * 4, cudaMemcpyDeviceToHost));
  long long bad = 0;
  for (int g = 0; g < 3; ++g)
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < 128; ++j) {
        int32_t s = 0;
        for (int ks = 0; ks < 4; ++ks)
          for (int kk = 0; kk < 32; ++kk) s += (int32_t)hA[i * 128 + ks * 32 + kk] * (int32_t)hB[j * 128 + ((ks + g) & 3) * 32 + kk];
        if (s != hD[(size_t)i * 384 + g * 128 + j]) {
          if (bad < 6) printf("  collector mismatch acc %d (%d,%d): got %d want %d\n", g, i, j, hD[(size_t)i * 384 + g * 128 + j], s);
          ++bad;
        }
      }
  printf("{\"test\": \"i8_mma_collector_correct\", \"a_reuse\": %d, \"mismatches\": %lld}\n", REUSE, bad);
  for (int grid : {1, sms}) {
    const int reps = 2048;
    i8_mma_coll_kernel<REUSE><<<grid, 128, smem>>>(dA, dB, nullptr, 64, dC);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaEventRecord(e0));
    i8_mma_coll_kernel<REUSE><<<grid, 128, smem>>>(dA, dB, nullptr, reps, dC);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> hc(grid);
    CK(cudaMemcpy(hc.data(), dC, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long cmax = 0;
    for (long long c : hc) cmax = c > cmax ? c : cmax;
    printf("{\"test\": \"i8_mma_collector_rate\", \"a_reuse\": %d, \"group\": 3, \"ctas\": %d, \"cycles_per_mma\": %.2f, \"mac_per_clk_per_sm\": %.1f, \"ms\": %.3f}\n", REUSE,
           grid, (double)cmax / (reps * 12), (double)reps * 12 * 128.0 * 128 * 32 / (double)cmax, ms);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dC);
  return bad == 0;
}

int main() {
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\"}\n", pr.name, pr.multiProcessorCount, pr.major, pr.minor);
  int ok = 1;
  ok &= run<64>(pr.multiProcessorCount, 0);
  ok &= run<128>(pr.multiProcessorCount, 0);
  ok &= run<256>(pr.multiProcessorCount, 0);
  ok &= run_2cta(pr.multiProcessorCount);
  ok &= run_ts(pr.multiProcessorCount);
  ok &= run_coll<0>(pr.multiProcessorCount);
  ok &= run_coll<1>(pr.multiProcessorCount);
  printf("{\"all_correct\": %s}\n", ok ? "true" : "false");
  return ok ? 0 : 1;
}
