// Probe for the Ozaki (INT8-slice) FP64 emulation of the posterior GEMM:
//  (1) one tcgen05.mma.kind::i8 (M128 x N x K32, K-major operands, no swizzle) checked against the CPU,
//  (2) issue-rate of back-to-back MMAs on resident shared-memory tiles for N = 64 / 128 / 256.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/i8_mma_probe tools/i8_mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE canonical layout: ((8, m), 2) : ((16 B, SBO), LBO)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // version = 1 (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ uint32_t make_idesc_s8(int M, int N) {
  uint32_t d = 0;
  d |= 2u << 4;                 // c_format = S32
  d |= 1u << 7;                 // a_format = signed 8 bit
  d |= 1u << 10;                // b_format = signed 8 bit
  d |= (uint32_t)(N >> 3) << 17;
  d |= (uint32_t)(M >> 4) << 24;
  return d;                     // K-major A and B, dense, no saturate
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0));
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// A [128 x KTOT] and B [N x KTOT] int8, K-major rows; smem layout per 16-byte K chunk c: rows in groups of 8 x 16 B
template <int N, int KTOT>
__global__ void __launch_bounds__(128) probe_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                    int32_t* __restrict__ out, int iters, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  constexpr int CH = KTOT / 16;              // 16-byte K chunks
  uint8_t* As = smem;                        // [CH][128 rows][16 B]
  uint8_t* Bs = smem + CH * 128 * 16;        // [CH][N rows][16 B]
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int idx = tid; idx < 128 * KTOT; idx += 128) {
    const int r = idx / KTOT, k = idx % KTOT;
    As[(k / 16) * (128 * 16) + r * 16 + (k % 16)] = (uint8_t)A[idx];
  }
  for (int idx = tid; idx < N * KTOT; idx += 128) {
    const int r = idx / KTOT, k = idx % KTOT;
    Bs[(k / 16) * (N * 16) + r * 16 + (k % 16)] = (uint8_t)B[idx];
  }
  if (tid == 0) mbar_init(smem_u32(&mbar), 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(N < 32 ? 32 : N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");   // generic-proxy smem writes -> visible to the MMA
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = make_idesc_s8(128, N);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int ks = 0; ks < KTOT / 32; ++ks) {
        // one MMA consumes 32 bytes of K = chunks 2ks, 2ks+1: LBO = distance between the two chunks, SBO = 8-row groups
        const uint64_t da = make_desc(smem_u32(As + (2 * ks) * (128 * 16)), 128 * 16, 128);
        const uint64_t db = make_desc(smem_u32(Bs + (2 * ks) * (N * 16)), N * 16, 128);
        mma_i8(tmem_base, da, db, idesc, (it > 0 || ks > 0) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&mbar));
  }
  mbar_wait(smem_u32(&mbar), 0);
  if (tid == 0) { t1 = clock64(); if (cycles) *cycles = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w .. 32w+31 (rows), 16 columns at a time
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(size_t)tid * N + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(N < 32 ? 32 : N));
}

__device__ __forceinline__ void mma_i8_ts(uint32_t tmem_c, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_c), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0));
}

// (3) A operand staged in TMEM: tcgen05.cp.128x256b (128 rows x 32 bytes) per K = 32 step, then A-from-TMEM MMAs.
// `reuse` MMAs are issued per copied A tile (the digit-plane GEMM reuses one A plane for up to 7 B planes).
template <int N, int KTOT>
__global__ void __launch_bounds__(128) probe_ts_kernel(const int8_t* __restrict__ A, const int8_t* __restrict__ B,
                                                       int32_t* __restrict__ out, int iters, int reuse, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t mbar;
  constexpr int CH = KTOT / 16;
  uint8_t* As = smem;
  uint8_t* Bs = smem + CH * 128 * 16;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int idx = tid; idx < 128 * KTOT; idx += 128) {
    const int r = idx / KTOT, k = idx % KTOT;
    As[(k / 16) * (128 * 16) + r * 16 + (k % 16)] = (uint8_t)A[idx];
  }
  for (int idx = tid; idx < N * KTOT; idx += 128) {
    const int r = idx / KTOT, k = idx % KTOT;
    Bs[(k / 16) * (N * 16) + r * 16 + (k % 16)] = (uint8_t)B[idx];
  }
  if (tid == 0) mbar_init(smem_u32(&mbar), 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t tmem_a = tmem_base + 448;     // 8 columns (32 bytes per lane) per K = 32 step
  const uint32_t idesc = make_idesc_s8(128, N);
  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      for (int ks = 0; ks < KTOT / 32; ++ks) {
        const uint64_t da = make_desc(smem_u32(As + (2 * ks) * (128 * 16)), 128 * 16, 128);
        const uint64_t db = make_desc(smem_u32(Bs + (2 * ks) * (N * 16)), N * 16, 128);
        asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(tmem_a), "l"(da));
        for (int rpt = 0; rpt < reuse; ++rpt) mma_i8_ts(tmem_base, tmem_a, db, idesc, (it > 0 || ks > 0 || rpt > 0) ? 1u : 0u);
      }
    }
    umma_commit(smem_u32(&mbar));
  }
  mbar_wait(smem_u32(&mbar), 0);
  if (tid == 0) { t1 = clock64(); if (cycles) *cycles = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
    for (int j = 0; j < 16; ++j) out[(size_t)tid * N + c0 + j] = (int32_t)v[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
}

template <int N, int KTOT>
static void run_ts(int iters, int reuse, bool check) {
  std::vector<int8_t> hA(128 * KTOT), hB(N * KTOT);
  srand(2);
  for (auto& v : hA) v = (int8_t)(rand() % 255 - 127);
  for (auto& v : hB) v = (int8_t)(rand() % 255 - 127);
  int8_t *dA, *dB; int32_t* dO; long long* dC;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dO, 128 * N * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  size_t smem = (size_t)(128 + N) * KTOT + 1024;
  CK(cudaFuncSetAttribute(probe_ts_kernel<N, KTOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_ts_kernel<N, KTOT><<<1, 128, smem>>>(dA, dB, dO, iters, reuse, dC);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hO(128 * N);
  long long cyc = 0;
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  if (check) {
    long long bad = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < N; ++j) {
        long long ref = 0;
        for (int k = 0; k < KTOT; ++k) ref += (long long)hA[i * KTOT + k] * hB[j * KTOT + k];
        ref *= (long long)iters * reuse;
        if ((long long)hO[i * N + j] != ref) { if (bad < 5) printf("  TS mismatch (%d,%d): got %d want %lld\n", i, j, hO[i * N + j], ref); ++bad; }
      }
    printf("TS  N=%d K=%d iters=%d reuse=%d: %lld mismatches of %d\n", N, KTOT, iters, reuse, bad, 128 * N);
  } else {
    const double mmas = (double)iters * (KTOT / 32) * reuse;
    printf("TS  N=%3d reuse=%d: %lld cycles for %.0f MMAs -> %.1f clk/MMA\n", N, reuse, cyc, mmas, cyc / mmas);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
}

template <int N, int KTOT>
static void run(int iters, bool check) {
  std::vector<int8_t> hA(128 * KTOT), hB(N * KTOT);
  srand(1);
  for (auto& v : hA) v = (int8_t)(rand() % 127 - 63);
  for (auto& v : hB) v = (int8_t)(rand() % 127 - 63);
  int8_t *dA, *dB; int32_t* dO; long long* dC;
  CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size())); CK(cudaMalloc(&dO, 128 * N * 4)); CK(cudaMalloc(&dC, 8));
  CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
  size_t smem = (size_t)(128 + N) * KTOT + 1024;
  CK(cudaFuncSetAttribute(probe_kernel<N, KTOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  probe_kernel<N, KTOT><<<1, 128, smem>>>(dA, dB, dO, iters, dC);
  CK(cudaDeviceSynchronize());
  std::vector<int32_t> hO(128 * N);
  long long cyc = 0;
  CK(cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
  if (check) {
    long long bad = 0;
    for (int i = 0; i < 128; ++i)
      for (int j = 0; j < N; ++j) {
        long long ref = 0;
        for (int k = 0; k < KTOT; ++k) ref += (long long)hA[i * KTOT + k] * hB[j * KTOT + k];
        ref *= iters;
        if ((long long)hO[i * N + j] != ref) { if (bad < 5) printf("  mismatch (%d,%d): got %d want %lld\n", i, j, hO[i * N + j], ref); ++bad; }
      }
    printf("N=%d K=%d iters=%d: %lld mismatches of %d\n", N, KTOT, iters, bad, 128 * N);
  } else {
    const double mmas = (double)iters * (KTOT / 32);
    const double macs = mmas * 128.0 * N * 32.0;
    printf("N=%3d: %lld cycles for %.0f MMAs -> %.1f clk/MMA, %.0f MAC/clk/SM (x148 SMs x 1.9 GHz x 2 = %.2f POPS)\n", N, cyc, mmas,
           cyc / mmas, macs / cyc, macs / cyc * 148 * 1.9e9 * 2 / 1e15);
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dO); cudaFree(dC);
}

int main() {
  run<64, 64>(1, true);
  run<128, 64>(1, true);
  run<64, 64>(3, true);
  run<64, 128>(2000, false);
  run<128, 128>(2000, false);
  run<256, 128>(2000, false);
  run_ts<64, 64>(1, 1, true);
  run_ts<64, 64>(2, 3, true);
  run_ts<64, 128>(500, 1, false);
  run_ts<64, 128>(500, 4, false);
  run_ts<64, 128>(500, 7, false);
  run_ts<128, 128>(500, 4, false);
  return 0;
}
