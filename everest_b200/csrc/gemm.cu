// K3 (GEMM part): FP64 tensor-pipe (DMMA m8n8k4) "NT" GEMMs.
//
//   gemm_nt_kernel        C[m,n] = alpha * sum_k A[m,k] B[n,k] + beta * C        (set-up work:
//                         blocked Cholesky updates, inverse root, baseline / pruning posteriors)
//   posterior_gemm_kernel the hot loop of AcquisitionFunction.forward: V = K(X*,X) L^-T against
//                         the cached LOWER-TRIANGULAR inverse root, fused with the reductions the
//                         posterior needs from V, so V itself never reaches HBM:
//                           Gqq[batch] = V_q V_q^T   (q x q block of K*X (K+s2 I)^-1 KX*)
//                           W          = V_q V_b^T   (cross block against the cached baseline rows)
//                           mu_raw     = K*X alpha
//
// This is what gpytorch's DefaultPredictionStrategy.exact_predictive_covar / exact_predictive_mean
// compute on CPU with MKL (reached from model.posterior, reference botorch.py:180,223):
// test_train_covar @ L^-T, then (.) @ (.)^T.  SURVEY.md 2.3 K3 / 8a rows a8.
//
// Both operands are K-contiguous ("NT"), so one 16-byte shared-memory load feeds TWO DMMAs
// (k = 2t and 2t+1 of an 8-wide k group; A and B use the same k permutation).  Tiles are staged
// with a 3-deep cp.async (LDGSTS) pipeline; row stride 24 doubles (192 B) makes every quarter-warp
// 128-bit fragment load hit 8 distinct 16-byte bank groups.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

#define GK 16       // k per pipeline stage
#define GLDS 24     // smem row stride in doubles

template <int ROWS, int NTHREADS>
__device__ __forceinline__ void load_tile_async(double* s, const double* __restrict__ g, int ld, int row0,
                                                int rows_valid, int k0, int tid) {
#pragma unroll
  for (int c = tid; c < ROWS * 8; c += NTHREADS) {
    int r = c >> 3, ch = c & 7;
    bool p = (row0 + r) < rows_valid;
    const double* src = g + (size_t)(p ? (row0 + r) : 0) * ld + k0 + ch * 2;
    cp_async16(s + r * GLDS + ch * 2, src, p);
  }
}

template <int TM, int TN>
__device__ __forceinline__ void mma_stage(const double* As, const double* Bs, int wm0, int wn0, int g, int t,
                                          double (&acc)[TM][TN][2]) {
#pragma unroll
  for (int kg = 0; kg < 2; ++kg) {
    double2 af[TM], bf[TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
      af[i] = *reinterpret_cast<const double2*>(As + (wm0 + i * 8 + g) * GLDS + kg * 8 + 2 * t);
#pragma unroll
    for (int j = 0; j < TN; ++j)
      bf[j] = *reinterpret_cast<const double2*>(Bs + (wn0 + j * 8 + g) * GLDS + kg * 8 + 2 * t);
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        mma_884(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
        mma_884(acc[i][j][0], acc[i][j][1], af[i].y, bf[j].y);
      }
  }
}

// ------------------------------------------------------------------------------------------------
// generic NT GEMM: 64 x 64 tile, 4 warps (2 x 2), warp tile 32 x 32
// ------------------------------------------------------------------------------------------------
#define G1_BM 64
#define G1_BN 64
#define G1_ST 3

__global__ void __launch_bounds__(128)
gemm_nt_kernel(int m, int n, int k, double alpha, const double* __restrict__ A, int lda,
               const double* __restrict__ B, int ldb, double beta, double* __restrict__ C, int ldc, int lower_only) {
  extern __shared__ __align__(16) double gsm[];
  const int m0 = blockIdx.y * G1_BM, n0 = blockIdx.x * G1_BN;
  if (lower_only && n0 > m0 + G1_BM - 1) return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm0 = (warp & 1) * 32, wn0 = (warp >> 1) * 32;
  double* As = gsm;
  double* Bs = gsm + G1_ST * G1_BM * GLDS;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  const int nk = (k + GK - 1) / GK;
  for (int s = 0; s < G1_ST - 1; ++s) {
    if (s < nk) {
      load_tile_async<G1_BM, 128>(As + s * G1_BM * GLDS, A, lda, m0, m, s * GK, tid);
      load_tile_async<G1_BN, 128>(Bs + s * G1_BN * GLDS, B, ldb, n0, n, s * GK, tid);
    }
    cp_async_commit();
  }
  for (int ks = 0; ks < nk; ++ks) {
    cp_async_wait<G1_ST - 2>();
    __syncthreads();
    int nx = ks + G1_ST - 1;
    if (nx < nk) {
      int st = nx % G1_ST;
      load_tile_async<G1_BM, 128>(As + st * G1_BM * GLDS, A, lda, m0, m, nx * GK, tid);
      load_tile_async<G1_BN, 128>(Bs + st * G1_BN * GLDS, B, ldb, n0, n, nx * GK, tid);
    }
    cp_async_commit();
    int st = ks % G1_ST;
    mma_stage<4, 4>(As + st * G1_BM * GLDS, Bs + st * G1_BN * GLDS, wm0, wn0, g, t, acc);
  }
  cp_async_wait<0>();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int r = m0 + wm0 + i * 8 + g;
    if (r >= m) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        int c = n0 + wn0 + j * 8 + 2 * t + e;
        if (c >= n) continue;
        if (lower_only && c > r) continue;
        double v = alpha * acc[i][j][e];
        if (beta != 0.0) v += beta * C[(size_t)r * ldc + c];
        C[(size_t)r * ldc + c] = v;
      }
  }
}

int launch_gemm_nt(int m, int n, int k, double alpha, const double* A, int lda, const double* B, int ldb, double beta,
                   double* C, int ldc, bool lower_only, cudaStream_t s, LaunchCounter* lc) {
  if (m <= 0 || n <= 0) return BO_OK;
  if ((lda % 2) || (ldb % 2) || ((uintptr_t)A % 16) || ((uintptr_t)B % 16)) {
    bo_set_error("gemm_nt: operands must be 16-byte aligned with even leading dimensions");
    return BO_ERR_INVALID;
  }
  // k is rounded up to a multiple of 16 inside the kernel: callers guarantee zero padding up to ld.
  if (((k + GK - 1) / GK) * GK > lda || ((k + GK - 1) / GK) * GK > ldb) {
    bo_set_error("gemm_nt: k=%d padded to 16 exceeds lda=%d / ldb=%d", k, lda, ldb);
    return BO_ERR_INVALID;
  }
  static PerDeviceOnce attr_once; bool& attr_set = *attr_once.slot();
  size_t smem = (size_t)G1_ST * (G1_BM + G1_BN) * GLDS * sizeof(double);
  if (!attr_set) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(gemm_nt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  dim3 grid((n + G1_BN - 1) / G1_BN, (m + G1_BM - 1) / G1_BM);
  gemm_nt_kernel<<<grid, 128, smem, s>>>(m, n, k, alpha, A, lda, B, ldb, beta, C, ldc, lower_only ? 1 : 0);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// posterior GEMM.  B operand = "LinvExt" [Rpad, ldk]: rows [0, N) the lower-triangular inverse root,
// row N the mean cache alpha, rows N+1 .. N+n_ext-1 the baseline rows K_bX (K + s2 I)^-1, zero padded
// to a multiple of 128 rows.  A 128-row CTA owns whole q-batches and walks every 128-column block:
// purely triangular blocks need k only up to their last row; blocks containing extra rows are dense.
// Columns n < N feed the Gram reduction Gqq = V_q V_q^T; columns n >= N are outputs
// (e = n - N: e == 0 -> mu_raw, e >= 1 -> W[:, e-1] = V_q V_b^T).
// ------------------------------------------------------------------------------------------------
#define PG_BM 128
#define PG_BN 128
#define PG_ST 3
#define PG_THREADS 256
#define PG_VLD (PG_BN + 1)
#define PG_MAXITEMS 5  // ceil(max_q items / 256): q=16 -> 8 batches * 136 pairs = 1088

size_t posterior_gemm_smem_bytes() {
  size_t pipe = (size_t)PG_ST * (PG_BM + PG_BN) * GLDS * sizeof(double);
  size_t stage = (size_t)PG_BM * PG_VLD * sizeof(double);
  return pipe > stage ? pipe : stage;
}

// v1 (any q <= 16): cp.async pipeline, V tile staged through shared memory for the Gram reduction.
__global__ void __launch_bounds__(PG_THREADS, 1) posterior_gemm_kernel(PostGemmArgs a) {
  extern __shared__ __align__(16) double psm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int wm0 = (warp & 1) * 64, wn0 = (warp >> 1) * 32;
  const int q = a.q;
  const int nbat_cta = PG_BM / q;           // whole q-batches per CTA
  const int rows_cta = nbat_cta * q;
  const int row0 = blockIdx.x * rows_cta;
  const int rows_valid = min(a.rows, row0 + rows_cta);  // rows >= rows_valid are zero-filled
  const int npairs = q * (q + 1) / 2;
  const int nitems = nbat_cta * npairs;

  double* As = psm;
  double* Bs = psm + PG_ST * PG_BM * GLDS;
  double* Vs = psm;  // reused after the pipeline drains

  double gacc[PG_MAXITEMS];
#pragma unroll
  for (int i = 0; i < PG_MAXITEMS; ++i) gacc[i] = 0.0;

  const int n_blocks = a.Rpad / PG_BN;
  const int kfull = a.ldk;  // multiple of 16, zero padded

  for (int jb = 0; jb < n_blocks; ++jb) {
    const int n0 = jb * PG_BN;
    const bool dense = (n0 + PG_BN > a.N);
    const int kmax = dense ? kfull : (n0 + PG_BN);
    const int nk = (kmax + GK - 1) / GK;
    const int first_ext = max(0, min(PG_BN, a.N - n0));  // columns >= first_ext of this block are extra rows

    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int s = 0; s < PG_ST - 1; ++s) {
      if (s < nk) {
        load_tile_async<PG_BM, PG_THREADS>(As + s * PG_BM * GLDS, a.Kx, a.ldk, row0, rows_valid, s * GK, tid);
        load_tile_async<PG_BN, PG_THREADS>(Bs + s * PG_BN * GLDS, a.B, a.ldk, n0, a.Rpad, s * GK, tid);
      }
      cp_async_commit();
    }
    for (int ks = 0; ks < nk; ++ks) {
      cp_async_wait<PG_ST - 2>();
      __syncthreads();
      int nx = ks + PG_ST - 1;
      if (nx < nk) {
        int st = nx % PG_ST;
        load_tile_async<PG_BM, PG_THREADS>(As + st * PG_BM * GLDS, a.Kx, a.ldk, row0, rows_valid, nx * GK, tid);
        load_tile_async<PG_BN, PG_THREADS>(Bs + st * PG_BN * GLDS, a.B, a.ldk, n0, a.Rpad, nx * GK, tid);
      }
      cp_async_commit();
      int st = ks % PG_ST;
      mma_stage<8, 4>(As + st * PG_BM * GLDS, Bs + st * PG_BN * GLDS, wm0, wn0, g, t, acc);
    }
    cp_async_wait<0>();
    __syncthreads();  // pipeline drained: stage buffers are free

#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int r = wm0 + i * 8 + g, c = wn0 + j * 8 + 2 * t;
        Vs[r * PG_VLD + c] = acc[i][j][0];
        Vs[r * PG_VLD + c + 1] = acc[i][j][1];
      }
    __syncthreads();

    if (first_ext > 0) {
#pragma unroll
      for (int it = 0; it < PG_MAXITEMS; ++it) {
        int item = tid + it * PG_THREADS;
        if (item < nitems) {
          int bat = item / npairs, p = item % npairs;
          int i = 0, rem = p;  // p -> (i <= j) in row-major upper-triangular order
          while (rem >= q - i) { rem -= q - i; ++i; }
          int j = i + rem;
          const double* vi = Vs + (bat * q + i) * PG_VLD;
          const double* vj = Vs + (bat * q + j) * PG_VLD;
          double s = 0.0;
#pragma unroll 8
          for (int c = 0; c < first_ext; ++c) s = fma(vi[c], vj[c], s);
          gacc[it] += s;
        }
      }
    }
    if (first_ext < PG_BN) {
      const int next = PG_BN - first_ext;
      for (int idx = tid; idx < rows_cta * next; idx += PG_THREADS) {
        int r = idx / next, c = first_ext + idx % next;
        int e = n0 + c - a.N;
        if (row0 + r < rows_valid && e < a.n_ext) {
          double v = Vs[r * PG_VLD + c];
          if (e == 0) a.mu_raw[row0 + r] = v;
          else a.W[(size_t)(row0 + r) * a.ldw + (e - 1)] = v;
        }
      }
    }
    __syncthreads();  // Vs is about to be overwritten by the next block's prologue loads
  }

#pragma unroll
  for (int it = 0; it < PG_MAXITEMS; ++it) {
    int item = tid + it * PG_THREADS;
    if (item < nitems) {
      int bat = item / npairs, p = item % npairs;
      int gb = blockIdx.x * nbat_cta + bat;
      if (gb * q < a.rows) {
        int i = 0, rem = p;
        while (rem >= q - i) { rem -= q - i; ++i; }
        int j = i + rem;
        a.Gqq[((size_t)gb * q + i) * q + j] = gacc[it];
        a.Gqq[((size_t)gb * q + j) * q + i] = gacc[it];
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// v2 (q in {1, 2, 4, 8}; all outputs of the model list in ONE launch, blockIdx.y = output):
// TMA-staged tiles + mbarrier ring, no CTA-wide barrier in the main loop, Gram on the tensor pipe.
//
//  * A (K(X*,X) rows) and B (LinvExt rows) tiles of 128 x 16 doubles arrive by
//    cp.async.bulk.tensor.2d with SWIZZLE_128B into a 6-deep ring (32 KB per stage); one elected lane
//    of warp 0 is the producer, every warp waits on full[stage] and arrives on empty[stage].
//  * Fragment row g of an 8-row MMA tile maps to tile row rho(g) = (g >> 1) | ((g & 1) << 2), which
//    makes each quarter-warp 128-bit load touch 8 distinct 16-byte chunks of the swizzled layout.
//  * V never leaves registers: per 8-row group the 8 x 8 Gram block is accumulated with
//    DMMA(a = acc, b = acc) (A[g][k=t] and B[k=t][n=g] are the same register for V V^T), so warps run
//    from one column block straight into the next without draining the pipeline.
// ------------------------------------------------------------------------------------------------
#include <cuda.h>

#define P2_ST 6
#define P2_TILE_BYTES (128 * 16 * 8)
#define P2_STAGE_BYTES (2 * P2_TILE_BYTES)
#define P2_DIST (P2_ST - 2)  // prefetch distance in stages
#define P2_MAXOUT 8

struct alignas(128) P2Item {
  CUtensorMap mapA;
  CUtensorMap mapB;
  PostGemmArgs a;
};
#define P2_MAXGROUPS 16
// n_groups > 1 (few q-batches): the column blocks are split into contiguous groups of similar work, one CTA per
// (row block, output, group); each group writes its partial Gram to gqq_part + group * gqq_stride and a tiny kernel
// adds the partials in a fixed order.  With one group the Gram goes straight to a.Gqq.
struct P2Batch {
  P2Item item[P2_MAXOUT];
  int n_groups;
  int interleave;  // 1: blockIdx.x = row_block * n_groups + group (CTAs sharing a K(X*,X) panel run together -> L2 hits)
  int gbeg[P2_MAXGROUPS + 1];
  double* gqq_part[P2_MAXOUT];
  long long gqq_stride;
};

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(256, 1) posterior_gemm_tma_kernel(const __grid_constant__ P2Batch batch) {
  extern __shared__ unsigned char p2raw[];
  const P2Item& item = batch.item[blockIdx.y];
  const PostGemmArgs& a = item.a;
  const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(p2raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;           // SWIZZLE_128B wants 1024-byte aligned tiles
  unsigned char* sm = p2raw + (base - raw_addr);
  const uint32_t bar_full = base + P2_ST * P2_STAGE_BYTES;      // 6 x 8 bytes
  const uint32_t bar_empty = bar_full + P2_ST * 8;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  // column group of the warp: warps w and w+4 share an SM sub-partition and get groups (0,3) / (1,2), so the
  // k-stages each may skip inside a diagonal block (L^-1 is lower triangular) are balanced per sub-partition
  const int wc = (warp < 4) ? (warp >> 1) : (3 - ((warp - 4) >> 1));
  const int wm0 = (warp & 1) * 64, wn0 = wc * 32;
  const int rho_g = (g >> 1) | ((g & 1) << 2);
  const int q = a.q, N = a.N, rows = a.rows;
  const int grp = batch.interleave ? (int)(blockIdx.x % batch.n_groups) : (int)blockIdx.z;
  const int row_block = batch.interleave ? (int)(blockIdx.x / batch.n_groups) : (int)blockIdx.x;
  const int row0 = row_block * PG_BM;

  if (tid == 0) {
    for (int s = 0; s < P2_ST; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 8); }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();

  const int jb_begin = batch.gbeg[grp], n_blocks = batch.gbeg[grp + 1];  // this CTA's column blocks
  const int kfull = a.ldk;
  auto nk_of = [&](int jb) {
    int kmax = ((jb + 1) * PG_BN > N) ? kfull : (jb + 1) * PG_BN;
    return (kmax + GK - 1) / GK;
  };

  // producer state (only meaningful in warp 0 / lane 0)
  int p_it = 0, p_jb = jb_begin, p_ks = 0, p_nk = nk_of(jb_begin);
  auto produce = [&]() {
    if (p_jb >= n_blocks) return;
    const int slot = p_it % P2_ST, fill = p_it / P2_ST;
    if (fill > 0) mbar_wait(bar_empty + 8 * slot, (fill - 1) & 1);
    const uint32_t dstA = base + slot * P2_STAGE_BYTES, dstB = dstA + P2_TILE_BYTES;
    mbar_expect_tx(bar_full + 8 * slot, P2_STAGE_BYTES);
    tma_load_2d(dstA, &item.mapA, p_ks * GK, row0, bar_full + 8 * slot);
    tma_load_2d(dstB, &item.mapB, p_ks * GK, p_jb * PG_BN, bar_full + 8 * slot);
    ++p_it;
    if (++p_ks == p_nk) { p_ks = 0; ++p_jb; if (p_jb < n_blocks) p_nk = nk_of(p_jb); }
  };
  if (warp == 0) {
    if (lane == 0)
      for (int s = 0; s < P2_DIST; ++s) produce();
    __syncwarp();
  }

  // swizzled fragment offsets: element (row r, k) sits at r*128 + (((k >> 1) ^ (r & 7)) << 4) + (k & 1)*8
  const int offk0 = ((0 + t) ^ rho_g) << 4, offk1 = ((4 + t) ^ rho_g) << 4;
  const int arow = (wm0 + rho_g) * 128, brow = (wn0 + rho_g) * 128;

  double gram[8][2];
#pragma unroll
  for (int i = 0; i < 8; ++i) gram[i][0] = gram[i][1] = 0.0;

  int it = 0;
  for (int jb = jb_begin; jb < n_blocks; ++jb) {
    const int n0 = jb * PG_BN;
    const int nk = nk_of(jb);
    // this warp's 32 columns are rows n0+wn0 .. n0+wn0+31 of LinvExt: if they are all L^-1 rows they vanish for
    // k beyond the last of them
    const int kneed = (n0 + wn0 + 32 <= N) ? (n0 + wn0 + 32) : kfull;
    double acc[8][4][2];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int ks = 0; ks < nk; ++ks, ++it) {
      if (warp == 0) {
        if (lane == 0) produce();
        __syncwarp();
      }
      const int slot = it % P2_ST;
      mbar_wait(bar_full + 8 * slot, (it / P2_ST) & 1);
      const unsigned char* As = sm + slot * P2_STAGE_BYTES;
      const unsigned char* Bs = As + P2_TILE_BYTES;
      if (ks * GK < kneed) {
#pragma unroll
      for (int kg = 0; kg < 2; ++kg) {
        const int off = kg ? offk1 : offk0;
        double2 af[8], bf[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) af[i] = *reinterpret_cast<const double2*>(As + arow + i * 1024 + off);
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = *reinterpret_cast<const double2*>(Bs + brow + j * 1024 + off);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_884(acc[i][j][0], acc[i][j][1], af[i].x, bf[j].x);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) mma_884(acc[i][j][0], acc[i][j][1], af[i].y, bf[j].y);
      }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * slot);
    }

    if (n0 + PG_BN > N) {
      // block holds extra rows: emit them and clear them out of the Gram operand
#pragma unroll
      for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int c2 = 2 * t + e;
          const int ext = n0 + wn0 + j * 8 + ((c2 >> 1) | ((c2 & 1) << 2)) - N;  // extra-row index, < 0 for L^-1 columns
          if (ext < 0) continue;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r = row0 + wm0 + i * 8 + rho_g;
            if (r < rows && ext < a.n_ext) {
              if (ext == 0) a.mu_raw[r] = acc[i][j][e];
              else a.W[(size_t)r * a.ldw + (ext - 1)] = acc[i][j][e];
            }
            acc[i][j][e] = 0.0;
          }
        }
    }
    if (n0 < N) {
      // Gram of each 8-row group over this warp's 32 columns, on the tensor pipe
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          mma_884(gram[i][0], gram[i][1], acc[i][j][0], acc[i][j][0]);
          mma_884(gram[i][0], gram[i][1], acc[i][j][1], acc[i][j][1]);
        }
    }
  }

  // reduce the 4 column-warps' partial Grams and emit Gqq (q divides 8: a q-batch never straddles an 8-row group)
  __syncthreads();  // every warp is past its last tile read: the ring can be reused
  double* Gs = reinterpret_cast<double*>(sm);  // [4][128][8]
  const int wcol = wc;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int c2 = 2 * t + e;
      Gs[(wcol * 128 + wm0 + i * 8 + rho_g) * 8 + ((c2 >> 1) | ((c2 & 1) << 2))] = gram[i][e];
    }
  __syncthreads();
  for (int idx = tid; idx < PG_BM * q; idx += PG_THREADS) {
    const int r = idx / q, a2 = idx % q;       // row in tile, partner index within its q-batch
    if (row0 + r >= rows) continue;
    const int a1 = r % q;
    const int r2 = r - a1 + a2;
    const int slot8 = r2 & 7;
    double s = Gs[(0 * 128 + r) * 8 + slot8];
    s += Gs[(1 * 128 + r) * 8 + slot8];
    s += Gs[(2 * 128 + r) * 8 + slot8];
    s += Gs[(3 * 128 + r) * 8 + slot8];
    double* dst = (batch.n_groups > 1) ? (batch.gqq_part[blockIdx.y] + (size_t)grp * batch.gqq_stride) : a.Gqq;
    dst[((size_t)(row0 + r) / q * q + a1) * q + a2] = s;
  }
}

__global__ void sum_gram_partials_kernel(const double* __restrict__ part, long long stride, int groups, long long n,
                                         double* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t = 0.0;
  for (int g = 0; g < groups; ++g) t += part[(size_t)g * stride + i];
  out[i] = t;
}

int launch_sum_gram_partials(const double* part, long long stride, int groups, double* out, cudaStream_t s, LaunchCounter* lc) {
  if (stride <= 0) return BO_OK;
  sum_gram_partials_kernel<<<(unsigned)((stride + 255) / 256), 256, 0, s>>>(part, stride, groups, stride, out);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_tile_map(CUtensorMap* map, const double* ptr, int rows, int ld) {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      bo_set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
      return BO_ERR_CUDA;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)(rows > 0 ? rows : 1)};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
  cuuint32_t box[2] = {16, 128};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { bo_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return BO_ERR_CUDA; }
  return BO_OK;
}

static bool use_v2(const PostGemmArgs& a) {
  static int force_v1 = -1;
  if (force_v1 < 0) { const char* e = getenv("EVEREST_GEMM_V1"); force_v1 = (e && e[0] == '1') ? 1 : 0; }
  const bool q_ok = (a.q == 1 || a.q == 2 || a.q == 4 || a.q == 8);
  return q_ok && !force_v1 && ((uintptr_t)a.Kx % 16) == 0 && ((uintptr_t)a.B % 16) == 0;
}

// All outputs of one forward in as few launches as possible (v2: up to 8 outputs per launch).
size_t posterior_gemm_partial_ws_doubles(int rows, int q, int n_out) {
  return (size_t)n_out * P2_MAXGROUPS * rows * q;
}

int launch_posterior_gemm_multi(const PostGemmArgs* args, int n_out, double* part_ws, cudaStream_t s, LaunchCounter* lc) {
  if (n_out <= 0 || args[0].rows <= 0) return BO_OK;
  for (int m = 0; m < n_out; ++m) {
    const PostGemmArgs& a = args[m];
    if (a.q < 1 || a.q > BO_MAX_Q || a.rows % a.q != 0) { bo_set_error("posterior_gemm: bad q"); return BO_ERR_INVALID; }
    if (a.ldk % GK != 0 || a.Rpad % PG_BN != 0 || a.N + a.n_ext > a.Rpad) { bo_set_error("posterior_gemm: padding violated"); return BO_ERR_INVALID; }
  }
  if (use_v2(args[0])) {
    static PerDeviceOnce attr_once; bool& attr_set = *attr_once.slot();
    const size_t smem = (size_t)P2_ST * P2_STAGE_BYTES + 2 * P2_ST * 8 + 1024;
    if (!attr_set) {
      CUDA_CHECK_RET(cudaFuncSetAttribute(posterior_gemm_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr_set = true;
    }
    // few q-batches: split the column blocks so that about two waves of CTAs exist (latency, not throughput, bound)
    static int n_sm = 0;
    if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
    const int row_blocks = (args[0].rows + PG_BM - 1) / PG_BM;
    const int n_blocks_total = args[0].Rpad / PG_BN;
    int groups = 1, interleave = 0;
    {
      // fewer than ~4 waves of CTAs: pick the group count that wastes the least of the last wave,
      // cost(G) = ceil(ctas * G / n_sm) / G  (in units of one full-length CTA)
      const int ctas = row_blocks * n_out;
      if (part_ws && ctas < 4 * n_sm) {
        double best = 1e30;
        const int gmax = (n_blocks_total < P2_MAXGROUPS) ? n_blocks_total : P2_MAXGROUPS;
        for (int G = 1; G <= gmax; ++G) {
          double cost = (double)((ctas * G + n_sm - 1) / n_sm) / (double)G;
          if (cost < best * 0.98) { best = cost; groups = G; }
        }
      } else if (part_ws) {
        // many q-batches: a CTA re-reads its 128-row K(X*,X) panel once per column block.  Splitting the column blocks
        // over G CTAs that are scheduled next to each other keeps the panels of all resident CTAs (n_sm / G panels) in L2,
        // so that the re-reads stop going to HBM.  Target: resident panels <= ~40 MB (L2 also holds the 34 MB factor).
        static int forced = -2;
        if (forced == -2) { const char* e = getenv("EVEREST_GEMM_GROUPS"); forced = e ? atoi(e) : -1; }
        const double panel_mb = (double)PG_BM * args[0].ldk * 8.0 / (1 << 20);
        int G = (int)((n_sm * panel_mb + 39.0) / 40.0);
        if (forced >= 1) G = forced;
        if (G > n_blocks_total) G = n_blocks_total;
        if (G > P2_MAXGROUPS) G = P2_MAXGROUPS;
        if (G > 1) { groups = G; interleave = 1; }
      }
    }
    for (int m0 = 0; m0 < n_out; m0 += P2_MAXOUT) {
      const int cnt = (n_out - m0 < P2_MAXOUT) ? (n_out - m0) : P2_MAXOUT;
      P2Batch batch;
      memset(&batch, 0, sizeof(batch));
      batch.n_groups = groups;
      batch.interleave = interleave;
      {
        // contiguous groups of similar work; work of block jb = its number of k stages
        const PostGemmArgs& a0 = args[m0];
        std::vector<long long> pre(n_blocks_total + 1, 0);
        for (int jb = 0; jb < n_blocks_total; ++jb) {
          int kmax = ((jb + 1) * PG_BN > a0.N) ? a0.ldk : (jb + 1) * PG_BN;
          pre[jb + 1] = pre[jb] + (kmax + GK - 1) / GK;
        }
        batch.gbeg[0] = 0;
        for (int g = 1; g < groups; ++g) {
          long long target = pre[n_blocks_total] * g / groups;
          int jb = batch.gbeg[g - 1] + 1;
          while (jb < n_blocks_total - (groups - g) && pre[jb] < target) ++jb;
          batch.gbeg[g] = jb;
        }
        batch.gbeg[groups] = n_blocks_total;
        batch.gqq_stride = (long long)a0.rows * a0.q;
        for (int i = 0; i < cnt; ++i) batch.gqq_part[i] = part_ws ? part_ws + (size_t)(m0 + i) * groups * batch.gqq_stride : nullptr;
      }
      for (int i = 0; i < cnt; ++i) {
        const PostGemmArgs& a = args[m0 + i];
        int rc;
        if ((rc = make_tile_map(&batch.item[i].mapA, a.Kx, a.rows, a.ldk)) != BO_OK) return rc;
        if ((rc = make_tile_map(&batch.item[i].mapB, a.B, a.Rpad, a.ldk)) != BO_OK) return rc;
        batch.item[i].a = a;
      }
      dim3 grid((args[m0].rows + PG_BM - 1) / PG_BM, cnt, groups);
      if (interleave) grid = dim3(((args[m0].rows + PG_BM - 1) / PG_BM) * groups, cnt, 1);
      posterior_gemm_tma_kernel<<<grid, PG_THREADS, smem, s>>>(batch);
      if (lc) lc->n++;
      if (groups > 1) {
        for (int i = 0; i < cnt; ++i) {
          const long long n = batch.gqq_stride;
          sum_gram_partials_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(batch.gqq_part[i], batch.gqq_stride, groups, n, args[m0 + i].Gqq);
          if (lc) lc->n++;
        }
      }
      CUDA_CHECK_RET(cudaGetLastError());
    }
    return BO_OK;
  }
  static PerDeviceOnce attr_once1; bool& attr_set1 = *attr_once1.slot();
  const size_t smem1 = posterior_gemm_smem_bytes();
  if (!attr_set1) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(posterior_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
    attr_set1 = true;
  }
  for (int m = 0; m < n_out; ++m) {
    const PostGemmArgs& a = args[m];
    int nbat_cta = PG_BM / a.q;
    int nbat = a.rows / a.q;
    int grid = (nbat + nbat_cta - 1) / nbat_cta;
    posterior_gemm_kernel<<<grid, PG_THREADS, smem1, s>>>(a);
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
  }
  return BO_OK;
}

int launch_posterior_gemm(const PostGemmArgs& a, cudaStream_t s, LaunchCounter* lc) {
  return launch_posterior_gemm_multi(&a, 1, nullptr, s, lc);
}

// ------------------------------------------------------------------------------------------------
// Skinny path for the refinement iterations (rows = restarts * q <= 64): the 128-row tensor-pipe tiles above would be
// 50-75 % padding and leave most SMs idle.  C[r][n] = sum_k A[r][k] B[n][k] with all rows of A in one CTA and 32 columns
// per CTA (68 CTAs per output at N = 2000), FP64 FMA from shared-memory tiles; rows n < tri_rows of B are lower
// triangular (L^-1), so their k loop stops at the end of the column tile.
// ------------------------------------------------------------------------------------------------
#define SK_BN 32
#define SK_BK 32
#define SK_MAXOUT 8
struct SkinnyBatch {
  SkinnyItem item[SK_MAXOUT];
  int rows, n, k, lda, ldb, ldc, tri_rows;
};

// leading dimension BK + 4 (= 4 mod 16 doubles): the 32 fragment loads of a warp fall into 2 conflict-free wavefronts
template <int RB, int BK>
__global__ void __launch_bounds__(128)
skinny_gemm_kernel(SkinnyBatch p) {
  // DMMA m8n8k4: warp w owns columns [n0 + 8w, n0 + 8w + 8) for all RB rows (RB / 8 accumulator pairs); both operands are
  // K-contiguous, staged by cp.async into a 2-stage ring.
  __shared__ __align__(16) double As[2][RB][(BK + 4)];
  __shared__ __align__(16) double Bs[2][SK_BN][(BK + 4)];
  const SkinnyItem it = p.item[blockIdx.y];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int n0 = blockIdx.x * SK_BN;
  int kmax = p.k;
  if (n0 + SK_BN <= p.tri_rows) kmax = min(p.k, n0 + SK_BN);  // purely triangular tile
  constexpr int MT = RB / 8;
  double c0[MT], c1[MT];
#pragma unroll
  for (int i = 0; i < MT; ++i) { c0[i] = 0.0; c1[i] = 0.0; }
  const int n_chunks = (kmax + BK - 1) / BK;
  auto stage = [&](int chunk, int buf) {
    const int k0 = chunk * BK;
    for (int idx = tid; idx < RB * (BK / 2); idx += 128) {
      const int r = idx / (BK / 2), kk = (idx % (BK / 2)) * 2;
      cp_async16(&As[buf][r][kk], it.A + (size_t)r * p.lda + k0 + kk, r < p.rows && k0 + kk < kmax);
    }
    for (int idx = tid; idx < SK_BN * (BK / 2); idx += 128) {
      const int c = idx / (BK / 2), kk = (idx % (BK / 2)) * 2;
      cp_async16(&Bs[buf][c][kk], it.B + (size_t)(n0 + c) * p.ldb + k0 + kk, n0 + c < p.n && k0 + kk < kmax);
    }
    cp_async_commit();
  };
  stage(0, 0);
  for (int ch = 0; ch < n_chunks; ++ch) {
    const int buf = ch & 1;
    if (ch + 1 < n_chunks) { stage(ch + 1, buf ^ 1); cp_async_wait<1>(); }
    else cp_async_wait<0>();
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk += 4) {
      const double bf = Bs[buf][warp * 8 + g][kk + t];
#pragma unroll
      for (int i = 0; i < MT; ++i) mma_884(c0[i], c1[i], As[buf][i * 8 + g][kk + t], bf);
    }
    __syncthreads();
  }
  const int col = n0 + warp * 8 + 2 * t;
#pragma unroll
  for (int i = 0; i < MT; ++i) {
    const int r = i * 8 + g;
    if (r < p.rows) {
      if (col < p.n) it.C[(size_t)r * p.ldc + col] = c0[i];
      if (col + 1 < p.n) it.C[(size_t)r * p.ldc + col + 1] = c1[i];
    }
  }
}

int launch_skinny_gemm_nt(const SkinnyItem* items, int n_items, int rows, int n, int k, int lda, int ldb, int ldc,
                          int tri_rows, cudaStream_t s, LaunchCounter* lc) {
  if (rows <= 0 || n <= 0 || n_items <= 0) return BO_OK;
  if (rows > 64) { bo_set_error("skinny_gemm: rows=%d > 64", rows); return BO_ERR_INVALID; }
  for (int i0 = 0; i0 < n_items; i0 += SK_MAXOUT) {
    SkinnyBatch p;
    const int cnt = (n_items - i0 < SK_MAXOUT) ? n_items - i0 : SK_MAXOUT;
    for (int i = 0; i < cnt; ++i) p.item[i] = items[i0 + i];
    p.rows = rows; p.n = n; p.k = k; p.lda = lda; p.ldb = ldb; p.ldc = ldc; p.tri_rows = tri_rows;
    dim3 grid((n + SK_BN - 1) / SK_BN, cnt);
    if (rows <= 32) skinny_gemm_kernel<32, 32><<<grid, 128, 0, s>>>(p);
    else skinny_gemm_kernel<64, 16><<<grid, 128, 0, s>>>(p);
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
  }
  return BO_OK;
}

// Gram / mean / cross terms from the materialised V = K*X LinvExt^T (ld = ldv): one CTA per (q-batch, output).
struct SmallGramBatch {
  const double* V[SK_MAXOUT];
  double* Gqq[SK_MAXOUT];
  double* W[SK_MAXOUT];
  double* mu_raw[SK_MAXOUT];
  int q, N, n_ext, ldv, ldw;
};

__global__ void __launch_bounds__(256)
small_gram_kernel(SmallGramBatch p) {
  const int batch = blockIdx.x, m = blockIdx.y;
  const int q = p.q, warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const double* V = p.V[m] + (size_t)batch * q * p.ldv;
  for (int pr = warp; pr < q * q; pr += nw) {
    const int i = pr / q, j = pr % q;
    if (j < i) continue;
    double s = 0.0;
    for (int n = lane; n < p.N; n += 32) s = fma(V[(size_t)i * p.ldv + n], V[(size_t)j * p.ldv + n], s);
    s = warp_sum(s);
    if (lane == 0) {
      p.Gqq[m][((size_t)batch * q + i) * q + j] = s;
      p.Gqq[m][((size_t)batch * q + j) * q + i] = s;
    }
  }
  for (int idx = threadIdx.x; idx < q * p.n_ext; idx += blockDim.x) {
    const int i = idx / p.n_ext, e = idx % p.n_ext;
    const double v = V[(size_t)i * p.ldv + p.N + e];
    if (e == 0) p.mu_raw[m][(size_t)batch * q + i] = v;
    else if (p.W[m]) p.W[m][((size_t)batch * q + i) * p.ldw + (e - 1)] = v;
  }
}

size_t posterior_small_ws_doubles(int rows, int Rpad, int n_out) { return (size_t)n_out * rows * Rpad; }

// Same outputs as launch_posterior_gemm_multi for rows <= 64; vws holds V for every output ([n_out][rows][Rpad]).
int launch_posterior_small(const PostGemmArgs* args, int n_out, double* vws, cudaStream_t s, LaunchCounter* lc) {
  if (n_out <= 0 || args[0].rows <= 0) return BO_OK;
  const PostGemmArgs& a0 = args[0];
  std::vector<SkinnyItem> items(n_out);
  for (int m = 0; m < n_out; ++m) {
    items[m].A = args[m].Kx; items[m].B = args[m].B; items[m].C = vws + (size_t)m * a0.rows * a0.Rpad;
    if (args[m].Rpad != a0.Rpad || args[m].rows != a0.rows) { bo_set_error("posterior_small: outputs must share the shapes"); return BO_ERR_INVALID; }
  }
  int rc = launch_skinny_gemm_nt(items.data(), n_out, a0.rows, a0.N + a0.n_ext, a0.N, a0.ldk, a0.ldk, a0.Rpad, a0.N, s, lc);
  if (rc != BO_OK) return rc;
  for (int m0 = 0; m0 < n_out; m0 += SK_MAXOUT) {
    SmallGramBatch g;
    const int cnt = (n_out - m0 < SK_MAXOUT) ? n_out - m0 : SK_MAXOUT;
    for (int i = 0; i < cnt; ++i) {
      g.V[i] = items[m0 + i].C; g.Gqq[i] = args[m0 + i].Gqq; g.W[i] = args[m0 + i].W; g.mu_raw[i] = args[m0 + i].mu_raw;
    }
    g.q = a0.q; g.N = a0.N; g.n_ext = a0.n_ext; g.ldv = a0.Rpad; g.ldw = a0.ldw;
    dim3 grid(a0.rows / a0.q, cnt);
    small_gram_kernel<<<grid, 256, 0, s>>>(g);
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
  }
  return BO_OK;
}
