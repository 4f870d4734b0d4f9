// FP64-accurate posterior GEMM on the 5th-generation tensor cores (tcgen05.mma.kind::i8, accumulators in TMEM):
// Ozaki-style error-free splitting of V = K(X*,X) LinvExt^T.
//
// Every operand row is scaled by a power of two and written as a 56-bit fixed-point number in SEVEN balanced base-256
// digits d_p in [-128, 127] (x = s * 2^-55 * sum_p d_p 256^p, |x / s| < 0.498).  A product of two digit planes is an exact
// INT8 x INT8 -> INT32 GEMM (|d d'| <= 2^14, K <= 2^11 per tile, <= 7 pairs per accumulator: < 2^28); the 28 plane pairs
// with p + p' >= 6 are accumulated on SEVEN weight levels (p + p' = 6 .. 12), each level in its own TMEM accumulator
// (7 x 64 columns of the 512), and recombined in FP64 in the epilogue:
//     V[r][n] = sA * sB[n] * 2^-14 * sum_L acc_L[r][n] * 2^(8 (L - 12)).
// The dropped pairs (p + p' <= 5) are random-signed and below 2^-51 of sA * sB: ~10x the rounding of a DGEMM, far inside
// the 1e-9 posterior bar.  Same outputs as posterior_gemm_tma_kernel: Gram V_q V_q^T per q-batch (partial
// sums per column group), mu_raw and W = V_q V_b^T from the extra columns.
//
// Tile: 128 candidate rows x 64 columns per CTA pass (TMEM capacity: 128 lanes x 512 columns of INT32), K in blocks of 64
// bytes.  One warp issues TMA (two 4-D boxes per K block: all 7 planes of A and of B), one warp issues the 56 MMAs of the
// block, four warps form the epilogue (tcgen05.ld, FP64 recombination, shuffle Gram).  Operand tiles are K-major in the
// canonical no-swizzle layout [plane][16-byte K chunk][row][16 B], which is exactly how the planes are stored in HBM, so
// a TMA box lands in shared memory ready for the UMMA descriptors.
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

#define OZ_LEVELS 7
#define OZ_BM 128
#define OZ_BN 64
#define OZ_BK 64                      // bytes of K per block = 4 chunks of 16 B = 2 MMAs of K = 32
#define OZ_CH (OZ_BK / 16)
#define OZ_ST 2
#define OZ_A_PLANE (OZ_CH * OZ_BM * 16)      // 8192 B
#define OZ_B_PLANE (OZ_CH * OZ_BN * 16)      // 4096 B
#define OZ_A_BYTES (OZ_PLANES * OZ_A_PLANE)  // 57344
#define OZ_B_BYTES (OZ_PLANES * OZ_B_PLANE)  // 28672
#define OZ_STAGE_BYTES (OZ_A_BYTES + OZ_B_BYTES)
#define OZ_THREADS 320                // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two warps per TMEM lane quarter)
#define OZ_EPI_THREADS 256
#define OZ_MAXOUT 8
#define OZ_MAXGROUPS 16

// ------------------------------------------------------------------------------------------------
// slicing: FP64 rows -> 7 balanced base-256 digit planes, layout [plane][K chunk][row][16 B]
// ------------------------------------------------------------------------------------------------
// per-row power-of-two scale: smallest 2^e with max|row| / 2^e < 0.49
__global__ void __launch_bounds__(256)
oz_row_scale_kernel(const double* __restrict__ X, int rows, int cols, int ld, double* __restrict__ scale) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= rows) return;
  double mx = 0.0;
  for (int k = lane; k < cols; k += 32) mx = fmax(mx, fabs(X[(size_t)warp * ld + k]));
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) {
    double s = 1.0;
    if (mx > 0.0) {
      int e;
      frexp(mx, &e);              // mx = f * 2^e, f in [0.5, 1)
      s = ldexp(1.0, e + 2);      // mx / s < 0.25 <= 0.49
    }
    scale[warp] = s;
  }
}

// X [rows, ld] (cols valid, zero beyond) -> planes; row_scale == nullptr: global scale `gscale`
__global__ void __launch_bounds__(256)
oz_slice_kernel(const double* __restrict__ X, int rows, int cols, int ld, const double* __restrict__ row_scale, double gscale,
                signed char* __restrict__ planes, int rows_alloc, int n_chunks) {
  // thread = (row, 16-byte chunk): reads 16 doubles, writes 16 B to each plane
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int r = (int)(idx % rows_alloc), c = (int)(idx / rows_alloc);
  if (c >= n_chunks) return;
  signed char dig[OZ_PLANES][16];
  const double inv = (r < rows) ? 1.0 / (row_scale ? row_scale[r] : gscale) : 0.0;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int k = c * 16 + t;
    const double v = (r < rows && k < cols) ? X[(size_t)r * ld + k] * inv : 0.0;
    signed char d[OZ_PLANES];
    oz_split_digits(v, d);
#pragma unroll
    for (int p = 0; p < OZ_PLANES; ++p) dig[p][t] = d[p];
  }
  const size_t plane_stride = (size_t)n_chunks * rows_alloc * 16;
#pragma unroll
  for (int p = 0; p < OZ_PLANES; ++p) {
    int4 w;
    memcpy(&w, dig[p], 16);
    *reinterpret_cast<int4*>(planes + p * plane_stride + ((size_t)c * rows_alloc + r) * 16) = w;
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 GEMM
// ------------------------------------------------------------------------------------------------
struct alignas(128) OzItem {
  CUtensorMap mapA;          // planes of K(X*,X): dims {rows_alloc * 16 B as u64, chunks, 7}
  CUtensorMap mapB;          // planes of LinvExt
  const double* scaleB;      // [Rpad] per-row scale of LinvExt
  double scaleA;             // global scale of K(X*,X)
  double* gqq_part;          // [groups][b, q, q]
  double* W;                 // [rows, ldw] or null
  double* mu_raw;            // [rows]
};
struct OzBatch {
  OzItem item[OZ_MAXOUT];
  int rows, q, N, n_ext, Rpad, ldw, n_chunks_k;   // n_chunks_k = ldk / 16
  int n_groups;
  int dbg;                                        // EVEREST_OZAKI_DBG bits: 1 = no MMAs, 2 = no epilogue math, 4 = no TMA (pair kernel), 16 = evict_last scratch
  int gbeg[OZ_MAXGROUPS + 1];                     // column-tile ranges (units of OZ_BN columns)
  long long gqq_stride;
};

__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void oz_tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}
// K-major, SWIZZLE_NONE canonical layout ((8, m), 2) : ((16 B, SBO), LBO)
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ void oz_mma_i8(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t"
      "}\n" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0));
}
__device__ __forceinline__ void oz_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Gram update of one element: this thread owns row i of its q-batch's G, the partners are the Q lanes of its group.  The
// shuffles must sit in straight-line code: under a condition the compiler cannot prove warp-uniform every __shfl_sync is
// wrapped in an elect / divergence-barrier sequence with local-memory spills (ncu: ~100 warp instructions per element).
template <int Q>
__device__ __forceinline__ void oz_gram_update(double v, int lane_base, double (&g)[8]) {
#pragma unroll
  for (int j = 0; j < Q; ++j) g[j] = fma(v, __shfl_sync(0xffffffffu, v, lane_base + j), g[j]);
}
template <int Q>
__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm_kernel(const __grid_constant__ OzBatch batch) {
  extern __shared__ unsigned char ozraw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bars[2 * OZ_ST + 2];   // full[ST], empty[ST], acc_full, acc_empty
  __shared__ double sB_s[8][32];                          // per-epilogue-warp column scales of the current tile
  const OzItem& item = batch.item[blockIdx.y];
  const uint32_t base = (oz_smem_u32(ozraw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = (int)(blockIdx.x % batch.n_groups), row_tile = (int)(blockIdx.x / batch.n_groups);
  const int row0 = row_tile * OZ_BM;
  // Column tiles are dealt to the groups in snake order (round r: tile r G + g, or r G + G - 1 - g on odd rounds): every
  // CTA gets a mix of short (early, triangular) and long tiles of similar total work, so that the FP64 epilogue of one tile
  // hides behind the MMAs of the next instead of piling up on a run of one-K-block tiles.
  const int n_tiles = batch.Rpad / OZ_BN, G = batch.n_groups;
  auto tile_at = [&](int i) { const int j = i * G + ((i & 1) ? (G - 1 - grp) : grp); return (j < n_tiles) ? j : -1; };
  const uint32_t full0 = oz_smem_u32(&bars[0]), empty0 = oz_smem_u32(&bars[OZ_ST]);
  const uint32_t acc_full = oz_smem_u32(&bars[2 * OZ_ST]), acc_empty = oz_smem_u32(&bars[2 * OZ_ST + 1]);
  if (tid == 0) {
    for (int s = 0; s < OZ_ST; ++s) { oz_mbar_init(full0 + 8 * s, 1); oz_mbar_init(empty0 + 8 * s, 1); }
    oz_mbar_init(acc_full, 1);
    oz_mbar_init(acc_empty, OZ_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  // K blocks of a column tile: purely triangular tiles (all 64 rows of L^-1 below N) stop at their last row
  auto k_blocks_of = [&](int jt) {
    const int n_end = (jt + 1) * OZ_BN;
    const int kbytes = (n_end <= batch.N) ? n_end : batch.n_chunks_k * 16;
    return (min(kbytes, batch.n_chunks_k * 16) + OZ_BK - 1) / OZ_BK;
  };

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int it = 0;
      for (int ti = 0, jt; (jt = tile_at(ti)) >= 0; ++ti) {
        const int nkb = k_blocks_of(jt);
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % OZ_ST;
          const uint32_t ph = (uint32_t)((it / OZ_ST) & 1);
          oz_mbar_wait(empty0 + 8 * s, ph ^ 1u);          // first use of a stage passes immediately
          const uint32_t fb = full0 + 8 * s;
          oz_mbar_expect_tx(fb, OZ_STAGE_BYTES);
          const uint32_t dst = base + (uint32_t)s * OZ_STAGE_BYTES;
          oz_tma_load_3d(dst, &item.mapA, row0 * 2, kb * OZ_CH, 0, fb);              // inner coordinate in 8-byte units
          oz_tma_load_3d(dst + OZ_A_BYTES, &item.mapB, jt * OZ_BN * 2, kb * OZ_CH, 0, fb);
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      uint32_t idesc = 0;
      idesc |= 2u << 4;                       // D = S32
      idesc |= 1u << 7;                       // A signed 8 bit
      idesc |= 1u << 10;                      // B signed 8 bit
      idesc |= (uint32_t)(OZ_BN >> 3) << 17;
      idesc |= (uint32_t)(OZ_BM >> 4) << 24;
      int it = 0, tile = 0;
      for (int ti = 0, jt; (jt = tile_at(ti)) >= 0; ++ti, ++tile) {
        const int nkb = k_blocks_of(jt);
        // the epilogue must have drained the accumulators of the previous tile
        oz_mbar_wait(acc_empty, (uint32_t)((tile & 1) ^ 1));
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % OZ_ST;
          const uint32_t ph = (uint32_t)((it / OZ_ST) & 1);
          oz_mbar_wait(full0 + 8 * s, ph);
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          const uint32_t sa = base + (uint32_t)s * OZ_STAGE_BYTES, sb = sa + OZ_A_BYTES;
          // descriptors differ only in the 14-bit start-address field: one base per operand, compile-time offsets
          // (the issuing thread must deliver one MMA every ~48 clocks)
          const uint64_t dA0 = oz_desc(sa, OZ_BM * 16, 128), dB0 = oz_desc(sb, OZ_BN * 16, 128);
          const uint32_t acc0 = (kb > 0) ? 1u : 0u;
          if (!(batch.dbg & 1))
#pragma unroll
          for (int pa = OZ_PLANES - 1; pa >= 0; --pa) {
#pragma unroll
            for (int pb = OZ_PLANES - 1; pb >= 0; --pb) {
              if (pa + pb < 6) continue;
              const int lvl = pa + pb - 6;
              // with pa descending from 6 the first pair of every level is (6, L - 6)
              const bool first = (pa == OZ_PLANES - 1);
#pragma unroll
              for (int ks = 0; ks < OZ_BK / 32; ++ks) {
                const uint64_t da = dA0 + (uint64_t)((pa * OZ_A_PLANE + ks * 2 * (OZ_BM * 16)) >> 4);
                const uint64_t db = dB0 + (uint64_t)((pb * OZ_B_PLANE + ks * 2 * (OZ_BN * 16)) >> 4);
                oz_mma_i8(tmem_base + (uint32_t)(lvl * OZ_BN), da, db, idesc, (first && ks == 0) ? acc0 : 1u);
              }
            }
          }
          oz_commit(empty0 + 8 * s);          // the stage is free once these MMAs have read it
        }
        oz_commit(acc_full);                  // accumulators of this tile complete
      }
    }
  } else {
    // ---------------- epilogue: warps 2..9; TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 ----------------
    const int quarter = warp & 3, half = (warp - 2) >> 2, ewarp = warp - 2;
    const int r_local = quarter * 32 + lane;
    const int row = row0 + r_local;
    constexpr int q = Q;
    const int lane_base = lane & ~(q - 1);
    double g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.0;
    const double sA = item.scaleA * 6.103515625e-05;   // 2^-14
    int tile = 0;
    for (int ti = 0, jt; (jt = tile_at(ti)) >= 0; ++ti, ++tile) {
      const int n0 = jt * OZ_BN + half * 32;           // this warp's 32 columns
      // per-column scales of the tile: one coalesced load per warp, broadcast from shared memory
      sB_s[ewarp][lane] = item.scaleB[n0 + lane] * sA;
      __syncwarp();
      oz_mbar_wait(acc_full, (uint32_t)(tile & 1));
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      // Drain all seven accumulators into two exact 64-bit integers per column (Horner in base 256: levels 6..3 and 2..0),
      // hand TMEM back to the MMA warp, and only then do the FP64 work -- it overlaps the next tile's MMAs.
      long long hi[32], lo[32];
#pragma unroll
      for (int lvl = OZ_LEVELS - 1; lvl >= 0; --lvl) {
        uint32_t v[32];
#pragma unroll
        for (int c0 = 0; c0 < 32; c0 += 16) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(lvl * OZ_BN + half * 32 + c0);
          asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                       : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
                         "=r"(v[c0 + 6]), "=r"(v[c0 + 7]), "=r"(v[c0 + 8]), "=r"(v[c0 + 9]), "=r"(v[c0 + 10]), "=r"(v[c0 + 11]),
                         "=r"(v[c0 + 12]), "=r"(v[c0 + 13]), "=r"(v[c0 + 14]), "=r"(v[c0 + 15])
                       : "r"(taddr));
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          if (lvl == OZ_LEVELS - 1) hi[c] = (long long)(int)v[c];
          else if (lvl >= 3) hi[c] = hi[c] * 256 + (long long)(int)v[c];
          else if (lvl == 2) lo[c] = (long long)(int)v[c];
          else lo[c] = lo[c] * 256 + (long long)(int)v[c];
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      oz_mbar_arrive(acc_empty);
      if (!(batch.dbg & 2)) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          // sum_lvl acc_lvl 2^(8 (lvl - 6)) = hi 2^-24 + lo 2^-48, both integers exact in FP64 (< 2^53)
          const double sum = fma((double)lo[c], 3.552713678800501e-15 /* 2^-48 */, (double)hi[c] * 5.960464477539063e-08 /* 2^-24 */);
          const int n = n0 + c;
          const double val = sum * sB_s[ewarp][c];
          // Gram of the q-batch (columns beyond N contribute zero: branch-free, see oz_gram_update)
          oz_gram_update<Q>((n < batch.N) ? val : 0.0, lane_base, g);
          if (n >= batch.N && row < batch.rows) {
            const int e = n - batch.N;
            if (e == 0) item.mu_raw[row] = val;
            else if (item.W && e < batch.n_ext) item.W[(size_t)row * batch.ldw + (e - 1)] = val;
          }
        }
      }
      __syncwarp();

    }
    if (row < batch.rows) {
      // partial Gram slot = (column group, column half): summed in a fixed order by sum_gram_partials_kernel
      double* dst = item.gqq_part + (size_t)(grp * 2 + half) * batch.gqq_stride + ((size_t)(row / q) * q + (row % q)) * q;
      for (int j = 0; j < q; ++j) dst[j] = g[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
}

// ------------------------------------------------------------------------------------------------
// Two-pass variant with 128 x 128 tiles.  An M128 x N64 MMA occupies the tensor pipe for 48 clocks where M128 x N128 needs
// 64 (tools/i8_mma_probe.cu), so N = 128 is worth 1.5x -- but seven 128-column accumulators do not fit the 512 TMEM
// columns.  The K loop of a tile therefore runs twice: pass 1 accumulates the four LOW levels (p + p' = 6..9, 22 plane
// pairs, all planes), the epilogue warps drain them into one exact 64-bit integer per element and park it in an
// L2-resident scratch slab (128 KB per SM); pass 2 re-streams only planes 4..6 of both operands for the three HIGH levels
// (6 pairs).  Operand traffic per MAC is the same as the one-pass kernel's (10/7 of the planes for twice the columns).
// ------------------------------------------------------------------------------------------------
#define O2_BN 128
#define O2_BK 32
#define O2_CH (O2_BK / 16)
#define O2_ST 3
#define O2_PLANE (O2_CH * 128 * 16)            // 4096 B: one plane of one operand per K block
#define O2_OP_BYTES (OZ_PLANES * O2_PLANE)     // 28672
#define O2_STAGE_BYTES (2 * O2_OP_BYTES)       // 57344
#define O2_HI_PLANES 3
#define O2_HI_OP_BYTES (O2_HI_PLANES * O2_PLANE)   // 12288
#define O2_MAXOUT 4
#define O2_SCRATCH_SLOTS 256   // >= %nsmid on every part seen so far (148 SMs on B200); the kernel traps beyond

struct alignas(128) O2Item {
  CUtensorMap mapA, mapB;      // box {128 rows, 2 chunks, 7 planes}
  CUtensorMap mapA3, mapB3;    // box {128 rows, 2 chunks, 3 planes} (planes 4..6)
  const double* scaleB;
  double scaleA;
  double* gqq_part;
  double* W;
  double* mu_raw;
};
struct O2Batch {
  O2Item item[O2_MAXOUT];
  int rows, q, N, n_ext, Rpad, ldw, n_chunks_k;
  int n_groups;
  int dbg;
  long long gqq_stride;
  long long* scratch;          // [O2_SCRATCH_SLOTS][8 warps][64 columns][32 lanes]
};

// Experiment kept behind EVEREST_OZAKI_DBG bit 16: marking the integer scratch slab (128 KB per SM, rewritten for every
// tile) evict_last keeps it out of HBM (DRAM writes per launch 1.13 GB -> 0.19 GB, ncu) but the pinned lines cost the
// neighbouring kernels more than the GEMM gains: the 16384-q-batch screen went from 9.51 to 9.63 ms (A/B on one box).
// Default: plain stores / loads.
__device__ __forceinline__ uint64_t oz_policy_evict_last() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void oz_st_keep(long long* p, long long v, uint64_t pol) {
  if (pol) asm volatile("st.global.L2::cache_hint.b64 [%0], %1, %2;\n" ::"l"(p), "l"(v), "l"(pol) : "memory");
  else *p = v;
}
__device__ __forceinline__ long long oz_ld_keep(const long long* p, uint64_t pol) {
  long long v;
  if (pol) asm volatile("ld.global.L2::cache_hint.b64 %0, [%1], %2;\n" : "=l"(v) : "l"(p), "l"(pol) : "memory");
  else v = *p;
  return v;
}

template <int Q>
__global__ void __launch_bounds__(OZ_THREADS, 1) ozaki_gemm2p_kernel(const __grid_constant__ O2Batch batch) {
  extern __shared__ unsigned char ozraw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bars[2 * O2_ST + 2];
  __shared__ double sB_s[8][64];
  const O2Item& item = batch.item[blockIdx.y];
  const uint32_t base = (oz_smem_u32(ozraw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int grp = (int)(blockIdx.x % batch.n_groups), row_tile = (int)(blockIdx.x / batch.n_groups);
  const int row0 = row_tile * OZ_BM;
  const int n_tiles = batch.Rpad / O2_BN, G = batch.n_groups;
  auto tile_at = [&](int i) { const int j = i * G + ((i & 1) ? (G - 1 - grp) : grp); return (j < n_tiles) ? j : -1; };
  const uint32_t full0 = oz_smem_u32(&bars[0]), empty0 = oz_smem_u32(&bars[O2_ST]);
  const uint32_t acc_full = oz_smem_u32(&bars[2 * O2_ST]), acc_empty = oz_smem_u32(&bars[2 * O2_ST + 1]);
  if (tid == 0) {
    for (int s = 0; s < O2_ST; ++s) { oz_mbar_init(full0 + 8 * s, 1); oz_mbar_init(empty0 + 8 * s, 1); }
    oz_mbar_init(acc_full, 1);
    oz_mbar_init(acc_empty, OZ_EPI_THREADS);
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  auto k_blocks_of = [&](int jt) {
    const int n_end = (jt + 1) * O2_BN;
    const int kmax = batch.n_chunks_k * 16;
    const int kbytes = (n_end <= batch.N) ? n_end : kmax;
    return (min(kbytes, kmax) + O2_BK - 1) / O2_BK;
  };

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (lane == 0) {
      int it = 0;
      for (int ti = 0, jt; (jt = tile_at(ti)) >= 0; ++ti) {
        const int nkb = k_blocks_of(jt);
        for (int pass = 0; pass < 2; ++pass) {
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % O2_ST;
            const uint32_t ph = (uint32_t)((it / O2_ST) & 1);
            oz_mbar_wait(empty0 + 8 * s, ph ^ 1u);
            const uint32_t fb = full0 + 8 * s;
            const uint32_t dst = base + (uint32_t)s * O2_STAGE_BYTES;
            if (pass == 0) {
              oz_mbar_expect_tx(fb, O2_STAGE_BYTES);
              oz_tma_load_3d(dst, &item.mapA, row0 * 2, kb * O2_CH, 0, fb);
              oz_tma_load_3d(dst + O2_OP_BYTES, &item.mapB, jt * O2_BN * 2, kb * O2_CH, 0, fb);
            } else {
              oz_mbar_expect_tx(fb, 2 * O2_HI_OP_BYTES);
              oz_tma_load_3d(dst, &item.mapA3, row0 * 2, kb * O2_CH, OZ_PLANES - O2_HI_PLANES, fb);
              oz_tma_load_3d(dst + O2_HI_OP_BYTES, &item.mapB3, jt * O2_BN * 2, kb * O2_CH, OZ_PLANES - O2_HI_PLANES, fb);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    if (lane == 0) {
      uint32_t idesc = 0;
      idesc |= 2u << 4;                       // D = S32
      idesc |= 1u << 7;                       // A signed 8 bit
      idesc |= 1u << 10;                      // B signed 8 bit
      idesc |= (uint32_t)(OZ_BM >> 4) << 24;
      const uint32_t idesc_base = idesc;
      int it = 0, ev = 0;
      for (int ti = 0, jt; (jt = tile_at(ti)) >= 0; ++ti) {
        const int nkb = k_blocks_of(jt);
        // the last column tile usually holds only a few real columns (N + 1 + n_b is padded to 128): issue narrower MMAs
        // there (N is any multiple of 16); the accumulator columns beyond keep stale integers the epilogue never uses
        const int n_eff = min(O2_BN, ((batch.N + batch.n_ext - jt * O2_BN + 15) >> 4) << 4);
        idesc = idesc_base | ((uint32_t)(max(n_eff, 16) >> 3) << 17);
        for (int pass = 0; pass < 2; ++pass, ++ev) {
          // pass 1 needs the previous tile's high levels drained, pass 2 this tile's low levels
          oz_mbar_wait(acc_empty, (uint32_t)((ev & 1) ^ 1));
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % O2_ST;
            const uint32_t ph = (uint32_t)((it / O2_ST) & 1);
            oz_mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t sa = base + (uint32_t)s * O2_STAGE_BYTES;
            const uint32_t acc0 = (kb > 0) ? 1u : 0u;
            if (pass == 0) {
              const uint64_t dA0 = oz_desc(sa, 128 * 16, 128), dB0 = oz_desc(sa + O2_OP_BYTES, 128 * 16, 128);
              if (!(batch.dbg & 1))
#pragma unroll
              for (int pa = OZ_PLANES - 1; pa >= 0; --pa) {
#pragma unroll
                for (int pb = OZ_PLANES - 1; pb >= 0; --pb) {
                  if (pa + pb < 6 || pa + pb > 9) continue;
                  const int lvl = pa + pb - 6;                         // 0..3; first pair of a level: pa == 6
                  oz_mma_i8(tmem_base + (uint32_t)(lvl * O2_BN), dA0 + (uint64_t)((pa * O2_PLANE) >> 4),
                            dB0 + (uint64_t)((pb * O2_PLANE) >> 4), idesc, (pa == OZ_PLANES - 1) ? acc0 : 1u);
                }
              }
            } else {
              const uint64_t dA0 = oz_desc(sa, 128 * 16, 128), dB0 = oz_desc(sa + O2_HI_OP_BYTES, 128 * 16, 128);
              if (!(batch.dbg & 1))
#pragma unroll
              for (int pa = OZ_PLANES - 1; pa >= OZ_PLANES - O2_HI_PLANES; --pa) {
#pragma unroll
                for (int pb = OZ_PLANES - 1; pb >= OZ_PLANES - O2_HI_PLANES; --pb) {
                  if (pa + pb < 10) continue;
                  const int lvl = pa + pb - 10;                        // 0..2 = levels 4..6
                  oz_mma_i8(tmem_base + (uint32_t)(lvl * O2_BN), dA0 + (uint64_t)(((pa - 4) * O2_PLANE) >> 4),
                            dB0 + (uint64_t)(((pb - 4) * O2_PLANE) >> 4), idesc, (pa == OZ_PLANES - 1) ? acc0 : 1u);
                }
              }
            }
            oz_commit(empty0 + 8 * s);
          }
          oz_commit(acc_full);
        }
      }
    }
  } else {
    // ---------------- epilogue: warps 2..9; TMEM lane quarter = warp % 4, 64-column half = (warp - 2) / 4 ----------------
    const int quarter = warp & 3, half = (warp - 2) >> 2, ewarp = warp - 2;
    const int r_local = quarter * 32 + lane;
    const int row = row0 + r_local;
    constexpr int q = Q;
    const int lane_base = lane & ~(q - 1);
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid >= O2_SCRATCH_SLOTS) __trap();
    const uint64_t scr_policy = (batch.dbg & 16) ? oz_policy_evict_last() : 0ull;
    long long* scr = batch.scratch + (((size_t)smid * 8 + ewarp) * 64) * 32 + lane;   // [column][lane]
    double g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.0;
    const double sA = item.scaleA * 6.103515625e-05;   // 2^-14
    const uint32_t tcol0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 64);
    int ev = 0;
    for (int ti = 0, jt; (jt = tile_at(ti)) >= 0; ++ti) {
      const int n0 = jt * O2_BN + half * 64;
      sB_s[ewarp][lane] = item.scaleB[n0 + lane] * sA;
      sB_s[ewarp][lane + 32] = item.scaleB[n0 + 32 + lane] * sA;
      __syncwarp();
      // ---- low levels (3..0) -> one exact integer per element, parked in the scratch slab ----
      oz_mbar_wait(acc_full, (uint32_t)(ev & 1)); ++ev;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        long long lo[32];
#pragma unroll
        for (int lvl = 3; lvl >= 0; --lvl) {
          uint32_t v[32];
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            const uint32_t taddr = tcol0 + (uint32_t)(lvl * O2_BN + h * 32 + c0);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                         : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
                           "=r"(v[c0 + 6]), "=r"(v[c0 + 7]), "=r"(v[c0 + 8]), "=r"(v[c0 + 9]), "=r"(v[c0 + 10]), "=r"(v[c0 + 11]),
                           "=r"(v[c0 + 12]), "=r"(v[c0 + 13]), "=r"(v[c0 + 14]), "=r"(v[c0 + 15])
                         : "r"(taddr));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
          for (int c = 0; c < 32; ++c) lo[c] = (lvl == 3) ? (long long)(int)v[c] : lo[c] * 256 + (long long)(int)v[c];
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) oz_st_keep(scr + (size_t)(h * 32 + c) * 32, lo[c], scr_policy);
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      oz_mbar_arrive(acc_empty);
      // ---- high levels (6..4) ----
      oz_mbar_wait(acc_full, (uint32_t)(ev & 1)); ++ev;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      long long hi[64];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int lvl = 2; lvl >= 0; --lvl) {
          uint32_t v[32];
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            const uint32_t taddr = tcol0 + (uint32_t)(lvl * O2_BN + h * 32 + c0);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                         : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
                           "=r"(v[c0 + 6]), "=r"(v[c0 + 7]), "=r"(v[c0 + 8]), "=r"(v[c0 + 9]), "=r"(v[c0 + 10]), "=r"(v[c0 + 11]),
                           "=r"(v[c0 + 12]), "=r"(v[c0 + 13]), "=r"(v[c0 + 14]), "=r"(v[c0 + 15])
                         : "r"(taddr));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
          for (int c = 0; c < 32; ++c)
            hi[h * 32 + c] = (lvl == 2) ? (long long)(int)v[c] : hi[h * 32 + c] * 256 + (long long)(int)v[c];
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      oz_mbar_arrive(acc_empty);
      if (!(batch.dbg & 2)) {
#pragma unroll
        for (int c8 = 0; c8 < 64; c8 += 8) {
          long long lo[8];                     // the parked low levels come back from L2 eight at a time
#pragma unroll
          for (int k = 0; k < 8; ++k) lo[k] = oz_ld_keep(scr + (size_t)(c8 + k) * 32, scr_policy);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = c8 + k;
            // sum_L acc_L 2^(8 (L - 6)) = hi 2^-16 + lo 2^-48
            const double sum = fma((double)lo[k], 3.552713678800501e-15 /* 2^-48 */, (double)hi[c] * 1.52587890625e-05 /* 2^-16 */);
            const int n = n0 + c;
            const double val = sum * sB_s[ewarp][c];
            oz_gram_update<Q>((n < batch.N) ? val : 0.0, lane_base, g);
            if (n >= batch.N && row < batch.rows) {
              const int e = n - batch.N;
              if (e == 0) item.mu_raw[row] = val;
              else if (item.W && e < batch.n_ext) item.W[(size_t)row * batch.ldw + (e - 1)] = val;
            }
          }
        }
      }
      __syncwarp();
    }
    if (row < batch.rows) {
      double* dst = item.gqq_part + (size_t)(grp * 2 + half) * batch.gqq_stride + ((size_t)(row / q) * q + (row % q)) * q;
      for (int j = 0; j < q; ++j) dst[j] = g[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant of the two-pass kernel (tcgen05 cta_group::2, M = 256 x N = 128 per MMA).  Shared-memory bandwidth
// (128 B/clk per SM) is what bounds the single-CTA kernels: an M128 x N128 x K32 MMA reads 4 KB of A and 4 KB of B while
// TMA writes another 2.6 KB per MMA into the same memory -> 83 clocks instead of the pipe's 64.  In a pair each SM reads
// its own 128 rows of A and only HALF of the B tile (the other half arrives from the peer SM) and TMA writes 1.9 KB per
// MMA: 62 clocks' worth, i.e. the tensor pipe becomes the bound.  Both CTAs load (TMA signalling the leader's barriers),
// the leader issues every MMA, commits are multicast to both CTAs, each CTA drains its own TMEM.
// ------------------------------------------------------------------------------------------------
#define O3_ST 4
#define O3_B_PLANE (O2_CH * 64 * 16)               // 2048 B: half of the B tile
#define O3_A_BYTES (OZ_PLANES * O2_PLANE)          // 28672
#define O3_B_BYTES (OZ_PLANES * O3_B_PLANE)        // 14336
#define O3_STAGE_BYTES (O3_A_BYTES + O3_B_BYTES)   // 43008
#define O3_HI_A_BYTES (O2_HI_PLANES * O2_PLANE)    // 12288
#define O3_HI_B_BYTES (O2_HI_PLANES * O3_B_PLANE)  // 6144

__device__ __forceinline__ uint32_t oz_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void oz_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// own shared memory as destination, the mbarrier lives in the pair's leader CTA
__device__ __forceinline__ void oz_tma_load_3d_2sm(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n"
               ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void oz_mma_i8_2sm(uint32_t tmem_c, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t"
      "}\n" ::"r"(tmem_c), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0), "r"(0));
}
__device__ __forceinline__ void oz_commit_2sm(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((unsigned short)3) : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];\n" ::"r"(bar_cluster) : "memory");
}

template <int Q>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(OZ_THREADS, 1) ozaki_gemm2c_kernel(const __grid_constant__ O2Batch batch) {
  extern __shared__ unsigned char ozraw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bars[2 * O3_ST + 2];
  __shared__ double sB_s[8][64];
  const O2Item& item = batch.item[blockIdx.y];
  const uint32_t base = (oz_smem_u32(ozraw) + 1023u) & ~1023u;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int pair = (int)(blockIdx.x >> 1);
  const int grp = pair % batch.n_groups, row_pair = pair / batch.n_groups;
  const int row0 = (row_pair * 2 + (int)rank) * OZ_BM;
  const int n_tiles = batch.Rpad / O2_BN, G = batch.n_groups;
  auto tile_at = [&](int i) { const int j = i * G + ((i & 1) ? (G - 1 - grp) : grp); return (j < n_tiles) ? j : -1; };
  const uint32_t full0 = oz_smem_u32(&bars[0]), empty0 = oz_smem_u32(&bars[O3_ST]);
  const uint32_t acc_full = oz_smem_u32(&bars[2 * O3_ST]), acc_empty = oz_smem_u32(&bars[2 * O3_ST + 1]);
  if (tid == 0) {
    for (int s = 0; s < O3_ST; ++s) { oz_mbar_init(full0 + 8 * s, 1); oz_mbar_init(empty0 + 8 * s, 1); }
    oz_mbar_init(acc_full, 1);
    oz_mbar_init(acc_empty, 16);            // one arrival per epilogue warp of both CTAs (leader's barrier only)
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(oz_smem_u32(&tmem_base_s)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  oz_cluster_sync();
  asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  auto k_blocks_of = [&](int jt) {
    const int n_end = (jt + 1) * O2_BN;
    const int kmax = batch.n_chunks_k * 16;
    const int kbytes = (n_end <= batch.N) ? n_end : kmax;
    return (min(kbytes, kmax) + O2_BK - 1) / O2_BK;
  };

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs; the leader also posts the byte count of the pair) ----------------
    if (lane == 0 && !(batch.dbg & 4)) {
      int it = 0;
      for (int ti = 0, jt; (jt = tile_at(ti)) >= 0; ++ti) {
        const int nkb = k_blocks_of(jt);
        const int brow = (jt * O2_BN + (int)rank * 64) * 2;
        for (int pass = 0; pass < 2; ++pass) {
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % O3_ST;
            const uint32_t ph = (uint32_t)((it / O3_ST) & 1);
            oz_mbar_wait(empty0 + 8 * s, ph ^ 1u);
            const uint32_t fb = oz_mapa(full0 + 8 * s, 0);
            const uint32_t dst = base + (uint32_t)s * O3_STAGE_BYTES;
            if (pass == 0) {
              if (rank == 0) oz_mbar_expect_tx(full0 + 8 * s, 2 * O3_STAGE_BYTES);
              oz_tma_load_3d_2sm(dst, &item.mapA, row0 * 2, kb * O2_CH, 0, fb);
              oz_tma_load_3d_2sm(dst + O3_A_BYTES, &item.mapB, brow, kb * O2_CH, 0, fb);
            } else {
              if (rank == 0) oz_mbar_expect_tx(full0 + 8 * s, 2 * (O3_HI_A_BYTES + O3_HI_B_BYTES));
              oz_tma_load_3d_2sm(dst, &item.mapA3, row0 * 2, kb * O2_CH, OZ_PLANES - O2_HI_PLANES, fb);
              oz_tma_load_3d_2sm(dst + O3_HI_A_BYTES, &item.mapB3, brow, kb * O2_CH, OZ_PLANES - O2_HI_PLANES, fb);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: leader CTA only ----------------
    if (lane == 0 && rank == 0) {
      uint32_t idesc = 0;
      idesc |= 2u << 4;                       // D = S32
      idesc |= 1u << 7;                       // A signed 8 bit
      idesc |= 1u << 10;                      // B signed 8 bit
      idesc |= (uint32_t)(O2_BN >> 3) << 17;
      idesc |= (uint32_t)(256 >> 4) << 24;    // M = 256 across the pair
      int it = 0, ev = 0;
      for (int ti = 0, jt; (jt = tile_at(ti)) >= 0; ++ti) {
        const int nkb = k_blocks_of(jt);
        for (int pass = 0; pass < 2; ++pass, ++ev) {
          oz_mbar_wait(acc_empty, (uint32_t)((ev & 1) ^ 1));
          asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
          for (int kb = 0; kb < nkb; ++kb, ++it) {
            const int s = it % O3_ST;
            const uint32_t ph = (uint32_t)((it / O3_ST) & 1);
            if (!(batch.dbg & 4)) oz_mbar_wait(full0 + 8 * s, ph);
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
            const uint32_t sa = base + (uint32_t)s * O3_STAGE_BYTES;
            const uint32_t acc0 = (kb > 0) ? 1u : 0u;
            if (pass == 0) {
              const uint64_t dA0 = oz_desc(sa, 128 * 16, 128), dB0 = oz_desc(sa + O3_A_BYTES, 64 * 16, 128);
              if (!(batch.dbg & 1))
#pragma unroll
              for (int pa = OZ_PLANES - 1; pa >= 0; --pa) {
#pragma unroll
                for (int pb = OZ_PLANES - 1; pb >= 0; --pb) {
                  if (pa + pb < 6 || pa + pb > 9) continue;
                  const int lvl = pa + pb - 6;
                  oz_mma_i8_2sm(tmem_base + (uint32_t)(lvl * O2_BN), dA0 + (uint64_t)((pa * O2_PLANE) >> 4),
                                dB0 + (uint64_t)((pb * O3_B_PLANE) >> 4), idesc, (pa == OZ_PLANES - 1) ? acc0 : 1u);
                }
              }
            } else {
              const uint64_t dA0 = oz_desc(sa, 128 * 16, 128), dB0 = oz_desc(sa + O3_HI_A_BYTES, 64 * 16, 128);
              if (!(batch.dbg & 1))
#pragma unroll
              for (int pa = OZ_PLANES - 1; pa >= OZ_PLANES - O2_HI_PLANES; --pa) {
#pragma unroll
                for (int pb = OZ_PLANES - 1; pb >= OZ_PLANES - O2_HI_PLANES; --pb) {
                  if (pa + pb < 10) continue;
                  const int lvl = pa + pb - 10;
                  oz_mma_i8_2sm(tmem_base + (uint32_t)(lvl * O2_BN), dA0 + (uint64_t)(((pa - 4) * O2_PLANE) >> 4),
                                dB0 + (uint64_t)(((pb - 4) * O3_B_PLANE) >> 4), idesc, (pa == OZ_PLANES - 1) ? acc0 : 1u);
                }
              }
            }
            if (!(batch.dbg & 4)) oz_commit_2sm(empty0 + 8 * s);    // frees the stage in both CTAs
          }
          oz_commit_2sm(acc_full);            // both CTAs' epilogues
        }
      }
    }
  } else {
    // ---------------- epilogue (both CTAs, own TMEM): warps 2..9 ----------------
    const int quarter = warp & 3, half = (warp - 2) >> 2, ewarp = warp - 2;
    const int r_local = quarter * 32 + lane;
    const int row = row0 + r_local;
    constexpr int q = Q;
    const int lane_base = lane & ~(q - 1);
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (smid >= O2_SCRATCH_SLOTS) __trap();
    const uint64_t scr_policy = (batch.dbg & 16) ? oz_policy_evict_last() : 0ull;
    long long* scr = batch.scratch + (((size_t)smid * 8 + ewarp) * 64) * 32 + lane;
    const uint32_t acc_empty_leader = oz_mapa(acc_empty, 0);
    double g[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) g[j] = 0.0;
    const double sA = item.scaleA * 6.103515625e-05;   // 2^-14
    const uint32_t tcol0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 64);
    int ev = 0;
    for (int ti = 0, jt; (jt = tile_at(ti)) >= 0; ++ti) {
      const int n0 = jt * O2_BN + half * 64;
      sB_s[ewarp][lane] = item.scaleB[n0 + lane] * sA;
      sB_s[ewarp][lane + 32] = item.scaleB[n0 + 32 + lane] * sA;
      __syncwarp();
      oz_mbar_wait(acc_full, (uint32_t)(ev & 1)); ++ev;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        long long lo[32];
#pragma unroll
        for (int lvl = 3; lvl >= 0; --lvl) {
          uint32_t v[32];
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            const uint32_t taddr = tcol0 + (uint32_t)(lvl * O2_BN + h * 32 + c0);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                         : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
                           "=r"(v[c0 + 6]), "=r"(v[c0 + 7]), "=r"(v[c0 + 8]), "=r"(v[c0 + 9]), "=r"(v[c0 + 10]), "=r"(v[c0 + 11]),
                           "=r"(v[c0 + 12]), "=r"(v[c0 + 13]), "=r"(v[c0 + 14]), "=r"(v[c0 + 15])
                         : "r"(taddr));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
          for (int c = 0; c < 32; ++c) lo[c] = (lvl == 3) ? (long long)(int)v[c] : lo[c] * 256 + (long long)(int)v[c];
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) oz_st_keep(scr + (size_t)(h * 32 + c) * 32, lo[c], scr_policy);
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) oz_mbar_arrive_cluster(acc_empty_leader);
      oz_mbar_wait(acc_full, (uint32_t)(ev & 1)); ++ev;
      asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
      long long hi[64];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int lvl = 2; lvl >= 0; --lvl) {
          uint32_t v[32];
#pragma unroll
          for (int c0 = 0; c0 < 32; c0 += 16) {
            const uint32_t taddr = tcol0 + (uint32_t)(lvl * O2_BN + h * 32 + c0);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
                         : "=r"(v[c0 + 0]), "=r"(v[c0 + 1]), "=r"(v[c0 + 2]), "=r"(v[c0 + 3]), "=r"(v[c0 + 4]), "=r"(v[c0 + 5]),
                           "=r"(v[c0 + 6]), "=r"(v[c0 + 7]), "=r"(v[c0 + 8]), "=r"(v[c0 + 9]), "=r"(v[c0 + 10]), "=r"(v[c0 + 11]),
                           "=r"(v[c0 + 12]), "=r"(v[c0 + 13]), "=r"(v[c0 + 14]), "=r"(v[c0 + 15])
                         : "r"(taddr));
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
#pragma unroll
          for (int c = 0; c < 32; ++c)
            hi[h * 32 + c] = (lvl == 2) ? (long long)(int)v[c] : hi[h * 32 + c] * 256 + (long long)(int)v[c];
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
      __syncwarp();
      if (lane == 0) oz_mbar_arrive_cluster(acc_empty_leader);
      if (!(batch.dbg & 2)) {
#pragma unroll
        for (int c8 = 0; c8 < 64; c8 += 8) {
          long long lo[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) lo[k] = oz_ld_keep(scr + (size_t)(c8 + k) * 32, scr_policy);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int c = c8 + k;
            const double sum = fma((double)lo[k], 3.552713678800501e-15 /* 2^-48 */, (double)hi[c] * 1.52587890625e-05 /* 2^-16 */);
            const int n = n0 + c;
            const double val = sum * sB_s[ewarp][c];
            oz_gram_update<Q>((n < batch.N) ? val : 0.0, lane_base, g);
            if (n >= batch.N && row < batch.rows) {
              const int e = n - batch.N;
              if (e == 0) item.mu_raw[row] = val;
              else if (item.W && e < batch.n_ext) item.W[(size_t)row * batch.ldw + (e - 1)] = val;
            }
          }
        }
      }
      __syncwarp();
    }
    if (row < batch.rows) {
      double* dst = item.gqq_part + (size_t)(grp * 2 + half) * batch.gqq_stride + ((size_t)(row / q) * q + (row % q)) * q;
      for (int j = 0; j < q; ++j) dst[j] = g[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
  __syncthreads();
  oz_cluster_sync();       // the peer may still read this CTA's shared memory / signal its barriers until here
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(tmem_base), "r"(512));
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiledOz)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int oz_make_map(CUtensorMap* map, const signed char* planes, int rows_alloc, int n_chunks, int box_rows,
                       int box_chunks = OZ_CH, int box_planes = OZ_PLANES) {
  static PFN_encodeTiledOz fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
      bo_set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
      return BO_ERR_CUDA;
    }
    fn = reinterpret_cast<PFN_encodeTiledOz>(p);
  }
  // The [row][16 B] slab of one (plane, K chunk) is contiguous: describe it as ONE inner dimension of 8-byte elements so
  // that a box row is a 1-2 KB contiguous request (a 16-byte inner dimension makes the TMA unit issue 16-byte requests and
  // caps the fill rate at ~13 B/clk per SM).
  cuuint64_t dims[3] = {(cuuint64_t)rows_alloc * 2, (cuuint64_t)n_chunks, OZ_PLANES};
  cuuint64_t strides[2] = {(cuuint64_t)rows_alloc * 16, (cuuint64_t)n_chunks * rows_alloc * 16};
  cuuint32_t box[3] = {(cuuint32_t)box_rows * 2, (cuuint32_t)box_chunks, (cuuint32_t)box_planes};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<signed char*>(planes), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { bo_set_error("cuTensorMapEncodeTiled (int8 planes) failed (%d)", (int)r); return BO_ERR_CUDA; }
  return BO_OK;
}

size_t ozaki_scratch_bytes() { return (size_t)O2_SCRATCH_SLOTS * 8 * 64 * 32 * sizeof(long long); }

size_t ozaki_plane_bytes(int rows_alloc, int ldk) { return (size_t)OZ_PLANES * (ldk / 16) * rows_alloc * 16; }

int launch_ozaki_row_scale(const double* X, int rows, int cols, int ld, double* scale, cudaStream_t s, LaunchCounter* lc) {
  if (rows <= 0) return BO_OK;
  oz_row_scale_kernel<<<(rows * 32 + 255) / 256, 256, 0, s>>>(X, rows, cols, ld, scale);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

int launch_ozaki_slice(const double* X, int rows, int cols, int ld, const double* row_scale, double gscale, signed char* planes,
                       int rows_alloc, int ldk, cudaStream_t s, LaunchCounter* lc) {
  const int n_chunks = ldk / 16;
  const long long n = (long long)rows_alloc * n_chunks;
  if (n <= 0) return BO_OK;
  oz_slice_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(X, rows, cols, ld, row_scale, gscale, planes, rows_alloc, n_chunks);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

size_t ozaki_partial_ws_doubles(int rows, int q, int n_out) { return (size_t)n_out * 2 * OZ_MAXGROUPS * rows * q; }

static int launch_ozaki_gemm2p(const OzakiArgs* args, int n_out, double* part_ws, int n_sm, bool pair, cudaStream_t s, LaunchCounter* lc) {
  const OzakiArgs& a0 = args[0];
  static PerDeviceOnce attr_once; bool& attr_set = *attr_once.slot();
  const size_t smem = pair ? (size_t)O3_ST * O3_STAGE_BYTES + 1024 : (size_t)O2_ST * O2_STAGE_BYTES + 1024;
  if (!attr_set) {
    const size_t smem = (size_t)O2_ST * O2_STAGE_BYTES + 1024;
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm2p_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm2p_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm2p_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm2p_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int smem_pair = O3_ST * O3_STAGE_BYTES + 1024;
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm2c_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pair));
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm2c_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pair));
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm2c_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pair));
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm2c_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_pair));
    attr_set = true;
  }
  // Integer scratch slab of the two-pass kernels, indexed by %smid (each such CTA allocates all 512 TMEM columns, so at most
  // one is resident per SM; only the 148 x 128 KB of the resident CTAs are ever touched: L2-resident).  It belongs to the
  // caller's handle (capi.cu), so that launches of different handles / streams never share a slab.
  long long* scratch = a0.scratch;
  if (!scratch) { bo_set_error("ozaki_gemm: the two-pass kernels need a scratch slab"); return BO_ERR_INVALID; }
  const int row_tiles = (a0.rows + OZ_BM - 1) / OZ_BM;
  const int n_tiles = a0.Rpad / O2_BN;
  int groups;
  {
    static int forced = -2;
    if (forced == -2) { const char* e = getenv("EVEREST_OZAKI_GROUPS"); forced = e ? atoi(e) : -1; }
    const double panel_mb = (double)OZ_PLANES * OZ_BM * a0.ldk / (1 << 20);
    groups = (int)((n_sm * panel_mb + 39.0) / 40.0);
    if (forced >= 1) groups = forced;
    groups = std::max(1, std::min(groups, std::min(n_tiles, OZ_MAXGROUPS)));
  }
  for (int m0 = 0; m0 < n_out; m0 += O2_MAXOUT) {
    const int cnt = std::min(O2_MAXOUT, n_out - m0);
    O2Batch batch;
    memset(&batch, 0, sizeof(batch));
    batch.rows = a0.rows; batch.q = a0.q; batch.N = a0.N; batch.n_ext = a0.n_ext; batch.Rpad = a0.Rpad; batch.ldw = a0.ldw;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("EVEREST_OZAKI_DBG"); dbg = e ? atoi(e) : 0; } batch.dbg = dbg; }
    batch.n_chunks_k = a0.ldk / 16; batch.n_groups = groups; batch.gqq_stride = (long long)a0.rows * a0.q;
    batch.scratch = scratch;
    for (int i = 0; i < cnt; ++i) {
      const OzakiArgs& a = args[m0 + i];
      int rc;
      if ((rc = oz_make_map(&batch.item[i].mapA, a.Aplanes, a.rows_alloc, a.ldk / 16, OZ_BM, O2_CH, OZ_PLANES)) != BO_OK) return rc;
      const int b_rows = pair ? O2_BN / 2 : O2_BN;   // in a CTA pair every CTA loads half of the B tile
      if ((rc = oz_make_map(&batch.item[i].mapB, a.Bplanes, a.Rpad, a.ldk / 16, b_rows, O2_CH, OZ_PLANES)) != BO_OK) return rc;
      if ((rc = oz_make_map(&batch.item[i].mapA3, a.Aplanes, a.rows_alloc, a.ldk / 16, OZ_BM, O2_CH, O2_HI_PLANES)) != BO_OK) return rc;
      if ((rc = oz_make_map(&batch.item[i].mapB3, a.Bplanes, a.Rpad, a.ldk / 16, b_rows, O2_CH, O2_HI_PLANES)) != BO_OK) return rc;
      batch.item[i].scaleB = a.scaleB; batch.item[i].scaleA = a.scaleA;
      batch.item[i].gqq_part = part_ws + (size_t)(m0 + i) * 2 * groups * batch.gqq_stride;
      batch.item[i].W = a.W; batch.item[i].mu_raw = a.mu_raw;
    }
    dim3 grid(row_tiles * groups, cnt);
    if (pair) {
      grid.x = 2 * ((row_tiles + 1) / 2) * groups;
      if (a0.q == 1) ozaki_gemm2c_kernel<1><<<grid, OZ_THREADS, smem, s>>>(batch);
      else if (a0.q == 2) ozaki_gemm2c_kernel<2><<<grid, OZ_THREADS, smem, s>>>(batch);
      else if (a0.q == 4) ozaki_gemm2c_kernel<4><<<grid, OZ_THREADS, smem, s>>>(batch);
      else ozaki_gemm2c_kernel<8><<<grid, OZ_THREADS, smem, s>>>(batch);
    } else
    if (a0.q == 1) ozaki_gemm2p_kernel<1><<<grid, OZ_THREADS, smem, s>>>(batch);
    else if (a0.q == 2) ozaki_gemm2p_kernel<2><<<grid, OZ_THREADS, smem, s>>>(batch);
    else if (a0.q == 4) ozaki_gemm2p_kernel<4><<<grid, OZ_THREADS, smem, s>>>(batch);
    else ozaki_gemm2p_kernel<8><<<grid, OZ_THREADS, smem, s>>>(batch);
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
    for (int i = 0; i < cnt; ++i) {
      int rc = launch_sum_gram_partials(batch.item[i].gqq_part, batch.gqq_stride, 2 * groups, args[m0 + i].Gqq, s, lc);
      if (rc != BO_OK) return rc;
    }
  }
  return BO_OK;
}

int launch_ozaki_gemm(const OzakiArgs* args, int n_out, double* part_ws, int tile, cudaStream_t s, LaunchCounter* lc) {
  if (n_out <= 0 || args[0].rows <= 0) return BO_OK;
  const OzakiArgs& a0 = args[0];
  if (!(a0.q == 1 || a0.q == 2 || a0.q == 4 || a0.q == 8) || a0.rows % a0.q) { bo_set_error("ozaki_gemm: q must be 1, 2, 4 or 8"); return BO_ERR_INVALID; }
  if (a0.ldk % 16 || a0.Rpad % OZ_BN || !part_ws) { bo_set_error("ozaki_gemm: padding / workspace violated"); return BO_ERR_INVALID; }
  static int n_sm = 0;
  if (!n_sm) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev); if (n_sm <= 0) n_sm = 148; }
  static PerDeviceOnce attr_once; bool& attr_set = *attr_once.slot();
  const size_t smem = (size_t)OZ_ST * OZ_STAGE_BYTES + 1024;
  if (!attr_set) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_CHECK_RET(cudaFuncSetAttribute(ozaki_gemm_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const int row_tiles = (a0.rows + OZ_BM - 1) / OZ_BM;
  static int tile_env = 0;
  if (!tile_env) { const char* e = getenv("EVEREST_OZAKI_TILE"); tile_env = e ? atoi(e) : 128; if (tile_env != 64 && tile_env != 256) tile_env = 128; }
  const int tile_n = tile ? tile : tile_env;
  if (tile_n >= 128 && a0.Rpad % O2_BN == 0) return launch_ozaki_gemm2p(args, n_out, part_ws, n_sm, tile_n == 256, s, lc);
  const int n_tiles = a0.Rpad / OZ_BN;
  // column groups: resident K(X*,X) digit panels (n_sm / G panels of 7 * 128 * ldk bytes) should stay in L2 (~40 MB)
  int groups;
  {
    static int forced = -2;
    if (forced == -2) { const char* e = getenv("EVEREST_OZAKI_GROUPS"); forced = e ? atoi(e) : -1; }
    const double panel_mb = (double)OZ_PLANES * OZ_BM * a0.ldk / (1 << 20);
    groups = (int)((n_sm * panel_mb + 39.0) / 40.0);
    if (forced >= 1) groups = forced;
    groups = std::max(1, std::min(groups, std::min(n_tiles, OZ_MAXGROUPS)));
  }
  for (int m0 = 0; m0 < n_out; m0 += OZ_MAXOUT) {
    const int cnt = std::min(OZ_MAXOUT, n_out - m0);
    OzBatch batch;
    memset(&batch, 0, sizeof(batch));
    batch.rows = a0.rows; batch.q = a0.q; batch.N = a0.N; batch.n_ext = a0.n_ext; batch.Rpad = a0.Rpad; batch.ldw = a0.ldw;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("EVEREST_OZAKI_DBG"); dbg = e ? atoi(e) : 0; } batch.dbg = dbg; }
    batch.n_chunks_k = a0.ldk / 16; batch.n_groups = groups; batch.gqq_stride = (long long)a0.rows * a0.q;
    {
      std::vector<long long> pre(n_tiles + 1, 0);
      for (int jt = 0; jt < n_tiles; ++jt) {
        const int n_end = (jt + 1) * OZ_BN;
        const int kbytes = (n_end <= a0.N) ? n_end : a0.ldk;
        pre[jt + 1] = pre[jt] + (std::min(kbytes, a0.ldk) + OZ_BK - 1) / OZ_BK;
      }
      batch.gbeg[0] = 0;
      for (int g = 1; g < groups; ++g) {
        long long target = pre[n_tiles] * g / groups;
        int jt = batch.gbeg[g - 1] + 1;
        while (jt < n_tiles - (groups - g) && pre[jt] < target) ++jt;
        batch.gbeg[g] = jt;
      }
      batch.gbeg[groups] = n_tiles;
    }
    for (int i = 0; i < cnt; ++i) {
      const OzakiArgs& a = args[m0 + i];
      int rc;
      if ((rc = oz_make_map(&batch.item[i].mapA, a.Aplanes, a.rows_alloc, a.ldk / 16, OZ_BM)) != BO_OK) return rc;
      if ((rc = oz_make_map(&batch.item[i].mapB, a.Bplanes, a.Rpad, a.ldk / 16, OZ_BN)) != BO_OK) return rc;
      batch.item[i].scaleB = a.scaleB; batch.item[i].scaleA = a.scaleA;
      batch.item[i].gqq_part = part_ws + (size_t)(m0 + i) * 2 * groups * batch.gqq_stride;
      batch.item[i].W = a.W; batch.item[i].mu_raw = a.mu_raw;
    }
    dim3 grid(row_tiles * groups, cnt);
    if (a0.q == 1) ozaki_gemm_kernel<1><<<grid, OZ_THREADS, smem, s>>>(batch);
    else if (a0.q == 2) ozaki_gemm_kernel<2><<<grid, OZ_THREADS, smem, s>>>(batch);
    else if (a0.q == 4) ozaki_gemm_kernel<4><<<grid, OZ_THREADS, smem, s>>>(batch);
    else ozaki_gemm_kernel<8><<<grid, OZ_THREADS, smem, s>>>(batch);
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
    for (int i = 0; i < cnt; ++i) {
      int rc = launch_sum_gram_partials(batch.item[i].gqq_part, batch.gqq_stride, 2 * groups, args[m0 + i].Gqq, s, lc);
      if (rc != BO_OK) return rc;
    }
  }
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// Per-row guard of the digit-plane product (automatic mode).  The fixed-point planes are accurate relative to the ROW
// SCALES, not to the entries, so the absolute error of one element of V = K(X*,X) LinvExt^T is
//     dV ~ 2^-56.2 sqrt(K) sA sB[n]      (6 dropped pairs of level p + p' = 5 with digit products of RMS 2^12.4, K terms of
//                                          random sign, plus the 2^-56 rounding of the operands to the fixed-point grid)
// whatever the size of V itself.  What the parity bar is stated on is the posterior variance k** - G_ii, G_ii = sum_n V_in^2:
//     |dG_ii| <~ 2 sqrt(G_ii) eps,   eps = kappa 2^-56 sqrt(N) sA max_n sB[n]   (kappa = 8: eight standard deviations),
// and the posterior mean (column N, its own row scale).  A q-batch whose rows do not satisfy
//     2 sqrt(G_ii) eps + N eps^2 <= tol (k** - G_ii)   and   eps_mu <= tol (1 + |mu|)          (tol = 1e-10)
// -- candidates next to training points, where the variance is many orders below the prior -- is flagged; the host redoes
// exactly those q-batches with the FP64 DMMA kernel (capi.cu).  Measured error / this estimate: profiles/r02_guard_calibration.txt.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
oz_scale_max_kernel(const double* __restrict__ scaleB, int N, double* __restrict__ out2) {
  // out2[0] = max_{n < N} scaleB[n], out2[1] = scaleB[N] (the mean-cache row)
  __shared__ double red[256];
  double mx = 0.0;
  for (int n = threadIdx.x; n < N; n += 256) mx = fmax(mx, scaleB[n]);
  red[threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 256; ++i) mx = fmax(mx, red[i]);
    out2[0] = mx;
    out2[1] = scaleB[N];
  }
}

__global__ void __launch_bounds__(256)
oz_guard_kernel(const double* __restrict__ Gqq, const double* __restrict__ mu_raw, int rows, int q, int N, double kmax,
                double scaleA, const double* __restrict__ sb2, double kappa, double tol, int* __restrict__ flags,
                int* __restrict__ list, int* __restrict__ count) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const int batch = r / q, j = r - batch * q;
  const double c0 = kappa * 1.3877787807814457e-17 /* 2^-56 */ * sqrt((double)N) * scaleA;
  const double eps = c0 * sb2[0], eps_mu = c0 * sb2[1];
  const double g = Gqq[((size_t)batch * q + j) * q + j];
  const double var = kmax - g;
  const double err = 2.0 * sqrt(fmax(g, 0.0)) * eps + (double)N * eps * eps;
  const bool ok = (err <= tol * var) && (eps_mu <= tol * (1.0 + fabs(mu_raw[r])));   // NaN / non-positive variance -> flagged
  if (!ok && atomicExch(flags + batch, 1) == 0) list[atomicAdd(count, 1)] = batch;
}

// q-batches `list[0..n)` of X [b, q, d] -> Xg [n, q, d].  With `count` (device) only the first *count entries of the list
// are real: the other slots are filled with a valid q-batch (the first flagged one, or q-batch 0) so that the FP64 chain
// can be launched with a fixed capacity before the host knows the count.
__global__ void __launch_bounds__(256)
oz_gather_x_kernel(const double* __restrict__ X, const int* __restrict__ list, const int* __restrict__ count, int n, int qd,
                   double* __restrict__ Xg) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * qd) return;
  const int t = (int)(idx / qd), e = (int)(idx - (long long)t * qd);
  int src = list[t];
  if (count) { const int c = *count; src = (t < c) ? list[t] : (c > 0 ? list[0] : 0); }
  Xg[idx] = X[(size_t)src * qd + e];
}

// redone Gram / W / mean rows of the flagged q-batches back into the full arrays
__global__ void __launch_bounds__(256)
oz_scatter_kernel(const int* __restrict__ list, const int* __restrict__ count, int n, int q, int n_w, int ldw,
                  const double* __restrict__ Gg, const double* __restrict__ Wg, const double* __restrict__ mug,
                  double* __restrict__ Gqq, double* __restrict__ W, double* __restrict__ mu_raw) {
  const int per = q * q + q * n_w + q;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)n * per) return;
  const int t = (int)(idx / per);
  if (count && t >= *count) return;
  int e = (int)(idx - (long long)t * per);
  const size_t src_b = (size_t)t, dst_b = (size_t)list[t];
  if (e < q * q) { Gqq[dst_b * q * q + e] = Gg[src_b * q * q + e]; return; }
  e -= q * q;
  if (e < q * n_w) {
    const int j = e / n_w, c = e - j * n_w;
    W[(dst_b * q + j) * ldw + c] = Wg[(src_b * q + j) * ldw + c];
    return;
  }
  e -= q * n_w;
  mu_raw[dst_b * q + e] = mug[src_b * q + e];
}

int launch_ozaki_scale_max(const double* scaleB, int N, double* out2, cudaStream_t s, LaunchCounter* lc) {
  oz_scale_max_kernel<<<1, 256, 0, s>>>(scaleB, N, out2);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

int launch_ozaki_guard(const double* Gqq, const double* mu_raw, int rows, int q, int N, double kmax, double scaleA,
                       const double* sb2, double kappa, double tol, int* flags, int* list, int* count, cudaStream_t s,
                       LaunchCounter* lc) {
  if (rows <= 0) return BO_OK;
  oz_guard_kernel<<<(rows + 255) / 256, 256, 0, s>>>(Gqq, mu_raw, rows, q, N, kmax, scaleA, sb2, kappa, tol, flags, list, count);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

int launch_ozaki_gather_x(const double* X, const int* list, const int* count_dev, int n, int qd, double* Xg, cudaStream_t s,
                          LaunchCounter* lc) {
  if (n <= 0) return BO_OK;
  const long long tot = (long long)n * qd;
  oz_gather_x_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(X, list, count_dev, n, qd, Xg);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

int launch_ozaki_scatter(const int* list, const int* count_dev, int n, int q, int n_w, int ldw, const double* Gg, const double* Wg,
                         const double* mug, double* Gqq, double* W, double* mu_raw, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return BO_OK;
  const long long tot = (long long)n * (q * q + q * n_w + q);
  oz_scatter_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, s>>>(list, count_dev, n, q, n_w, ldw, Gg, Wg, mug, Gqq, W, mu_raw);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}
