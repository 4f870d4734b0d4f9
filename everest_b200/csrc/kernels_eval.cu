// K1 + K2: input transforms / packing and the fused cross-covariance kernel.
//
// Replaces (reference, CPU float64 via BoTorch/GPyTorch):
//   * Normalize / OneHotToNumeric input transforms built in bofire/surrogates/utils.py:103-164 and
//     surrogates/mixed_single_task_gp.py:82-88  -> prep_points_kernel
//   * kernel evaluation K(X*, X) for RBF / Matern / Hamming / Tanimoto and Scale/Add/Mul trees
//     (bofire/kernels/mapper.py:31-253, kernels/categorical.py:43-70,
//     fingerprint_kernels/base_fingerprint_kernel.py:36-53)     -> crosscov_kernel
//
// crosscov_kernel: 64 x 64 output tile per CTA, 8 warps, each warp a 16 x 32 sub-tile made of
// 2 x 4 DMMA m8n8k4 accumulators for the a.b contraction of continuous leaves (FP64 tensor pipe);
// Tanimoto uses AND + POPC over bit-packed words, Hamming compares integer codes.  Distances,
// lengthscales (pre-divided coordinates), outputscales (term coefficients) and the kernel function
// are applied in the same pass; HBM traffic is one 8-byte store per element.
#include <string.h>

#include <algorithm>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------
// prep: one warp per point
// ------------------------------------------------------------------------------------------------
__global__ void prep_points_kernel(ModelD md, const double* __restrict__ X, int n, int d, PrepD prep) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (warp >= n) return;
  const double* x = X + (size_t)warp * d;
  for (int l = 0; l < md.n_leaves; ++l) {
    const LeafD& L = md.leaf[l];
    if (L.kind <= BO_LEAF_MATERN52) {
      double* o = prep.Xs[l] + (size_t)warp * L.dpad;
      double acc = 0.0;
      for (int k = lane; k < L.dpad; k += 32) {
        double u = 0.0;
        if (k < L.nd) {
          double t = (x[L.col[k]] - L.in_off[k]) / L.in_scl[k];
          u = (t - L.center[k]) / L.ls[k];
        }
        o[k] = u;
        acc = fma(u, u, acc);
      }
      acc = warp_sum(acc);
      if (lane == 0) prep.n2[l][warp] = acc;
    } else if (L.kind == BO_LEAF_HAMMING) {
      for (int f = lane; f < L.nd; f += 32) {
        int s = L.col[f], c = L.card[f];
        int best = 0;
        double bv = x[s];
        for (int j = 1; j < c; ++j) {
          double v = x[s + j];
          if (v > bv) { bv = v; best = j; }
        }
        prep.codes[l][(size_t)warp * L.nd + f] = best;
      }
    } else {  // Tanimoto: pack 64 columns per word with two ballots; the same bits as 0 / 1 bytes behind the words
      int total = 0;
      const int rb = tanimoto_row_bytes(L.dpad);
      unsigned char* by = const_cast<unsigned char*>(tanimoto_bytes(prep.bits[l], n, L.dpad)) + (size_t)warp * rb;
      for (int w = 0; w < L.dpad; ++w) {
        int k0 = w * 64 + lane, k1 = k0 + 32;
        bool b0 = (k0 < L.nd) && (x[L.col[k0]] != 0.0);
        bool b1 = (k1 < L.nd) && (x[L.col[k1]] != 0.0);
        unsigned lo = __ballot_sync(0xffffffffu, b0);
        unsigned hi = __ballot_sync(0xffffffffu, b1);
        u64 word = ((u64)hi << 32) | (u64)lo;
        total += __popcll(word);
        if (lane == 0) prep.bits[l][(size_t)warp * L.dpad + w] = word;
        by[k0] = b0 ? 1 : 0;
        by[k1] = b1 ? 1 : 0;
      }
      for (int k = L.dpad * 64 + lane; k < rb; k += 32) by[k] = 0;
      if (lane == 0) prep.pc[l][warp] = total;
    }
  }
}

int launch_prep_points(const ModelD& md, const double* X, int n, int d, PrepD prep, cudaStream_t s, LaunchCounter* lc) {
  if (n <= 0) return BO_OK;
  int threads = 256;
  int blocks = (n * 32 + threads - 1) / threads;
  prep_points_kernel<<<blocks, threads, 0, s>>>(md, X, n, d, prep);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// Packed wire format of the host entry points (capi.cu forward_host_impl) -> dense rows: column c of row r comes from
// dense[r, src[c]] (src[c] >= 0) or from bit -(src[c] + 1) of the row's fingerprint words.
__global__ void unpack_rows_kernel(const double* __restrict__ dense, const u64* __restrict__ bits, int nd, int W,
                                   const int* __restrict__ src, int rows, int d, double* __restrict__ X) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (size_t)rows * d) return;
  const size_t r = i / d;
  const int c = (int)(i - r * d);
  const int sidx = src[c];
  double v;
  if (sidx >= 0) v = dense[r * nd + sidx];
  else {
    const int k = -(sidx + 1);
    v = (double)((bits[r * W + (k >> 6)] >> (k & 63)) & 1ull);
  }
  X[i] = v;
}

int launch_unpack_rows(const double* dense, const u64* bits, int nd, int W, const int* src, int rows, int d, double* X,
                       cudaStream_t s, LaunchCounter* lc) {
  if (rows <= 0) return BO_OK;
  const size_t n = (size_t)rows * d;
  unpack_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dense, bits, nd, W, src, rows, d, X);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// ------------------------------------------------------------------------------------------------
// crosscov
// ------------------------------------------------------------------------------------------------
#define CC_TILE 64
#define CC_KC 32
#define CC_LDS 36  // CC_KC + 4 -> (stride mod 16) == 4: conflict-free 64-bit fragment loads

struct ColSides { LeafSide s[BO_MAX_LEAVES]; };
#define CC2_DSM (2 * 2 * CC_TILE * 272)   // crosscov_kernel2: two buffers of (A | B) x 64 rows x 272 bytes

__global__ void __launch_bounds__(256, 2)
crosscov_kernel2(ModelD md, PrepD rows, ColSides cols, int n_cols, double* __restrict__ out, int ld, int same_set) {
  __shared__ __align__(16) double smem[2 * CC_TILE * CC_LDS];
  extern __shared__ __align__(16) unsigned char dsm[];   // Tanimoto chunk ring (launcher: CC2_DSM bytes when the tree has such a leaf)
  double* As = smem;
  double* Bs = smem + CC_TILE * CC_LDS;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp & 3, wc = warp >> 2;  // 4 row groups of 16, 2 col groups of 32
  const int row0 = blockIdx.y * CC_TILE, col0 = blockIdx.x * CC_TILE;
  const int n_rows = rows.n;

  double total[16], prod[16];
#pragma unroll
  for (int e = 0; e < 16; ++e) total[e] = 0.0;

  // Evaluation order.  Kernel trees of the form  sum of single leaves (+ ONE product of leaves)  -- SingleTaskGP, Additive,
  // MixedSingleTaskGP / MixedTanimotoGP (mixed_tanimoto_gp.py:165-215: (s1 Kc + s2 Km + s3 Kh) + (s4 Kc s5 Km s6 Kh)) -- are
  // evaluated LEAF-major: every leaf (2048-bit Tanimoto: 64 POPCs per pair) is computed once and feeds both its own term
  // and the product.  Anything else keeps the term-major order (a leaf is re-evaluated in every term that contains it).
  bool leaf_major = true;
  int multi = -1, n_fac_total = 0;
  for (int t2 = 0; t2 < md.n_terms; ++t2) {
    n_fac_total += md.nfac[t2];
    if (md.nfac[t2] != 1) {
      if (multi >= 0 || md.nfac[t2] < 1) leaf_major = false;
      multi = t2;
    }
  }
  if (leaf_major && multi >= 0)
    for (int f1 = 0; f1 < md.nfac[multi]; ++f1)
      for (int f2 = f1 + 1; f2 < md.nfac[multi]; ++f2)
        if (md.fac[multi][f1] == md.fac[multi][f2]) leaf_major = false;
  const int n_items = leaf_major ? md.n_leaves : n_fac_total;
  // role of every leaf in the leaf-major order, found once per CTA (it was re-derived by every thread for every leaf):
  // coefficient of the single-leaf terms that hold it, and whether the product term holds it
  __shared__ double s_n2[2 * CC_TILE];      // squared norms of the tile's rows | columns (current continuous leaf)
  __shared__ int s_pc[2 * CC_TILE];         // popcounts of the tile's rows | columns (current Tanimoto leaf)
  __shared__ double s_csingle[BO_MAX_LEAVES];
  __shared__ int s_role[BO_MAX_LEAVES];   // bit 0: has a single-leaf term, bit 1: factor of the product
  if (leaf_major && tid < md.n_leaves) {
    double c = 0.0;
    int role = 0;
    for (int t2 = 0; t2 < md.n_terms; ++t2)
      if (md.nfac[t2] == 1 && md.fac[t2][0] == tid) { c += md.coef[t2]; role |= 1; }
    if (multi >= 0)
      for (int f2 = 0; f2 < md.nfac[multi]; ++f2)
        if (md.fac[multi][f2] == tid) role |= 2;
    s_csingle[tid] = c;
    s_role[tid] = role;
  }
  __syncthreads();
  int term = 0, f = 0;
  while (!leaf_major && term < md.n_terms && md.nfac[term] < 1) {   // (a constant term: no factors)
#pragma unroll
    for (int e = 0; e < 16; ++e) total[e] += md.coef[term];
    ++term;
  }
#pragma unroll
  for (int e = 0; e < 16; ++e)
    prod[e] = leaf_major ? ((multi >= 0) ? md.coef[multi] : 0.0) : ((term < md.n_terms) ? md.coef[term] : 0.0);

  for (int it = 0; it < n_items; ++it) {
    int l;
    double c_single = 0.0;
    bool single = false, in_prod = false;
    if (leaf_major) {
      l = it;
      c_single = s_csingle[l];
      single = (s_role[l] & 1) != 0;
      in_prod = (s_role[l] & 2) != 0;
      if (!single && !in_prod) continue;
    } else {
      l = md.fac[term][f];
    }
    {
      const LeafD& L = md.leaf[l];
      double lv[16];
      if (L.kind <= BO_LEAF_MATERN52) {
        double acc[2][4][2];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        const double* Ag = rows.Xs[l];
        const double* Bg = cols.s[l].Xs;
        for (int k0 = 0; k0 < L.dpad; k0 += CC_KC) {
          const int kc = min(CC_KC, L.dpad - k0);
          __syncthreads();
          // the norms of the tile's points: one coalesced load per CTA instead of 10 scattered global loads per thread behind
          // the MMA loop (written behind the barrier above -- the previous leaf may still have been reading its own --
          // and published by the barrier that ends the staging)
          if (k0 == 0 && tid < 2 * CC_TILE) {
            const int r = tid & (CC_TILE - 1);
            s_n2[tid] = (tid < CC_TILE) ? ((row0 + r < n_rows) ? rows.n2[l][row0 + r] : 0.0)
                                        : ((col0 + r < n_cols) ? cols.s[l].n2[col0 + r] : 0.0);
          }
          // 16-byte pieces of the kc (a multiple of 4) columns the MMA loop reads; nothing beyond kc is touched
          const int kh = kc >> 1;
          for (int idx = tid; idx < CC_TILE * kh; idx += 256) {
            const int r = (kh == CC_KC / 2) ? (idx >> 4) : (idx / kh), k = (idx - r * kh) * 2;
            double2 va = make_double2(0.0, 0.0), vb = make_double2(0.0, 0.0);
            if (row0 + r < n_rows) va = *reinterpret_cast<const double2*>(Ag + (size_t)(row0 + r) * L.dpad + k0 + k);
            if (col0 + r < n_cols) vb = *reinterpret_cast<const double2*>(Bg + (size_t)(col0 + r) * L.dpad + k0 + k);
            *reinterpret_cast<double2*>(As + r * CC_LDS + k) = va;
            *reinterpret_cast<double2*>(Bs + r * CC_LDS + k) = vb;
          }
          __syncthreads();
          for (int kk = 0; kk < kc; kk += 4) {
            double a[2], b[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) a[i] = As[(wr * 16 + i * 8 + g) * CC_LDS + kk + t];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(wc * 32 + j * 8 + g) * CC_LDS + kk + t];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
              for (int j = 0; j < 4; ++j) mma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
          }
        }
        // squared distances first, then ONE switch over the kernel function around the 16 values: with the switch inside the
        // value loop the executed path hopped over 16 scattered copies of the other four functions (ncu: 0.86 `no_instruction`
        // stalls per issue in this 13 K-instruction kernel)
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          int r = row0 + wr * 16 + i * 8 + g;
          double na = s_n2[wr * 16 + i * 8 + g];
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              int c = col0 + wc * 32 + j * 8 + 2 * t + e;
              double nb = s_n2[CC_TILE + wc * 32 + j * 8 + 2 * t + e];
              double stat = fmax(na + nb - 2.0 * acc[i][j][e], 0.0);
              if (same_set && r == c) stat = 0.0;
              lv[(i * 4 + j) * 2 + e] = stat;
            }
        }
        switch (L.kind) {
          case BO_LEAF_RBF:
#pragma unroll
            for (int e = 0; e < 16; ++e) lv[e] = leaf_value_from_stat(BO_LEAF_RBF, lv[e]);
            break;
          case BO_LEAF_MATERN12:
#pragma unroll
            for (int e = 0; e < 16; ++e) lv[e] = leaf_value_from_stat(BO_LEAF_MATERN12, lv[e]);
            break;
          case BO_LEAF_MATERN32:
#pragma unroll
            for (int e = 0; e < 16; ++e) lv[e] = leaf_value_from_stat(BO_LEAF_MATERN32, lv[e]);
            break;
          default:
#pragma unroll
            for (int e = 0; e < 16; ++e) lv[e] = leaf_value_from_stat(BO_LEAF_MATERN52, lv[e]);
            break;
        }
      } else if (L.kind == BO_LEAF_HAMMING) {
        int* Ac = reinterpret_cast<int*>(smem);
        int* Bc = Ac + CC_TILE * BO_MAX_GROUPS;
        __syncthreads();
        for (int idx = tid; idx < CC_TILE * L.nd; idx += 256) {
          int r = idx / L.nd, f = idx % L.nd;
          Ac[r * BO_MAX_GROUPS + f] = (row0 + r < n_rows) ? rows.codes[l][(size_t)(row0 + r) * L.nd + f] : 0;
          Bc[r * BO_MAX_GROUPS + f] = (col0 + r < n_cols) ? cols.s[l].codes[(size_t)(col0 + r) * L.nd + f] : 0;
        }
        // A leaf over nd groups takes only 2^nd distinct values (one per mismatch pattern): for nd <= 6 they are tabulated
        // once per CTA with exactly the per-pair arithmetic (weights added in group order, the matching groups add 0.0), and
        // a pair costs nd compares and one shared-memory load instead of an FP64 division (slow path whenever all groups
        // match: 0 / nd) and an exp.
        double* Ht = reinterpret_cast<double*>(Bc + CC_TILE * BO_MAX_GROUPS);
        int* Apk = reinterpret_cast<int*>(Ht + 64);     // codes of a point as nibbles of one word: the mismatch pattern of a
        int* Bpk = Apk + CC_TILE;                       // pair is read off the XOR of two registers
        const bool tabulated = L.nd <= 6;
        if (tabulated && tid < (1 << L.nd)) {
          double acc = 0.0;
          for (int f = 0; f < L.nd; ++f) acc += ((tid >> f) & 1) ? L.wls[f] : 0.0;
          Ht[tid] = exp_nonpos(-(acc / (double)L.nd));
        }
        int wide_code = 0;                              // a code that does not fit a nibble (cardinality > 16)
        if (tabulated && tid < 2 * CC_TILE) {
          const int r = tid & (CC_TILE - 1);
          const bool rowside = tid < CC_TILE;
          const int gr = rowside ? row0 + r : col0 + r;
          const bool in = rowside ? (gr < n_rows) : (gr < n_cols);
          const int* cp = (rowside ? rows.codes[l] : cols.s[l].codes) + (size_t)(in ? gr : 0) * L.nd;
          int pk = 0;
          for (int f = 0; f < L.nd; ++f) {
            const int c = in ? cp[f] : 0;
            wide_code |= (c > 15);
            pk |= (c & 15) << (4 * f);
          }
          (rowside ? Apk : Bpk)[r] = pk;
        }
        const int any_wide = __syncthreads_or(wide_code);   // (also the barrier that publishes the staged codes and tables)
        const bool packed = tabulated && !any_wide;
        if (packed) {
          const int ra[2] = {Apk[wr * 16 + g], Apk[wr * 16 + 8 + g]};
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              const int cb = Bpk[wc * 32 + j * 8 + 2 * t + e];
#pragma unroll
              for (int i = 0; i < 2; ++i) {
                // bit f of the mismatch pattern = (nibble f of the XOR is not zero), without a loop over the groups: fold each
                // nibble onto its lowest bit, gather the bits of two neighbouring nibbles per byte, then the three bytes
                const unsigned x = (unsigned)(ra[i] ^ cb);
                const unsigned y = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x111111u;
                const unsigned y2 = (y | (y >> 3)) & 0x030303u;
                const int mask = (int)((y2 | (y2 >> 6) | (y2 >> 12)) & 0x3fu);
                lv[(i * 4 + j) * 2 + e] = Ht[mask];
              }
            }
        } else if (tabulated) {
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                int rl = wr * 16 + i * 8 + g, cl = wc * 32 + j * 8 + 2 * t + e;
                int mask = 0;
                for (int f = 0; f < L.nd; ++f)
                  mask |= (Ac[rl * BO_MAX_GROUPS + f] != Bc[cl * BO_MAX_GROUPS + f]) ? (1 << f) : 0;
                lv[(i * 4 + j) * 2 + e] = Ht[mask];
              }
        } else {
#pragma unroll
          for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                int rl = wr * 16 + i * 8 + g, cl = wc * 32 + j * 8 + 2 * t + e;
                double acc = 0.0;
                for (int f = 0; f < L.nd; ++f)
                  acc += (Ac[rl * BO_MAX_GROUPS + f] != Bc[cl * BO_MAX_GROUPS + f]) ? L.wls[f] : 0.0;
                lv[(i * 4 + j) * 2 + e] = exp_nonpos(-(acc / (double)L.nd));
              }
        }
      } else {  // Tanimoto
        // <x, x'> over 0 / 1 fingerprints = an exact integer GEMM: u8 x u8 -> s32 tensor-core MMAs (m16n8k32) on the byte
        // copies of the fingerprints instead of AND + POPC over the packed words (POPC issues at 1/8 rate: the POPC loop was
        // 41 % of this kernel's instructions and held it at 0.22 of even the POPC-pipe bound).  Chunks of 256 columns per
        // row; row stride 272 bytes keeps the ldmatrix rows (16 bytes each) on distinct banks.
        // The chunks arrive by cp.async into two buffers: chunk k + 1 is in flight while the MMAs of chunk k run, one
        // barrier per chunk (the copy into a buffer is issued after the barrier that ends its previous use).
        const int TB_CH = 256, TB_LD = 272;
        unsigned char* Tb = reinterpret_cast<unsigned char*>(dsm);           // [2 buffers][A | B][64 rows][TB_LD]
        const int rb = tanimoto_row_bytes(L.dpad);
        const unsigned char* Ag = tanimoto_bytes(rows.bits[l], n_rows, L.dpad);
        const unsigned char* Bg = tanimoto_bytes(cols.s[l].bits, n_cols, L.dpad);
        int dot[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) dot[e] = 0;
        // ldmatrix lane addressing (8 x 8 matrices of 16-bit = 8 rows x 16 bytes): matrix m = lane >> 3, row = lane & 7
        const int lm = lane >> 3, lr = lane & 7;
        const unsigned a_off = (unsigned)((wr * 16 + (lm & 1) * 8 + lr) * TB_LD + (lm >> 1) * 16);   // (rows 0-7 | 8-15) x (k 0-15 | 16-31)
        const unsigned b_off = (unsigned)((wc * 32 + (lm >> 1) * 8 + lr) * TB_LD + (lm & 1) * 16);  // (n 0-7: k lo, k hi | n 8-15: k lo, k hi)
        const unsigned Tb_u = (unsigned)__cvta_generic_to_shared(Tb);
        const unsigned buf_bytes = 2 * CC_TILE * TB_LD;
        // this thread's 16-byte pieces of a chunk: piece idx = tid + 256 * i -> row idx / 16, piece idx % 16 (4 per matrix)
        auto issue_chunk = [&](int k0, int buf) {
          unsigned char* Ab = Tb + (size_t)buf * buf_bytes;
          unsigned char* Bb = Ab + CC_TILE * TB_LD;
#pragma unroll
          for (int i = 0; i < CC_TILE * (TB_CH / 16) / 256; ++i) {
            const int idx = tid + 256 * i, r = idx >> 4, c = idx & 15;
            const bool pa = row0 + r < n_rows, pb = col0 + r < n_cols;
            cp_async16(Ab + r * TB_LD + c * 16, Ag + (size_t)(pa ? row0 + r : 0) * rb + k0 + c * 16, pa);
            cp_async16(Bb + r * TB_LD + c * 16, Bg + (size_t)(pb ? col0 + r : 0) * rb + k0 + c * 16, pb);
          }
          cp_async_commit();
        };
        __syncthreads();               // the previous leaf is done with the static tiles; dsm is only used here
        if (tid < 2 * CC_TILE) {       // popcounts of the tile's points (published by the first barrier of the chunk loop)
          const int r = tid & (CC_TILE - 1);
          s_pc[tid] = (tid < CC_TILE) ? ((row0 + r < n_rows) ? rows.pc[l][row0 + r] : 0)
                                      : ((col0 + r < n_cols) ? cols.s[l].pc[col0 + r] : 0);
        }
        issue_chunk(0, 0);
        int buf = 0;
        for (int k0 = 0; k0 < rb; k0 += TB_CH, buf ^= 1) {
          cp_async_wait<0>();
          __syncthreads();             // chunk k0 visible to all; everyone has finished the MMAs of the chunk before it
          if (k0 + TB_CH < rb) issue_chunk(k0 + TB_CH, buf ^ 1);
          const unsigned As_u = Tb_u + buf * buf_bytes, Bs_u = As_u + CC_TILE * TB_LD;
#pragma unroll
          for (int ks = 0; ks < TB_CH; ks += 32) {
            unsigned a0, a1, a2, a3;
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                         : "=r"(a0), "=r"(a1), "=r"(a2), "=r"(a3) : "r"(As_u + a_off + ks));
#pragma unroll
            for (int jp = 0; jp < 2; ++jp) {      // two n8 groups per ldmatrix.x4
              unsigned b0, b1, b2, b3;
              asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
                           : "=r"(b0), "=r"(b1), "=r"(b2), "=r"(b3) : "r"(Bs_u + b_off + jp * 16 * TB_LD + ks));
              // accumulators of an m16n8 tile: (row g, cols 2t, 2t+1), (row g + 8, cols 2t, 2t+1) = lv index (i*4+j)*2+e
              int* d0 = dot + (0 * 4 + jp * 2) * 2;
              int* d1 = dot + (1 * 4 + jp * 2) * 2;
              asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                           : "+r"(d0[0]), "+r"(d0[1]), "+r"(d1[0]), "+r"(d1[1]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
              asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                           : "+r"(d0[2]), "+r"(d0[3]), "+r"(d1[2]), "+r"(d1[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b2), "r"(b3));
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          int pa = s_pc[wr * 16 + i * 8 + g];
#pragma unroll
          for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              int pb = s_pc[CC_TILE + wc * 32 + j * 8 + 2 * t + e];
              lv[(i * 4 + j) * 2 + e] = tanimoto_value(dot[(i * 4 + j) * 2 + e], pa, pb);
            }
        }
      }
      if (leaf_major) {
        if (single) {
#pragma unroll
          for (int e = 0; e < 16; ++e) total[e] += c_single * lv[e];
        }
        if (in_prod) {
#pragma unroll
          for (int e = 0; e < 16; ++e) prod[e] *= lv[e];
        }
      } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) prod[e] *= lv[e];
        if (++f == md.nfac[term]) {
#pragma unroll
          for (int e = 0; e < 16; ++e) total[e] += prod[e];
          ++term; f = 0;
          while (term < md.n_terms && md.nfac[term] < 1) {
#pragma unroll
            for (int e = 0; e < 16; ++e) total[e] += md.coef[term];
            ++term;
          }
          if (term < md.n_terms) {
#pragma unroll
            for (int e = 0; e < 16; ++e) prod[e] = md.coef[term];
          }
        }
      }
    }
  }
  if (leaf_major && multi >= 0) {
#pragma unroll
    for (int e = 0; e < 16; ++e) total[e] += prod[e];
  }

  // store (two adjacent doubles per thread -> 16-byte stores); zero the padding columns [n_cols, ld)
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    int r = row0 + wr * 16 + i * 8 + g;
    if (r >= n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int c = col0 + wc * 32 + j * 8 + 2 * t;
      double v0 = (c < n_cols) ? total[(i * 4 + j) * 2] : 0.0;
      double v1 = (c + 1 < n_cols) ? total[(i * 4 + j) * 2 + 1] : 0.0;
      if (c + 1 < ld) {
        *reinterpret_cast<double2*>(out + (size_t)r * ld + c) = make_double2(v0, v1);
      } else if (c < ld) {
        out[(size_t)r * ld + c] = v0;
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Fast path: one term, one continuous leaf with <= 32 (padded) dims -- the SingleTaskGP default
// (RBF-ARD, or Scale(Matern)).  A CTA keeps its 64 query rows in shared memory and streams a strip of
// CC3_CT column tiles of training rows through a cp.async double buffer; the kernel function is a
// template parameter, so the per-element epilogue is branch-free.
// ------------------------------------------------------------------------------------------------
#define CC3_CT 8

template <int ROWS>
__device__ __forceinline__ void cc3_load_tile(double* dst, const double* __restrict__ src, int dpad, int row0, int n_valid,
                                              int tid) {
  const int chunks = dpad >> 1;  // 16-byte chunks per row
  const int sh = ((chunks & (chunks - 1)) == 0) ? __ffs(chunks) - 1 : -1;   // power of two (dpad = 4, 8, 16, 32): no division
  for (int c = tid; c < ROWS * chunks; c += 256) {
    int r = (sh >= 0) ? (c >> sh) : (c / chunks), ch = c - r * chunks;
    bool p = (row0 + r) < n_valid;
    cp_async16(dst + r * CC_LDS + ch * 2, src + (size_t)(p ? row0 + r : 0) * dpad + ch * 2, p);
  }
}

#define CC3_CLD 66  // C tile row stride (doubles): 16-byte aligned rows, staggered banks

template <int KIND>
__global__ void __launch_bounds__(256, 2)
crosscov_fast_kernel(const double* __restrict__ Aq, const double* __restrict__ n2a_g, int n_rows,
                     const double* __restrict__ Bt, const double* __restrict__ n2b_g, int n_cols, int dpad,
                     double coef, double* __restrict__ out, int ld, int same_set, OzPlanesOut oz, int ct_per_cta) {
  extern __shared__ __align__(16) double cc3sm[];
  double* As = cc3sm;
  double* Bs[2] = {cc3sm + CC_TILE * CC_LDS, cc3sm + 2 * CC_TILE * CC_LDS};
  double* Cs = cc3sm + 3 * CC_TILE * CC_LDS;  // [64][CC3_CLD] raw a.b dot products
  double* n2a_s = Cs + CC_TILE * CC3_CLD;     // [64]
  double* n2b_s = n2a_s + CC_TILE;            // [64]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wr = warp & 3, wc = warp >> 2;
  const int row0 = blockIdx.y * CC_TILE;
  const int ct0 = blockIdx.x * ct_per_cta;
  const int n_ct_total = (ld + CC_TILE - 1) / CC_TILE;
  const int n_ct = min(ct_per_cta, n_ct_total - ct0);

  __shared__ double2 exptab[32];
  exp_tab_load(exptab);                 // visible after the first __syncthreads() of the column-tile loop
  cc3_load_tile<CC_TILE>(As, Aq, dpad, row0, n_rows, tid);
  cc3_load_tile<CC_TILE>(Bs[0], Bt, dpad, ct0 * CC_TILE, n_cols, tid);
  cp_async_commit();
  if (tid < CC_TILE) n2a_s[tid] = (row0 + tid < n_rows) ? n2a_g[row0 + tid] : 0.0;
  // not unrolled: the compiler otherwise replicates the body for the 8 column tiles (9248 instructions = 148 KB of code) and
  // the kernel stalls on instruction fetch (ncu r02: 1.75 `no_instruction` stalls per issue)
#pragma unroll 1
  for (int ci = 0; ci < n_ct; ++ci) {
    const int col0 = (ct0 + ci) * CC_TILE;
    cp_async_wait<0>();
    __syncthreads();  // tile ci landed; previous epilogue finished with Cs / n2b_s
    if (ci + 1 < n_ct) cc3_load_tile<CC_TILE>(Bs[(ci + 1) & 1], Bt, dpad, col0 + CC_TILE, n_cols, tid);
    cp_async_commit();
    if (tid < CC_TILE) n2b_s[tid] = (col0 + tid < n_cols) ? n2b_g[col0 + tid] : 0.0;
    const double* B = Bs[ci & 1];
    double acc[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int kk = 0; kk < dpad; kk += 4) {
      double a[2], b[4];
#pragma unroll
      for (int i = 0; i < 2; ++i) a[i] = As[(wr * 16 + i * 8 + g) * CC_LDS + kk + t];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = B[(wc * 32 + j * 8 + g) * CC_LDS + kk + t];
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) mma_884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        *reinterpret_cast<double2*>(Cs + (wr * 16 + i * 8 + g) * CC3_CLD + wc * 32 + j * 8 + 2 * t) =
            make_double2(acc[i][j][0], acc[i][j][1]);
    __syncthreads();
    if (oz.planes && !oz.write_fp64) {
      // Screens (digit planes only): kernel function and slicing in ONE pass over the tile -- thread = (row, 16-column chunk)
      // reads its 16 raw dot products, evaluates the kernel four values at a time (four independent exp chains) and splits
      // them straight into its 7 x 16 bytes.  The two-pass form below stored the kernel values back to the tile and read them
      // again behind another barrier (ncu r02: shared-memory pipe 45 % busy, issue 42 %).
      const int r = tid & 63, c = tid >> 6;
      const int gr = row0 + r, chunk = (col0 >> 4) + c;
      const double naR = n2a_s[r];
      const bool okR = gr < n_rows;
      signed char dig[OZ_PLANES][16];
      // interior chunks (all 16 columns inside the matrix, no diagonal to pin): no per-value tests
      const bool plain = okR && !same_set && col0 + c * 16 + 16 <= n_cols;
#pragma unroll
      for (int t4 = 0; t4 < 16; t4 += 4) {
        double st[4], lv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int cc = c * 16 + t4 + u;
          st[u] = fmax(naR + n2b_s[cc] - 2.0 * Cs[r * CC3_CLD + cc], 0.0);
          if (!plain && same_set && gr == col0 + cc) st[u] = 0.0;
        }
        leaf_value_from_stat_tab_n<4>(KIND, st, lv, exptab);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double v = (plain || (okR && col0 + c * 16 + t4 + u < n_cols)) ? coef * lv[u] : 0.0;
          signed char d[OZ_PLANES];
          oz_split_digits(v * oz.inv_scale, d);
#pragma unroll
          for (int p = 0; p < OZ_PLANES; ++p) dig[p][t4 + u] = d[p];
        }
      }
      if (gr < oz.rows_cover && chunk < oz.n_chunks) {
#pragma unroll
        for (int p = 0; p < OZ_PLANES; ++p) {
          int4 w;
          memcpy(&w, dig[p], 16);
          *reinterpret_cast<int4*>(oz.planes + (size_t)p * oz.plane_stride +
                                   ((size_t)chunk * oz.rows_alloc + oz.row_offset + gr) * 16) = w;
        }
      }
      continue;
    }
    // compact, coalesced epilogue: one instance of the kernel function in the instruction stream; two row slices (4 kernel
    // values) per iteration so that four independent exp chains are in flight per thread
#pragma unroll 1
    for (int idx = tid; idx < CC_TILE * (CC_TILE / 2); idx += 512) {
      const int c2 = (idx & 31) * 2, gc = col0 + c2;
      const int rA = idx >> 5, rB = rA + 8;
      const int grA = row0 + rA, grB = row0 + rB;
      const double2 dA = *reinterpret_cast<const double2*>(Cs + rA * CC3_CLD + c2);
      const double2 dB = *reinterpret_cast<const double2*>(Cs + rB * CC3_CLD + c2);
      const double nb0 = n2b_s[c2], nb1 = n2b_s[c2 + 1], naA = n2a_s[rA], naB = n2a_s[rB];
      double st[4] = {fmax(naA + nb0 - 2.0 * dA.x, 0.0), fmax(naA + nb1 - 2.0 * dA.y, 0.0),
                      fmax(naB + nb0 - 2.0 * dB.x, 0.0), fmax(naB + nb1 - 2.0 * dB.y, 0.0)};
      if (same_set) {
        if (grA == gc) st[0] = 0.0;
        if (grA == gc + 1) st[1] = 0.0;
        if (grB == gc) st[2] = 0.0;
        if (grB == gc + 1) st[3] = 0.0;
      }
      double lv[4];
      leaf_value_from_stat_tab_n<4>(KIND, st, lv, exptab);
      const bool c0 = gc < n_cols, c1 = gc + 1 < n_cols, okA = grA < n_rows, okB = grB < n_rows;
      const double vA0 = (okA && c0) ? coef * lv[0] : 0.0, vA1 = (okA && c1) ? coef * lv[1] : 0.0;
      const double vB0 = (okB && c0) ? coef * lv[2] : 0.0, vB1 = (okB && c1) ? coef * lv[3] : 0.0;
      if ((!oz.planes || oz.write_fp64) && gc < ld) {
        if (gc + 1 < ld) {
          if (okA) *reinterpret_cast<double2*>(out + (size_t)grA * ld + gc) = make_double2(vA0, vA1);
          if (okB) *reinterpret_cast<double2*>(out + (size_t)grB * ld + gc) = make_double2(vB0, vB1);
        } else {
          if (okA) out[(size_t)grA * ld + gc] = vA0;
          if (okB) out[(size_t)grB * ld + gc] = vB0;
        }
      }
      if (oz.planes) {   // values for the slicing pass (zeros outside the matrix)
        *reinterpret_cast<double2*>(Cs + rA * CC3_CLD + c2) = make_double2(vA0, vA1);
        *reinterpret_cast<double2*>(Cs + rB * CC3_CLD + c2) = make_double2(vB0, vB1);
      }
    }
    if (oz.planes) {
      // fused slicing for the INT8 digit-plane GEMM (ozaki.cu): thread = (row, 16-column chunk) of the 64 x 64 tile,
      // 7 x 16 bytes per thread, consecutive threads -> consecutive rows of one [chunk][row][16 B] slab
      __syncthreads();
      const int r = tid & 63, c = tid >> 6;
      const int gr = row0 + r, chunk = (col0 >> 4) + c;
      if (gr < oz.rows_cover && chunk < oz.n_chunks) {
        signed char dig[OZ_PLANES][16];
#pragma unroll
        for (int t2 = 0; t2 < 16; ++t2) {
          signed char d[OZ_PLANES];
          oz_split_digits(Cs[r * CC3_CLD + c * 16 + t2] * oz.inv_scale, d);
#pragma unroll
          for (int p = 0; p < OZ_PLANES; ++p) dig[p][t2] = d[p];
        }
#pragma unroll
        for (int p = 0; p < OZ_PLANES; ++p) {
          int4 w;
          memcpy(&w, dig[p], 16);
          *reinterpret_cast<int4*>(oz.planes + (size_t)p * oz.plane_stride +
                                   ((size_t)chunk * oz.rows_alloc + oz.row_offset + gr) * 16) = w;
        }
      }
    }
  }
}

int launch_crosscov(const ModelD& md, PrepD rows, PrepD colsOrTrain, bool cols_are_train, int n_cols, double* out,
                    int ld, bool same_set, cudaStream_t s, LaunchCounter* lc) {
  return launch_crosscov_ex(md, rows, colsOrTrain, cols_are_train, n_cols, out, ld, same_set, nullptr, nullptr, s, lc);
}

int launch_crosscov_ex(const ModelD& md, PrepD rows, PrepD colsOrTrain, bool cols_are_train, int n_cols, double* out,
                       int ld, bool same_set, const OzPlanesOut* ozp, bool* fused, cudaStream_t s, LaunchCounter* lc) {
  if (fused) *fused = false;
  OzPlanesOut oz;
  memset(&oz, 0, sizeof(oz));
  if (rows.n <= 0 || n_cols <= 0) return BO_OK;
  if (ld % 2 != 0) { bo_set_error("crosscov: ld must be even"); return BO_ERR_INVALID; }
  ColSides cs;
  for (int l = 0; l < BO_MAX_LEAVES; ++l) {
    if (l < md.n_leaves) {
      if (cols_are_train) {
        const LeafD& L = md.leaf[l];
        cs.s[l].Xs = L.Xs; cs.s[l].n2 = L.n2; cs.s[l].codes = L.codes; cs.s[l].bits = L.bits; cs.s[l].pc = L.pc;
      } else {
        cs.s[l].Xs = colsOrTrain.Xs[l]; cs.s[l].n2 = colsOrTrain.n2[l]; cs.s[l].codes = colsOrTrain.codes[l];
        cs.s[l].bits = colsOrTrain.bits[l]; cs.s[l].pc = colsOrTrain.pc[l];
      }
    } else {
      cs.s[l] = LeafSide{nullptr, nullptr, nullptr, nullptr, nullptr};
    }
  }
  if (md.n_terms == 1 && md.nfac[0] == 1 && md.leaf[md.fac[0][0]].kind <= BO_LEAF_MATERN52 &&
      md.leaf[md.fac[0][0]].dpad <= CC_KC) {
    const int l = md.fac[0][0];
    const LeafD& L = md.leaf[l];
    const int n_ct = (ld + CC_TILE - 1) / CC_TILE;
    // column tiles per CTA: 8 amortise the A tile for the big screens; a refinement step has one or two row tiles, and 8 tiles
    // in sequence per CTA left 4 CTAs working for 47 us -- one tile per CTA until the grid covers the SMs
    const int row_tiles = (std::max(rows.n, 1) + CC_TILE - 1) / CC_TILE;
    int ctp = CC3_CT;
    while (ctp > 1 && (long long)row_tiles * ((n_ct + ctp - 1) / ctp) < 296) ctp >>= 1;
    dim3 gridf((n_ct + ctp - 1) / ctp, (rows.n + CC_TILE - 1) / CC_TILE);
    if (ozp && ozp->planes) {
      oz = *ozp;
      if (oz.rows_cover <= 0) oz.rows_cover = oz.rows_alloc - oz.row_offset;
      gridf.y = (oz.rows_cover + CC_TILE - 1) / CC_TILE;   // the padding rows of the planes are written (zeros) too
      if (fused) *fused = true;
    }
    const double* Aq = rows.Xs[l];
    const double* n2a = rows.n2[l];
    const size_t smf = ((size_t)3 * CC_TILE * CC_LDS + (size_t)CC_TILE * CC3_CLD + 2 * CC_TILE) * sizeof(double);
    static PerDeviceOnce attr_once; bool& attr_set = *attr_once.slot();
    if (!attr_set) {
      CUDA_CHECK_RET(cudaFuncSetAttribute(crosscov_fast_kernel<BO_LEAF_RBF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf));
      CUDA_CHECK_RET(cudaFuncSetAttribute(crosscov_fast_kernel<BO_LEAF_MATERN12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf));
      CUDA_CHECK_RET(cudaFuncSetAttribute(crosscov_fast_kernel<BO_LEAF_MATERN32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf));
      CUDA_CHECK_RET(cudaFuncSetAttribute(crosscov_fast_kernel<BO_LEAF_MATERN52>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smf));
      attr_set = true;
    }
    switch (L.kind) {
      case BO_LEAF_RBF:
        crosscov_fast_kernel<BO_LEAF_RBF><<<gridf, 256, smf, s>>>(Aq, n2a, rows.n, cs.s[l].Xs, cs.s[l].n2, n_cols, L.dpad, md.coef[0], out, ld, same_set ? 1 : 0, oz, ctp);
        break;
      case BO_LEAF_MATERN12:
        crosscov_fast_kernel<BO_LEAF_MATERN12><<<gridf, 256, smf, s>>>(Aq, n2a, rows.n, cs.s[l].Xs, cs.s[l].n2, n_cols, L.dpad, md.coef[0], out, ld, same_set ? 1 : 0, oz, ctp);
        break;
      case BO_LEAF_MATERN32:
        crosscov_fast_kernel<BO_LEAF_MATERN32><<<gridf, 256, smf, s>>>(Aq, n2a, rows.n, cs.s[l].Xs, cs.s[l].n2, n_cols, L.dpad, md.coef[0], out, ld, same_set ? 1 : 0, oz, ctp);
        break;
      default:
        crosscov_fast_kernel<BO_LEAF_MATERN52><<<gridf, 256, smf, s>>>(Aq, n2a, rows.n, cs.s[l].Xs, cs.s[l].n2, n_cols, L.dpad, md.coef[0], out, ld, same_set ? 1 : 0, oz, ctp);
        break;
    }
    if (lc) lc->n++;
    CUDA_CHECK_RET(cudaGetLastError());
    return BO_OK;
  }
  dim3 grid((ld + CC_TILE - 1) / CC_TILE, (rows.n + CC_TILE - 1) / CC_TILE);
  size_t dsm = 0;
  for (int l = 0; l < md.n_leaves; ++l)
    if (md.leaf[l].kind == BO_LEAF_TANIMOTO) dsm = CC2_DSM;   // Tanimoto leaf: two chunk buffers
  if (dsm) {
    static PerDeviceOnce attr2_once; bool& attr2_set = *attr2_once.slot();
    if (!attr2_set) {
      CUDA_CHECK_RET(cudaFuncSetAttribute(crosscov_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CC2_DSM));
      attr2_set = true;
    }
  }
  crosscov_kernel2<<<grid, 256, dsm, s>>>(md, rows, cs, n_cols, out, ld, same_set ? 1 : 0);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}
