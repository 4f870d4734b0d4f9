// K3 (factorisation part): blocked FP64 Cholesky and triangular inverse on the device.
//
// Replaces the CPU LAPACK calls behind gpytorch's prediction-strategy caches and
// linear_operator's psd_safe_cholesky (reached from model.posterior, reference
// strategies/predictives/botorch.py:180,223): training Gram K + s2 I = L L^T and
// the inverse root L^-1 (gpytorch caches L^-T), plus the joint posterior covariance factor used by
// baseline pruning and the cached baseline root.
//
// Right-looking blocked algorithm, block 64: the diagonal block is factored in shared memory by one
// CTA (which also inverts it), the panel solve and the trailing update are DMMA "NT" GEMMs
// (gemm.cu).  info follows LAPACK: 0 or (1-based) index of the first non-positive pivot.
#include "common.cuh"

#define CB 64
#define CB_LD 65

__global__ void __launch_bounds__(256)
potrf_diag_kernel(double* __restrict__ A, int ld, int kb, int k0, double* __restrict__ dinv, int* __restrict__ info) {
  extern __shared__ double potrf_sm[];
  double* a = potrf_sm;
  double* x = potrf_sm + CB * CB_LD;
  __shared__ int bad;
  const int tid = threadIdx.x;
  if (tid == 0) bad = (*info != 0) ? -1 : 0;  // an earlier block already failed: leave everything untouched
  for (int idx = tid; idx < CB * CB; idx += 256) {
    int r = idx / CB, c = idx % CB;
    a[r * CB_LD + c] = (r < kb && c <= r) ? A[(size_t)r * ld + c] : 0.0;
    x[r * CB_LD + c] = 0.0;
  }
  __syncthreads();
  if (bad != 0) return;
  for (int j = 0; j < kb; ++j) {
    if (tid == 0) {
      double d = a[j * CB_LD + j];
      if (!(d > 0.0)) bad = k0 + j + 1;
      else a[j * CB_LD + j] = sqrt(d);
    }
    __syncthreads();
    if (bad != 0) break;
    const double dj = a[j * CB_LD + j];
    for (int i = j + 1 + tid; i < kb; i += 256) a[i * CB_LD + j] /= dj;
    __syncthreads();
    const int rem = kb - j - 1;
    for (int idx = tid; idx < rem * rem; idx += 256) {
      int i = j + 1 + idx / rem, c = j + 1 + idx % rem;
      if (c <= i) a[i * CB_LD + c] -= a[i * CB_LD + j] * a[c * CB_LD + j];
    }
    __syncthreads();
  }
  if (bad != 0) {
    if (tid == 0) *info = bad;
    return;
  }
  // inverse of the lower-triangular diagonal block, one column per thread
  if (tid < kb) {
    const int c = tid;
    for (int i = c; i < kb; ++i) {
      double s = (i == c) ? 1.0 : 0.0;
      for (int l = c; l < i; ++l) s -= a[i * CB_LD + l] * x[l * CB_LD + c];
      x[i * CB_LD + c] = s / a[i * CB_LD + i];
    }
  }
  __syncthreads();
  for (int idx = tid; idx < CB * CB; idx += 256) {
    int r = idx / CB, c = idx % CB;
    dinv[idx] = x[r * CB_LD + c];
    if (r < kb && c < kb) A[(size_t)r * ld + c] = (c <= r) ? a[r * CB_LD + c] : 0.0;
  }
}

// zero the strict upper triangle (outside the diagonal blocks, which potrf_diag already cleaned)
__global__ void zero_upper_kernel(double* __restrict__ A, int ld, int n) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int r = blockIdx.y;
  if (r < n && c < n && c > r) A[(size_t)r * ld + c] = 0.0;
}

int chol_blocked(double* A, int ld, int n, double* work_dinv, int* info_dev, cudaStream_t s, LaunchCounter* lc) {
  // info_dev must be zeroed by the caller; work_dinv holds ceil(n/64) blocks of 64*64 doubles.
  const size_t potrf_smem = (size_t)2 * CB * CB_LD * sizeof(double);
  static PerDeviceOnce attr_once; bool& attr_set = *attr_once.slot();
  if (!attr_set) {
    CUDA_CHECK_RET(cudaFuncSetAttribute(potrf_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)potrf_smem));
    attr_set = true;
  }
  for (int k0 = 0; k0 < n; k0 += CB) {
    int kb = (n - k0 < CB) ? (n - k0) : CB;
    double* dinv = work_dinv + (size_t)(k0 / CB) * CB * CB;
    potrf_diag_kernel<<<1, 256, potrf_smem, s>>>(A + (size_t)k0 * ld + k0, ld, kb, k0, dinv, info_dev);
    if (lc) lc->n++;
    int n2 = n - k0 - kb;
    if (n2 > 0) {
      double* A21 = A + (size_t)(k0 + kb) * ld + k0;
      // panel: L21 = A21 * L11^-T  (in place: each CTA owns its 64 rows and has finished all
      // global reads of them before its epilogue writes)
      int rc = launch_gemm_nt(n2, CB, CB, 1.0, A21, ld, dinv, CB, 0.0, A21, ld, false, s, lc);
      if (rc) return rc;
      // trailing update, lower triangle only: A22 -= L21 L21^T
      double* A22 = A + (size_t)(k0 + kb) * ld + (k0 + kb);
      rc = launch_gemm_nt(n2, n2, CB, -1.0, A21, ld, A21, ld, 1.0, A22, ld, true, s, lc);
      if (rc) return rc;
    }
  }
  dim3 grid((n + 255) / 256, n);
  zero_upper_kernel<<<grid, 256, 0, s>>>(A, ld, n);
  if (lc) lc->n++;
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

// X[i0:i0+ib, 0:i0] = -Dinv * P ;  X[i0:i0+ib, i0:i0+ib] = Dinv ; mirrored into XT.
__global__ void __launch_bounds__(256)
triinv_finish_kernel(const double* __restrict__ dinv, const double* __restrict__ P, int ldp, int i0, int ib,
                     double* __restrict__ X, double* __restrict__ XT, int ldx) {
  __shared__ double dsm[CB * CB_LD];
  for (int idx = threadIdx.x; idx < CB * CB; idx += 256) dsm[(idx / CB) * CB_LD + (idx % CB)] = dinv[idx];
  __syncthreads();
  const int c = blockIdx.x * 64 + (threadIdx.x & 63);
  const int rg = threadIdx.x >> 6;  // 4 row groups of 16
  if (c >= i0 + ib) return;
  if (c < i0) {
    double acc[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc[r] = 0.0;
    for (int l = 0; l < ib; ++l) {
      double p = P[(size_t)l * ldp + c];
#pragma unroll
      for (int r = 0; r < 16; ++r) acc[r] = fma(dsm[(rg * 16 + r) * CB_LD + l], p, acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      int rr = rg * 16 + r;
      if (rr < ib) {
        X[(size_t)(i0 + rr) * ldx + c] = -acc[r];
        XT[(size_t)c * ldx + i0 + rr] = -acc[r];
      }
    }
  } else {
    int cc = c - i0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
      int rr = rg * 16 + r;
      if (rr < ib) {
        double v = dsm[rr * CB_LD + cc];
        X[(size_t)(i0 + rr) * ldx + c] = v;
        XT[(size_t)c * ldx + i0 + rr] = v;
      }
    }
  }
}

int tri_inverse_blocked(const double* L, int ld, int n, const double* dinv, double* X, double* XT, int ldx, double* tmp,
                        cudaStream_t s, LaunchCounter* lc) {
  // X, XT: [rows >= n, ldx] zero-initialised by the caller. tmp: [64, ldx].
  for (int i0 = 0; i0 < n; i0 += CB) {
    int ib = (n - i0 < CB) ? (n - i0) : CB;
    if (i0 > 0) {
      int rc = launch_gemm_nt(ib, i0, i0, 1.0, L + (size_t)i0 * ld, ld, XT, ldx, 0.0, tmp, ldx, false, s, lc);
      if (rc) return rc;
    }
    int cols = i0 + ib;
    triinv_finish_kernel<<<(cols + 63) / 64, 256, 0, s>>>(dinv + (size_t)(i0 / CB) * CB * CB, tmp, ldx, i0, ib, X, XT,
                                                          ldx);
    if (lc) lc->n++;
  }
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}
