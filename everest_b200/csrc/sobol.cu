// Base samples of the MC acquisition functions on the device: scrambled Sobol points -> standard normal draws.
// Replaces (reference, CPU): botorch SobolQMCNormalSampler -> torch.quasirandom.SobolEngine(scramble=True, seed).draw(S)
// -> v = 0.5 + (1 - 1e-10)(u - 0.5) -> sqrt(2) erfinv(2v - 1), reached from every _get_acqfs (qnehvi.py:39-52,
// mobo.py:72-90, sobo.py:64-89) and from prune_inferior_points (2048 samples over N * M dimensions: the largest
// host-side cost of the acquisition set-up).  The integer pipeline is bit-identical to torch's engine:
//   scramble:  every direction number v (30 bits, MSB first) is multiplied over GF(2) by the unit lower-triangular
//              matrix of its dimension (torch._sobol_engine_scramble_);
//   draw:      x_n = shift XOR (XOR of the direction numbers selected by the Gray code of n)   (Antonov-Saleev).
// The random bits (shift, matrices) still come from torch's CPU generator so that a seed means the same as in BoTorch.
#include "common.cuh"

#define SOBOL_MAXBIT 30

// rows[d][p]: bit (29 - k) set iff L_d[p][k] = 1 (unit diagonal included)
__global__ void sobol_scramble_kernel(long long* __restrict__ ss, const long long* __restrict__ rows, int dim) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)dim * SOBOL_MAXBIT) return;
  const int d = (int)(idx / SOBOL_MAXBIT);
  const unsigned v = (unsigned)ss[idx];
  unsigned out = 0;
#pragma unroll
  for (int p = 0; p < SOBOL_MAXBIT; ++p) {
    const unsigned r = (unsigned)rows[(size_t)d * SOBOL_MAXBIT + p];
    out |= (unsigned)(__popc(r & v) & 1) << (SOBOL_MAXBIT - 1 - p);
  }
  ss[idx] = (long long)out;
}

// out[(s * n_points + i) * M + m] for Sobol dimension m * n_points + i (non-interleaved multi-output layout)
__global__ void sobol_normal_kernel(const long long* __restrict__ ss, const long long* __restrict__ shift, int n_points,
                                    int M, int S, double* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int dim = n_points * M;
  if (idx >= (long long)S * dim) return;
  const int dflat = (int)(idx % dim), s = (int)(idx / dim);
  unsigned x = (unsigned)shift[dflat];
  unsigned gray = (unsigned)s ^ ((unsigned)s >> 1);
  while (gray) {
    const int b = __ffs(gray) - 1;
    x ^= (unsigned)ss[(size_t)dflat * SOBOL_MAXBIT + b];
    gray &= gray - 1;
  }
  // torch quirk kept for bit parity: the FIRST point is formed in float32 (SobolEngine._first_point = quasi / 2**30 in the
  // default dtype), every later one in the requested dtype
  const double u = (s == 0) ? (double)(__uint2float_rn(x) * 9.313225746154785e-10f) : (double)x * 9.313225746154785e-10;  // 2^-30
  const double v = 0.5 + (1.0 - 1e-10) * (u - 0.5);
  const double z = erfinv(2.0 * v - 1.0) * 1.4142135623730951;
  const int m = dflat / n_points, i = dflat % n_points;
  out[((size_t)s * n_points + i) * M + m] = z;
}

// uniform points out[s, dim] in [0, 1): SobolEngine.draw(S, dtype=float64) (raw samples of optimize_acqf,
// [UPSTREAM] draw_sobol_samples)
__global__ void sobol_uniform_kernel(const long long* __restrict__ ss, const long long* __restrict__ shift, int dim, int S,
                                     double* __restrict__ out) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)S * dim) return;
  const int dflat = (int)(idx % dim), s = (int)(idx / dim);
  unsigned x = (unsigned)shift[dflat];
  unsigned gray = (unsigned)s ^ ((unsigned)s >> 1);
  while (gray) {
    const int b = __ffs(gray) - 1;
    x ^= (unsigned)ss[(size_t)dflat * SOBOL_MAXBIT + b];
    gray &= gray - 1;
  }
  out[idx] = (s == 0) ? (double)(__uint2float_rn(x) * 9.313225746154785e-10f) : (double)x * 9.313225746154785e-10;
}

extern "C" int bo_sobol_uniform(const int64_t* sobolstate_dev, const int64_t* shift_dev, int32_t dim, int32_t S,
                                double* out_dev, void* stream) {
  if (!sobolstate_dev || !shift_dev || !out_dev || dim < 1 || S < 1 || S > (1 << SOBOL_MAXBIT)) {
    bo_set_error("sobol_uniform: bad arguments");
    return BO_ERR_INVALID;
  }
  const long long n = (long long)S * dim;
  sobol_uniform_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(sobolstate_dev), reinterpret_cast<const long long*>(shift_dev), dim, S, out_dev);
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

extern "C" int bo_sobol_scramble(int64_t* sobolstate_dev, const int64_t* ltm_rows_dev, int32_t dim, void* stream) {
  if (!sobolstate_dev || !ltm_rows_dev || dim < 1) { bo_set_error("sobol_scramble: bad arguments"); return BO_ERR_INVALID; }
  const long long n = (long long)dim * SOBOL_MAXBIT;
  sobol_scramble_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<long long*>(sobolstate_dev), reinterpret_cast<const long long*>(ltm_rows_dev), dim);
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}

extern "C" int bo_sobol_normal(const int64_t* sobolstate_dev, const int64_t* shift_dev, int32_t n_points, int32_t M,
                               int32_t S, double* out_dev, void* stream) {
  if (!sobolstate_dev || !shift_dev || !out_dev || n_points < 1 || M < 1 || S < 1 || S > (1 << SOBOL_MAXBIT)) {
    bo_set_error("sobol_normal: bad arguments");
    return BO_ERR_INVALID;
  }
  const long long n = (long long)S * n_points * M;
  sobol_normal_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const long long*>(sobolstate_dev), reinterpret_cast<const long long*>(shift_dev), n_points, M, S, out_dev);
  CUDA_CHECK_RET(cudaGetLastError());
  return BO_OK;
}
