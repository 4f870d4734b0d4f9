// Small device helpers shared by the MC acquisition kernels (acqf.cu) and their adjoints (grad.cu).
#pragma once
#include "common.cuh"

__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    t = warp_sum(t);
  }
  return t;  // valid in warp 0
}

// [UPSTREAM] botorch.utils.safe_math (restated in oracle/bo_oracle.py: log_softplus, log_fatplus, fatmax)
__device__ __forceinline__ double softplus_d(double x) {  // torch softplus(beta=1, threshold=32)
  return (x > 32.0) ? x : log1p(exp(x));
}
__device__ __forceinline__ double log_softplus_d(double x) {
  return (x > -35.0) ? log(softplus_d(x)) : x;  // tau = 1: x / tau + log(tau)
}
__device__ __forceinline__ double logaddexp_d(double a, double b) {
  double mx = fmax(a, b), mn = fmin(a, b);
  if (isinf(mx) && mx < 0) return mx;
  return mx + log1p(exp(mn - mx));
}
// log_fatplus(x, tau) = log(tau) + logaddexp( log_softplus(z), log(0.1) - log1p(z^2) ),  z = x / tau.
// Far from the kink (|z| large -- the usual case with tau = 1e-6) the same expression needs one or two logarithms:
//   z > 32  : softplus(z) = z (torch threshold), so the value is log z + log1p(0.1 / (z (1 + z^2)))
//   z < -750: exp(z - B) underflows, the value is B = log(0.1) - log1p(z^2)
// Both are algebraic identities of the generic formula (rounding differs at 1e-16).
__device__ __forceinline__ double log_fatplus_lt(double x, double tau, double log_tau) {
  const double z = x / tau;
  if (z > 32.0) {
    const double r = 0.1 / (z * (1.0 + z * z));
    return log_tau + log(z) + ((r < 1e-9) ? r : log1p(r));
  }
  const double B = -2.302585092994046 - log1p(z * z);
  if (z < -750.0) return log_tau + B;
  return log_tau + logaddexp_d(log_softplus_d(z), B);
}
__device__ __forceinline__ double log_fatplus_d(double x, double tau) { return log_fatplus_lt(x, tau, log(tau)); }
// d log_fatplus(x, tau) / dx
__device__ __forceinline__ double log_fatplus_grad_d(double x, double tau) {
  const double z = x / tau;
  const double dB = -2.0 * z / (1.0 + z * z);
  if (z > 32.0) {
    const double r = 0.1 / (z * (1.0 + z * z));     // exp(B - A)
    return ((1.0 / z) + r * dB) / ((1.0 + r) * tau);
  }
  if (z < -750.0) return dB / tau;
  const double A = log_softplus_d(z), B = -2.302585092994046 - log1p(z * z);
  const double lae = logaddexp_d(A, B);
  double dA;
  if (z > -35.0) {
    const double sp = softplus_d(z);
    const double sg = 1.0 / (1.0 + exp(-z));
    dA = sg / sp;
  } else {
    dA = 1.0;
  }
  return (exp(A - lae) * dA + exp(B - lae) * dB) / tau;
}

// pareto(x) = 2 / (2 + 2x + x^2): the alpha = 2 kernel of fatmax / fatmin; derivative -P^2 (1 + x)
__device__ __forceinline__ double pareto2_d(double x) { return 2.0 / (2.0 + 2.0 * x + x * x); }
__device__ __forceinline__ double log1mexp_d(double x) {  // log(1 - exp(x)), x < 0
  return (x > -0.6931471805599453) ? log(-expm1(x)) : log1p(-exp(x));
}
// fatmoid: twice differentiable step with an O(1/x^2) tail (log-space feasibility weights, fat = True)
__device__ __forceinline__ double fatmoid_d(double x) {
  const double m = 0.5773502691896258;
  if (x < 0.0) { const double t = x - m; return (2.0 / 3.0) / (1.0 + t * t); }
  const double t = x + m;
  return 1.0 - (2.0 / 3.0) / (1.0 + t * t);
}
__device__ __forceinline__ double fatmoid_grad_d(double x) {
  const double m = 0.5773502691896258;
  const double t = (x < 0.0) ? x - m : x + m;
  const double den = 1.0 + t * t;
  const double g = (2.0 / 3.0) * 2.0 * t / (den * den);
  return (x < 0.0) ? -g : g;
}
// log feasibility weight sum_c log fatmoid(-c(y) / eta) and (optionally) its gradient w.r.t. the outputs (added into dy)
__device__ __forceinline__ double log_feas_fat(const ObjD& od, const double* y, double scale, double* dy) {
  double lf = 0.0;
  for (int c = 0; c < od.n_cons; ++c) {
    const double x = -od.con[c].sign * (y[od.con[c].out_idx] - od.con[c].tp) / od.con[c].eta;
    const double f = fatmoid_d(x);
    lf += log(f);
    if (dy) dy[od.con[c].out_idx] += scale * (fatmoid_grad_d(x) / f) * (-od.con[c].sign / od.con[c].eta);
  }
  return lf;
}
// feasibility weight prod_c sigmoid(-c(y) / eta) (fat = False) and (optionally) scale * d weight / d y added into dy
__device__ __forceinline__ double feas_sigmoid(const ObjD& od, const double* y, double scale, double* dy) {
  double sg[BO_MAX_CONSTRAINTS];
  double w = 1.0;
  for (int c = 0; c < od.n_cons; ++c) {
    const double cv = od.con[c].sign * (y[od.con[c].out_idx] - od.con[c].tp);
    sg[c] = 1.0 / (1.0 + exp(cv / od.con[c].eta));
    w *= sg[c];
  }
  if (dy)
    for (int c = 0; c < od.n_cons; ++c) {
      double rest = 1.0;
      for (int c2 = 0; c2 < od.n_cons; ++c2)
        if (c2 != c) rest *= sg[c2];
      dy[od.con[c].out_idx] += scale * rest * sg[c] * (1.0 - sg[c]) * (-od.con[c].sign / od.con[c].eta);
    }
  return w;
}

// d objective / d y[op.out_idx]  (utils/torch_tools.py:384-450; forward: objective_apply in common.cuh)
__device__ __forceinline__ double objective_grad(const bo_objective_op& op, const double* y) {
  const double v = y[op.out_idx];
  switch (op.kind) {
    case BO_OBJ_MAX:
      return 1.0 / (op.p1 - op.p0);
    case BO_OBJ_MIN:
      return -1.0 / (op.p1 - op.p0);
    case BO_OBJ_CLOSE_TO_TARGET: {
      const double t = v - op.p0;
      if (t == 0.0) return 0.0;
      return -op.p1 * pow(fabs(t), op.p1 - 1.0) * (t > 0.0 ? 1.0 : -1.0);
    }
    case BO_OBJ_MIN_SIGMOID: {
      const double sg = 1.0 / (1.0 + exp(-1.0 * op.p0 * (v - op.p1)));
      return -op.p0 * sg * (1.0 - sg);
    }
    case BO_OBJ_MAX_SIGMOID: {
      const double sg = 1.0 / (1.0 + exp(-1.0 * op.p0 * (v - op.p1)));
      return op.p0 * sg * (1.0 - sg);
    }
    case BO_OBJ_TARGET: {
      const double s1 = 1.0 / (1.0 + exp(-1.0 * op.p2 * (v - (op.p0 - op.p1))));
      const double s2 = 1.0 / (1.0 + exp(-1.0 * op.p2 * (v - (op.p0 + op.p1))));
      return op.p2 * s1 * (1.0 - s1) * (1.0 - s2) - s1 * op.p2 * s2 * (1.0 - s2);
    }
  }
  return 0.0;
}

// Scalarised objective of qLogEI (single / additive / multiplicative, utils/torch_tools.py:662-697) and, when
// `dy` is given, its gradient with respect to the M model outputs.
__device__ __forceinline__ double scalar_objective_apply(const ObjD& od, const double* y, int M, double* dy) {
  if (dy) for (int m = 0; m < M; ++m) dy[m] = 0.0;
  if (od.combine == BO_COMBINE_SINGLE) {
    if (dy) dy[od.op[0].out_idx] = objective_grad(od.op[0], y);
    return objective_apply(od.op[0], y);
  }
  if (od.combine == BO_COMBINE_ADDITIVE) {
    double o = 0.0;
    for (int k = 0; k < od.n_obj; ++k) {
      o = o + objective_apply(od.op[k], y) * od.op[k].w;
      if (dy) dy[od.op[k].out_idx] += objective_grad(od.op[k], y) * od.op[k].w;
    }
    return o;
  }
  double o = 1.0;
  for (int k = 0; k < od.n_obj; ++k) o = o * pow(objective_apply(od.op[k], y), od.op[k].w);
  if (dy) {
    for (int k = 0; k < od.n_obj; ++k) {
      double rest = 1.0;
      for (int k2 = 0; k2 < od.n_obj; ++k2)
        if (k2 != k) rest *= pow(objective_apply(od.op[k2], y), od.op[k2].w);
      const double f = objective_apply(od.op[k], y);
      dy[od.op[k].out_idx] += rest * od.op[k].w * pow(f, od.op[k].w - 1.0) * objective_grad(od.op[k], y);
    }
  }
  return o;
}
