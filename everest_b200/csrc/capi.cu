// C-ABI entry points (include/everest_b200.h) and host-side orchestration of the kernels.
#include <stdarg.h>
#include <string.h>

#include <math.h>

#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <atomic>
#include <memory>
#include <functional>
#include <vector>

#include "acqf.cuh"
#include "common.cuh"
#include "host_pack.h"
#include "lbfgs.cuh"

static thread_local char g_err[1024] = "";

void bo_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static inline int round_up(int v, int m) { return ((v + m - 1) / m) * m; }

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  int ensure(size_t need, bool zero = false) {
    if (need == 0) need = 16;
    if (bytes < need) {
      if (p) cudaFree(p);
      p = nullptr;
      bytes = 0;
      cudaError_t e = cudaMalloc(&p, need);
      if (e != cudaSuccess) {
        bo_set_error("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e));
        return BO_ERR_CUDA;
      }
      bytes = need;
      zero = true;
    }
    if (zero) {
      // set-up paths only (the forward pass never zeroes): order the clear against every stream
      cudaError_t e = cudaDeviceSynchronize();
      if (e == cudaSuccess) e = cudaMemset(p, 0, bytes);
      if (e == cudaSuccess) e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { bo_set_error("cudaMemset failed: %s", cudaGetErrorString(e)); return BO_ERR_CUDA; }
    }
    return BO_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; bytes = 0; }
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

#define RC(expr) do { int _rc = (expr); if (_rc != BO_OK) return _rc; } while (0)

// EVEREST_SETUP_TRACE=1: wall-clock checkpoints (stream-synchronised) of the set-up entry points on stderr
struct SetupTrace {
  const char* name; cudaStream_t s; bool on;
  std::chrono::steady_clock::time_point t0;
  SetupTrace(const char* n, cudaStream_t st) : name(n), s(st) {
    static const bool enabled = getenv("EVEREST_SETUP_TRACE") != nullptr;
    on = enabled;
    if (on) { cudaStreamSynchronize(s); t0 = std::chrono::steady_clock::now(); }
  }
  void mark(const char* what) {
    if (!on) return;
    cudaStreamSynchronize(s);
    fprintf(stderr, "[setup] %s: %s at %.3f ms\n", name, what,
            std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
  }
};

struct PrepBuf {
  DevBuf Xs[BO_MAX_LEAVES], n2[BO_MAX_LEAVES], codes[BO_MAX_LEAVES], bits[BO_MAX_LEAVES], pc[BO_MAX_LEAVES];
  int ensure(const ModelD& md, int n, PrepD* out) {
    memset(out, 0, sizeof(PrepD));
    out->n = n;
    size_t nn = (size_t)std::max(n, 1);
    for (int l = 0; l < md.n_leaves; ++l) {
      const LeafD& L = md.leaf[l];
      if (L.kind <= BO_LEAF_MATERN52) {
        RC(Xs[l].ensure(nn * L.dpad * sizeof(double)));
        RC(n2[l].ensure(nn * sizeof(double)));
        out->Xs[l] = Xs[l].as<double>();
        out->n2[l] = n2[l].as<double>();
      } else if (L.kind == BO_LEAF_HAMMING) {
        RC(codes[l].ensure(nn * L.nd * sizeof(int)));
        out->codes[l] = codes[l].as<int>();
      } else {
        // packed words, followed by the same fingerprints as one 0 / 1 byte per column (rows padded to 256 bytes): the
        // operands of the u8 tensor-core inner products in crosscov_kernel2
        RC(bits[l].ensure(nn * L.dpad * sizeof(u64) + nn * (size_t)tanimoto_row_bytes(L.dpad)));
        RC(pc[l].ensure(nn * sizeof(int)));
        out->bits[l] = bits[l].as<u64>();
        out->pc[l] = pc[l].as<int>();
      }
    }
    return BO_OK;
  }
  void release() {
    for (int l = 0; l < BO_MAX_LEAVES; ++l) { Xs[l].release(); n2[l].release(); codes[l].release(); bits[l].release(); pc[l].release(); }
  }
};

struct OutputH {
  ModelD md;
  std::vector<void*> owned;  // small parameter arrays
  PrepBuf train_prep;
  PrepD train_prepd;
  DevBuf L, Linv, LinvT, LinvExt, alpha_row, dinv, resid, tvec;
  DevBuf Kinv;  // (K + s2 I)^-1, built lazily by the first backward pass
  bool kinv_ready = false;
  DevBuf ozB, ozScaleB;  // INT8 digit planes of LinvExt + per-row scales (ozaki.cu), rebuilt when LinvExt changes
  DevBuf ozSb2;          // [max_{n<N} scaleB[n], scaleB[N]]: what the per-row guard of the INT8 path needs
  bool oz_ready = false;
  PrepBuf g_prep;        // prepared points of the q-batches the guard sends back through the FP64 kernel
  PrepD g_prepd;
  int Rpad = 0;  // rows of LinvExt
  double jitter = 0.0;
  // acquisition state
  PrepBuf base_prep;
  PrepD base_prepd;
  DevBuf Lb, Sbb, LbInv, LbInvT;
  // per-forward query prep
  PrepBuf q_prep;
  PrepD q_prepd;
  // host copies of every leaf's columns (Hamming: start column + cardinality of each group): wire-format layout
  std::vector<int> leaf_cols[BO_MAX_LEAVES], leaf_card[BO_MAX_LEAVES];
  std::vector<double> y_host;   // raw targets: the residual is rebuilt when the constant mean changes (bo_state_set_hyperparameters)
};

struct TimingRec { std::string name; cudaEvent_t a, b; };

struct bo_state {
  int N = 0, d = 0, M = 0, ldk = 0, Nr = 0;
  bool factorized = false;
  DevBuf X_train;
  std::vector<OutputH> out;
  LaunchCounter lc{0};
  // workspaces
  DevBuf wsKx, wsV, wsGqq, wsW, wsMuRaw, wsRoot, wsMu, wsZqT, wsTmp, wsInfo, wsCov, wsMean, wsF, wsZM, wsObj, wsFeas,
      wsFront, wsCounts, wsJit, wsPart;
  // acquisition
  int acqf_kind = 0;  // 0 none, 1 nehvi, 2 ehvi, 3 logei
  int nb = 0, S = 0, ldlb = 0, cap = 0;
  ObjD od;
  double best_f = 0.0;
  int scalar_variant = 0;      // bo_scalar_acqf
  double vparam = 0.0;         // qUCB beta / qPI tau
  bool noisy_scalar = false;   // qNEI / qLogNEI: per-sample incumbent in best_f_s
  bool baseline_f_valid = false;   // wsF / mean_b still hold the baseline samples of the prepared noisy scalar acqf
  DevBuf best_f_s;
  int log_hvi = 0;
  int ozaki = 0;               // posterior GEMM of large batches on the INT8 tensor cores (ozaki.cu)
  int ozaki_tile = 0;          // kernel variant of the INT8 GEMM (0 = default)
  int oz_calib = 0;            // automatic mode: 0 = no large call yet for this prepared state, 1 = INT8 path in use,
                               // -1 = the guard flagged most rows of a call: FP64 kernel from then on
  long long oz_flagged_last = 0, oz_batches_last = 0;   // q-batches the guard redid in FP64 / scored in the last forward
  long long oz_flagged_total = 0, oz_batches_total = 0; // ... since the last prepare
  double oz_kappa = 8.0, oz_tol = 1e-10;
  double part_alpha = 0.0;     // approximate box decomposition threshold of the next prepare (0 = exact)
  DevBuf wsOzFlags, wsOzXg, wsOzKxG, wsOzOutG;
  int* pin_count = nullptr;    // pinned: the guard's counter of flagged q-batches
  cudaEvent_t oz_event = nullptr;
  DevBuf wsLbfgs, wsLbBounds, wsLbGrad;   // on-device multi-start refinement (lbfgs.cu)
  DevBuf Xb_raw, wsXfull, wsMeanJ;        // joint re-sampling fallback: baseline points, (baseline, q-batch), joint mean
  int force_fallback = 0;                 // test switch: treat every q-batch as flagged
  int auto_fallback = 1;                  // forward: re-score flagged q-batches from the joint posterior (BoTorch's fallback)
  int fb_resampled_last = 0;              // q-batches the last forward call re-scored
  DevBuf wsInfoOut, wsFbCount;
  int* pin_fb = nullptr;
  cudaEvent_t fb_event = nullptr;
  DevBuf wsJointRoot, wsJointCov, wsJointDinv;   // joint posterior root / covariance of the pruning passes
  PrepBuf jointPrep;
  int* pin_lb = nullptr;                  // pinned: progress counters of bo_acqf_optimize
  std::vector<cudaEvent_t> lb_events;
  cudaStream_t side2 = nullptr;           // odd outputs of a refinement step (acqf_run)
  cudaEvent_t side2_fork = nullptr, side2_join = nullptr;
  DevBuf wsDXm;
  cudaStream_t side_stream = nullptr;     // forked work of the refinement steps (acqf_run)
  cudaEvent_t side_fork = nullptr, side_join = nullptr;
  int lb_launches_per_step = 0;
  bool lb_graph_used = false;
  double tau_relu = 1e-6, tau_max = 1e-2;
  DevBuf wsOzA;  // INT8 digit planes of K(X*,X) (all outputs)
  DevBuf wsOzScratch;  // integer slab of the two-pass INT8 GEMM (ozaki.cu), one per handle
  DevBuf wsDF, wsDRoot, wsDMu, wsEG, wsEW, wsEmu, wsU;  // adjoint workspaces (grad.cu)
  DevBuf wsGramPart, wsObjW, zbT, zbM, cell_lo, cell_up, ncells, front_idx, ref_dev, mean_b, obj_b, samples_b, wsBL, wsFp, wsPartial;
  int max_cells = 0;
  int cells_shared = 0;
  // host staging for the HOST-buffer entry point
  void* pin_in = nullptr; size_t pin_in_bytes = 0;
  // packed wire format of the host entry points: fingerprint columns as bits
  bool pack_ready = false, pack_src_ready = false;
  std::vector<int> pack_bit_cols, pack_dense_cols;
  DevBuf stage_pk, pack_src;
  void* pin_out = nullptr; size_t pin_out_bytes = 0;
  DevBuf stage_in, stage_out;
  cudaStream_t copy_stream = nullptr;
  std::vector<cudaEvent_t> copy_events;
  // host entry point, piece-wise input (forward_host_impl -> acqf_run): X of the next forward call arrives in row pieces;
  // piece_wait(i) blocks until piece i's copy has been enqueued and makes the call's stream wait for it.  Only the point
  // preparation and K(X*,X) of output 0 are launched piece by piece; everything after them sees the whole batch.
  std::vector<int> piece_row_end;                       // cumulative candidate rows per piece (empty: X is complete)
  std::function<int(int)> piece_wait;
  // timing
  bool timing = false;
  std::vector<TimingRec> recs;
};

static int upload(void** dst, const void* src, size_t bytes, std::vector<void*>& owned) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, std::max<size_t>(bytes, 16));
  if (e != cudaSuccess) { bo_set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return BO_ERR_CUDA; }
  if (bytes) {
    e = cudaMemcpy(p, src, bytes, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { bo_set_error("cudaMemcpy failed: %s", cudaGetErrorString(e)); return BO_ERR_CUDA; }
  }
  owned.push_back(p);
  *dst = p;
  return BO_OK;
}

extern "C" int bo_version(void) { return 100; }
extern "C" const char* bo_last_error(void) { return g_err; }

extern "C" void bo_state_destroy(bo_state* st) {
  if (!st) return;
  for (auto& o : st->out) {
    for (void* p : o.owned) cudaFree(p);
    o.train_prep.release(); o.base_prep.release(); o.q_prep.release(); o.g_prep.release();
    DevBuf* bs[] = {&o.ozB, &o.ozScaleB, &o.ozSb2, &o.Kinv, &o.L, &o.Linv, &o.LinvT, &o.LinvExt, &o.alpha_row, &o.dinv, &o.resid, &o.tvec, &o.Lb, &o.Sbb, &o.LbInv, &o.LbInvT};
    for (DevBuf* b : bs) b->release();
  }
  DevBuf* bs[] = {&st->X_train, &st->wsKx, &st->wsV, &st->wsGqq, &st->wsW, &st->wsMuRaw, &st->wsRoot, &st->wsMu,
                  &st->wsZqT, &st->wsTmp, &st->wsInfo, &st->wsCov, &st->wsMean, &st->wsF, &st->wsZM, &st->wsObj,
                  &st->wsFeas, &st->wsFront, &st->wsCounts, &st->wsJit, &st->wsPart, &st->zbT, &st->cell_lo,
                  &st->cell_up, &st->ncells, &st->front_idx, &st->wsGramPart, &st->wsObjW, &st->zbM, &st->wsBL, &st->wsFp, &st->wsPartial, &st->ref_dev, &st->mean_b, &st->obj_b, &st->samples_b,
                  &st->stage_in, &st->stage_out, &st->stage_pk, &st->pack_src, &st->wsDF, &st->wsDRoot, &st->wsDMu, &st->wsEG, &st->wsEW, &st->wsEmu, &st->wsU, &st->best_f_s, &st->wsOzA, &st->wsOzFlags, &st->wsOzXg, &st->wsOzKxG, &st->wsOzOutG, &st->wsOzScratch};
  for (DevBuf* b : bs) b->release();
  if (st->pin_count) cudaFreeHost(st->pin_count);
  if (st->oz_event) cudaEventDestroy(st->oz_event);
  st->wsLbfgs.release(); st->wsLbBounds.release(); st->wsLbGrad.release();
  st->Xb_raw.release(); st->wsXfull.release(); st->wsMeanJ.release(); st->wsInfoOut.release(); st->wsFbCount.release();
  if (st->side2) { cudaStreamDestroy(st->side2); cudaEventDestroy(st->side2_fork); cudaEventDestroy(st->side2_join); }
  st->wsDXm.release();
  if (st->side_stream) { cudaStreamDestroy(st->side_stream); cudaEventDestroy(st->side_fork); cudaEventDestroy(st->side_join); }
  if (st->pin_fb) cudaFreeHost(st->pin_fb);
  if (st->fb_event) cudaEventDestroy(st->fb_event);
  st->wsJointRoot.release(); st->wsJointCov.release(); st->wsJointDinv.release(); st->jointPrep.release();
  if (st->pin_lb) cudaFreeHost(st->pin_lb);
  for (auto& e : st->lb_events) cudaEventDestroy(e);
  if (st->pin_in) cudaFreeHost(st->pin_in);
  if (st->pin_out) cudaFreeHost(st->pin_out);
  for (auto& e : st->copy_events) cudaEventDestroy(e);
  if (st->copy_stream) cudaStreamDestroy(st->copy_stream);
  for (auto& r : st->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  delete st;
}

extern "C" int bo_state_create(const bo_state_config* cfg, bo_state** out_state) {
  if (!cfg || !out_state) { bo_set_error("null argument"); return BO_ERR_INVALID; }
  if (cfg->N < 1 || cfg->d < 1 || cfg->M < 1 || cfg->M > 2 * BO_MAX_OBJECTIVES) {
    bo_set_error("bad sizes N=%d d=%d M=%d (M <= %d)", cfg->N, cfg->d, cfg->M, 2 * BO_MAX_OBJECTIVES);
    return BO_ERR_INVALID;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    bo_set_error("no CUDA device: everest_b200 has no CPU fallback");
    return BO_ERR_CUDA;
  }
  bo_state* st = new bo_state();
  st->N = cfg->N; st->d = cfg->d; st->M = cfg->M;
  { const char* e = getenv("EVEREST_OZAKI"); st->ozaki = e ? std::max(0, std::min(2, atoi(e))) : 1; }
  st->ldk = round_up(cfg->N, 16);
  st->Nr = round_up(cfg->N, 128);
  const int N = cfg->N, d = cfg->d;
  int rc = st->X_train.ensure((size_t)N * d * sizeof(double));
  if (rc) { bo_state_destroy(st); return rc; }
  if (!cfg->X_train || !cfg->outputs) { bo_set_error("X_train / outputs is NULL"); bo_state_destroy(st); return BO_ERR_INVALID; }
  {
    cudaError_t ce = cudaMemcpy(st->X_train.p, cfg->X_train, (size_t)N * d * sizeof(double), cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { bo_set_error("create: copy of X_train failed: %s", cudaGetErrorString(ce)); bo_state_destroy(st); return BO_ERR_CUDA; }
  }
  st->out.resize(cfg->M);
  for (int m = 0; m < cfg->M; ++m) {
    const bo_output_model& om = cfg->outputs[m];
    OutputH& o = st->out[m];
    memset(&o.md, 0, sizeof(ModelD));
    if (om.n_leaves < 1 || om.n_leaves > BO_MAX_LEAVES || om.n_terms < 1 || om.n_terms > BO_MAX_TERMS || !om.leaves || !om.terms) {
      bo_set_error("output %d: n_leaves=%d n_terms=%d out of range", m, om.n_leaves, om.n_terms);
      bo_state_destroy(st); return BO_ERR_INVALID;
    }
    if (!(om.y_std > 0.0) || !(om.noise >= 0.0)) { bo_set_error("output %d: y_std must be > 0 and noise >= 0", m); bo_state_destroy(st); return BO_ERR_INVALID; }
    o.md.n_leaves = om.n_leaves; o.md.n_terms = om.n_terms;
    o.md.mean_const = om.mean_const; o.md.noise = om.noise; o.md.y_mean = om.y_mean; o.md.y_std = om.y_std;
    for (int t = 0; t < om.n_terms; ++t) {
      const bo_kernel_term& kt = om.terms[t];
      if (kt.n_factors < 1 || kt.n_factors > BO_MAX_FACTORS) { bo_set_error("term %d: bad n_factors", t); bo_state_destroy(st); return BO_ERR_INVALID; }
      o.md.coef[t] = kt.coef; o.md.nfac[t] = kt.n_factors;
      for (int f = 0; f < kt.n_factors; ++f) {
        if (kt.factors[f] < 0 || kt.factors[f] >= om.n_leaves) { bo_set_error("term %d: bad leaf index", t); bo_state_destroy(st); return BO_ERR_INVALID; }
        o.md.fac[t][f] = kt.factors[f];
      }
    }
    for (int l = 0; l < om.n_leaves; ++l) {
      const bo_kernel_leaf& kl = om.leaves[l];
      LeafD& L = o.md.leaf[l];
      L.kind = kl.kind; L.nd = kl.n_dims;
      if (kl.n_dims < 1 || !kl.dims) { bo_set_error("leaf %d: n_dims < 1 or dims is NULL", l); bo_state_destroy(st); return BO_ERR_INVALID; }
      if (kl.kind == BO_LEAF_HAMMING && (!kl.cardinality || !kl.lengthscale)) {
        bo_set_error("leaf %d: a Hamming leaf needs cardinality and lengthscale", l); bo_state_destroy(st); return BO_ERR_INVALID;
      }
      for (int k = 0; k < kl.n_dims; ++k) {
        int width = (kl.kind == BO_LEAF_HAMMING) ? kl.cardinality[k] : 1;
        if (kl.dims[k] < 0 || kl.dims[k] + width > d) { bo_set_error("leaf %d: column out of range", l); bo_state_destroy(st); return BO_ERR_INVALID; }
      }
      std::vector<int> col(kl.dims, kl.dims + kl.n_dims);
      o.leaf_cols[l] = col;
      if (kl.kind == BO_LEAF_HAMMING) o.leaf_card[l].assign(kl.cardinality, kl.cardinality + kl.n_dims);
      void* p;
      rc = upload(&p, col.data(), col.size() * sizeof(int), o.owned); if (rc) { bo_state_destroy(st); return rc; }
      L.col = (const int*)p;
      if (kl.kind <= BO_LEAF_MATERN52) {
        if (kl.kind < 0 || (kl.n_ls != 1 && kl.n_ls != kl.n_dims) || !kl.lengthscale) { bo_set_error("leaf %d: bad lengthscale spec", l); bo_state_destroy(st); return BO_ERR_INVALID; }
        L.dpad = round_up(kl.n_dims, 4);
        std::vector<double> off(kl.n_dims), scl(kl.n_dims), ls(kl.n_dims), cen(kl.n_dims);
        for (int k = 0; k < kl.n_dims; ++k) {
          off[k] = om.in_offset ? om.in_offset[kl.dims[k]] : 0.0;
          scl[k] = om.in_scale ? om.in_scale[kl.dims[k]] : 1.0;
          ls[k] = kl.lengthscale[kl.n_ls == 1 ? 0 : k];
          if (!(ls[k] > 0.0) || scl[k] == 0.0) { bo_set_error("leaf %d: non-positive lengthscale / zero scale", l); bo_state_destroy(st); return BO_ERR_INVALID; }
          double acc = 0.0;
          for (int i = 0; i < N; ++i) acc += (cfg->X_train[(size_t)i * d + kl.dims[k]] - off[k]) / scl[k];
          cen[k] = acc / (double)N;
        }
        rc = upload(&p, off.data(), off.size() * 8, o.owned); if (rc) { bo_state_destroy(st); return rc; } L.in_off = (const double*)p;
        rc = upload(&p, scl.data(), scl.size() * 8, o.owned); if (rc) { bo_state_destroy(st); return rc; } L.in_scl = (const double*)p;
        rc = upload(&p, cen.data(), cen.size() * 8, o.owned); if (rc) { bo_state_destroy(st); return rc; } L.center = (const double*)p;
        rc = upload(&p, ls.data(), ls.size() * 8, o.owned); if (rc) { bo_state_destroy(st); return rc; } L.ls = (const double*)p;
      } else if (kl.kind == BO_LEAF_HAMMING) {
        if (kl.n_dims > BO_MAX_GROUPS || !kl.cardinality || !kl.lengthscale || (kl.n_ls != 1 && kl.n_ls < kl.n_dims)) {
          bo_set_error("leaf %d: bad Hamming spec (<= %d groups)", l, BO_MAX_GROUPS); bo_state_destroy(st); return BO_ERR_INVALID;
        }
        L.dpad = kl.n_dims;
        std::vector<int> card(kl.cardinality, kl.cardinality + kl.n_dims);
        std::vector<double> wls(kl.n_dims);
        for (int k = 0; k < kl.n_dims; ++k) wls[k] = 1.0 / kl.lengthscale[kl.n_ls == 1 ? 0 : k];
        rc = upload(&p, card.data(), card.size() * sizeof(int), o.owned); if (rc) { bo_state_destroy(st); return rc; } L.card = (const int*)p;
        rc = upload(&p, wls.data(), wls.size() * 8, o.owned); if (rc) { bo_state_destroy(st); return rc; } L.wls = (const double*)p;
      } else if (kl.kind == BO_LEAF_TANIMOTO) {
        L.dpad = (kl.n_dims + 63) / 64;
        for (int i = 0; i < N; ++i)
          for (int k = 0; k < kl.n_dims; ++k) {
            double v = cfg->X_train[(size_t)i * d + kl.dims[k]];
            if (v != 0.0 && v != 1.0) {
              bo_set_error("leaf %d: Tanimoto columns must hold 0/1 fingerprints (count features are not supported)", l);
              bo_state_destroy(st); return BO_ERR_INVALID;
            }
          }
      } else { bo_set_error("leaf %d: unknown kind %d", l, kl.kind); bo_state_destroy(st); return BO_ERR_INVALID; }
    }
    // prepared training side
    rc = o.train_prep.ensure(o.md, N, &o.train_prepd); if (rc) { bo_state_destroy(st); return rc; }
    rc = launch_prep_points(o.md, st->X_train.as<double>(), N, d, o.train_prepd, 0, &st->lc); if (rc) { bo_state_destroy(st); return rc; }
    for (int l = 0; l < om.n_leaves; ++l) {
      LeafD& L = o.md.leaf[l];
      L.Xs = o.train_prepd.Xs[l]; L.n2 = o.train_prepd.n2[l]; L.codes = o.train_prepd.codes[l];
      L.bits = o.train_prepd.bits[l]; L.pc = o.train_prepd.pc[l];
    }
    // residual r = (y - y_mean) / y_std - mean_const, zero padded row of length ldk
    if (!om.y) { bo_set_error("output %d: y is NULL", m); bo_state_destroy(st); return BO_ERR_INVALID; }
    o.y_host.assign(om.y, om.y + N);
    std::vector<double> r(st->ldk, 0.0);
    for (int i = 0; i < N; ++i) r[i] = (om.y[i] - om.y_mean) / om.y_std - om.mean_const;
    rc = o.resid.ensure((size_t)st->ldk * 8); if (rc) { bo_state_destroy(st); return rc; }
    {
      cudaError_t ce = cudaMemcpy(o.resid.p, r.data(), (size_t)st->ldk * 8, cudaMemcpyHostToDevice);
      if (ce != cudaSuccess) { bo_set_error("create: copy of the residual failed: %s", cudaGetErrorString(ce)); bo_state_destroy(st); return BO_ERR_CUDA; }
    }
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { bo_set_error("create: %s", cudaGetErrorString(e)); bo_state_destroy(st); return BO_ERR_CUDA; }
  *out_state = st;
  return BO_OK;
}

// New hyper-parameter VALUES for output m of an existing state (same kernel tree, same columns, same transforms and
// targets): what a step of the marginal-likelihood optimiser changes.  Only the small parameter arrays are rewritten and the
// training side is prepared again; every large buffer of the handle is reused by the following bo_state_factorize -- a fit of
// N = 2000 points spent most of its time creating and destroying states (~50 ms per evaluation for a 15 ms factorisation).
extern "C" int bo_state_set_hyperparameters(bo_state* st, int32_t m, const bo_output_model* om, void* stream) {
  if (!st || !om) { bo_set_error("set_hyperparameters: null argument"); return BO_ERR_INVALID; }
  if (m < 0 || m >= st->M) { bo_set_error("set_hyperparameters: bad output index"); return BO_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  OutputH& o = st->out[m];
  if (om->n_leaves != o.md.n_leaves || om->n_terms != o.md.n_terms || !om->leaves || !om->terms) {
    bo_set_error("set_hyperparameters: the kernel tree must keep its shape (%d leaves, %d terms)", o.md.n_leaves, o.md.n_terms);
    return BO_ERR_INVALID;
  }
  if (!(om->noise >= 0.0)) { bo_set_error("set_hyperparameters: noise must be >= 0"); return BO_ERR_INVALID; }
  for (int t = 0; t < om->n_terms; ++t) {
    const bo_kernel_term& kt = om->terms[t];
    if (kt.n_factors != o.md.nfac[t]) { bo_set_error("set_hyperparameters: term %d changed its factors", t); return BO_ERR_INVALID; }
    for (int f = 0; f < kt.n_factors; ++f)
      if (kt.factors[f] != o.md.fac[t][f]) { bo_set_error("set_hyperparameters: term %d changed its factors", t); return BO_ERR_INVALID; }
  }
  for (int l = 0; l < om->n_leaves; ++l) {
    const bo_kernel_leaf& kl = om->leaves[l];
    const LeafD& L = o.md.leaf[l];
    if (kl.kind != L.kind || kl.n_dims != L.nd) { bo_set_error("set_hyperparameters: leaf %d changed kind / dims", l); return BO_ERR_INVALID; }
    if (kl.kind == BO_LEAF_TANIMOTO) continue;
    if (!kl.lengthscale || (kl.n_ls != 1 && kl.n_ls < kl.n_dims)) { bo_set_error("set_hyperparameters: leaf %d: bad lengthscale spec", l); return BO_ERR_INVALID; }
    std::vector<double> v(kl.n_dims);
    for (int k = 0; k < kl.n_dims; ++k) {
      const double ls = kl.lengthscale[kl.n_ls == 1 ? 0 : k];
      if (!(ls > 0.0)) { bo_set_error("set_hyperparameters: leaf %d: non-positive lengthscale", l); return BO_ERR_INVALID; }
      v[k] = (kl.kind == BO_LEAF_HAMMING) ? 1.0 / ls : ls;
    }
    double* dst = const_cast<double*>(kl.kind == BO_LEAF_HAMMING ? L.wls : L.ls);
    CUDA_CHECK_RET(cudaMemcpyAsync(dst, v.data(), v.size() * 8, cudaMemcpyHostToDevice, s));
    CUDA_CHECK_RET(cudaStreamSynchronize(s));      // v lives on this stack frame
  }
  for (int t = 0; t < om->n_terms; ++t) o.md.coef[t] = om->terms[t].coef;
  const bool mean_changed = om->mean_const != o.md.mean_const;
  o.md.noise = om->noise; o.md.mean_const = om->mean_const;
  RC(launch_prep_points(o.md, st->X_train.as<double>(), st->N, st->d, o.train_prepd, s, &st->lc));
  if (mean_changed) {
    std::vector<double> r(st->ldk, 0.0);
    for (int i = 0; i < st->N; ++i) r[i] = (o.y_host[i] - o.md.y_mean) / o.md.y_std - o.md.mean_const;
    CUDA_CHECK_RET(cudaMemcpyAsync(o.resid.p, r.data(), (size_t)st->ldk * 8, cudaMemcpyHostToDevice, s));
    CUDA_CHECK_RET(cudaStreamSynchronize(s));
  }
  st->factorized = false;
  st->acqf_kind = 0;
  o.kinv_ready = false;
  o.oz_ready = false;
  return BO_OK;
}

// psd-safe blocked Cholesky of `src` (n x n, ld) into `dst`: jitter ladder 1e-8 .. 1e-3 on the whole diagonal.
static int psd_safe_chol(bo_state* st, const double* src, double* dst, int ld, int n, DevBuf& dinv, int* info_out,
                         double* jitter_out, cudaStream_t s) {
  RC(dinv.ensure((size_t)((n + 63) / 64) * 64 * 64 * 8));
  RC(st->wsInfo.ensure(64));
  int* info_dev = st->wsInfo.as<int>();
  double applied = 0.0;
  for (int attempt = 0; attempt <= 6; ++attempt) {
    CUDA_CHECK_RET(cudaMemcpyAsync(dst, src, (size_t)n * ld * 8, cudaMemcpyDeviceToDevice, s));
    if (attempt > 0) {
      applied = 1e-8 * pow(10.0, attempt - 1);
      RC(launch_add_diag(dst, ld, n, applied, s, &st->lc));
    }
    CUDA_CHECK_RET(cudaMemsetAsync(info_dev, 0, sizeof(int), s));
    RC(chol_blocked(dst, ld, n, dinv.as<double>(), info_dev, s, &st->lc));
    int info = 0;
    CUDA_CHECK_RET(cudaMemcpyAsync(&info, info_dev, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK_RET(cudaStreamSynchronize(s));
    *info_out = info;
    if (info == 0) { if (jitter_out) *jitter_out = applied; return BO_OK; }
  }
  if (jitter_out) *jitter_out = applied;
  return BO_OK;  // caller inspects info
}

// LinvExt = [L^-1 (N rows) ; alpha ; n_b baseline rows (filled by the caller) ; zero padding to 128 rows]
static int build_linv_ext(bo_state* st, OutputH& o, int nb, cudaStream_t s) {
  const int N = st->N, ldk = st->ldk;
  o.Rpad = round_up(N + 1 + nb, 128);
  o.oz_ready = false;
  st->oz_calib = 0;   // a new factor: the INT8 path gets a fresh chance, its per-row guard decides again
  st->oz_flagged_total = 0; st->oz_batches_total = 0;
  RC(o.LinvExt.ensure((size_t)o.Rpad * ldk * 8, true));
  CUDA_CHECK_RET(cudaMemcpyAsync(o.LinvExt.p, o.Linv.p, (size_t)N * ldk * 8, cudaMemcpyDeviceToDevice, s));
  CUDA_CHECK_RET(cudaMemcpyAsync(o.LinvExt.as<double>() + (size_t)N * ldk, o.alpha_row.p, (size_t)ldk * 8, cudaMemcpyDeviceToDevice, s));
  return BO_OK;
}

extern "C" int bo_state_factorize(bo_state* st, int32_t* info, double* jitter, void* stream) {
  if (!st) { bo_set_error("null state"); return BO_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  const int N = st->N, ldk = st->ldk, Nr = st->Nr;
  bool ok = true;
  for (int m = 0; m < st->M; ++m) {
    OutputH& o = st->out[m];
    RC(st->wsCov.ensure((size_t)Nr * ldk * 8, true));
    double* G = st->wsCov.as<double>();
    RC(launch_crosscov(o.md, o.train_prepd, o.train_prepd, true, N, G, ldk, true, s, &st->lc));
    RC(launch_add_diag(G, ldk, N, o.md.noise, s, &st->lc));
    RC(o.L.ensure((size_t)Nr * ldk * 8, true));
    int inf = 0; double jit = 0.0;
    RC(psd_safe_chol(st, G, o.L.as<double>(), ldk, N, o.dinv, &inf, &jit, s));
    if (info) info[m] = inf;
    if (jitter) jitter[m] = jit;
    o.jitter = jit;
    o.kinv_ready = false;
    if (inf != 0) { ok = false; continue; }
    RC(o.Linv.ensure((size_t)Nr * ldk * 8, true));
    RC(o.LinvT.ensure((size_t)Nr * ldk * 8, true));
    RC(st->wsTmp.ensure((size_t)64 * ldk * 8, true));
    RC(tri_inverse_blocked(o.L.as<double>(), ldk, N, o.dinv.as<double>(), o.Linv.as<double>(), o.LinvT.as<double>(), ldk,
                           st->wsTmp.as<double>(), s, &st->lc));
    // alpha = L^-T (L^-1 r)
    RC(o.tvec.ensure((size_t)ldk * 8, true));
    RC(o.alpha_row.ensure((size_t)ldk * 8, true));
    RC(launch_gemm_nt(N, 1, N, 1.0, o.Linv.as<double>(), ldk, o.resid.as<double>(), ldk, 0.0, o.tvec.as<double>(), 1, false, s, &st->lc));
    RC(launch_gemm_nt(N, 1, N, 1.0, o.LinvT.as<double>(), ldk, o.tvec.as<double>(), ldk, 0.0, o.alpha_row.as<double>(), 1, false, s, &st->lc));
    RC(build_linv_ext(st, o, 0, s));
  }
  CUDA_CHECK_RET(cudaStreamSynchronize(s));
  st->factorized = ok;
  if (!ok) { bo_set_error("training Gram matrix not positive definite after jitter 1e-3"); return BO_ERR_NOT_PSD; }
  return BO_OK;
}

// mean [n, M] (ldo = M), optional cov_m written by callback-free path below
static int posterior_blocks(bo_state* st, int m, const double* X_dev, int n, PrepBuf& pb, PrepD* pd, double* Kx,
                            double* V, double* mu_raw, cudaStream_t s) {
  OutputH& o = st->out[m];
  RC(pb.ensure(o.md, n, pd));
  RC(launch_prep_points(o.md, X_dev, n, st->d, *pd, s, &st->lc));
  RC(launch_crosscov(o.md, *pd, o.train_prepd, true, st->N, Kx, st->ldk, false, s, &st->lc));
  if (V) RC(launch_gemm_nt(n, st->N, st->N, 1.0, Kx, st->ldk, o.Linv.as<double>(), st->ldk, 0.0, V, st->ldk, false, s, &st->lc));
  if (mu_raw) RC(launch_gemm_nt(n, 1, st->N, 1.0, Kx, st->ldk, o.alpha_row.as<double>(), st->ldk, 0.0, mu_raw, 1, false, s, &st->lc));
  return BO_OK;
}

extern "C" int bo_posterior_joint(bo_state* st, const double* X_dev, int32_t n, double* mean_dev, double* cov_dev,
                                  void* stream) {
  if (!st || !st->factorized) { bo_set_error("state not factorized"); return BO_ERR_STATE; }
  cudaStream_t s = (cudaStream_t)stream;
  const int ldk = st->ldk, ldn = round_up(n, 16);
  RC(st->wsKx.ensure((size_t)n * ldk * 8));
  RC(st->wsV.ensure((size_t)n * ldk * 8, true));
  RC(st->wsMuRaw.ensure((size_t)n * 8));
  RC(st->wsCov.ensure((size_t)n * ldn * 8));
  for (int m = 0; m < st->M; ++m) {
    OutputH& o = st->out[m];
    RC(posterior_blocks(st, m, X_dev, n, o.q_prep, &o.q_prepd, st->wsKx.as<double>(), st->wsV.as<double>(), st->wsMuRaw.as<double>(), s));
    RC(launch_finish_mean(st->wsMuRaw.as<double>(), n, o.md.mean_const, o.md.y_std, o.md.y_mean, mean_dev, st->M, m, s, &st->lc));
    if (cov_dev) {
      RC(launch_crosscov(o.md, o.q_prepd, o.q_prepd, false, n, st->wsCov.as<double>(), ldn, true, s, &st->lc));
      RC(launch_gemm_nt(n, n, st->N, -1.0, st->wsV.as<double>(), ldk, st->wsV.as<double>(), ldk, 1.0, st->wsCov.as<double>(), ldn, false, s, &st->lc));
      RC(launch_copy_scale(st->wsCov.as<double>(), ldn, cov_dev + (size_t)m * n * n, n, n, n, o.md.y_std * o.md.y_std, s, &st->lc));
    }
  }
  return BO_OK;
}

extern "C" int bo_posterior_marginal(bo_state* st, const double* X_dev, int32_t n, int32_t observation_noise,
                                     double* mean_dev, double* var_dev, void* stream) {
  if (!st || !st->factorized) { bo_set_error("state not factorized"); return BO_ERR_STATE; }
  cudaStream_t s = (cudaStream_t)stream;
  const int ldk = st->ldk;
  RC(st->wsKx.ensure((size_t)n * ldk * 8));
  RC(st->wsMuRaw.ensure((size_t)n * 8));
  RC(st->wsGqq.ensure((size_t)n * 8));
  for (int m = 0; m < st->M; ++m) {
    OutputH& o = st->out[m];
    RC(posterior_blocks(st, m, X_dev, n, o.q_prep, &o.q_prepd, st->wsKx.as<double>(), nullptr, nullptr, s));
    PostGemmArgs a;
    a.Kx = st->wsKx.as<double>(); a.rows = n; a.ldk = ldk; a.B = o.LinvExt.as<double>();
    a.N = st->N; a.Rpad = o.Rpad; a.n_ext = 1; a.q = 1; a.Gqq = st->wsGqq.as<double>(); a.W = nullptr; a.ldw = 0;
    a.mu_raw = st->wsMuRaw.as<double>();
    RC(launch_posterior_gemm(a, s, &st->lc));
    RC(launch_finish_mean(st->wsMuRaw.as<double>(), n, o.md.mean_const, o.md.y_std, o.md.y_mean, mean_dev, st->M, m, s, &st->lc));
    RC(launch_finish_var(o.md, o.q_prepd, st->wsGqq.as<double>(), n, observation_noise, var_dev, st->M, m, s, &st->lc));
  }
  return BO_OK;
}

static int fill_objd(ObjD* od, const bo_objective_op* obj, int n_obj, const bo_constraint_op* cons, int n_cons, int combine,
                     int M) {
  if (n_obj < 1 || n_obj > BO_MAX_OBJECTIVES || n_cons < 0 || n_cons > BO_MAX_CONSTRAINTS) {
    bo_set_error("n_obj=%d / n_cons=%d out of range", n_obj, n_cons);
    return BO_ERR_INVALID;
  }
  memset(od, 0, sizeof(ObjD));
  od->n_obj = n_obj; od->n_cons = n_cons; od->combine = combine;
  for (int i = 0; i < n_obj; ++i) {
    if (obj[i].out_idx < 0 || obj[i].out_idx >= M || obj[i].kind < 0 || obj[i].kind > BO_OBJ_TARGET) { bo_set_error("objective %d invalid", i); return BO_ERR_INVALID; }
    od->op[i] = obj[i];
  }
  for (int i = 0; i < n_cons; ++i) {
    if (cons[i].out_idx < 0 || cons[i].out_idx >= M || !(cons[i].eta > 0.0)) { bo_set_error("constraint %d invalid", i); return BO_ERR_INVALID; }
    od->con[i] = cons[i];
  }
  return BO_OK;
}

// Joint posterior at X (n points): mean_dev [n, M]; per output the lower Cholesky root of the
// un-standardised covariance into roots[m] ([n, ldn]); V rows optionally kept in Vkeep[m] ([n, ldk]).
static int joint_root(bo_state* st, int m, const double* X_dev, int n, int ldn, PrepBuf& pb, PrepD* pd, double* mean_dev,
                      DevBuf& rootbuf, DevBuf& covcopy, double* Vdst, int* info, double* jit, cudaStream_t s,
                      DevBuf* dinv_keep = nullptr) {
  OutputH& o = st->out[m];
  const int ldk = st->ldk;
  RC(st->wsKx.ensure((size_t)n * ldk * 8));
  RC(st->wsMuRaw.ensure((size_t)n * 8));
  RC(posterior_blocks(st, m, X_dev, n, pb, pd, st->wsKx.as<double>(), Vdst, st->wsMuRaw.as<double>(), s));
  RC(launch_finish_mean(st->wsMuRaw.as<double>(), n, o.md.mean_const, o.md.y_std, o.md.y_mean, mean_dev, st->M, m, s, &st->lc));
  RC(covcopy.ensure((size_t)n * ldn * 8, true));
  RC(rootbuf.ensure((size_t)n * ldn * 8, true));
  double* C = covcopy.as<double>();
  RC(launch_crosscov(o.md, *pd, *pd, false, n, C, ldn, true, s, &st->lc));
  RC(launch_gemm_nt(n, n, st->N, -1.0, Vdst, ldk, Vdst, ldk, 1.0, C, ldn, false, s, &st->lc));
  RC(launch_scale_matrix(C, ldn, n, n, o.md.y_std * o.md.y_std, s, &st->lc));
  DevBuf& dinv = dinv_keep ? *dinv_keep : st->wsJointDinv;
  return psd_safe_chol(st, C, rootbuf.as<double>(), ldn, n, dinv, info, jit, s);
}

extern "C" int bo_prune_counts(bo_state* st, const double* X_dev, int32_t n, const double* z_dev, int32_t S,
                               const bo_objective_op* obj, int32_t n_obj, const bo_constraint_op* cons, int32_t n_cons,
                               const double* ref_point, int32_t* counts_dev, int32_t* info, void* stream) {
  if (!st || !st->factorized) { bo_set_error("state not factorized"); return BO_ERR_STATE; }
  cudaStream_t s = (cudaStream_t)stream;
  ObjD od;
  RC(fill_objd(&od, obj, n_obj, cons, n_cons, 0, st->M));
  st->baseline_f_valid = false;   // wsF is reused below
  const int M = st->M, ldn = round_up(n, 16), ldk = st->ldk;
  SetupTrace tr("prune_counts", s);
  RC(st->wsMean.ensure((size_t)n * M * 8));
  RC(st->wsZM.ensure((size_t)M * S * ldn * 8, true));
  RC(st->wsF.ensure((size_t)M * S * ldn * 8));
  RC(st->wsV.ensure((size_t)n * ldk * 8, true));
  tr.mark("workspaces");
  RC(launch_transpose_base_samples(z_dev, S, n, M, nullptr, st->wsZM.as<double>(), ldn, s, &st->lc));
  // persistent temporaries of the handle: allocating and freeing 2 x N^2 doubles per call cost up to 0.5 s of cudaFree
  DevBuf& root = st->wsJointRoot; DevBuf& cov = st->wsJointCov;
  PrepBuf& pb = st->jointPrep; PrepD pd;
  int rcode = BO_OK;
  for (int m = 0; m < M && rcode == BO_OK; ++m) {
    int inf = 0; double jit = 0;
    rcode = joint_root(st, m, X_dev, n, ldn, pb, &pd, st->wsMean.as<double>(), root, cov, st->wsV.as<double>(), &inf, &jit, s);
    tr.mark("joint_root");
    if (info) info[m] = inf;
    if (rcode == BO_OK && inf != 0) { bo_set_error("posterior covariance at the baseline not p.d. (output %d)", m); rcode = BO_ERR_NOT_PSD; }
    if (rcode == BO_OK)
      rcode = launch_gemm_nt(S, n, n, 1.0, st->wsZM.as<double>() + (size_t)m * S * ldn, ldn, root.as<double>(), ldn, 0.0,
                             st->wsF.as<double>() + (size_t)m * S * ldn, ldn, false, s, &st->lc);
    tr.mark("sample gemm");
  }
  if (rcode == BO_OK) rcode = st->wsObj.ensure((size_t)S * n * n_obj * 8);
  if (rcode == BO_OK) rcode = st->wsFeas.ensure((size_t)S * n);
  if (rcode == BO_OK) rcode = st->ref_dev.ensure(BO_MAX_OBJECTIVES * 8);
  if (rcode == BO_OK) {
    cudaMemcpyAsync(st->ref_dev.p, ref_point, n_obj * 8, cudaMemcpyHostToDevice, s);
    rcode = launch_baseline_objective(st->wsF.as<double>(), ldn, S, n, M, st->wsMean.as<double>(), od, st->wsObj.as<double>(),
                                      st->wsFeas.as<unsigned char>(), nullptr, s, &st->lc);
  }
  if (rcode == BO_OK) {
    cudaMemsetAsync(counts_dev, 0, (size_t)n * sizeof(int), s);
    rcode = launch_front(st->wsObj.as<double>(), st->wsFeas.as<unsigned char>(), S, n, n_obj, st->ref_dev.as<double>(), 0,
                         nullptr, counts_dev, s, &st->lc);
  }
  cudaStreamSynchronize(s);
  tr.mark("objective + front");
  return rcode;
}

static int build_cells(bo_state* st, const double* obj, const unsigned char* feas, int n, int S, int Mo, int* max_cells,
                       cudaStream_t s) {
  RC(st->wsFront.ensure((size_t)S * std::max(n, 1)));
  RC(launch_front(obj, feas, S, n, Mo, st->ref_dev.as<double>(), 1, st->wsFront.as<unsigned char>(), nullptr, s, &st->lc));
  RC(st->ncells.ensure((size_t)S * sizeof(int)));
  if (Mo == 2) {
    st->cap = n + 1;
    RC(st->cell_lo.ensure((size_t)st->cap * Mo * S * 8));
    RC(st->cell_up.ensure((size_t)st->cap * Mo * S * 8));
    RC(st->front_idx.ensure((size_t)S * st->cap * sizeof(int)));
    CUDA_CHECK_RET(cudaMemsetAsync(st->front_idx.p, 0xff, (size_t)S * st->cap * sizeof(int), s));
    RC(launch_partition2d(obj, st->wsFront.as<unsigned char>(), n, S, st->cap, st->ref_dev.as<double>(),
                          st->cell_lo.as<double>(), st->cell_up.as<double>(), st->ncells.as<int>(), st->front_idx.as<int>(), s, &st->lc));
  } else {
    // capacity: start generous, double on overflow
    int cap = std::max(64, 8 * (n + 1) * Mo);
    for (;;) {
      st->cap = cap;
      RC(st->cell_lo.ensure((size_t)cap * Mo * S * 8));
      RC(st->cell_up.ensure((size_t)cap * Mo * S * 8));
      RC(st->wsInfo.ensure(64));
      CUDA_CHECK_RET(cudaMemsetAsync(st->wsInfo.p, 0, sizeof(int), s));
      if (st->part_alpha > 0.0) {
        // approximate partitioning (BoFire's `alpha`): binary partitioning, cells below the volume threshold are dropped
        RC(st->wsPart.ensure((size_t)S * partition_binary_work_stride(n, Mo)));
        RC(launch_partition_binary(obj, st->wsFront.as<unsigned char>(), n, S, Mo, cap, st->ref_dev.as<double>(), st->part_alpha,
                                   st->wsPart.as<double>(), st->cell_lo.as<double>(), st->cell_up.as<double>(), st->ncells.as<int>(),
                                   st->wsInfo.as<int>(), s, &st->lc));
      } else {
        RC(st->wsPart.ensure((size_t)S * 2 * cap * (Mo + Mo * Mo) * 8));
        RC(launch_partition_nd(obj, st->wsFront.as<unsigned char>(), n, S, Mo, cap, st->ref_dev.as<double>(), st->wsPart.as<double>(),
                               st->cell_lo.as<double>(), st->cell_up.as<double>(), st->ncells.as<int>(), st->wsInfo.as<int>(), s, &st->lc));
      }
      int overflow = 0;
      CUDA_CHECK_RET(cudaMemcpyAsync(&overflow, st->wsInfo.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK_RET(cudaStreamSynchronize(s));
      if (!overflow) break;
      cap *= 2;
      if ((size_t)cap * (Mo + Mo * Mo) * S * 16 > ((size_t)48 << 30)) { bo_set_error("box decomposition exceeds the workspace budget"); return BO_ERR_INVALID; }
    }
  }
  std::vector<int> nc(S);
  CUDA_CHECK_RET(cudaMemcpyAsync(nc.data(), st->ncells.p, (size_t)S * sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK_RET(cudaStreamSynchronize(s));
  int mx = 0;
  for (int v : nc) mx = std::max(mx, v);
  if (max_cells) *max_cells = mx;
  return BO_OK;
}

// Cached-root state of a baseline point set (shared by qNEHVI / qLogNEHVI and qNEI / qLogNEI): posterior root L_b and
// its inverse per output, the extra rows K_bX (K + s2 I)^-1 of LinvExt, the transposed base samples and the baseline
// samples F = z_b L_b^T (wsF, [M][S][ldlb]) with their mean (mean_b).
static int prepare_baseline(bo_state* st, const double* Xb_dev, int nb, const double* zb_dev, int S, int32_t* info,
                            cudaStream_t s) {
  const int M = st->M, ldk = st->ldk, ldlb = round_up(std::max(nb, 1), 16);
  st->nb = nb; st->S = S; st->ldlb = ldlb;
  st->baseline_f_valid = false;
  // raw baseline points: the joint re-sampling fallback scores (baseline, q-batch) together
  RC(st->Xb_raw.ensure((size_t)std::max(nb, 1) * st->d * 8));
  if (nb > 0) CUDA_CHECK_RET(cudaMemcpyAsync(st->Xb_raw.p, Xb_dev, (size_t)nb * st->d * 8, cudaMemcpyDeviceToDevice, s));
  RC(st->mean_b.ensure((size_t)std::max(nb, 1) * M * 8));
  RC(st->zbT.ensure((size_t)std::max(nb, 1) * M * S * 8));
  RC(st->zbM.ensure((size_t)M * S * ldlb * 8, true));
  RC(st->wsF.ensure((size_t)M * S * ldlb * 8));
  if (nb > 0) RC(launch_transpose_base_samples(zb_dev, S, nb, M, st->zbT.as<double>(), st->zbM.as<double>(), ldlb, s, &st->lc));
  for (int m = 0; m < M; ++m) {
    OutputH& o = st->out[m];
    if (nb > 0) {
      int inf = 0; double jit = 0;
      RC(st->wsV.ensure((size_t)nb * ldk * 8, true));
      DevBuf dinv_b;
      int rcj = joint_root(st, m, Xb_dev, nb, ldlb, o.base_prep, &o.base_prepd, st->mean_b.as<double>(), o.Lb, o.Sbb, st->wsV.as<double>(), &inf, &jit, s, &dinv_b);
      if (info) info[m] = inf;
      if (rcj != BO_OK) { dinv_b.release(); return rcj; }
      if (inf != 0) { dinv_b.release(); bo_set_error("baseline posterior covariance not p.d. (output %d)", m); return BO_ERR_NOT_PSD; }
      // inverse of the cached baseline root: the per-q-batch triangular solve becomes a small product
      int rci = o.LbInv.ensure((size_t)round_up(nb, 64) * ldlb * 8, true);
      if (rci == BO_OK) rci = o.LbInvT.ensure((size_t)round_up(nb, 64) * ldlb * 8, true);
      if (rci == BO_OK) rci = st->wsTmp.ensure((size_t)64 * std::max(ldlb, st->ldk) * 8, true);
      if (rci == BO_OK) rci = tri_inverse_blocked(o.Lb.as<double>(), ldlb, nb, dinv_b.as<double>(), o.LbInv.as<double>(), o.LbInvT.as<double>(), ldlb, st->wsTmp.as<double>(), s, &st->lc);
      cudaStreamSynchronize(s);
      dinv_b.release();
      if (rci != BO_OK) return rci;
      // extra rows = V_b L^-1 = K_bX (K + s2 I)^-1, so that K*X (.)^T = V_q V_b^T
      RC(build_linv_ext(st, o, nb, s));
      RC(launch_gemm_nt(nb, st->N, st->N, 1.0, st->wsV.as<double>(), ldk, o.LinvT.as<double>(), ldk, 0.0,
                        o.LinvExt.as<double>() + (size_t)(st->N + 1) * ldk, ldk, false, s, &st->lc));
      RC(launch_gemm_nt(S, nb, nb, 1.0, st->zbM.as<double>() + (size_t)m * S * ldlb, ldlb, o.Lb.as<double>(), ldlb, 0.0,
                        st->wsF.as<double>() + (size_t)m * S * ldlb, ldlb, false, s, &st->lc));
    } else {
      RC(o.base_prep.ensure(o.md, 0, &o.base_prepd));
      RC(o.Lb.ensure(16));
      RC(build_linv_ext(st, o, 0, s));
      if (info) info[m] = 0;
    }
  }
  return BO_OK;
}

extern "C" int bo_nehvi_prepare(bo_state* st, const double* Xb_dev, int32_t n_b, const double* zb_dev, int32_t S,
                                const bo_objective_op* obj, int32_t n_obj, const bo_constraint_op* cons, int32_t n_cons,
                                const double* ref_point, int32_t* info, int32_t* max_cells, void* stream) {
  if (!st || !st->factorized) { bo_set_error("state not factorized"); return BO_ERR_STATE; }
  if (n_obj < 2) { bo_set_error("qNEHVI needs at least two objectives"); return BO_ERR_INVALID; }
  if (S < 1 || n_b < 0) { bo_set_error("bad S / n_b"); return BO_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  st->acqf_kind = 0; st->log_hvi = 0; st->noisy_scalar = false;
  RC(fill_objd(&st->od, obj, n_obj, cons, n_cons, 0, st->M));
  const int M = st->M, nb = n_b;
  st->cells_shared = 0;
  RC(st->ref_dev.ensure(BO_MAX_OBJECTIVES * 8));
  CUDA_CHECK_RET(cudaMemcpyAsync(st->ref_dev.p, ref_point, n_obj * 8, cudaMemcpyHostToDevice, s));
  SetupTrace tr("nehvi_prepare", s);
  RC(prepare_baseline(st, Xb_dev, nb, zb_dev, S, info, s));
  tr.mark("prepare_baseline");
  const int ldlb = st->ldlb;
  RC(st->obj_b.ensure((size_t)S * std::max(nb, 1) * n_obj * 8));
  RC(st->samples_b.ensure((size_t)S * std::max(nb, 1) * M * 8));
  RC(st->wsFeas.ensure((size_t)S * std::max(nb, 1)));
  RC(launch_baseline_objective(st->wsF.as<double>(), ldlb, S, nb, M, st->mean_b.as<double>(), st->od, st->obj_b.as<double>(),
                               st->wsFeas.as<unsigned char>(), st->samples_b.as<double>(), s, &st->lc));
  RC(build_cells(st, st->obj_b.as<double>(), st->wsFeas.as<unsigned char>(), nb, S, n_obj, &st->max_cells, s));
  tr.mark("cells");
  if (max_cells) *max_cells = st->max_cells;
  st->acqf_kind = 1;
  return BO_OK;
}

extern "C" int bo_ehvi_prepare(bo_state* st, const double* Yobj_dev, int32_t n, int32_t S, const bo_objective_op* obj,
                               int32_t n_obj, const double* ref_point, int32_t* max_cells, void* stream) {
  if (!st || !st->factorized) { bo_set_error("state not factorized"); return BO_ERR_STATE; }
  cudaStream_t s = (cudaStream_t)stream;
  st->acqf_kind = 0; st->log_hvi = 0; st->noisy_scalar = false;
  RC(fill_objd(&st->od, obj, n_obj, nullptr, 0, 0, st->M));
  st->nb = 0; st->S = S; st->ldlb = 16; st->cells_shared = 1;
  RC(st->ref_dev.ensure(BO_MAX_OBJECTIVES * 8));
  CUDA_CHECK_RET(cudaMemcpyAsync(st->ref_dev.p, ref_point, n_obj * 8, cudaMemcpyHostToDevice, s));
  RC(st->wsFeas.ensure((size_t)std::max(n, 1)));
  CUDA_CHECK_RET(cudaMemsetAsync(st->wsFeas.p, 1, (size_t)std::max(n, 1), s));
  RC(st->zbT.ensure(16));
  for (int m = 0; m < st->M; ++m) {
    OutputH& o = st->out[m];
    RC(o.base_prep.ensure(o.md, 0, &o.base_prepd));
    RC(o.Lb.ensure(16));
    RC(build_linv_ext(st, o, 0, s));
  }
  RC(build_cells(st, Yobj_dev, st->wsFeas.as<unsigned char>(), n, 1, n_obj, &st->max_cells, s));
  if (max_cells) *max_cells = st->max_cells;
  st->acqf_kind = 2;
  return BO_OK;
}

extern "C" int bo_scalar_prepare(bo_state* st, int32_t variant, double param, int32_t S, int32_t combine,
                                 const bo_objective_op* obj, int32_t n_obj, const bo_constraint_op* cons, int32_t n_cons,
                                 double best_f, const double* Xb_dev, int32_t n_b, const double* zb_dev, int32_t* info,
                                 void* stream) {
  if (!st || !st->factorized) { bo_set_error("state not factorized"); return BO_ERR_STATE; }
  cudaStream_t s = (cudaStream_t)stream;
  st->acqf_kind = 0; st->log_hvi = 0; st->noisy_scalar = false;
  if (variant < BO_ACQF_QLOGEI || variant > BO_ACQF_QPI) { bo_set_error("unknown scalar acquisition variant %d", variant); return BO_ERR_INVALID; }
  if (combine == BO_COMBINE_SINGLE && n_obj != 1) { bo_set_error("single objective needs n_obj == 1"); return BO_ERR_INVALID; }
  if (S < 1 || n_b < 0) { bo_set_error("bad S / n_b"); return BO_ERR_INVALID; }
  if (n_b > 0 && variant != BO_ACQF_QEI && variant != BO_ACQF_QLOGEI) { bo_set_error("a baseline (noisy variant) exists for qNEI / qLogNEI only"); return BO_ERR_INVALID; }
  if (n_cons > 0 && (variant == BO_ACQF_QSR || variant == BO_ACQF_QUCB)) { bo_set_error("output constraints need a non-negative utility (not qSR / qUCB)"); return BO_ERR_INVALID; }
  if (variant == BO_ACQF_QPI && !(param > 0.0)) { bo_set_error("qPI needs tau > 0"); return BO_ERR_INVALID; }
  if (variant == BO_ACQF_QUCB && !(param >= 0.0)) { bo_set_error("qUCB needs beta >= 0"); return BO_ERR_INVALID; }
  RC(fill_objd(&st->od, obj, n_obj, cons, n_cons, combine, st->M));
  st->cells_shared = 0; st->best_f = best_f; st->scalar_variant = variant; st->vparam = param;
  if (n_b > 0) {
    RC(prepare_baseline(st, Xb_dev, n_b, zb_dev, S, info, s));
    RC(st->best_f_s.ensure((size_t)S * 8));
    RC(launch_baseline_best(st->wsF.as<double>(), st->ldlb, S, n_b, st->M, st->mean_b.as<double>(), st->od, -INFINITY,
                            st->best_f_s.as<double>(), nullptr, nullptr, s, &st->lc));
    st->baseline_f_valid = true;
    st->noisy_scalar = true;
  } else {
    st->nb = 0; st->S = S; st->ldlb = 16;
    RC(st->zbT.ensure(16));
    for (int m = 0; m < st->M; ++m) {
      OutputH& o = st->out[m];
      RC(o.base_prep.ensure(o.md, 0, &o.base_prepd));
      RC(o.Lb.ensure(16));
      RC(build_linv_ext(st, o, 0, s));
      if (info) info[m] = 0;
    }
  }
  st->acqf_kind = 3;
  return BO_OK;
}

extern "C" int bo_scalar_baseline_best(bo_state* st, double infeasible_value, int32_t* n_all_infeasible, void* stream) {
  if (!st || st->acqf_kind != 3 || !st->noisy_scalar || !st->baseline_f_valid) {
    bo_set_error("scalar_baseline_best: call directly after bo_scalar_prepare of a noisy variant (n_b > 0)");
    return BO_ERR_STATE;
  }
  cudaStream_t s = (cudaStream_t)stream;
  RC(st->wsInfo.ensure(64));
  CUDA_CHECK_RET(cudaMemsetAsync(st->wsInfo.p, 0, sizeof(int), s));
  RC(launch_baseline_best(st->wsF.as<double>(), st->ldlb, st->S, st->nb, st->M, st->mean_b.as<double>(), st->od, infeasible_value,
                          st->best_f_s.as<double>(), nullptr, st->wsInfo.as<int>(), s, &st->lc));
  int cnt = 0;
  CUDA_CHECK_RET(cudaMemcpyAsync(&cnt, st->wsInfo.p, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK_RET(cudaStreamSynchronize(s));
  if (n_all_infeasible) *n_all_infeasible = cnt;
  return BO_OK;
}

extern "C" int bo_logei_prepare(bo_state* st, int32_t S, int32_t combine, const bo_objective_op* obj, int32_t n_obj,
                                double best_f, void* stream) {
  return bo_scalar_prepare(st, BO_ACQF_QLOGEI, 0.0, S, combine, obj, n_obj, nullptr, 0, best_f, nullptr, 0, nullptr, nullptr, stream);
}

extern "C" int bo_prune_counts_scalar(bo_state* st, const double* X_dev, int32_t n, const double* z_dev, int32_t S,
                                      int32_t combine, const bo_objective_op* obj, int32_t n_obj,
                                      const bo_constraint_op* cons, int32_t n_cons, int32_t* counts_dev, int32_t* info,
                                      void* stream) {
  if (!st || !st->factorized) { bo_set_error("state not factorized"); return BO_ERR_STATE; }
  if (n < 1 || S < 1) { bo_set_error("bad n / S"); return BO_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  ObjD od;
  RC(fill_objd(&od, obj, n_obj, cons, n_cons, combine, st->M));
  st->baseline_f_valid = false;   // wsF is reused below
  const int M = st->M, ldn = round_up(n, 16), ldk = st->ldk;
  RC(st->wsMean.ensure((size_t)n * M * 8));
  RC(st->wsZM.ensure((size_t)M * S * ldn * 8, true));
  RC(st->wsF.ensure((size_t)M * S * ldn * 8));
  RC(st->wsV.ensure((size_t)n * ldk * 8, true));
  RC(launch_transpose_base_samples(z_dev, S, n, M, nullptr, st->wsZM.as<double>(), ldn, s, &st->lc));
  // persistent temporaries of the handle: allocating and freeing 2 x N^2 doubles per call cost up to 0.5 s of cudaFree
  DevBuf& root = st->wsJointRoot; DevBuf& cov = st->wsJointCov;
  PrepBuf& pb = st->jointPrep; PrepD pd;
  int rcode = BO_OK;
  for (int m = 0; m < M && rcode == BO_OK; ++m) {
    int inf = 0; double jit = 0;
    rcode = joint_root(st, m, X_dev, n, ldn, pb, &pd, st->wsMean.as<double>(), root, cov, st->wsV.as<double>(), &inf, &jit, s);
    if (info) info[m] = inf;
    if (rcode == BO_OK && inf != 0) { bo_set_error("posterior covariance at the baseline not p.d. (output %d)", m); rcode = BO_ERR_NOT_PSD; }
    if (rcode == BO_OK)
      rcode = launch_gemm_nt(S, n, n, 1.0, st->wsZM.as<double>() + (size_t)m * S * ldn, ldn, root.as<double>(), ldn, 0.0,
                             st->wsF.as<double>() + (size_t)m * S * ldn, ldn, false, s, &st->lc);
  }
  if (rcode == BO_OK) {
    cudaMemsetAsync(counts_dev, 0, (size_t)n * sizeof(int), s);
    rcode = launch_baseline_best(st->wsF.as<double>(), ldn, S, n, M, st->wsMean.as<double>(), od, -INFINITY, nullptr, counts_dev,
                                 nullptr, s, &st->lc);
  }
  cudaStreamSynchronize(s);
  return rcode;
}

extern "C" int bo_acqf_set_option(bo_state* st, const char* name, double value) {
  if (!st || !name) { bo_set_error("null argument"); return BO_ERR_INVALID; }
  std::string nm(name);
  if (nm == "ozaki") {
    if (value != 0.0 && value != 1.0 && value != 2.0) { bo_set_error("ozaki must be 0 (off), 1 (automatic) or 2 (always)"); return BO_ERR_INVALID; }
    st->ozaki = (int)value;
    st->oz_calib = 0;
  } else if (nm == "ozaki_tile") {
    // 0 = default (EVEREST_OZAKI_TILE or 128); 64 = one-pass 128x64 tiles, 128 = two-pass 128x128, 256 = two-pass on CTA pairs
    if (value != 0.0 && value != 64.0 && value != 128.0 && value != 256.0) { bo_set_error("ozaki_tile must be 0, 64, 128 or 256"); return BO_ERR_INVALID; }
    st->ozaki_tile = (int)value;
    st->oz_calib = 0;
  } else if (nm == "ozaki_guard_kappa") {
    if (!(value > 0.0)) { bo_set_error("ozaki_guard_kappa must be > 0"); return BO_ERR_INVALID; }
    st->oz_kappa = value;
  } else if (nm == "ozaki_guard_tol") {
    if (!(value > 0.0)) { bo_set_error("ozaki_guard_tol must be > 0"); return BO_ERR_INVALID; }
    st->oz_tol = value;
  } else if (nm == "log_hvi") {
    if (st->acqf_kind != 1 && st->acqf_kind != 2) { bo_set_error("log_hvi applies to a prepared qNEHVI / qEHVI"); return BO_ERR_STATE; }
    st->log_hvi = value != 0.0;
  } else if (nm == "tau_relu") {
    if (!(value > 0.0)) { bo_set_error("tau_relu must be > 0"); return BO_ERR_INVALID; }
    st->tau_relu = value;
  } else if (nm == "joint_fallback") {
    st->auto_fallback = value != 0.0;    // 0: flagged q-batches keep their NaN (info = 1), nothing is re-scored
  } else if (nm == "force_joint_fallback") {
    st->force_fallback = value != 0.0;   // tests: bo_acqf_resample_flagged re-scores every q-batch
  } else if (nm == "partition_alpha") {
    // BoFire's `alpha` (data_models/strategies/predictives/qnehvi.py:19): read by the NEXT prepare call.  0 = exact
    // decomposition (local upper bounds); > 0 = approximate binary partitioning for more than two objectives (two objectives
    // are always decomposed exactly, as in BoTorch); -1 = exact binary partitioning (BoTorch's NondominatedPartitioning cells)
    if (!(value >= 0.0 && value <= 0.5) && value != -1.0) { bo_set_error("partition_alpha must be in [0, 0.5] (or -1)"); return BO_ERR_INVALID; }
    st->part_alpha = (value == -1.0) ? 1e-300 : value;
  } else if (nm == "tau_max") {
    if (!(value > 0.0)) { bo_set_error("tau_max must be > 0"); return BO_ERR_INVALID; }
    st->tau_max = value;
  } else { bo_set_error("unknown option '%s'", name); return BO_ERR_INVALID; }
  return BO_OK;
}

static void rec_begin(bo_state* st, const char* name, cudaStream_t s) {
  if (!st->timing) return;
  TimingRec r;
  r.name = name;
  cudaEventCreate(&r.a); cudaEventCreate(&r.b);
  cudaEventRecord(r.a, s);
  st->recs.push_back(r);
}
static void rec_end(bo_state* st, cudaStream_t s) {
  if (!st->timing) return;
  cudaEventRecord(st->recs.back().b, s);
}

// (K + s2 I)^-1 = L^-T L^-1, needed only by the backward pass (U = K*X (K + s2 I)^-1)
static int ensure_kinv(bo_state* st, OutputH& o, cudaStream_t s) {
  if (o.kinv_ready) return BO_OK;
  RC(o.Kinv.ensure((size_t)st->Nr * st->ldk * 8, true));
  RC(launch_gemm_nt(st->N, st->N, st->N, 1.0, o.LinvT.as<double>(), st->ldk, o.LinvT.as<double>(), st->ldk, 0.0,
                    o.Kinv.as<double>(), st->ldk, false, s, &st->lc));
  o.kinv_ready = true;
  return BO_OK;
}

extern "C" int bo_acqf_resample_flagged(bo_state* st, const double* X_dev, int32_t b, int32_t q, const double* zq_dev,
                                        double* out_dev, int32_t* info_dev, int32_t* n_resampled, void* stream);

// forward(X[b, q, d]) -> out[b]; when dX_dev is given also d out[i] / d X[i] (each value depends on its own q-batch only)
static int acqf_run(bo_state* st, const double* X_dev, int32_t b, int32_t q, const double* zq_dev, double* out_dev,
                    double* dX_dev, int32_t* info_dev, void* stream) {
  if (!st || !st->factorized || st->acqf_kind == 0) { bo_set_error("forward before prepare"); return BO_ERR_STATE; }
  if (b == 0 && q >= 1 && q <= BO_MAX_Q) return BO_OK;   // empty t-batch: nothing to score (BoTorch returns an empty tensor)
  if (b < 1 || q < 1 || q > BO_MAX_Q) { bo_set_error("bad b=%d / q=%d (q <= %d)", b, q, BO_MAX_Q); return BO_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  if (st->timing) {
    for (auto& r : st->recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    st->recs.clear();
  }
  const int M = st->M, ldk = st->ldk, nb = st->nb, S = st->S, nr = nb + q;
  const int ldw = std::max(nb, 1);
  if (!dX_dev) st->fb_resampled_last = 0;
  // chunk the q-batches so that the K(X*,X) workspace (all outputs) stays below ~4 GiB
  long long max_rows = std::max<long long>(((long long)4 << 30) / ((long long)ldk * 8 * M), (long long)q);
  int bchunk = (int)std::min<long long>(b, std::max<long long>(1, max_rows / q));
  RC(st->wsZqT.ensure((size_t)q * M * S * 8));
  RC(launch_transpose_base_samples(zq_dev, S, q, M, st->wsZqT.as<double>(), nullptr, 0, s, &st->lc));
  const size_t rows_max = (size_t)bchunk * q;
  RC(st->wsKx.ensure(rows_max * ldk * 8 * M));
  RC(st->wsGqq.ensure(rows_max * q * 8 * M));
  RC(st->wsW.ensure(rows_max * ldw * 8 * M));
  RC(st->wsMuRaw.ensure(rows_max * 8 * M));
  RC(st->wsRoot.ensure((size_t)bchunk * M * q * nr * 8));
  RC(st->wsMu.ensure(rows_max * M * 8));
  RC(st->wsJit.ensure((size_t)bchunk * M * sizeof(int)));
  const bool sample_gemm = (nb > 0 && st->acqf_kind == 1);
  const int ldbl = st->ldlb;
  if (sample_gemm) {
    RC(st->wsBL.ensure(rows_max * ldbl * 8 * M));
    RC(st->wsFp.ensure(rows_max * (size_t)S * 8 * M));
  }
  RC(st->wsPartial.ensure((size_t)mc_hvi_partial_groups(S) * bchunk * 8));
  std::vector<PostGemmArgs> pg(M);
  for (int b0 = 0; b0 < b; b0 += bchunk) {
    const int bc = std::min(bchunk, b - b0), rows = bc * q;
    const double* Xc = X_dev + (size_t)b0 * q * st->d;
    // [UPSTREAM] sample_cached_cholesky's fallback applies to the forward pass of acquisition functions with a cached root
    const bool fb_active = st->auto_fallback && !dX_dev && nb > 0;
    int* info_chunk = info_dev ? info_dev + b0 : nullptr;
    if (fb_active) {
      if (!info_chunk) { RC(st->wsInfoOut.ensure((size_t)bchunk * sizeof(int))); info_chunk = st->wsInfoOut.as<int>(); }
      RC(st->wsFbCount.ensure(64));
      if (!st->pin_fb) {
        CUDA_CHECK_RET(cudaHostAlloc(reinterpret_cast<void**>(&st->pin_fb), 64, cudaHostAllocDefault));
        CUDA_CHECK_RET(cudaEventCreateWithFlags(&st->fb_event, cudaEventDisableTiming));
      }
    }
    static const bool no_skinny = getenv("EVEREST_NO_SKINNY") != nullptr;   // debugging switch, read once
    const bool small_rows = rows <= 64 && !no_skinny;
    // INT8 digit-plane GEMM: 0 = off, 1 = automatic (large problems: the slicing pass and the 448-column TMEM tiles only
    // pay off when the GEMM dominates), 2 = always (tests)
    // N <= 16384: 7 plane pairs x N x 2^14 per S32 accumulator stays below 2^31
    const bool oz_shape = !small_rows && (q == 1 || q == 2 || q == 4 || q == 8) && st->N <= 16384;
    const bool use_ozaki = oz_shape && (st->ozaki == 2 || (st->ozaki == 1 && st->oz_calib >= 0 && (long long)rows * st->N >= (1ll << 22) && st->N >= 512));
    // automatic mode: every INT8 call is followed by the per-row guard (ozaki.cu); flagged q-batches are redone in FP64
    const bool oz_guard = use_ozaki && st->ozaki == 1;
    const int oz_rows_alloc = round_up(rows, 128);
    const size_t oz_pa = use_ozaki ? ozaki_plane_bytes(oz_rows_alloc, ldk) : 0;
    std::vector<double> oz_scaleA(M, 1.0);
    std::vector<char> oz_fused(M, 0);
    if (use_ozaki) RC(st->wsOzA.ensure(oz_pa * M));
    // Refinement steps: the outputs are independent until the MC kernel, and every kernel of a 32-point step is latency
    // bound (a few CTAs): the chains of the odd outputs run on a second stream next to those of the even ones.
    const bool par = small_rows && M > 1 && !st->timing;
    // big screens: the conditional roots and sample GEMMs of the outputs (0.18 + 0.13 ms each on config 3) are medium-sized
    // kernels that do not fill the GPU either: forked as well (9.00 -> 8.87 ms per config-3 screen, same bits;
    // EVEREST_TAIL_FORK=0 switches it off)
    static const bool tail_fork_env = []() { const char* e = getenv("EVEREST_TAIL_FORK"); return !e || atoi(e) != 0; }();
    const bool par_tail = par || (tail_fork_env && M > 1 && !st->timing && !dX_dev);
    if ((par || par_tail) && !st->side2) {
      CUDA_CHECK_RET(cudaStreamCreateWithFlags(&st->side2, cudaStreamNonBlocking));
      CUDA_CHECK_RET(cudaEventCreateWithFlags(&st->side2_fork, cudaEventDisableTiming));
      CUDA_CHECK_RET(cudaEventCreateWithFlags(&st->side2_join, cudaEventDisableTiming));
    }
    auto fork2 = [&]() -> int {
      CUDA_CHECK_RET(cudaEventRecord(st->side2_fork, s));
      CUDA_CHECK_RET(cudaStreamWaitEvent(st->side2, st->side2_fork, 0));
      return BO_OK;
    };
    auto join2 = [&]() -> int {
      CUDA_CHECK_RET(cudaEventRecord(st->side2_join, st->side2));
      CUDA_CHECK_RET(cudaStreamWaitEvent(s, st->side2_join, 0));
      return BO_OK;
    };
    cudaStream_t const s_main = s;
    // Piece-wise input (host entry point, forward_host_impl): usable when this call is ONE chunk over all rows and every
    // output takes the fused single-leaf cross-covariance path into the digit planes.  Pieces outermost, outputs inside: the
    // K(X*,X) work of ALL outputs (1.5 ms on config 3) is what the 15.7 MB of host rows (1.3 ms over PCIe on this pool)
    // hide behind; everything after this stage sees the whole batch.  Otherwise all pieces are awaited first.
    bool pieces_done = false;
    if (!st->piece_row_end.empty()) {
      bool ok = use_ozaki && !dX_dev && b0 == 0 && bc == b && st->piece_row_end.back() == rows;
      for (int m = 0; m < M && ok; ++m) {
        const ModelD& md = st->out[m].md;
        ok = md.n_terms == 1 && md.nfac[0] == 1 && md.leaf[md.fac[0][0]].kind <= BO_LEAF_MATERN52 && md.leaf[md.fac[0][0]].dpad <= 32;
      }
      if (ok) {
        for (int m = 0; m < M; ++m) {
          OutputH& o = st->out[m];
          RC(o.q_prep.ensure(o.md, rows, &o.q_prepd));
          double kmax = 0.0;
          for (int t = 0; t < o.md.n_terms; ++t) kmax += fabs(o.md.coef[t]);   // every leaf is <= 1
          oz_scaleA[m] = ldexp(1.0, (int)ceil(log2(std::max(kmax, 1e-300) / 0.49)));
        }
        rec_begin(st, "crosscov", s);
        for (int i = 0, r0 = 0; i < (int)st->piece_row_end.size(); ++i) {
          const int r1 = st->piece_row_end[i];
          const bool last = i + 1 == (int)st->piece_row_end.size();
          RC(st->piece_wait(i));
          for (int m = 0; m < M; ++m) {
            OutputH& o = st->out[m];
            const int l = o.md.fac[0][0];
            const int dpad = o.md.leaf[l].dpad;
            PrepD view = o.q_prepd;                            // rows [r0, r1) of the prepared candidate points
            view.n = r1 - r0;
            view.Xs[l] = o.q_prepd.Xs[l] + (size_t)r0 * dpad;
            view.n2[l] = o.q_prepd.n2[l] + r0;
            RC(launch_prep_points(o.md, Xc + (size_t)r0 * st->d, r1 - r0, st->d, view, s, &st->lc));
            OzPlanesOut po;
            po.planes = st->wsOzA.as<signed char>() + (size_t)m * oz_pa; po.plane_stride = (long long)(ldk / 16) * oz_rows_alloc * 16;
            po.rows_alloc = oz_rows_alloc; po.n_chunks = ldk / 16; po.inv_scale = 1.0 / oz_scaleA[m]; po.write_fp64 = 0;
            po.row_offset = r0; po.rows_cover = last ? oz_rows_alloc - r0 : r1 - r0;
            bool fz = false;
            double* Kxm = st->wsKx.as<double>() + (size_t)m * rows_max * ldk;
            RC(launch_crosscov_ex(o.md, view, o.train_prepd, true, st->N, Kxm + (size_t)r0 * ldk, ldk, false, &po, &fz, s, &st->lc));
            if (!fz) { bo_set_error("piece-wise K(X*,X): fused path expected"); return BO_ERR_STATE; }
          }
          r0 = r1;
        }
        rec_end(st, s);
        pieces_done = true;
      } else {
        for (int i = 0; i < (int)st->piece_row_end.size(); ++i) RC(st->piece_wait(i));
      }
    }
    if (par) RC(fork2());
    for (int m = 0; m < M; ++m) {
      cudaStream_t s = (par && (m & 1)) ? st->side2 : s_main;      // shadows the call's stream inside this loop body
      OutputH& o = st->out[m];
      double* Kx = st->wsKx.as<double>() + (size_t)m * rows_max * ldk;
      RC(o.q_prep.ensure(o.md, rows, &o.q_prepd));
      if (pieces_done) {
        oz_fused[m] = 1;     // prepared and sliced piece by piece above
      } else {
      rec_begin(st, "prep", s);
      RC(launch_prep_points(o.md, Xc, rows, st->d, o.q_prepd, s, &st->lc));
      rec_end(st, s);
      rec_begin(st, "crosscov", s);
      if (use_ozaki) {
        // the cross-covariance kernel emits the INT8 digit planes directly (and the FP64 matrix only for the adjoint)
        double kmax = 0.0;
        for (int t = 0; t < o.md.n_terms; ++t) kmax += fabs(o.md.coef[t]);   // every leaf is <= 1
        oz_scaleA[m] = ldexp(1.0, (int)ceil(log2(std::max(kmax, 1e-300) / 0.49)));
        OzPlanesOut po;
        po.planes = st->wsOzA.as<signed char>() + (size_t)m * oz_pa; po.plane_stride = (long long)(ldk / 16) * oz_rows_alloc * 16;
        po.rows_alloc = oz_rows_alloc; po.n_chunks = ldk / 16; po.inv_scale = 1.0 / oz_scaleA[m]; po.write_fp64 = dX_dev ? 1 : 0;
        po.row_offset = 0; po.rows_cover = 0;
        bool fz = false;
        RC(launch_crosscov_ex(o.md, o.q_prepd, o.train_prepd, true, st->N, Kx, ldk, false, &po, &fz, s, &st->lc));
        oz_fused[m] = fz ? 1 : 0;
      } else {
        RC(launch_crosscov(o.md, o.q_prepd, o.train_prepd, true, st->N, Kx, ldk, false, s, &st->lc));
      }
      rec_end(st, s);
      }
      PostGemmArgs& a = pg[m];
      a.Kx = Kx; a.rows = rows; a.ldk = ldk; a.B = o.LinvExt.as<double>(); a.N = st->N; a.n_ext = nb + 1; a.Rpad = o.Rpad;
      a.q = q; a.Gqq = st->wsGqq.as<double>() + (size_t)m * rows_max * q; a.W = st->wsW.as<double>() + (size_t)m * rows_max * ldw;
      a.ldw = ldw; a.mu_raw = st->wsMuRaw.as<double>() + (size_t)m * rows_max;
    }
    if (par) RC(join2());
    int *oz_flags = nullptr, *oz_list = nullptr, *oz_count = nullptr;
    int oz_cap = 0;
    // FP64 chain over `n` q-batches of the guard's list (count_dev != NULL: fixed capacity, real count on the device):
    // gather the q-batches, prepare them, K(X*,X) in FP64, FP64 posterior GEMM, scatter Gram / W / mean rows back
    auto redo_flagged = [&](int n, const int* count_dev) -> int {
      const int rg = n * q;
      const size_t per_out = (size_t)rg * q + (size_t)rg * ldw + rg;
      RC(st->wsOzXg.ensure((size_t)rg * st->d * 8));
      RC(st->wsOzKxG.ensure((size_t)rg * ldk * 8 * M));
      RC(st->wsOzOutG.ensure(per_out * 8 * M));
      RC(launch_ozaki_gather_x(Xc, oz_list, count_dev, n, q * st->d, st->wsOzXg.as<double>(), s, &st->lc));
      std::vector<PostGemmArgs> gg(M);
      for (int m = 0; m < M; ++m) {
        OutputH& o = st->out[m];
        double* KxG = st->wsOzKxG.as<double>() + (size_t)m * rg * ldk;
        RC(o.g_prep.ensure(o.md, rg, &o.g_prepd));
        RC(launch_prep_points(o.md, st->wsOzXg.as<double>(), rg, st->d, o.g_prepd, s, &st->lc));
        RC(launch_crosscov(o.md, o.g_prepd, o.train_prepd, true, st->N, KxG, ldk, false, s, &st->lc));
        gg[m] = pg[m];
        gg[m].Kx = KxG; gg[m].rows = rg;
        gg[m].Gqq = st->wsOzOutG.as<double>() + (size_t)m * per_out;
        gg[m].W = gg[m].Gqq + (size_t)rg * q;
        gg[m].mu_raw = gg[m].W + (size_t)rg * ldw;
      }
      if (rg <= 64) {
        RC(st->wsV.ensure(posterior_small_ws_doubles(rg, st->out[0].Rpad, M) * 8));
        RC(launch_posterior_small(gg.data(), M, st->wsV.as<double>(), s, &st->lc));
      } else {
        size_t pw = posterior_gemm_partial_ws_doubles(rg, q, M);
        if (pw) RC(st->wsGramPart.ensure(pw * 8));
        RC(launch_posterior_gemm_multi(gg.data(), M, st->wsGramPart.as<double>(), s, &st->lc));
      }
      for (int m = 0; m < M; ++m)
        RC(launch_ozaki_scatter(oz_list, count_dev, n, q, nb, ldw, gg[m].Gqq, gg[m].W, gg[m].mu_raw, pg[m].Gqq, pg[m].W, pg[m].mu_raw,
                                s, &st->lc));
      return BO_OK;
    };
    // Refinement steps (<= 64 candidate points, gradient wanted): U = K*X (K + s2 I)^-1 of the adjoint depends on K(X*,X)
    // only, so all outputs' products are launched NOW as one multi-item skinny GEMM on a forked stream and run next to the
    // forward chain (posterior GEMM, conditional roots, MC adjoint); they were 0.10 ms of a 0.51 ms step.
    bool u_forked = false;
    if (small_rows && dX_dev && !st->timing) {
      for (int m = 0; m < M; ++m) RC(ensure_kinv(st, st->out[m], s));
      RC(st->wsU.ensure((size_t)M * rows_max * ldk * 8, false));
      if (!st->side_stream) {
        CUDA_CHECK_RET(cudaStreamCreateWithFlags(&st->side_stream, cudaStreamNonBlocking));
        CUDA_CHECK_RET(cudaEventCreateWithFlags(&st->side_fork, cudaEventDisableTiming));
        CUDA_CHECK_RET(cudaEventCreateWithFlags(&st->side_join, cudaEventDisableTiming));
      }
      CUDA_CHECK_RET(cudaEventRecord(st->side_fork, s));
      CUDA_CHECK_RET(cudaStreamWaitEvent(st->side_stream, st->side_fork, 0));
      std::vector<SkinnyItem> items(M);
      for (int m = 0; m < M; ++m)
        items[m] = SkinnyItem{pg[m].Kx, st->out[m].Kinv.as<double>(), st->wsU.as<double>() + (size_t)m * rows_max * ldk};
      RC(launch_skinny_gemm_nt(items.data(), M, rows, st->N, st->N, ldk, ldk, ldk, 0, st->side_stream, &st->lc));
      CUDA_CHECK_RET(cudaEventRecord(st->side_join, st->side_stream));
      u_forked = true;
    }
    if (small_rows) {
      RC(st->wsV.ensure(posterior_small_ws_doubles(rows, st->out[0].Rpad, M) * 8));
      rec_begin(st, "posterior_gemm", s);
      RC(launch_posterior_small(pg.data(), M, st->wsV.as<double>(), s, &st->lc));
      rec_end(st, s);
    } else if (use_ozaki) {
      // FP64-accurate GEMM on the INT8 tensor cores: digit planes of K(X*,X) (every call) and of LinvExt (once)
      const int rows_alloc = oz_rows_alloc;
      const size_t pa = oz_pa;
      RC(st->wsGramPart.ensure(ozaki_partial_ws_doubles(rows, q, M) * 8));
      RC(st->wsOzScratch.ensure(ozaki_scratch_bytes()));
      std::vector<OzakiArgs> oa(M);
      rec_begin(st, "ozaki_slice", s);
      for (int m = 0; m < M; ++m) {
        OutputH& o = st->out[m];
        if (!o.oz_ready) {
          RC(o.ozScaleB.ensure((size_t)o.Rpad * 8));
          RC(o.ozB.ensure(ozaki_plane_bytes(o.Rpad, ldk)));
          RC(launch_ozaki_row_scale(o.LinvExt.as<double>(), o.Rpad, st->N, ldk, o.ozScaleB.as<double>(), s, &st->lc));
          RC(launch_ozaki_slice(o.LinvExt.as<double>(), o.Rpad, st->N, ldk, o.ozScaleB.as<double>(), 0.0,
                                o.ozB.as<signed char>(), o.Rpad, ldk, s, &st->lc));
          RC(o.ozSb2.ensure(16));
          RC(launch_ozaki_scale_max(o.ozScaleB.as<double>(), st->N, o.ozSb2.as<double>(), s, &st->lc));
          o.oz_ready = true;
        }
        const double scaleA = oz_scaleA[m];
        signed char* Ap = st->wsOzA.as<signed char>() + (size_t)m * pa;
        if (!oz_fused[m]) RC(launch_ozaki_slice(pg[m].Kx, rows, st->N, ldk, nullptr, scaleA, Ap, rows_alloc, ldk, s, &st->lc));
        OzakiArgs& a = oa[m];
        a.Aplanes = Ap; a.rows = rows; a.rows_alloc = rows_alloc; a.ldk = ldk; a.Bplanes = o.ozB.as<signed char>();
        a.scaleB = o.ozScaleB.as<double>(); a.scaleA = scaleA; a.N = st->N; a.n_ext = nb + 1; a.Rpad = o.Rpad; a.q = q;
        a.Gqq = pg[m].Gqq; a.W = pg[m].W; a.ldw = ldw; a.mu_raw = pg[m].mu_raw;
        a.scratch = st->wsOzScratch.as<long long>();
      }
      rec_end(st, s);
      rec_begin(st, "posterior_gemm", s);
      RC(launch_ozaki_gemm(oa.data(), M, st->wsGramPart.as<double>(), st->ozaki_tile, s, &st->lc));
      rec_end(st, s);
      st->oz_flagged_last = 0; st->oz_batches_last = bc;
      if (oz_guard) {
        // ---- per-row guard: which q-batches can the digit planes not guarantee to oz_tol?  The host must not stall the
        // launch queue waiting for the answer, so the FP64 chain is launched NOW with a fixed capacity of oz_cap q-batches
        // (<= 64 rows: the skinny FP64 kernels) and takes the real count from device memory; the counter is read back after
        // the rest of the chunk has been queued, and only a call that flags more q-batches than that is redone.
        rec_begin(st, "ozaki_guard", s);
        RC(st->wsOzFlags.ensure(((size_t)2 * bc + 4) * sizeof(int)));
        oz_flags = st->wsOzFlags.as<int>();
        oz_list = oz_flags + bc;
        oz_count = oz_list + bc;
        CUDA_CHECK_RET(cudaMemsetAsync(oz_flags, 0, ((size_t)2 * bc + 4) * sizeof(int), s));
        for (int m = 0; m < M; ++m) {
          double kmax = 0.0;
          for (int t = 0; t < st->out[m].md.n_terms; ++t) kmax += st->out[m].md.coef[t];
          RC(launch_ozaki_guard(pg[m].Gqq, pg[m].mu_raw, rows, q, st->N, kmax, oz_scaleA[m], st->out[m].ozSb2.as<double>(),
                                st->oz_kappa, st->oz_tol, oz_flags, oz_list, oz_count, s, &st->lc));
        }
        oz_cap = std::min(bc, 64 / q);
        RC(redo_flagged(oz_cap, oz_count));
        if (!st->pin_count) {
          CUDA_CHECK_RET(cudaHostAlloc(reinterpret_cast<void**>(&st->pin_count), 64, cudaHostAllocDefault));
          CUDA_CHECK_RET(cudaEventCreateWithFlags(&st->oz_event, cudaEventDisableTiming));
        }
        CUDA_CHECK_RET(cudaMemcpyAsync(st->pin_count, oz_count, sizeof(int), cudaMemcpyDeviceToHost, s));
        CUDA_CHECK_RET(cudaEventRecord(st->oz_event, s));
        rec_end(st, s);
        if (st->oz_calib == 0) st->oz_calib = 1;
      }
    } else {
      size_t pw = posterior_gemm_partial_ws_doubles(rows, q, M);
      if (pw) RC(st->wsGramPart.ensure(pw * 8));
      rec_begin(st, "posterior_gemm", s);
      RC(launch_posterior_gemm_multi(pg.data(), M, st->wsGramPart.as<double>(), s, &st->lc));
      rec_end(st, s);
    }
    // everything after the posterior GEMM of the chunk: conditional roots, baseline part of the samples, MC acquisition
    // value (or the adjoint chain).  A lambda because the INT8 guard may have to run it a second time (see below).
    auto run_tail = [&]() -> int {
    if (par_tail) RC(fork2());
    for (int m = 0; m < M; ++m) {
      cudaStream_t s = (par_tail && (m & 1)) ? st->side2 : s_main;
      OutputH& o = st->out[m];
      CondRootArgs c;
      c.md = o.md; c.prep_q = o.q_prepd; c.prep_b = o.base_prepd; c.b = bc; c.q = q; c.nb = nb; c.M = M; c.m = m;
      c.Gqq = pg[m].Gqq; c.W = pg[m].W; c.ldw = ldw; c.mu_raw = pg[m].mu_raw;
      c.Lb = o.Lb.as<double>(); c.LbInv = o.LbInv.as<double>(); c.LbInvT = o.LbInvT.as<double>(); c.ldlb = st->ldlb; c.root = st->wsRoot.as<double>(); c.mu = st->wsMu.as<double>();
      c.info = st->wsJit.as<int>(); c.jitter = nullptr;
      c.BL = sample_gemm ? st->wsBL.as<double>() + (size_t)m * rows_max * ldbl : nullptr; c.ldbl = ldbl;
      rec_begin(st, "cond_root", s);
      RC(launch_cond_root(c, s, &st->lc));
      rec_end(st, s);
      if (sample_gemm) {
        // baseline part of every MC sample, F'[row, s] = sum_e bl[row, e] z_b[s, e], on the tensor pipe
        rec_begin(st, "sample_gemm", s);
        RC(launch_gemm_nt(rows, S, nb, 1.0, c.BL, ldbl, st->zbM.as<double>() + (size_t)m * S * ldbl, ldbl, 0.0,
                          st->wsFp.as<double>() + (size_t)m * rows_max * S, S, false, s, &st->lc));
        rec_end(st, s);
      }
    }
    if (par_tail) RC(join2());
    if (fb_active) {
      // joint re-sampling fallback: how many (q-batch, output) conditional roots exhausted the jitter ladder?  The counter
      // travels to the host while the MC kernels below still run, so reading it costs no GPU time
      CUDA_CHECK_RET(cudaMemsetAsync(st->wsFbCount.p, 0, sizeof(int), s));
      RC(launch_count_nonzero(st->wsJit.as<int>(), bc * M, st->wsFbCount.as<int>(), s, &st->lc));
      CUDA_CHECK_RET(cudaMemcpyAsync(st->pin_fb, st->wsFbCount.p, sizeof(int), cudaMemcpyDeviceToHost, s));
      CUDA_CHECK_RET(cudaEventRecord(st->fb_event, s));
    }
    McArgs ma;
    ma.b = bc; ma.q = q; ma.nb = nb; ma.M = M; ma.S = S; ma.od = st->od; ma.root = st->wsRoot.as<double>();
    ma.mu = st->wsMu.as<double>(); ma.zbT = st->zbT.as<double>(); ma.zqT = st->wsZqT.as<double>();
    ma.cell_lo = st->cell_lo.as<double>(); ma.cell_up = st->cell_up.as<double>(); ma.ncells = st->ncells.as<int>();
    ma.cells_shared = st->cells_shared; ma.best_f = st->best_f; ma.out = out_dev + b0;
    ma.variant = st->acqf_kind == 3 ? st->scalar_variant : st->log_hvi; ma.vparam = st->vparam;
    ma.best_f_s = st->noisy_scalar ? st->best_f_s.as<double>() : nullptr; ma.tau_relu = st->tau_relu; ma.tau_max = st->tau_max;
    ma.info_in = st->wsJit.as<int>(); ma.info_out = info_chunk;
    ma.Fp = sample_gemm ? st->wsFp.as<double>() : nullptr; ma.fp_stride = rows_max * (size_t)S;
    ma.partial = st->wsPartial.as<double>();
    if (dX_dev) {
      // ---- backward (grad.cu) ----
      const size_t dfs = rows_max * (size_t)S;
      RC(st->wsDF.ensure(dfs * 8 * M));
      RC(st->wsDRoot.ensure((size_t)bchunk * M * q * nr * 8));
      RC(st->wsDMu.ensure(rows_max * M * 8));
      const size_t eg_n = (size_t)bchunk * q * q, ew_n = rows_max * ldw, emu_n = rows_max, dx_n = rows_max * st->d;
      const int n_buf = par ? M : 1;           // forked outputs need their own adjoint scratch
      RC(st->wsEG.ensure(eg_n * 8 * n_buf));
      RC(st->wsEW.ensure(ew_n * 8 * n_buf));
      RC(st->wsEmu.ensure(emu_n * 8 * n_buf));
      if (par) RC(st->wsDXm.ensure(dx_n * 8 * M));
      if (!u_forked) RC(st->wsU.ensure(rows_max * ldk * 8, false));
      rec_begin(st, "mc_grad", s);
      if (st->acqf_kind == 3) RC(launch_mc_scalar_grad(ma, st->wsDF.as<double>(), dfs, s, &st->lc));
      else if (st->log_hvi) {
        // the partial-sum workspace doubles as the [b][S] per-sample values of the sample-split log-space adjoint
        RC(st->wsPartial.ensure((size_t)bchunk * S * 8));
        ma.partial = st->wsPartial.as<double>();
        RC(launch_mc_loghvi_grad(ma, st->wsDF.as<double>(), dfs, st->wsPartial.as<double>(), s, &st->lc));
      }
      else RC(launch_mc_hvi_grad(ma, st->max_cells, st->wsDF.as<double>(), dfs, s, &st->lc));
      rec_end(st, s);
      rec_begin(st, "grad_reduce", s);
      RC(launch_grad_reduce(st->wsDF.as<double>(), dfs, st->zbT.as<double>(), st->wsZqT.as<double>(), S, nb, q, M, rows,
                            st->wsDRoot.as<double>(), st->wsDMu.as<double>(), s, &st->lc));
      rec_end(st, s);
      if (u_forked) CUDA_CHECK_RET(cudaStreamWaitEvent(s, st->side_join, 0));     // U of every output (forked above)
      if (par) RC(fork2());
      for (int m = 0; m < M; ++m) {
        cudaStream_t s = (par && (m & 1)) ? st->side2 : s_main;
        OutputH& o = st->out[m];
        RC(ensure_kinv(st, o, s));
        const size_t mb = par ? (size_t)m : 0;
        double* EGm = st->wsEG.as<double>() + mb * eg_n;
        double* EWm = st->wsEW.as<double>() + mb * ew_n;
        double* Emum = st->wsEmu.as<double>() + mb * emu_n;
        CondRootBwdArgs cb;
        cb.b = bc; cb.q = q; cb.nb = nb; cb.M = M; cb.m = m; cb.root = st->wsRoot.as<double>(); cb.droot = st->wsDRoot.as<double>();
        cb.dmu = st->wsDMu.as<double>(); cb.LbInv = o.LbInv.as<double>(); cb.ldlb = st->ldlb; cb.y_std = o.md.y_std;
        cb.EG = EGm; cb.EW = EWm; cb.ldw = ldw; cb.Emu = Emum;
        rec_begin(st, "cond_root_bwd", s);
        RC(launch_cond_root_bwd(cb, s, &st->lc));
        rec_end(st, s);
        rec_begin(st, "u_gemm", s);
        double* Um = st->wsU.as<double>() + (u_forked ? (size_t)m * rows_max * ldk : 0);
        if (u_forked) {
          // launched on the side stream right after K(X*,X); joined before this loop
        } else if (small_rows) {
          SkinnyItem it{pg[m].Kx, o.Kinv.as<double>(), st->wsU.as<double>()};
          RC(launch_skinny_gemm_nt(&it, 1, rows, st->N, st->N, ldk, ldk, ldk, 0, s, &st->lc));
        } else {
          RC(launch_gemm_nt(rows, st->N, st->N, 1.0, pg[m].Kx, ldk, o.Kinv.as<double>(), ldk, 0.0, st->wsU.as<double>(), ldk, false, s, &st->lc));
        }
        rec_end(st, s);
        KernelGradArgs kg;
        kg.md = o.md; kg.prep_q = o.q_prepd; kg.prep_b = o.base_prepd; kg.rows = rows; kg.q = q; kg.nb = nb; kg.N = st->N;
        kg.ldk = ldk; kg.d = st->d; kg.alpha = o.alpha_row.as<double>();
        kg.Aext = o.LinvExt.as<double>() + (size_t)(st->N + 1) * ldk; kg.U = Um;
        kg.EG = EGm; kg.EW = EWm; kg.ldw = ldw; kg.Emu = Emum;
        double* dX_out = dX_dev + (size_t)b0 * q * st->d;
        // forked outputs write their own gradient block; the blocks are added in output order after the join (fixed order:
        // bit-identical to the sequential accumulation)
        kg.dX = (par && m > 0) ? st->wsDXm.as<double>() + (size_t)m * dx_n : dX_out;
        kg.accumulate = (!par && m > 0) ? 1 : 0; kg.Kx = pg[m].Kx;
        rec_begin(st, "kernel_grad", s);
        RC(launch_kernel_grad(kg, s, &st->lc));
        rec_end(st, s);
      }
      if (par) {
        RC(join2());
        for (int m = 1; m < M; ++m)
          RC(launch_add_inplace(dX_dev + (size_t)b0 * q * st->d, st->wsDXm.as<double>() + (size_t)m * dx_n, (size_t)rows * st->d, s, &st->lc));
      }
      return BO_OK;
    }
    rec_begin(st, "mc_acqf", s);
    if (st->acqf_kind == 3) RC(launch_mc_scalar(ma, s, &st->lc));
    else if (st->log_hvi) RC(launch_mc_loghvi(ma, s, &st->lc));
    else {
      size_t ow = mc_hvi_obj_ws_bytes(ma, st->max_cells);
      if (ow) RC(st->wsObjW.ensure(ow));
      RC(launch_mc_hvi(ma, st->max_cells, ow ? st->wsObjW.as<double>() : nullptr, s, &st->lc));
    }
    rec_end(st, s);
    return BO_OK;
    };
    RC(run_tail());
    if (oz_count) {
      // the guard's counter: by now the rest of the chunk is queued behind the GEMM, so waiting for it costs no GPU time
      CUDA_CHECK_RET(cudaEventSynchronize(st->oz_event));
      const int n_flag = *st->pin_count;
      st->oz_flagged_last = n_flag;
      st->oz_flagged_total += n_flag; st->oz_batches_total += bc;
      if (n_flag > oz_cap) {
        rec_begin(st, "ozaki_redo", s);
        if ((long long)n_flag * 4 > bc) {
          // most of the call is out of the INT8 path's reach (candidates next to the training data): redo everything with
          // the FP64 kernel and stay there until the next prepare
          st->oz_calib = -1;
          for (int m = 0; m < M; ++m)
            RC(launch_crosscov(st->out[m].md, st->out[m].q_prepd, st->out[m].train_prepd, true, st->N, const_cast<double*>(pg[m].Kx),
                               ldk, false, s, &st->lc));
          size_t pw = posterior_gemm_partial_ws_doubles(rows, q, M);
          if (pw) RC(st->wsGramPart.ensure(pw * 8));
          RC(launch_posterior_gemm_multi(pg.data(), M, st->wsGramPart.as<double>(), s, &st->lc));
        } else {
          RC(redo_flagged(n_flag, nullptr));
        }
        rec_end(st, s);
        RC(run_tail());
      }
    }
    if (fb_active) {
      CUDA_CHECK_RET(cudaEventSynchronize(st->fb_event));
      if (*st->pin_fb > 0) {
        int n_re = 0;
        RC(bo_acqf_resample_flagged(st, Xc, bc, q, zq_dev, out_dev + b0, info_chunk, &n_re, stream));
        st->fb_resampled_last += n_re;
        // the fallback used the per-call workspaces: the base samples of the next chunk are transposed again
        RC(launch_transpose_base_samples(zq_dev, S, q, M, st->wsZqT.as<double>(), nullptr, 0, s, &st->lc));
      }
    }
  }
  return BO_OK;
}

extern "C" int bo_mll_forward_backward(bo_state* st, int32_t m, double* mll_out, double* d_noise, double* d_mean_const,
                                       double* d_lengthscale, int32_t n_lengthscale, double* d_coef, int32_t n_coef,
                                       void* stream) {
  if (!st || !st->factorized) { bo_set_error("state not factorized"); return BO_ERR_STATE; }
  if (m < 0 || m >= st->M || !mll_out) { bo_set_error("mll: bad output index / null result"); return BO_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  OutputH& o = st->out[m];
  const int N = st->N, ldk = st->ldk;
  // lengthscale slots: ARD dims of the continuous leaves, groups of the Hamming leaves, in leaf order
  int ls_offset[BO_MAX_LEAVES], n_ls = 0;
  for (int l = 0; l < o.md.n_leaves; ++l) {
    if (o.md.leaf[l].kind <= BO_LEAF_HAMMING) { ls_offset[l] = n_ls; n_ls += o.md.leaf[l].nd; }
    else ls_offset[l] = -1;
  }
  if ((d_lengthscale && n_lengthscale != n_ls) || (d_coef && n_coef != o.md.n_terms)) {
    bo_set_error("mll: expected %d lengthscale slots and %d term coefficients", n_ls, o.md.n_terms);
    return BO_ERR_INVALID;
  }
  RC(ensure_kinv(st, o, s));
  const int n_params = n_ls + o.md.n_terms;
  RC(st->wsTmp.ensure(((size_t)N * n_params + n_params + 8) * 8));
  double* part = st->wsTmp.as<double>();
  double* outp = part + (size_t)N * n_params;
  double* out5 = outp + n_params;
  RC(launch_mll_scalars(o.resid.as<double>(), o.alpha_row.as<double>(), o.L.as<double>(), o.Kinv.as<double>(), N, ldk, out5, s, &st->lc));
  const bool want_grad = d_lengthscale || d_coef;
  if (want_grad) RC(launch_mll_grad(o.md, o.train_prepd, N, ldk, o.alpha_row.as<double>(), o.Kinv.as<double>(), ls_offset, n_ls, part, outp, s, &st->lc));
  std::vector<double> host(n_params + 5);
  CUDA_CHECK_RET(cudaMemcpyAsync(host.data(), outp, (size_t)(n_params + 5) * 8, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK_RET(cudaStreamSynchronize(s));
  const double* h5 = host.data() + n_params;
  *mll_out = -0.5 * h5[0] - h5[1] - 0.5 * (double)N * 1.8378770664093453;  // log(2 pi)
  if (d_mean_const) *d_mean_const = h5[2];
  if (d_noise) *d_noise = 0.5 * (h5[3] - h5[4]);
  if (want_grad) {
    if (d_lengthscale) for (int k = 0; k < n_ls; ++k) d_lengthscale[k] = host[k];
    if (d_coef) for (int t = 0; t < o.md.n_terms; ++t) d_coef[t] = host[n_ls + t];
  }
  return BO_OK;
}

extern "C" int bo_acqf_forward(bo_state* st, const double* X_dev, int32_t b, int32_t q, const double* zq_dev,
                               double* out_dev, int32_t* info_dev, void* stream) {
  return acqf_run(st, X_dev, b, q, zq_dev, out_dev, nullptr, info_dev, stream);
}

extern "C" int bo_acqf_forward_backward(bo_state* st, const double* X_dev, int32_t b, int32_t q, const double* zq_dev,
                                        double* out_dev, double* dX_dev, int32_t* info_dev, void* stream) {
  if (!dX_dev) { bo_set_error("forward_backward: dX_dev is NULL"); return BO_ERR_INVALID; }
  return acqf_run(st, X_dev, b, q, zq_dev, out_dev, dX_dev, info_dev, stream);
}

// [UPSTREAM] sample_cached_cholesky's fallback: when the q x q conditional root of a q-batch cannot be
// factorised even with jitter 1e-3 (info = 1 from bo_acqf_forward, value NaN), BoTorch catches the NotPSDError / NanError, warns
// and samples the JOINT posterior over (X_baseline, X) instead -- psd_safe_cholesky of the full (n_b + q) x (n_b + q)
// covariance with its own jitter ladder on the whole diagonal.  This entry point does that for the flagged q-batches of the
// last forward call: it reads info_dev back (one synchronisation; the flagged set is expected to be empty), and for every
// flagged q-batch factorises the joint covariance per output, takes the last q rows of the root as [bl | br], and re-scores
// the q-batch with the same base samples and the cached cells.  info stays 1 for a re-scored q-batch (the BoTorch warning),
// becomes 2 when the joint factorisation fails as well (BoTorch: NotPSDError propagates; the value stays NaN).
// n_resampled (HOST, may be NULL) receives the number of re-scored q-batches.
extern "C" int bo_acqf_resample_flagged(bo_state* st, const double* X_dev, int32_t b, int32_t q, const double* zq_dev,
                                        double* out_dev, int32_t* info_dev, int32_t* n_resampled, void* stream) {
  if (!st || !st->factorized || st->acqf_kind == 0) { bo_set_error("resample_flagged before prepare"); return BO_ERR_STATE; }
  if (!X_dev || !out_dev || !info_dev || !zq_dev) { bo_set_error("resample_flagged: null argument"); return BO_ERR_INVALID; }
  if (n_resampled) *n_resampled = 0;
  if (b < 1) return BO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int nb = st->nb, M = st->M, d = st->d, S = st->S;
  if (nb == 0 && !st->force_fallback) return BO_OK;   // no cached root: the ladder already WAS the joint factorisation
  std::vector<int> info((size_t)b);
  CUDA_CHECK_RET(cudaMemcpyAsync(info.data(), info_dev, (size_t)b * sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK_RET(cudaStreamSynchronize(s));
  std::vector<int> flagged;
  for (int i = 0; i < b; ++i) if (info[i] != 0 || st->force_fallback) flagged.push_back(i);
  if (flagged.empty()) return BO_OK;
  const int n = nb + q, ldn = round_up(n, 16), nr = nb + q;
  RC(st->wsXfull.ensure((size_t)n * d * 8));
  RC(st->wsMeanJ.ensure((size_t)n * M * 8));
  RC(st->wsV.ensure((size_t)n * st->ldk * 8, true));
  RC(st->wsZqT.ensure((size_t)q * M * S * 8));
  RC(launch_transpose_base_samples(zq_dev, S, q, M, st->wsZqT.as<double>(), nullptr, 0, s, &st->lc));
  RC(st->wsRoot.ensure((size_t)M * q * nr * 8));
  RC(st->wsMu.ensure((size_t)q * M * 8));
  RC(st->wsJit.ensure((size_t)M * sizeof(int)));
  RC(st->wsPartial.ensure((size_t)mc_hvi_partial_groups(S) * 8));
  if (nb > 0) CUDA_CHECK_RET(cudaMemcpyAsync(st->wsXfull.p, st->Xb_raw.p, (size_t)nb * d * 8, cudaMemcpyDeviceToDevice, s));
  int done = 0;
  for (int i : flagged) {
    CUDA_CHECK_RET(cudaMemcpyAsync(st->wsXfull.as<double>() + (size_t)nb * d, X_dev + (size_t)i * q * d, (size_t)q * d * 8,
                                   cudaMemcpyDeviceToDevice, s));
    bool ok = true;
    for (int m = 0; m < M && ok; ++m) {
      PrepD pd;
      int inf = 0; double jit = 0.0;
      RC(joint_root(st, m, st->wsXfull.as<double>(), n, ldn, st->jointPrep, &pd, st->wsMeanJ.as<double>(), st->wsJointRoot, st->wsJointCov,
                    st->wsV.as<double>(), &inf, &jit, s));
      if (inf != 0) { ok = false; break; }
      RC(launch_joint_rows_to_root(st->wsJointRoot.as<double>(), ldn, st->wsMeanJ.as<double>(), nb, q, M, m, 0,
                                   st->wsRoot.as<double>(), st->wsMu.as<double>(), s, &st->lc));
    }
    int new_info = 1;
    if (!ok) new_info = 2;
    else {
      CUDA_CHECK_RET(cudaMemsetAsync(st->wsJit.p, 0, (size_t)M * sizeof(int), s));
      McArgs ma;
      ma.b = 1; ma.q = q; ma.nb = nb; ma.M = M; ma.S = S; ma.od = st->od; ma.root = st->wsRoot.as<double>();
      ma.mu = st->wsMu.as<double>(); ma.zbT = st->zbT.as<double>(); ma.zqT = st->wsZqT.as<double>();
      ma.cell_lo = st->cell_lo.as<double>(); ma.cell_up = st->cell_up.as<double>(); ma.ncells = st->ncells.as<int>();
      ma.cells_shared = st->cells_shared; ma.best_f = st->best_f; ma.out = out_dev + i;
      ma.variant = st->acqf_kind == 3 ? st->scalar_variant : st->log_hvi; ma.vparam = st->vparam;
      ma.best_f_s = st->noisy_scalar ? st->best_f_s.as<double>() : nullptr; ma.tau_relu = st->tau_relu; ma.tau_max = st->tau_max;
      ma.info_in = st->wsJit.as<int>(); ma.info_out = nullptr;
      ma.Fp = nullptr; ma.fp_stride = 0;
      ma.partial = st->wsPartial.as<double>();
      if (st->acqf_kind == 3) RC(launch_mc_scalar(ma, s, &st->lc));
      else if (st->log_hvi) RC(launch_mc_loghvi(ma, s, &st->lc));
      else {
        size_t ow = mc_hvi_obj_ws_bytes(ma, st->max_cells);
        if (ow) RC(st->wsObjW.ensure(ow));
        RC(launch_mc_hvi(ma, st->max_cells, ow ? st->wsObjW.as<double>() : nullptr, s, &st->lc));
      }
      ++done;
    }
    CUDA_CHECK_RET(cudaMemcpyAsync(info_dev + i, &new_info, sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_CHECK_RET(cudaStreamSynchronize(s));    // new_info lives on this stack frame
  }
  if (n_resampled) *n_resampled = done;
  return BO_OK;
}

extern "C" int32_t bo_acqf_last_resampled(bo_state* st) { return st ? st->fb_resampled_last : 0; }

extern "C" int bo_acqf_optimize(bo_state* st, double* X_dev, int32_t r, int32_t q_tot, int32_t q_free, const double* lb,
                                const double* ub, const double* zq_dev, int32_t maxiter, int32_t history, double pgtol,
                                double ftol, double* out_dev, int32_t* stats, void* stream) {
  if (!st || !X_dev || !lb || !ub || !out_dev) { bo_set_error("optimize: null argument"); return BO_ERR_INVALID; }
  if (r < 1 || q_free < 1 || q_free > q_tot || q_tot > BO_MAX_Q) { bo_set_error("optimize: bad r=%d / q_free=%d / q_tot=%d", r, q_free, q_tot); return BO_ERR_INVALID; }
  if (maxiter < 1 || history < 1 || history > LB_MAX_HIST) { bo_set_error("optimize: maxiter >= 1, 1 <= history <= %d", LB_MAX_HIST); return BO_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  const int d = st->d, n = q_free * d;
  for (int j = 0; j < d; ++j)
    if (!(lb[j] <= ub[j])) { bo_set_error("optimize: lb[%d] > ub[%d]", j, j); return BO_ERR_INVALID; }
  SetupTrace tr("acqf_optimize", s);
  RC(st->wsLbfgs.ensure(lbfgs_ws_bytes(r, n, history)));
  RC(st->wsLbBounds.ensure((size_t)2 * d * 8));
  RC(st->wsLbGrad.ensure((size_t)r * q_tot * d * 8));
  if (!st->pin_lb) CUDA_CHECK_RET(cudaHostAlloc(reinterpret_cast<void**>(&st->pin_lb), 64 * sizeof(int), cudaHostAllocDefault));
  while (st->lb_events.size() < 2) {
    cudaEvent_t e;
    CUDA_CHECK_RET(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    st->lb_events.push_back(e);
  }
  CUDA_CHECK_RET(cudaMemcpyAsync(st->wsLbBounds.p, lb, (size_t)d * 8, cudaMemcpyHostToDevice, s));
  CUDA_CHECK_RET(cudaMemcpyAsync(st->wsLbBounds.as<double>() + d, ub, (size_t)d * 8, cudaMemcpyHostToDevice, s));
  LbArgs a;
  a.q_tot = q_tot; a.q_free = q_free; a.d = d; a.hist = history; a.maxiter = maxiter; a.pgtol = pgtol; a.ftol = ftol;
  a.lb = st->wsLbBounds.as<double>(); a.ub = a.lb + d;
  a.X = X_dev; a.dX = st->wsLbGrad.as<double>(); a.vals = out_dev;
  lbfgs_carve(a, st->wsLbfgs.p, r, n, history);
  CUDA_CHECK_RET(cudaMemcpyAsync(a.n_running, &r, sizeof(int), cudaMemcpyHostToDevice, s));   // pageable 4-byte copy: staged at once
  tr.mark("workspaces");
  // every evaluation: the forward + adjoint chain of the r q-batches, then one state-machine step of every restart
  RC(acqf_run(st, X_dev, r, q_tot, zq_dev, out_dev, st->wsLbGrad.as<double>(), nullptr, stream));
  RC(launch_lbfgs_step(a, r, true, s, &st->lc));
  tr.mark("first evaluation");
  // L-BFGS-B allows ~1.25 evaluations per iteration on average (scipy: maxfun = 15000 for maxiter = 15000); the budget here
  // is 2 per iteration plus the 20 trials a failing line search may take
  const long long max_evals = 2ll * maxiter + 20;
  const int check_every = 8;
  long long n_eval = 1;
  int slot = 0, pending[2] = {0, 0};
  bool done = false;
  // One evaluation + state-machine step = ~20 small launches (0.66 ms of GPU time on config 3; the host enqueues them in
  // about half of that).  EVEREST_LBFGS_GRAPH=1 captures the step ONCE into a CUDA graph and replays it, which shields the
  // loop from a busy host (every workspace was sized by the first evaluation above; nothing inside allocates, synchronises
  // or copies from pageable memory any more).  Measured on an idle host: 0.160 s per 242 steps either way, but the first
  // instantiation in a process costs ~1 s -- more than an ask() -- so plain launches are the default.
  static const bool want_graph = []() { const char* e = getenv("EVEREST_LBFGS_GRAPH"); return e && atoi(e) != 0; }();
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  if (want_graph && s != nullptr) {          // the legacy default stream cannot be captured
    const long long lc0 = st->lc.n;
    if (cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      int rc1 = acqf_run(st, X_dev, r, q_tot, zq_dev, out_dev, st->wsLbGrad.as<double>(), nullptr, stream);
      if (rc1 == BO_OK) rc1 = launch_lbfgs_step(a, r, false, s, nullptr);
      cudaError_t ce = cudaStreamEndCapture(s, &graph);
      if (rc1 != BO_OK || ce != cudaSuccess || !graph || cudaGraphInstantiate(&gexec, graph, 0) != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        graph = nullptr; gexec = nullptr;
        cudaGetLastError();                  // clear the sticky-free error of the failed capture
      }
    } else cudaGetLastError();
    st->lb_launches_per_step = (int)(st->lc.n - lc0) + 1;
    st->lc.n = lc0;                          // captured launches are counted when the graph is replayed
  }
  while (!done && n_eval < max_evals) {
    for (int k = 0; k < check_every && n_eval < max_evals; ++k, ++n_eval) {
      if (gexec) {
        CUDA_CHECK_RET(cudaGraphLaunch(gexec, s));
        st->lc.n += st->lb_launches_per_step;
      } else {
        RC(acqf_run(st, X_dev, r, q_tot, zq_dev, out_dev, st->wsLbGrad.as<double>(), nullptr, stream));
        RC(launch_lbfgs_step(a, r, false, s, &st->lc));
      }
    }
    // progress counter of THIS batch goes to the host asynchronously; the one of the PREVIOUS batch is inspected now, so the
    // launch queue never runs dry while the host waits
    CUDA_CHECK_RET(cudaMemcpyAsync(st->pin_lb + slot, a.n_running, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK_RET(cudaEventRecord(st->lb_events[slot], s));
    pending[slot] = 1;
    const int prev = slot ^ 1;
    if (pending[prev]) {
      CUDA_CHECK_RET(cudaEventSynchronize(st->lb_events[prev]));
      pending[prev] = 0;
      if (st->pin_lb[prev] <= 0) done = true;
    }
    slot ^= 1;
  }
  tr.mark("loop");
  st->lb_graph_used = gexec != nullptr;
  if (gexec) { cudaGraphExecDestroy(gexec); cudaGraphDestroy(graph); }
  RC(launch_lbfgs_finish(a, r, s, &st->lc));
  // values at the returned points (a restart that stopped on a rejected trial returns its last accepted point)
  RC(acqf_run(st, X_dev, r, q_tot, zq_dev, out_dev, nullptr, nullptr, stream));
  if (stats) {
    std::vector<LbScalars> sc(r);
    CUDA_CHECK_RET(cudaMemcpyAsync(sc.data(), a.sc, (size_t)r * sizeof(LbScalars), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK_RET(cudaStreamSynchronize(s));
    stats[0] = (int)n_eval + 1; stats[1] = 0; stats[2] = 0; stats[3] = 0;
    for (int i = 0; i < r; ++i) {
      stats[1] = std::max(stats[1], sc[i].n_iter);
      if (sc[i].status == 1 || sc[i].status == 2) stats[2]++;
      if (sc[i].status == 4 || sc[i].status == 6) stats[3]++;
    }
  }
  return BO_OK;
}

// ---- HOST-buffer entry points -----------------------------------------------------------------------------------
// Columns read ONLY by Tanimoto leaves hold 0/1 fingerprints (create() checks the training rows): they cross PCIe as bits.
// A candidate row of BASELINE config 5 is 2060 float64 columns = 16.5 KB; packed it is 12 doubles + 32 words = 352 bytes.
enum HostMode { HOST_DENSE = 0, HOST_PACK = 1, HOST_PREPACKED = 2 };

static void pack_layout_init(bo_state* st) {
  if (st->pack_ready) return;
  const int d = st->d;
  std::vector<char> tani(d, 0), other(d, 0);
  for (int m = 0; m < st->M; ++m) {
    const OutputH& o = st->out[m];
    for (int l = 0; l < o.md.n_leaves; ++l) {
      const std::vector<int>& cols = o.leaf_cols[l];
      const int kind = o.md.leaf[l].kind;
      for (size_t k = 0; k < cols.size(); ++k) {
        if (kind == BO_LEAF_TANIMOTO) tani[cols[k]] = 1;
        else if (kind == BO_LEAF_HAMMING) { for (int j = 0; j < o.leaf_card[l][k]; ++j) other[cols[k] + j] = 1; }
        else other[cols[k]] = 1;
      }
    }
  }
  st->pack_bit_cols.clear(); st->pack_dense_cols.clear();
  for (int c = 0; c < d; ++c) {
    if (tani[c] && !other[c]) st->pack_bit_cols.push_back(c);
    else st->pack_dense_cols.push_back(c);
  }
  st->pack_ready = true;
}

extern "C" int bo_pack_layout(bo_state* st, int32_t* n_dense, int32_t* n_bits, int32_t* dense_cols, int32_t* bit_cols) {
  if (!st || !n_dense || !n_bits) { bo_set_error("pack_layout: null argument"); return BO_ERR_INVALID; }
  pack_layout_init(st);
  *n_dense = (int32_t)st->pack_dense_cols.size();
  *n_bits = (int32_t)st->pack_bit_cols.size();
  if (dense_cols) std::copy(st->pack_dense_cols.begin(), st->pack_dense_cols.end(), dense_cols);
  if (bit_cols) std::copy(st->pack_bit_cols.begin(), st->pack_bit_cols.end(), bit_cols);
  return BO_OK;
}

static bool pack_rows(const bo_state* st, const double* X, size_t r0, size_t r1, double* dense, unsigned long long* bits) {
  return everest_pack_rows(X, r0, r1, st->d, st->pack_dense_cols.data(), (int)st->pack_dense_cols.size(), st->pack_bit_cols.data(),
                           (int)st->pack_bit_cols.size(), dense, bits);
}

extern "C" int bo_pack_rows_host(bo_state* st, const double* X_host, int64_t rows, double* dense_out, uint64_t* bits_out) {
  if (!st || !X_host || !dense_out || !bits_out) { bo_set_error("pack_rows_host: null argument"); return BO_ERR_INVALID; }
  pack_layout_init(st);
  const int n_thr = (int)std::min<int64_t>(8, std::max<int64_t>(1, rows / 256));
  std::vector<std::thread> th;
  std::vector<char> okv(n_thr, 1);
  const size_t per = ((size_t)rows + n_thr - 1) / n_thr;
  for (int t = 0; t < n_thr; ++t) {
    const size_t r0 = std::min((size_t)rows, t * per), r1 = std::min((size_t)rows, r0 + per);
    th.emplace_back([=, &okv]() { okv[t] = pack_rows(st, X_host, r0, r1, dense_out, reinterpret_cast<unsigned long long*>(bits_out)) ? 1 : 0; });
  }
  for (auto& t : th) t.join();
  for (char c : okv) if (!c) { bo_set_error("fingerprint (Tanimoto) columns must hold 0 / 1"); return BO_ERR_INVALID; }
  return BO_OK;
}

static int forward_host_impl(bo_state* st, HostMode mode, const double* X_host, const double* dense_host, const uint64_t* bits_host,
                             int32_t b, int32_t q, const double* zq_dev, double* out_host, void* stream) {
  if (!st) { bo_set_error("null state"); return BO_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  if (b == 0) return BO_OK;
  const bool packed = mode != HOST_DENSE;
  if (packed) pack_layout_init(st);
  const int nd = packed ? (int)st->pack_dense_cols.size() : 0, nbits = packed ? (int)st->pack_bit_cols.size() : 0;
  const int W = (nbits + 63) / 64;
  const size_t n_rows = (size_t)b * q;
  const size_t in_bytes = n_rows * st->d * 8, out_bytes = (size_t)b * 8;
  const size_t row_wire = packed ? ((size_t)nd + W) * 8 : (size_t)st->d * 8;     // bytes per candidate point on the wire
  const size_t wire_bytes = n_rows * row_wire;
  if (st->pin_in_bytes < wire_bytes) {
    if (st->pin_in) cudaFreeHost(st->pin_in);
    st->pin_in = nullptr; st->pin_in_bytes = 0;
    CUDA_CHECK_RET(cudaHostAlloc(&st->pin_in, wire_bytes, cudaHostAllocDefault));
    st->pin_in_bytes = wire_bytes;
  }
  if (st->pin_out_bytes < out_bytes) {
    if (st->pin_out) cudaFreeHost(st->pin_out);
    CUDA_CHECK_RET(cudaHostAlloc(&st->pin_out, out_bytes, cudaHostAllocDefault));
    st->pin_out_bytes = out_bytes;
  }
  RC(st->stage_in.ensure(in_bytes));
  RC(st->stage_out.ensure(out_bytes));
  if (packed) {
    RC(st->stage_pk.ensure(wire_bytes));
    if (!st->pack_src_ready) {
      std::vector<int> src(st->d, 0);
      for (int k = 0; k < nd; ++k) src[st->pack_dense_cols[k]] = k;
      for (int k = 0; k < nbits; ++k) src[st->pack_bit_cols[k]] = -(k + 1);
      RC(st->pack_src.ensure((size_t)st->d * sizeof(int)));
      CUDA_CHECK_RET(cudaMemcpy(st->pack_src.p, src.data(), (size_t)st->d * sizeof(int), cudaMemcpyHostToDevice));
      st->pack_src_ready = true;
    }
  }
  // Pipeline over chunks of q-batches: the host stages chunk c+1 (pageable -> pinned, up to 4 threads) and the copy
  // engine moves it (second stream) while the kernels of chunk c run; only the first chunk's staging + H2D is exposed.
  if (!st->copy_stream) CUDA_CHECK_RET(cudaStreamCreateWithFlags(&st->copy_stream, cudaStreamNonBlocking));
  // Chunk plan from the ratio r of the copy time (bytes per candidate point at ~25 GB/s) to the compute time
  // (~ M N^2 flops per point at ~30 TFLOP/s):  r small (config 3: 0.04) -> a small first chunk so that the kernels start
  // early, then everything else in one efficient launch sequence;  r large (2048-bit fingerprints as float64: 0.8; the host
  // pass over the float64 rows costs about as much when they are packed here) -> equal chunks so that every copy but the
  // first hides behind a compute.
  std::vector<int> chunk_b;
  {
    const double host_bytes = (mode == HOST_PREPACKED) ? (double)row_wire : (double)st->d * 8.0;
    const double r = (host_bytes / 25e9) / ((double)st->M * (double)st->N * (double)st->N / 30e12 + 1e-12);
    const int min_b = std::max(1, 2048 / q);
    // inputs below 8 MiB cross PCIe in ~0.3 ms: nothing worth hiding, and whole-batch launches fill the GPU better
    if (b <= 2 * min_b || (double)n_rows * host_bytes < (double)((size_t)8 << 20)) chunk_b.push_back(b);
    else if (mode == HOST_PACK) {
      // the packer reads the float64 rows at ~70 GB/s with 8 threads (270 MB of config 5 in 4 ms, measured), a third of the
      // compute time, and small chunks cost GPU efficiency (2048 q-batches: 2.0 ms instead of 1.3): two chunks, the first a
      // quarter, so that the rest is packed by the time the first has been scored (8 equal chunks: 16.7 ms per config-5
      // screen; this plan: tools/probe_host_mixed.py)
      const int first = std::max(min_b, b / 4);
      chunk_b.push_back(first);
      chunk_b.push_back(b - first);
    } else if (r < 0.15) {
      // EVEREST_HOST_FIRST_DIV=n: first chunk = b / n (default 8); EVEREST_HOST_THREE=1: a third, intermediate
      // chunk of three times the first (experiment switches, read once).  Measured on config 3 (tools/exp_host.sh):
      // b/8 9.46 ms per screen, b/4 9.62, b/16 9.60, b/16 + 3b/16 + rest 9.78, b/32 three-way 10.27: b/8 stays.
      static const int first_div = []() { const char* e = getenv("EVEREST_HOST_FIRST_DIV"); int v = e ? atoi(e) : 8; return v >= 2 ? v : 8; }();
      static const int three = []() { const char* e = getenv("EVEREST_HOST_THREE"); return e ? atoi(e) : 0; }();
      const int first = std::max(min_b, b / first_div);
      chunk_b.push_back(first);
      if (three && b - first > 4 * first) chunk_b.push_back(3 * first);
      chunk_b.push_back(b - first - (chunk_b.size() > 1 ? 3 * first : 0));
    } else {
      const int n = (int)std::min<long long>(8, std::max<long long>(1, b / min_b));
      for (int i = 0, left = b; i < n; ++i) {
        const int take = (i == n - 1) ? left : (b + n - 1) / n;
        chunk_b.push_back(std::min(take, left));
        left -= chunk_b.back();
      }
    }
  }
  // Piece mode (float64 rows, every output a single continuous leaf, input above 8 MiB: configs 3 and 4): the kernels run ONCE
  // over the whole batch -- no small, inefficient first chunk -- and only the point preparation and K(X*,X) of the first
  // output are launched piece by piece as the copies land (acqf_run, `piecewise`).  Measured on config 3: the two-chunk plan
  // cost 1.65 + 7.88 ms against 8.73 ms for one launch sequence.  EVEREST_HOST_PIECES=0 restores the chunk plan.
  static const bool pieces_env = []() { const char* e = getenv("EVEREST_HOST_PIECES"); return !e || atoi(e) != 0; }();
  bool piece_mode = pieces_env && mode == HOST_DENSE && chunk_b.size() > 1 && st->ozaki == 1 && st->oz_calib >= 0;
  for (int m = 0; m < st->M && piece_mode; ++m) {
    const ModelD& md = st->out[m].md;
    piece_mode = md.n_terms == 1 && md.nfac[0] == 1 && md.leaf[md.fac[0][0]].kind <= BO_LEAF_MATERN52 && md.leaf[md.fac[0][0]].dpad <= 32;
  }
  std::vector<size_t> piece_rows;
  if (piece_mode) {
    // a first piece of ~1 MiB (the kernels start as soon as it has landed), then ~2 MiB pieces (the K(X*,X) work on the last
    // piece is what stays exposed); whole 128-row blocks, at most 16 pieces
    static const int piece_mib = []() { const char* e = getenv("EVEREST_HOST_PIECE_MIB"); int v = e ? atoi(e) : 5; return v >= 1 ? v : 5; }();
    size_t per = std::max<size_t>(128, (((size_t)piece_mib << 20) / row_wire) / 128 * 128);
    if ((n_rows + per - 1) / per > 15) per = ((n_rows + 14) / 15 + 127) / 128 * 128;
    const size_t first = std::min(n_rows, std::max<size_t>(128, (((size_t)1 << 20) / row_wire) / 128 * 128));
    piece_rows.push_back(first);
    for (size_t r = first; r < n_rows; r += per) piece_rows.push_back(std::min(per, n_rows - r));
  }
  const int n_chunks = piece_mode ? (int)piece_rows.size() : (int)chunk_b.size();
  while ((int)st->copy_events.size() < n_chunks) {
    cudaEvent_t e;
    CUDA_CHECK_RET(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    st->copy_events.push_back(e);
  }
  // the staging buffer may still be read by kernels of an earlier call on `s`: order the first copy after them
  CUDA_CHECK_RET(cudaEventRecord(st->copy_events[0], s));
  CUDA_CHECK_RET(cudaStreamWaitEvent(st->copy_stream, st->copy_events[0], 0));
  char* pin = reinterpret_cast<char*>(st->pin_in);
  // wire layout of the packed modes: [dense of all rows][bits of all rows], on the host and in stage_pk alike
  double* pin_dense = reinterpret_cast<double*>(pin);
  unsigned long long* pin_bits = reinterpret_cast<unsigned long long*>(pin + n_rows * (size_t)nd * 8);
  char* dev_pk = reinterpret_cast<char*>(st->stage_pk.p);
  // A staging thread walks the chunks (pageable -> pinned with up to 4 helper threads, then the H2D on the copy stream) and
  // publishes each chunk with an event; this thread launches the kernels of a chunk as soon as its event exists.  The
  // launch thread may block inside a chunk (the per-row guard of the INT8 GEMM reads one counter back), the staging thread
  // keeps the copy engine busy meanwhile.
  std::vector<size_t> c_row0(n_chunks), c_rows(n_chunks);
  std::vector<int> c_b0(n_chunks);
  if (piece_mode) {
    size_t r = 0;
    for (int c = 0; c < n_chunks; r += piece_rows[c], ++c) { c_b0[c] = 0; c_row0[c] = r; c_rows[c] = piece_rows[c]; }
  } else {
    for (int c = 0, b0 = 0; c < n_chunks; b0 += chunk_b[c], ++c) {
      c_b0[c] = b0;
      c_row0[c] = (size_t)b0 * q;
      c_rows[c] = (size_t)chunk_b[c] * q;
    }
  }
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  std::mutex mtx;
  std::condition_variable cv;
  int published = 0;                  // chunks whose H2D has been enqueued and whose event is recorded
  cudaError_t stage_err = cudaSuccess;
  bool pack_ok = true;
  auto stage_chunk = [&](int c) -> cudaError_t {
    const size_t r0 = c_row0[c], nr = c_rows[c];
    const size_t host_bytes = nr * (mode == HOST_PREPACKED ? row_wire : (size_t)st->d * 8);
    const size_t piece = piece_mode ? (size_t)1 << 20 : (size_t)4 << 20;
    const int n_thr = (int)std::max<size_t>(1, std::min<size_t>(mode == HOST_PACK ? 8 : 4, (host_bytes + piece - 1) / piece));
    const size_t per = (nr + n_thr - 1) / n_thr;
    std::vector<std::thread> th;
    std::vector<char> okv(n_thr, 1);
    auto work = [&](int t) {
      const size_t a = std::min(r0 + nr, r0 + (size_t)t * per), e = std::min(r0 + nr, a + per);
      if (e <= a) return;
      if (mode == HOST_DENSE) memcpy(pin + a * row_wire, reinterpret_cast<const char*>(X_host) + a * row_wire, (e - a) * row_wire);
      else if (mode == HOST_PACK) okv[t] = pack_rows(st, X_host, a, e, pin_dense, pin_bits) ? 1 : 0;
      else {
        memcpy(pin_dense + a * nd, dense_host + a * nd, (e - a) * (size_t)nd * 8);
        memcpy(pin_bits + a * W, bits_host + a * W, (e - a) * (size_t)W * 8);
      }
    };
    for (int t = 1; t < n_thr; ++t) th.emplace_back(work, t);
    work(0);
    for (auto& t : th) t.join();
    for (char ok : okv) if (!ok) pack_ok = false;
    cudaError_t e;
    if (!packed) {
      e = cudaMemcpyAsync(reinterpret_cast<char*>(st->stage_in.p) + r0 * row_wire, pin + r0 * row_wire, nr * row_wire,
                          cudaMemcpyHostToDevice, st->copy_stream);
    } else {
      e = cudaSuccess;
      if (nd) e = cudaMemcpyAsync(dev_pk + r0 * (size_t)nd * 8, pin_dense + r0 * nd, nr * (size_t)nd * 8, cudaMemcpyHostToDevice, st->copy_stream);
      if (e == cudaSuccess && W)
        e = cudaMemcpyAsync(dev_pk + n_rows * (size_t)nd * 8 + r0 * (size_t)W * 8, pin_bits + r0 * W, nr * (size_t)W * 8,
                            cudaMemcpyHostToDevice, st->copy_stream);
    }
    if (e == cudaSuccess) e = cudaEventRecord(st->copy_events[c], st->copy_stream);
    return e;
  };
  // Piece mode: a team of staging threads spawned once per call; thread t copies slice t of EVERY piece (pageable -> pinned)
  // and moves on; whoever completes a piece last enqueues its H2D + event and publishes it.  (Spawning helpers per piece cost
  // 0.4 ms per 4 MiB piece: the copies then arrived more slowly than K(X*,X) consumed them.)
  std::vector<std::thread> team;
  std::vector<char> piece_ready(n_chunks, 0);
  std::unique_ptr<std::atomic<int>[]> piece_cnt;
  if (piece_mode) {
    const int T = 6;
    piece_cnt.reset(new std::atomic<int>[n_chunks]);
    for (int c = 0; c < n_chunks; ++c) piece_cnt[c].store(0);
    for (int t = 0; t < T; ++t)
      team.emplace_back([&, t]() {
        cudaSetDevice(dev_id);
        for (int c = 1; c < n_chunks; ++c) {                 // (piece 0 is staged by the calling thread, below)
          const size_t r0 = c_row0[c], nr = c_rows[c];
          const size_t a = r0 + nr * (size_t)t / T, e = r0 + nr * (size_t)(t + 1) / T;
          if (e > a) memcpy(pin + a * row_wire, reinterpret_cast<const char*>(X_host) + a * row_wire, (e - a) * row_wire);
          if (piece_cnt[c].fetch_add(1) + 1 == T) {
            cudaError_t er = cudaMemcpyAsync(reinterpret_cast<char*>(st->stage_in.p) + r0 * row_wire, pin + r0 * row_wire,
                                             nr * row_wire, cudaMemcpyHostToDevice, st->copy_stream);
            if (er == cudaSuccess) er = cudaEventRecord(st->copy_events[c], st->copy_stream);
            {
              std::lock_guard<std::mutex> lk(mtx);
              if (er != cudaSuccess && stage_err == cudaSuccess) stage_err = er;
              piece_ready[c] = 1;
            }
            cv.notify_all();
          }
        }
      });
  }
  // the first chunk is staged here (nothing to overlap with yet), the others by the staging thread
  if (!piece_mode) {
    stage_err = stage_chunk(0);
    published = 1;
  } else {
    // the small first piece: copied and sent by this thread while the team starts on the others
    const size_t nr = c_rows[0];
    memcpy(pin, X_host, nr * row_wire);
    cudaError_t er = cudaMemcpyAsync(st->stage_in.p, pin, nr * row_wire, cudaMemcpyHostToDevice, st->copy_stream);
    if (er == cudaSuccess) er = cudaEventRecord(st->copy_events[0], st->copy_stream);
    std::lock_guard<std::mutex> lk(mtx);
    if (er != cudaSuccess && stage_err == cudaSuccess) stage_err = er;
    piece_ready[0] = 1;
  }
  std::thread stager;
  if (!piece_mode && n_chunks > 1 && stage_err == cudaSuccess) {
    stager = std::thread([&]() {
      cudaSetDevice(dev_id);
      for (int c = 1; c < n_chunks; ++c) {
        cudaError_t e = stage_chunk(c);
        {
          std::lock_guard<std::mutex> lk(mtx);
          if (e != cudaSuccess && stage_err == cudaSuccess) stage_err = e;
          published = c + 1;
        }
        cv.notify_all();
        if (e != cudaSuccess) break;
      }
    });
  }
  int rc_fwd = BO_OK;
  static const bool trace = getenv("EVEREST_HOST_TRACE") != nullptr;
  auto t_start = std::chrono::steady_clock::now();
  auto ms_since = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
  if (piece_mode) {
    st->piece_row_end.clear();
    for (int c = 0; c < n_chunks; ++c) st->piece_row_end.push_back((int)(c_row0[c] + c_rows[c]));
    st->piece_wait = [&](int c) -> int {
      {
        std::unique_lock<std::mutex> lk(mtx);
        cv.wait(lk, [&]() { return piece_ready[c] || stage_err != cudaSuccess; });
        if (stage_err != cudaSuccess) { bo_set_error("forward_host: staging failed: %s", cudaGetErrorString(stage_err)); return BO_ERR_CUDA; }
      }
      if (trace) fprintf(stderr, "[host] piece %d published at %.3f ms\n", c, ms_since());
      CUDA_CHECK_RET(cudaStreamWaitEvent(s, st->copy_events[c], 0));
      return BO_OK;
    };
    rc_fwd = bo_acqf_forward(st, st->stage_in.as<double>(), b, q, zq_dev, st->stage_out.as<double>(), nullptr, s);
    st->piece_row_end.clear();
    st->piece_wait = nullptr;
    if (trace) fprintf(stderr, "[host] piece-wise forward returned at %.3f ms\n", ms_since());
  }
  for (int c = 0; c < n_chunks && rc_fwd == BO_OK && !piece_mode; ++c) {
    {
      std::unique_lock<std::mutex> lk(mtx);
      cv.wait(lk, [&]() { return published > c || stage_err != cudaSuccess; });
      if (stage_err != cudaSuccess) break;
    }
    if (trace) fprintf(stderr, "[host] chunk %d (%d q-batches) published at %.3f ms\n", c, chunk_b[c], ms_since());
    cudaError_t e = cudaStreamWaitEvent(s, st->copy_events[c], 0);
    if (e != cudaSuccess) { std::lock_guard<std::mutex> lk(mtx); stage_err = e; break; }
    double* Xc = st->stage_in.as<double>() + c_row0[c] * st->d;
    if (packed)
      rc_fwd = launch_unpack_rows(reinterpret_cast<const double*>(dev_pk) + c_row0[c] * nd,
                                  reinterpret_cast<const u64*>(dev_pk + n_rows * (size_t)nd * 8) + c_row0[c] * W, nd, W,
                                  st->pack_src.as<int>(), (int)c_rows[c], st->d, Xc, s, &st->lc);
    if (rc_fwd == BO_OK)
      rc_fwd = bo_acqf_forward(st, Xc, chunk_b[c], q, zq_dev, st->stage_out.as<double>() + c_b0[c], nullptr, s);
    if (trace) fprintf(stderr, "[host] chunk %d launched, forward returned at %.3f ms\n", c, ms_since());
  }
  if (stager.joinable()) stager.join();
  for (auto& t : team) t.join();
  if (stage_err != cudaSuccess) { bo_set_error("forward_host: staging failed: %s", cudaGetErrorString(stage_err)); return BO_ERR_CUDA; }
  RC(rc_fwd);
  CUDA_CHECK_RET(cudaMemcpyAsync(st->pin_out, st->stage_out.p, out_bytes, cudaMemcpyDeviceToHost, s));
  CUDA_CHECK_RET(cudaStreamSynchronize(s));
  if (!pack_ok) { bo_set_error("fingerprint (Tanimoto) columns must hold 0 / 1"); return BO_ERR_INVALID; }
  memcpy(out_host, st->pin_out, out_bytes);
  if (trace) fprintf(stderr, "[host] done at %.3f ms\n", ms_since());
  return BO_OK;
}

extern "C" int bo_acqf_forward_host(bo_state* st, const double* X_host, int32_t b, int32_t q, const double* zq_dev,
                                    double* out_host, void* stream) {
  if (!st) { bo_set_error("null state"); return BO_ERR_INVALID; }
  if (!X_host || !out_host) { bo_set_error("forward_host: null buffer"); return BO_ERR_INVALID; }
  // wide fingerprint blocks go over PCIe as bits (EVEREST_HOST_PACK=0 keeps the float64 wire format)
  static const int pack_env = []() { const char* e = getenv("EVEREST_HOST_PACK"); return e ? atoi(e) : 1; }();
  pack_layout_init(st);
  const bool pack = pack_env != 0 && st->pack_bit_cols.size() >= 256;
  return forward_host_impl(st, pack ? HOST_PACK : HOST_DENSE, X_host, nullptr, nullptr, b, q, zq_dev, out_host, stream);
}

extern "C" int bo_acqf_forward_host_packed(bo_state* st, const double* dense_host, const uint64_t* bits_host, int32_t b, int32_t q,
                                           const double* zq_dev, double* out_host, void* stream) {
  if (!st) { bo_set_error("null state"); return BO_ERR_INVALID; }
  pack_layout_init(st);
  if ((!dense_host && !st->pack_dense_cols.empty()) || (!bits_host && !st->pack_bit_cols.empty()) || !out_host) {
    bo_set_error("forward_host_packed: null buffer"); return BO_ERR_INVALID;
  }
  return forward_host_impl(st, HOST_PREPACKED, nullptr, dense_host, bits_host, b, q, zq_dev, out_host, stream);
}

// ---- stand-alone multi-objective utilities (no GP state) ---------------------------------------------
struct TempBufs {
  std::vector<void*> ptrs;
  ~TempBufs() { for (void* p : ptrs) cudaFree(p); }
  template <typename T> int alloc(T** out, size_t count) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, std::max<size_t>(count * sizeof(T), 16));
    if (e != cudaSuccess) { bo_set_error("cudaMalloc failed: %s", cudaGetErrorString(e)); return BO_ERR_CUDA; }
    ptrs.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return BO_OK;
  }
};

static int front_of(const double* Y_dev, int n, int m, const double* ref_host, int dedup, TempBufs& tb, unsigned char** front,
                    unsigned char** feas, double** ref_dev, cudaStream_t s) {
  RC(tb.alloc(front, (size_t)std::max(n, 1)));
  RC(tb.alloc(feas, (size_t)std::max(n, 1)));
  RC(tb.alloc(ref_dev, (size_t)BO_MAX_OBJECTIVES));
  std::vector<double> ref(BO_MAX_OBJECTIVES, -INFINITY);
  if (ref_host) for (int o = 0; o < m; ++o) ref[o] = ref_host[o];
  CUDA_CHECK_RET(cudaMemcpyAsync(*ref_dev, ref.data(), BO_MAX_OBJECTIVES * 8, cudaMemcpyHostToDevice, s));
  CUDA_CHECK_RET(cudaMemsetAsync(*feas, 1, (size_t)std::max(n, 1), s));
  RC(launch_front(Y_dev, *feas, 1, n, m, *ref_dev, dedup, *front, nullptr, s, nullptr));
  return BO_OK;
}

extern "C" int bo_pareto_mask(const double* Y_dev, int32_t n, int32_t m, int32_t deduplicate, int32_t* mask_dev, void* stream) {
  if (n < 0 || m < 1 || m > BO_MAX_OBJECTIVES) { bo_set_error("pareto_mask: bad n=%d / m=%d", n, m); return BO_ERR_INVALID; }
  if (n == 0) return BO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  TempBufs tb;
  unsigned char *front, *feas;
  double* ref_dev;
  RC(front_of(Y_dev, n, m, nullptr, deduplicate ? 1 : 0, tb, &front, &feas, &ref_dev, s));
  RC(launch_front_to_mask(front, n, mask_dev, s, nullptr));
  CUDA_CHECK_RET(cudaStreamSynchronize(s));
  return BO_OK;
}

extern "C" int bo_hypervolume(const double* Y_dev, int32_t n, int32_t m, const double* ref_point, double* hv_out, void* stream) {
  if (n < 0 || m < 2 || m > BO_MAX_OBJECTIVES || !ref_point || !hv_out) { bo_set_error("hypervolume: bad arguments"); return BO_ERR_INVALID; }
  *hv_out = 0.0;
  if (n == 0) return BO_OK;
  cudaStream_t s = (cudaStream_t)stream;
  TempBufs tb;
  unsigned char *front, *feas;
  double* ref_dev;
  RC(front_of(Y_dev, n, m, ref_point, 1, tb, &front, &feas, &ref_dev, s));
  double *lo, *up, *work, *hv_dev;
  int *ncells, *overflow;
  RC(tb.alloc(&ncells, 1));
  RC(tb.alloc(&overflow, 1));
  RC(tb.alloc(&hv_dev, 1));
  int cap = (m == 2) ? n + 1 : std::max(64, 8 * (n + 1) * m);
  for (;;) {
    RC(tb.alloc(&lo, (size_t)cap * m));
    RC(tb.alloc(&up, (size_t)cap * m));
    if (m == 2) {
      RC(launch_partition2d(Y_dev, front, n, 1, cap, ref_dev, lo, up, ncells, nullptr, s, nullptr));
      break;
    }
    RC(tb.alloc(&work, (size_t)2 * cap * (m + m * m)));
    CUDA_CHECK_RET(cudaMemsetAsync(overflow, 0, sizeof(int), s));
    RC(launch_partition_nd(Y_dev, front, n, 1, m, cap, ref_dev, work, lo, up, ncells, overflow, s, nullptr));
    int ov = 0;
    CUDA_CHECK_RET(cudaMemcpyAsync(&ov, overflow, sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_CHECK_RET(cudaStreamSynchronize(s));
    if (!ov) break;
    cap *= 2;
    if ((size_t)cap * (m + m * m) * 16 > ((size_t)8 << 30)) { bo_set_error("hypervolume: decomposition too large"); return BO_ERR_INVALID; }
  }
  RC(launch_hypervolume_from_cells(Y_dev, front, n, m, ref_dev, lo, up, ncells, hv_dev, s, nullptr));
  CUDA_CHECK_RET(cudaMemcpyAsync(hv_out, hv_dev, sizeof(double), cudaMemcpyDeviceToHost, s));
  CUDA_CHECK_RET(cudaStreamSynchronize(s));
  return BO_OK;
}

extern "C" int64_t bo_launch_count(const bo_state* st) { return st ? st->lc.n : 0; }

extern "C" int bo_set_timing(bo_state* st, int32_t enabled) {
  if (!st) return BO_ERR_INVALID;
  st->timing = enabled != 0;
  return BO_OK;
}

extern "C" int bo_last_timing(const bo_state* st, const char* name, double* ms) {
  if (!st || !ms) return BO_ERR_INVALID;
  double total = 0.0;
  int cnt = 0;
  for (auto& r : st->recs) {
    if (r.name != name) continue;
    if (cudaEventSynchronize(r.b) != cudaSuccess) { bo_set_error("event sync failed"); return BO_ERR_CUDA; }
    float t = 0.f;
    cudaEventElapsedTime(&t, r.a, r.b);
    total += t;
    ++cnt;
  }
  *ms = total;
  return cnt;
}

extern "C" int bo_debug_get(bo_state* st, const char* name, int32_t m, double* out_dev, int64_t capacity,
                            int64_t* n_written, void* stream) {
  if (!st || !name) return BO_ERR_INVALID;
  cudaStream_t s = (cudaStream_t)stream;
  std::string nm(name);
  const double* src = nullptr;
  int64_t n = 0;
  auto chk_m = [&]() { return m >= 0 && m < st->M; };
  if (nm == "L" && chk_m()) { src = st->out[m].L.as<double>(); n = (int64_t)st->N * st->ldk; }
  else if (nm == "Linv" && chk_m()) { src = st->out[m].Linv.as<double>(); n = (int64_t)st->N * st->ldk; }
  else if (nm == "alpha" && chk_m()) { src = st->out[m].alpha_row.as<double>(); n = st->N; }
  else if (nm == "baseline_L" && chk_m()) { src = st->out[m].Lb.as<double>(); n = (int64_t)st->nb * st->ldlb; }
  else if (nm == "Kx") { src = st->wsKx.as<double>(); n = std::min<int64_t>(capacity, (int64_t)st->wsKx.bytes / 8); }
  else if (nm == "root") { src = st->wsRoot.as<double>(); n = std::min<int64_t>(capacity, (int64_t)st->wsRoot.bytes / 8); }
  else if (nm == "mu") { src = st->wsMu.as<double>(); n = std::min<int64_t>(capacity, (int64_t)st->wsMu.bytes / 8); }
  else if (nm == "cell_lo") { src = st->cell_lo.as<double>(); n = (int64_t)st->cap * st->od.n_obj * (st->cells_shared ? 1 : st->S); }
  else if (nm == "cell_up") { src = st->cell_up.as<double>(); n = (int64_t)st->cap * st->od.n_obj * (st->cells_shared ? 1 : st->S); }
  else if (nm == "samples_b") { src = st->samples_b.as<double>(); n = (int64_t)st->S * st->nb * st->M; }
  else if (nm == "obj_b") { src = st->obj_b.as<double>(); n = (int64_t)st->S * st->nb * st->od.n_obj; }
  else if (nm == "best_f_s" && st->noisy_scalar) { src = st->best_f_s.as<double>(); n = st->S; }
  else if (nm == "ozaki_check") {
    // [state (0 = no INT8 call yet since the last prepare, 1 = INT8 path with per-row guard, -1 = guard sent the state to FP64),
    //  q-batches redone in FP64 by the guard in the last forward, q-batches of the last forward, the same two since the last
    //  prepare, kappa, tol]
    if (capacity < 7) { bo_set_error("debug_get: capacity too small"); return BO_ERR_INVALID; }
    double h[7] = {(double)st->oz_calib, (double)st->oz_flagged_last, (double)st->oz_batches_last, (double)st->oz_flagged_total,
                   (double)st->oz_batches_total, st->oz_kappa, st->oz_tol};
    CUDA_CHECK_RET(cudaMemcpyAsync(out_dev, h, sizeof(h), cudaMemcpyHostToDevice, s));
    CUDA_CHECK_RET(cudaStreamSynchronize(s));
    if (n_written) *n_written = 7;
    return BO_OK;
  }
  else if (nm == "Gqq") { src = st->wsGqq.as<double>(); n = std::min<int64_t>(capacity, (int64_t)st->wsGqq.bytes / 8); }
  else if (nm == "mu_raw") { src = st->wsMuRaw.as<double>(); n = std::min<int64_t>(capacity, (int64_t)st->wsMuRaw.bytes / 8); }
  else if (nm == "ncells" || nm == "front_idx") {
    // integer buffers are returned through the same byte pipe (caller views them as int32)
    const void* isrc = (nm == "ncells") ? st->ncells.p : st->front_idx.p;
    int64_t bytes = (nm == "ncells") ? (int64_t)(st->cells_shared ? 1 : st->S) * 4 : (int64_t)st->S * st->cap * 4;
    if (bytes > capacity * 8) { bo_set_error("debug_get: capacity too small"); return BO_ERR_INVALID; }
    CUDA_CHECK_RET(cudaMemcpyAsync(out_dev, isrc, bytes, cudaMemcpyDeviceToDevice, s));
    if (n_written) *n_written = bytes / 4;
    return BO_OK;
  } else { bo_set_error("debug_get: unknown buffer '%s'", name); return BO_ERR_INVALID; }
  if (n > capacity) { bo_set_error("debug_get: capacity %lld < %lld", (long long)capacity, (long long)n); return BO_ERR_INVALID; }
  CUDA_CHECK_RET(cudaMemcpyAsync(out_dev, src, n * 8, cudaMemcpyDeviceToDevice, s));
  if (n_written) *n_written = n;
  return BO_OK;
}
