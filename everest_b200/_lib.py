"""ctypes binding of libeverest_b200.so (C ABI: include/everest_b200.h).

There is NO CPU fallback: if the shared library is missing or no CUDA device is present the
product path raises.  Build in-tree with ``python -m everest_b200.build`` (nvcc, sm_100a).
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libeverest_b200.so")

BO_MAX_FACTORS = 4
BO_MAX_Q = 16
BO_MAX_OBJECTIVES = 8
BO_MAX_CONSTRAINTS = 8

LEAF_RBF, LEAF_MATERN12, LEAF_MATERN32, LEAF_MATERN52, LEAF_HAMMING, LEAF_TANIMOTO = range(6)
OBJ_MAX, OBJ_MIN, OBJ_CLOSE_TO_TARGET, OBJ_MIN_SIGMOID, OBJ_MAX_SIGMOID, OBJ_TARGET = range(6)
COMBINE_SINGLE, COMBINE_ADDITIVE, COMBINE_MULTIPLICATIVE = range(3)
ACQF_QLOGEI, ACQF_QEI, ACQF_QSR, ACQF_QUCB, ACQF_QPI = range(5)

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int32)


class KernelLeaf(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_dims", C.c_int32), ("dims", c_int_p), ("cardinality", c_int_p),
                ("lengthscale", c_double_p), ("n_ls", C.c_int32)]


class KernelTerm(C.Structure):
    _fields_ = [("coef", C.c_double), ("n_factors", C.c_int32), ("factors", C.c_int32 * BO_MAX_FACTORS)]


class OutputModel(C.Structure):
    _fields_ = [("n_leaves", C.c_int32), ("leaves", C.POINTER(KernelLeaf)), ("n_terms", C.c_int32),
                ("terms", C.POINTER(KernelTerm)), ("in_offset", c_double_p), ("in_scale", c_double_p),
                ("mean_const", C.c_double), ("noise", C.c_double), ("y_mean", C.c_double), ("y_std", C.c_double),
                ("y", c_double_p)]


class StateConfig(C.Structure):
    _fields_ = [("N", C.c_int32), ("d", C.c_int32), ("M", C.c_int32), ("X_train", c_double_p),
                ("outputs", C.POINTER(OutputModel))]


class ObjectiveOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("out_idx", C.c_int32), ("p0", C.c_double), ("p1", C.c_double),
                ("p2", C.c_double), ("w", C.c_double)]


class ConstraintOp(C.Structure):
    _fields_ = [("out_idx", C.c_int32), ("sign", C.c_double), ("tp", C.c_double), ("eta", C.c_double)]


# every symbol include/everest_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "bo_version": (C.c_int, []),
    "bo_last_error": (C.c_char_p, []),
    "bo_state_create": (C.c_int, [C.POINTER(StateConfig), C.POINTER(C.c_void_p)]),
    "bo_state_destroy": (None, [C.c_void_p]),
    "bo_state_factorize": (C.c_int, [C.c_void_p, c_int_p, c_double_p, C.c_void_p]),
    "bo_state_set_hyperparameters": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "bo_posterior_marginal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bo_posterior_joint": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "bo_prune_counts": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(ObjectiveOp),
                                  C.c_int32, C.POINTER(ConstraintOp), C.c_int32, c_double_p, C.c_void_p, c_int_p,
                                  C.c_void_p]),
    "bo_nehvi_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.POINTER(ObjectiveOp),
                                   C.c_int32, C.POINTER(ConstraintOp), C.c_int32, c_double_p, c_int_p, c_int_p,
                                   C.c_void_p]),
    "bo_ehvi_prepare": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(ObjectiveOp), C.c_int32,
                                  c_double_p, c_int_p, C.c_void_p]),
    "bo_logei_prepare": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(ObjectiveOp), C.c_int32, C.c_double,
                                   C.c_void_p]),
    "bo_scalar_prepare": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_int32, C.c_int32, C.POINTER(ObjectiveOp), C.c_int32,
                                    C.POINTER(ConstraintOp), C.c_int32, C.c_double, C.c_void_p, C.c_int32, C.c_void_p,
                                    c_int_p, C.c_void_p]),
    "bo_prune_counts_scalar": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                         C.POINTER(ObjectiveOp), C.c_int32, C.POINTER(ConstraintOp), C.c_int32, C.c_void_p,
                                         c_int_p, C.c_void_p]),
    "bo_scalar_baseline_best": (C.c_int, [C.c_void_p, C.c_double, c_int_p, C.c_void_p]),
    "bo_acqf_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_double]),
    "bo_acqf_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p]),
    "bo_acqf_forward_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_void_p]),
    "bo_acqf_resample_flagged": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.POINTER(C.c_int32), C.c_void_p]),
    "bo_acqf_last_resampled": (C.c_int32, [C.c_void_p]),
    "bo_acqf_optimize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, c_double_p, c_double_p, C.c_void_p,
                                   C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "bo_acqf_forward_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "bo_pack_layout": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_void_p, C.c_void_p]),
    "bo_pack_rows_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "bo_acqf_forward_host_packed": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                              C.c_void_p]),
    "bo_pareto_mask": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bo_hypervolume": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p, C.c_void_p]),
    "bo_mll_forward_backward": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int32,
                                          c_double_p, C.c_int32, C.c_void_p]),
    "bo_sobol_scramble": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]),
    "bo_sobol_uniform": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bo_sobol_normal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "bo_debug_get": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int32, C.c_void_p, C.c_int64, C.POINTER(C.c_int64),
                               C.c_void_p]),
    "bo_launch_count": (C.c_int64, [C.c_void_p]),
    "bo_last_timing": (C.c_int, [C.c_void_p, C.c_char_p, c_double_p]),
    "bo_set_timing": (C.c_int, [C.c_void_p, C.c_int32]),
}

_lib = None


class EverestError(RuntimeError):
    pass


class NotPSDError(EverestError):
    """Mirrors linear_operator.utils.errors.NotPSDError."""


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EverestError(
            f"{LIB_PATH} is missing: build it with `python -m everest_b200.build` (nvcc, sm_100a). "
            "everest_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the ABI drifted
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == 0:
        return
    msg = load().bo_last_error().decode()
    if rc == -1:
        raise ValueError(msg)
    if rc == -3:
        raise NotPSDError(msg)
    raise EverestError(f"everest_b200 error {rc}: {msg}")
