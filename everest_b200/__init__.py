"""everest_b200 -- B200-native acquisition evaluation for BoFire's predictive strategies.

Host-side mirror of the reference's interface for this one path; all arithmetic runs
in hand-written sm_100a CUDA kernels behind the C ABI in include/everest_b200.h.  No CPU fallback.
"""
from . import kernels, multiobjective, objectives, sampling  # noqa: F401
from ._lib import EverestError, NotPSDError  # noqa: F401
from .acquisition import (  # noqa: F401
    get_acquisition_function,
    qExpectedHypervolumeImprovement,
    qExpectedImprovement,
    qLogExpectedHypervolumeImprovement,
    qLogExpectedImprovement,
    qLogNoisyExpectedHypervolumeImprovement,
    qLogNoisyExpectedImprovement,
    qNoisyExpectedHypervolumeImprovement,
    qNoisyExpectedImprovement,
    qProbabilityOfImprovement,
    qSimpleRegret,
    qUpperConfidenceBound,
)
from .model import DeviceGPState, SingleTaskGPSpec, normalize_bounds, standardize_stats  # noqa: F401
from .optim import (calc_acquisition, gen_batch_initial_conditions, gen_candidates_scipy, initialize_q_batch,
                    optimize_acqf, optimize_acqf_discrete, optimize_acqf_mixed)  # noqa: F401

__version__ = "0.1.0"
