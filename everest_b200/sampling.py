"""Base samples for the MC acquisition functions ([UPSTREAM] botorch.sampling.qmc / SobolQMCNormalSampler: torch
SobolEngine(scramble=True, seed).draw(S) -> v = 0.5 + (1 - 1e-10)(u - 0.5) -> sqrt(2) erfinv(2v - 1)).

`base_samples` is the host version (what the reference does on CPU); `base_samples_device` runs the scramble, the
Gray-code draw and the inverse CDF in sm_100a kernels (csrc/sobol.cu) and reproduces torch's integer pipeline bit for
bit -- only the random bits (shift vector, scramble matrices) still come from torch's CPU generator so that a seed means
exactly what it means in BoTorch."""
import ctypes as C
import math

import torch

MAXBIT = 30


def draw_sobol_normal_samples(d: int, n: int, seed: int) -> torch.Tensor:
    eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=seed)
    u = eng.draw(n, dtype=torch.double)
    v = 0.5 + (1 - 1e-10) * (u - 0.5)
    return torch.erfinv(2 * v - 1) * math.sqrt(2)


MAXDIM = torch.quasirandom.SobolEngine.MAXDIM   # 21201


def iid_normal_samples(d: int, n: int, seed: int) -> torch.Tensor:
    """[UPSTREAM] botorch.sampling.get_sampler falls back from SobolQMCNormalSampler to IIDNormalSampler when the joint
    sample dimension exceeds SobolEngine.MAXDIM (e.g. pruning a baseline of more than 10600 points with two outputs):
    seeded i.i.d. N(0, 1) draws."""
    g = torch.Generator().manual_seed(int(seed))
    return torch.randn(n, d, dtype=torch.double, generator=g)


def base_samples(n_points: int, n_outputs: int, n_samples: int, seed: int) -> torch.Tensor:
    """z[S, n_points, M]; flat Sobol dimension = m * n_points + i (non-interleaved MTMVN layout)."""
    if n_points == 0:
        return torch.zeros(n_samples, 0, n_outputs, dtype=torch.double)
    dim = n_points * n_outputs
    z = draw_sobol_normal_samples(dim, n_samples, seed) if dim <= MAXDIM else iid_normal_samples(dim, n_samples, seed)
    return z.view(n_samples, n_outputs, n_points).transpose(1, 2).contiguous()


def sobol_scramble_inputs(dim: int, seed: int):
    """The host part of SobolEngine.__init__ / _scramble: unscrambled direction numbers, the random shift and the random
    unit-lower-triangular matrices packed one 30-bit row per integer (bit 29 - k of row p = L[p][k])."""
    if dim < 1 or dim > torch.quasirandom.SobolEngine.MAXDIM:
        raise ValueError(f"Supported range of dimensionality for SobolEngine is [1, {torch.quasirandom.SobolEngine.MAXDIM}]")
    ss = torch.zeros(dim, MAXBIT, dtype=torch.long)
    torch._sobol_engine_initialize_state_(ss, dim)
    g = torch.Generator()
    g.manual_seed(int(seed))
    shift_ints = torch.randint(2, (dim, MAXBIT), generator=g)
    pw = torch.pow(2, torch.arange(0, MAXBIT))
    shift = torch.mv(shift_ints, pw)
    ltm = torch.randint(2, (dim, MAXBIT, MAXBIT), generator=g).tril(-1) + torch.eye(MAXBIT, dtype=torch.long)
    rows = (ltm * pw.flip(0)).sum(dim=-1)          # k-th column weighs 2^(29 - k)
    return ss, shift, rows


def base_samples_device(n_points: int, n_outputs: int, n_samples: int, seed: int, device) -> torch.Tensor:
    """Same contract as `base_samples`, generated on `device` (a CUDA device)."""
    from . import _lib as L

    device = torch.device(device)
    if n_points == 0:
        return torch.zeros(n_samples, 0, n_outputs, dtype=torch.double, device=device)
    dim = n_points * n_outputs
    if dim > MAXDIM:      # IIDNormalSampler fallback (see iid_normal_samples); same layout as the Sobol draw
        return base_samples(n_points, n_outputs, n_samples, seed).to(device)
    ss, shift, rows = sobol_scramble_inputs(dim, seed)
    lib = L.load()
    ss_d, shift_d, rows_d = ss.to(device), shift.to(device), rows.to(device)
    out = torch.empty(n_samples, n_points, n_outputs, dtype=torch.double, device=device)
    with torch.cuda.device(device):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.check(lib.bo_sobol_scramble(C.c_void_p(ss_d.data_ptr()), C.c_void_p(rows_d.data_ptr()), dim, stream))
        L.check(lib.bo_sobol_normal(C.c_void_p(ss_d.data_ptr()), C.c_void_p(shift_d.data_ptr()), n_points, n_outputs,
                                    n_samples, C.c_void_p(out.data_ptr()), stream))
    return out


def sobol_uniform_device(dim: int, n: int, seed: int, device) -> torch.Tensor:
    """SobolEngine(dim, scramble=True, seed).draw(n, dtype=float64) on `device`, bit-identical to torch's engine."""
    from . import _lib as L

    device = torch.device(device)
    ss, shift, rows = sobol_scramble_inputs(dim, seed)
    lib = L.load()
    ss_d, shift_d, rows_d = ss.to(device), shift.to(device), rows.to(device)
    out = torch.empty(n, dim, dtype=torch.double, device=device)
    with torch.cuda.device(device):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        L.check(lib.bo_sobol_scramble(C.c_void_p(ss_d.data_ptr()), C.c_void_p(rows_d.data_ptr()), dim, stream))
        L.check(lib.bo_sobol_uniform(C.c_void_p(ss_d.data_ptr()), C.c_void_p(shift_d.data_ptr()), dim, n,
                                     C.c_void_p(out.data_ptr()), stream))
    return out
