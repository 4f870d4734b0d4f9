"""Base samples for the MC acquisition functions (host side, torch SobolEngine -- the same generator
BoTorch's SobolQMCNormalSampler uses on CPU in the reference; [UPSTREAM] botorch.sampling.qmc)."""
import math

import torch


def draw_sobol_normal_samples(d: int, n: int, seed: int) -> torch.Tensor:
    eng = torch.quasirandom.SobolEngine(d, scramble=True, seed=seed)
    u = eng.draw(n, dtype=torch.double)
    v = 0.5 + (1 - 1e-10) * (u - 0.5)
    return torch.erfinv(2 * v - 1) * math.sqrt(2)


def base_samples(n_points: int, n_outputs: int, n_samples: int, seed: int) -> torch.Tensor:
    """z[S, n_points, M]; flat Sobol dimension = m * n_points + i (non-interleaved MTMVN layout)."""
    if n_points == 0:
        return torch.zeros(n_samples, 0, n_outputs, dtype=torch.double)
    z = draw_sobol_normal_samples(n_points * n_outputs, n_samples, seed)
    return z.view(n_samples, n_outputs, n_points).transpose(1, 2).contiguous()
