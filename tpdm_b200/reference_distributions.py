"""Drop-in for /root/reference/src/models/reference_distributions.py (KL prior of the RLOO rollout; scalar host math)."""
import math

import torch

EPSILON = 1e-3
CONCENTRATION = 20
ex = math.exp(1)


def get_ref_beta(sigmas_1, num_steps=28):
    """(alpha, beta) of the concentration-20 Beta whose mode reproduces one step of the 28-step shifted schedule
    (reference_distributions.py:9-19)."""
    t_1 = sigmas_1 / (ex + (1 - ex) * sigmas_1)
    t_2 = torch.clamp(t_1 - 1.0 / num_steps, EPSILON)
    sigmas_2 = ex / (ex + 1 / t_2 - 1)
    mode = sigmas_2 / sigmas_1
    alpha = mode * (CONCENTRATION - 2) + 1
    beta = (1 - mode) * (CONCENTRATION - 2) + 1
    return alpha, beta
