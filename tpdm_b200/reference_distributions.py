"""Reference schedule prior of the RLOO rollout (module name and `get_ref_beta` signature as in the reference's
src/models/reference_distributions.py, which `modeling_sd3_pnt.py:26` imports).

One step of the 28-step *shifted* flow-matching schedule is used as the prior for the TimePredictor: with the time-shift
map  sigma(t) = e t / (1 + (e - 1) t)  and its inverse  t(sigma) = sigma / (e + (1 - e) sigma),  a step of size 1/num_steps
in t from the current sigma lands at sigma'; the prior is the Beta distribution with concentration 20 whose mode is the ratio
sigma' / sigma.  The device version (KL against this prior) lives in csrc/tpm_train.cu::rollout_shaping_kernel; this host
function exists for callers of the reference API and is checked bit for bit against the pinned restatement in the tests."""
import math

import torch

_SHIFT = math.e          # time-shift factor of the schedule
_CONCENTRATION = 20.0    # alpha + beta of the prior
_T_FLOOR = 1e-3          # smallest t the schedule steps to


def _mode_of_reference_step(sigma: torch.Tensor, num_steps: int) -> torch.Tensor:
    t_now = sigma / (_SHIFT + (1 - _SHIFT) * sigma)                       # invert the shift
    t_next = torch.clamp(t_now - 1.0 / num_steps, _T_FLOOR)              # one uniform step in t, floored
    sigma_next = _SHIFT / (_SHIFT + 1 / t_next - 1)                       # shift again
    return sigma_next / sigma


def get_ref_beta(sigmas_1, num_steps=28):
    """-> (alpha, beta), same shape as `sigmas_1`: Beta(mode * 18 + 1, (1 - mode) * 18 + 1)."""
    mode = _mode_of_reference_step(sigmas_1, num_steps)
    spread = _CONCENTRATION - 2
    return mode * spread + 1, (1 - mode) * spread + 1
