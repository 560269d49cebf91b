"""CPU model of the fast attention kernel's softmax bookkeeping (tpdm_b200/csrc/attention_tcgen05.cu, attn_cta<DP, true>): the
reference m_ref is the row maximum of the FIRST 128-key tile only, and from then on it is guarded by the row sum -- at the top of a
tile, l > 2^32 moves the reference up by floor(log2 l) and rescales O and l by that power of two; l > 2^64 / inf / NaN (or an
argument > 127 in a polynomial slot) flags the tile for the exact pass.  The model mirrors that control flow in fp32 tile by tile
and checks (i) that whenever no flag is raised the result equals softmax(S) V, however the scores are ordered, and (ii) that the
flag is raised exactly for the inputs the guard cannot keep in range.  No GPU, no library: this pins the ALGORITHM; the kernel
itself is compared with fp32 SDPA in tests/test_parity_gpu.py."""
import math

import pytest
import torch

KT = 128
SOFT, HARD, POLY_MAX = 2.0 ** 32, 2.0 ** 64, 127.0


def guarded_softmax_v(scores: torch.Tensor, v: torch.Tensor, scale_log2: float):
    """scores [rows, S] fp32 (raw q.k), v [S, d].  Returns (out [rows, d], flagged: bool, n_renorm: int)."""
    rows, S = scores.shape
    x_all = scores.float() * scale_log2
    n_kv = (S + KT - 1) // KT
    m_ref = x_all[:, :KT].max(dim=1).values                      # row maximum of the first key tile
    l = torch.zeros(rows)
    o = torch.zeros(rows, v.shape[1])
    pmax = torch.full((rows,), -math.inf)
    hard = False
    n_renorm = 0
    for j in range(n_kv):
        bad = ~(l <= SOFT) | (pmax > POLY_MAX)                   # guard at the top of the tile
        if bool(bad.any()):
            if bool((~(l <= HARD) | (pmax > POLY_MAX)).any()):
                hard = True
            need = (l > SOFT) & torch.isfinite(l)
            e = torch.where(need, torch.floor(torch.log2(l.clamp_min(1.0))), torch.zeros(rows))
            alpha = torch.exp2(-e)
            m_ref = m_ref + e
            l = l * alpha
            o = o * alpha[:, None]
            pmax = torch.full((rows,), -math.inf)
            n_renorm += int(need.sum())
        x = x_all[:, j * KT:(j + 1) * KT] - m_ref[:, None]
        pmax = torch.maximum(pmax, x[:, 3::4].max(dim=1).values)  # every 4th pair runs through the polynomial: tracked
        p = torch.exp2(x)                                         # fp32: overflows to inf above 128
        l = l + p.sum(dim=1)
        o = o + p.to(torch.bfloat16).float() @ v[j * KT:(j + 1) * KT].float()
    if bool((~(l <= HARD) | (pmax > POLY_MAX)).any()):
        hard = True
    return o / l[:, None], hard, n_renorm


def reference(scores, v, scale_log2):
    return torch.softmax(scores.double() * scale_log2 * math.log(2.0), dim=1) @ v.double()


@pytest.mark.parametrize("order", ["random", "ascending", "descending"])
def test_ordinary_scores_need_no_renormalisation(order):
    g = torch.Generator().manual_seed(0)
    s = torch.randn(64, 1000, generator=g) * 20.0
    if order != "random":
        s = s.sort(dim=1, descending=(order == "descending")).values
    v = torch.randn(1000, 16, generator=g)
    out, hard, n = guarded_softmax_v(s, v, 1.4427 / 8)
    assert not hard
    if order != "ascending":
        assert n == 0
    assert float((out.double() - reference(s, v, 1.4427 / 8)).abs().max()) < 2e-2


def test_steadily_growing_scores_are_renormalised_in_place():
    """+40 log2 units per tile: every tile boundary moves the reference, nothing is flagged, the result stays exact."""
    g = torch.Generator().manual_seed(1)
    S = 8 * KT
    ramp = torch.arange(S).float() / KT * 40.0 / (1.4427 / 8)
    s = torch.randn(32, S, generator=g) * 4.0 + ramp[None, :]
    v = torch.randn(S, 8, generator=g)
    out, hard, n = guarded_softmax_v(s, v, 1.4427 / 8)
    assert not hard and n >= 32 * 5
    assert float((out.double() - reference(s, v, 1.4427 / 8)).abs().max()) < 2e-2


@pytest.mark.parametrize("jump_log2", [70.0, 200.0, 1e4])
def test_a_jump_the_guard_cannot_absorb_is_flagged(jump_log2):
    """A score more than 2^64 above everything the row has seen (here in a later tile, in a MUFU slot and in a polynomial slot):
    l leaves the range before the next guard -> the tile must be flagged (the kernel then reruns it with per-chunk maxima)."""
    g = torch.Generator().manual_seed(2)
    for col in (3 * KT + 5, 3 * KT + 7):                       # column 7 of a tile is a polynomial slot (every 4th pair)
        s = torch.randn(8, 5 * KT, generator=g)
        s[2, col] += jump_log2 / (1.4427 / 8)
        v = torch.randn(5 * KT, 8, generator=g)
        out, hard, _ = guarded_softmax_v(s, v, 1.4427 / 8)
        assert hard


def test_thirty_bit_jumps_stay_inside_the_range():
    """Jumps of 2^30 per tile never overflow and never need the exact pass; the reference follows with at most one tile's delay."""
    g = torch.Generator().manual_seed(3)
    S = 6 * KT
    s = torch.randn(16, S, generator=g)
    for t in range(1, 6):
        s[:, t * KT + 11] += t * 30.0 / (1.4427 / 8)
    v = torch.randn(S, 8, generator=g)
    out, hard, n = guarded_softmax_v(s, v, 1.4427 / 8)
    assert not hard
    assert float((out.double() - reference(s, v, 1.4427 / 8)).abs().max()) < 2e-2
