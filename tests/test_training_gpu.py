"""GPU tests of the training half (rows R1-R3): native TimePredictor backward vs autograd on the oracle, PPO-clip loss vs
the restated trainer math, fused clip + AdamW vs torch.optim.AdamW, and one RLOO update end to end."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _pair(in_channels, g, ns, seed=0, sensitive=True):
    from oracle import sd3_oracle as O
    from tpdm_b200.modeling_sd3_pnt import TimePredictor
    from tpdm_b200.tpm_training import TimePredictorTrainer

    torch.manual_seed(seed)
    ora = O.OracleTimePredictor(128, in_channels).cuda()
    with torch.no_grad():
        if sensitive:   # the reference init is bias dominated; make the head depend on its input
            ora.fc2.weight.mul_(15)
            ora.fc1.weight.mul_(6)
            ora.conv2.weight.mul_(4)
            ora.norm1.linear.weight.mul_(3)
        ora.conv1.weight.copy_(ora.conv1.weight.bfloat16().float())   # tensor cores read conv1 in bf16
    tp = TimePredictor(128, in_channels, device="cuda", dtype=torch.float32)
    tp.load_state_dict(ora.state_dict())
    tr = TimePredictorTrainer(tp, grid=g, max_samples=ns)
    x = torch.randn(ns, in_channels, g, g, device="cuda").bfloat16().float()
    temb = torch.randn(ns, in_channels // 2, device="cuda")
    return ora, tp, tr, x, temb


@pytest.mark.parametrize("in_channels,g,ns", [(256, 16, 3), (768, 16, 4), (3072, 32, 2)])
def test_tpm_backward_matches_autograd(in_channels, g, ns):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    ora, tp, tr, x, temb = _pair(in_channels, g, ns)
    import torch.nn.functional as F

    # OracleTimePredictor.forward up to the fc2 output z (alpha, beta = exp(z) + 1), so that d loss / d z is exactly w
    h = F.silu(ora.norm1(ora.conv1(x), temb))
    h = F.adaptive_max_pool2d(F.adaptive_avg_pool2d(ora.conv2(h), (16, 16)), (1, 1)).view(ns, -1)
    z = ora.fc2(F.silu(ora.fc1(h)))
    y = torch.exp(z) + 1.0
    w = torch.tensor([[0.7, -1.3]], device="cuda") * torch.arange(1, ns + 1, device="cuda")[:, None]
    (z * w).sum().backward()
    ab = tr.forward(x.permute(0, 2, 3, 1).contiguous().bfloat16(), temb)
    assert torch.allclose(ab, y.detach(), rtol=3e-3)
    tr.backward(w.expand(ns, 2).contiguous())
    torch.cuda.synchronize()
    got = tr.grad_dict()
    for name, p in ora.named_parameters():
        tol = 1.5e-2 if name in ("conv1.weight", "conv1.bias") else 5e-3   # dY is rounded to bf16 for the wgrad GEMM
        assert rel(got[name], p.grad) < tol, (name, rel(got[name], p.grad))


def test_ppo_clip_loss_and_dz_match_restated_trainer_math():
    from oracle import sd3_oracle as O
    from tpdm_b200 import _lib as L

    lib = L.load()
    torch.manual_seed(1)
    mb, T = 6, 5
    z = (torch.randn(mb * T, 2, device="cuda", dtype=torch.float64) * 0.5 + 1.0).requires_grad_(True)
    ab = torch.exp(z) + 1.0
    ratios = torch.rand(mb, T, device="cuda", dtype=torch.float64) * 0.5 + 0.3
    ratios[2, 3:] = 0.01                                   # sample 2 falls below min_sigma -> masked tail
    sig = torch.cumprod(ratios, dim=1)
    min_sigma = 0.01
    prev = torch.cat([torch.ones(mb, 1, device="cuda", dtype=torch.float64), sig[:, :-1]], 1)
    mask = prev < min_sigma
    r = torch.clamp(sig / prev, 1e-3, 1 - 1e-3)
    a, b = ab[:, 0].reshape(mb, T), ab[:, 1].reshape(mb, T)
    new_lp = torch.where(mask, torch.ones_like(r), O.beta_log_prob(a, b, r))
    old_lp = (new_lp.detach() + torch.randn(mb, T, device="cuda", dtype=torch.float64) * 0.2).masked_fill(mask, 1.0)
    adv = torch.randn(mb, device="cuda", dtype=torch.float64)
    loss = O.ppo_clip_loss(new_lp, old_lp, adv, 0.2)
    loss.backward()
    f = lambda t: t.detach().float().contiguous()
    out_lp, dz, stats = torch.empty(mb, T, device="cuda"), torch.empty(mb * T, 2, device="cuda"), torch.empty(4, device="cuda")
    ab32, sig32, old32, adv32 = f(ab), f(sig), f(old_lp), f(adv)      # keep the fp32 copies alive across the async launch
    tail = torch.full((2,), 7.0, device="cuda")
    L.check(lib.tpdm_ppo_clip_loss(L.ptr(ab32), L.ptr(sig32), L.ptr(old32), L.ptr(adv32), mb, T, min_sigma, 1e-3, 1, 0, 0.2, 1.0,
                                   L.ptr(out_lp), L.ptr(dz), L.ptr(stats), L.ptr(tail), None))
    torch.cuda.synchronize()
    assert float(tail[0]) == float(stats[0]) and float(tail[1]) == 0.0      # {loss, non-finite flag} for the all-reduce tail
    assert torch.allclose(out_lp.double(), new_lp.detach(), atol=2e-4)
    assert abs(float(stats[0]) - float(loss)) < 2e-4 * max(1.0, abs(float(loss)))
    assert rel(dz, z.grad) < 2e-3
    assert float(dz.reshape(mb, T, 2)[2, 4:].abs().max()) == 0.0     # masked steps carry no gradient


def test_adamw_step_matches_torch():
    from tpdm_b200 import _lib as L

    lib = L.load()
    torch.manual_seed(2)
    n = 100_003
    p0 = torch.randn(n, device="cuda")
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.AdamW([ref_p], lr=1e-3, betas=(0.9, 0.99), eps=1e-5, weight_decay=0.01)
    p, m, v = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    sumsq = torch.zeros(1, device="cuda", dtype=torch.float64)
    copy = torch.zeros(1000, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        g = torch.randn(n, device="cuda") * (3.0 if step == 2 else 0.001)
        ref_p.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([ref_p], 1.0)
        opt.step()
        L.check(lib.tpdm_adamw_step(L.ptr(p), L.ptr(g), L.ptr(m), L.ptr(v), n, 1e-3, 0.9, 0.99, 1e-5, 0.01, 1.0, step, 1.0, L.ptr(sumsq),
                                    L.ptr(copy), 1000, None, None))
        torch.cuda.synchronize()
        assert abs(float(sumsq.sqrt()) - float(g.norm())) < 1e-3 * float(g.norm())
        assert torch.allclose(p, ref_p.detach(), rtol=1e-5, atol=1e-6)
    assert torch.equal(copy, p[:1000].bfloat16())
    bad = torch.full((n,), float("nan"), device="cuda")      # NaN gradient: update skipped (rloo_trainer.py:518-520)
    before = p.clone()
    L.check(lib.tpdm_adamw_step(L.ptr(p), L.ptr(bad), L.ptr(m), L.ptr(v), n, 1e-3, 0.9, 0.99, 1e-5, 0.01, 1.0, 4, 1.0, L.ptr(sumsq), None, 0, None, None))
    torch.cuda.synchronize()
    assert torch.equal(p, before)
    # a raised non-finite-loss flag (summed over ranks by the gradient all-reduce) skips the step as well (rloo_trainer.py:497-500)
    good, flag = torch.randn(n, device="cuda") * 0.001, torch.tensor([2.0], device="cuda")
    L.check(lib.tpdm_adamw_step(L.ptr(p), L.ptr(good), L.ptr(m), L.ptr(v), n, 1e-3, 0.9, 0.99, 1e-5, 0.01, 1.0, 4, 1.0, L.ptr(sumsq), None, 0,
                                L.ptr(flag), None))
    torch.cuda.synchronize()
    assert torch.equal(p, before)
    flag.zero_()
    L.check(lib.tpdm_adamw_step(L.ptr(p), L.ptr(good), L.ptr(m), L.ptr(v), n, 1e-3, 0.9, 0.99, 1e-5, 0.01, 1.0, 4, 1.0, L.ptr(sumsq), None, 0,
                                L.ptr(flag), None))
    torch.cuda.synchronize()
    assert not torch.equal(p, before)


@pytest.mark.parametrize("prediction_type,relative", [("alpha_beta", True), ("mode_concentration", True), ("alpha_beta", False)])
def test_beta_logprob_and_dz_match_autograd(prediction_type, relative):
    """tpdm_beta_logprob / the PPO kernel against autograd through the restated log-prob, incl. the mode_concentration
    transform (modeling_sd3_pnt.py:559-563) the rollout applies: replay and PPO gradient must score the SAME distribution."""
    from oracle import sd3_oracle as O
    from tpdm_b200 import _lib as L

    lib = L.load()
    torch.manual_seed(4)
    mb, T, min_sigma, eps = 5, 6, 0.02, 1e-3
    tpm_eps = 1.0 if prediction_type == "alpha_beta" else 0.0
    z = torch.randn(mb * T, 2, device="cuda", dtype=torch.float64) * 0.3
    if prediction_type == "mode_concentration":
        z[:, 0] = z[:, 0] * 0.5 - 0.6          # mode = exp(z0) in (0, 1)
        z[:, 1] = z[:, 1] + 2.5                # concentration = exp(z1) > 2
    z.requires_grad_(True)
    p = torch.exp(z) + tpm_eps
    p1, p2 = p[:, 0].reshape(mb, T), p[:, 1].reshape(mb, T)
    a, b = (p1, p2) if prediction_type == "alpha_beta" else (p1 * (p2 - 2) + 1, (1 - p1) * (p2 - 2) + 1)
    if relative:
        sig = torch.cumprod(torch.rand(mb, T, device="cuda", dtype=torch.float64) * 0.4 + 0.3, dim=1)
    else:
        sig = 1.0 - torch.cumsum(torch.rand(mb, T, device="cuda", dtype=torch.float64) * 0.25, dim=1).clamp(max=0.999)
    prev = torch.cat([torch.ones(mb, 1, device="cuda", dtype=torch.float64), sig[:, :-1]], 1)
    mask = prev < min_sigma
    r = torch.clamp(sig / prev if relative else prev - sig, eps, 1 - eps)
    lp_ref = torch.where(mask, torch.ones_like(r), O.beta_log_prob(a, b, r))
    w = torch.randn(mb, T, device="cuda", dtype=torch.float64)
    (lp_ref * w).sum().backward(retain_graph=True)
    f = lambda t: t.detach().float().contiguous()
    ab32, sig32 = f(p), f(sig)
    lp, dlp = torch.empty(mb, T, device="cuda"), torch.empty(mb * T, 2, device="cuda")
    L.check(lib.tpdm_beta_logprob(L.ptr(ab32), L.ptr(sig32), mb, T, min_sigma, eps, int(relative), 0 if prediction_type == "alpha_beta" else 1,
                                  tpm_eps, L.ptr(lp), L.ptr(dlp), None))
    torch.cuda.synchronize()
    assert torch.allclose(lp.double(), lp_ref.detach(), atol=3e-4, rtol=1e-5)
    assert rel(dlp * f(w).reshape(-1, 1), z.grad) < 2e-3
    # the PPO kernel applies the same transform (it used the raw head outputs)
    z.grad = None
    old_lp = (lp_ref.detach() + 0.1 * torch.randn(mb, T, device="cuda", dtype=torch.float64)).masked_fill(mask, 1.0)
    adv = torch.randn(mb, device="cuda", dtype=torch.float64)
    loss = O.ppo_clip_loss(lp_ref, old_lp, adv, 0.2)
    loss.backward()
    out_lp, dz, stats = torch.empty(mb, T, device="cuda"), torch.empty(mb * T, 2, device="cuda"), torch.empty(4, device="cuda")
    old32, adv32 = f(old_lp), f(adv)
    L.check(lib.tpdm_ppo_clip_loss(L.ptr(ab32), L.ptr(sig32), L.ptr(old32), L.ptr(adv32), mb, T, min_sigma, eps, int(relative),
                                   0 if prediction_type == "alpha_beta" else 1, 0.2, tpm_eps, L.ptr(out_lp), L.ptr(dz), L.ptr(stats), None, None))
    torch.cuda.synchronize()
    assert abs(float(stats[0]) - float(loss)) < 3e-4 * max(1.0, abs(float(loss)))
    assert rel(dz, z.grad) < 2e-3


def test_wrapper_logprobs_is_differentiable_like_the_reference():
    """rloo_trainer.py:485-501 on the drop-in: new_logprobs = model.logprobs(...); loss.backward() deposits time_predictor.*.grad.
    The gradients equal autograd through the fp32 oracle TimePredictor on the same recorded inputs."""
    from oracle import sd3_oracle as O
    from tpdm_b200.modeling_sd3_pnt import SD3PredictNextTimeStepModelRLOOWrapper

    tiny = dict(sample_size=32, patch_size=2, in_channels=16, num_layers=2, attention_head_dim=96, num_attention_heads=4,
                joint_attention_dim=4096, caption_projection_dim=384, pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=96)
    torch.manual_seed(31)
    wrapper = SD3PredictNextTimeStepModelRLOOWrapper(transformer_config=tiny, torch_dtype=torch.float32, device="cuda", min_sigma=0.05,
                                                     max_inference_steps=6)
    tp = wrapper.agent_model.time_predictor
    with torch.no_grad():
        tp.fc2.weight.mul_(10)
        tp.fc1.weight.mul_(4)
        tp.conv1.weight.copy_(tp.conv1.weight.bfloat16().float())        # tensor cores read conv1 in bf16
    assert all(p.requires_grad for p in tp.parameters())
    g = torch.Generator().manual_seed(5)
    mk = lambda *s: torch.randn(*s, generator=g).cuda()
    data = dict(prompt_embeds=mk(3, 333, 4096), negative_prompt_embeds=mk(3, 333, 4096), pooled_prompt_embeds=mk(3, 2048),
                negative_pooled_prompt_embeds=mk(3, 2048), latents=mk(3, 16, 32, 32), predict=False, generator=torch.Generator().manual_seed(9))
    outputs = wrapper.sample(dict(data))
    old = outputs["logprobs"]
    assert not old.requires_grad and outputs["hidden_states_combineds"].shape[:2] == old.shape
    # micro-batch subset exactly as the trainer does (:480-485)
    idx = torch.tensor([2, 0], device="cuda")
    mb_out = wrapper.subset_outputs(outputs, idx)
    new = wrapper.logprobs(None, mb_out)
    assert new.requires_grad and new.shape == (2, old.shape[1])
    assert float((new.detach() - old[idx]).abs().max()) < 5e-3          # same parameters: the replay reproduces the rollout
    adv = torch.tensor([0.7, -1.3], device="cuda")
    ratio = torch.exp(new.sum(1) - old[idx].sum(1))
    loss = torch.max(-adv * ratio, -adv * torch.clamp(ratio, 0.8, 1.2)).mean()
    loss.backward()
    grads = {n: p.grad.clone() for n, p in tp.named_parameters()}
    assert all(gr is not None and torch.isfinite(gr).all() for gr in grads.values())
    # the same through autograd on the oracle head
    ora = O.OracleTimePredictor(128, 768).cuda()
    ora.load_state_dict(tp.state_dict())
    pipe = O.OraclePipeline(O.tiny_config(), min_sigma=0.05)
    pipe.time_predictor = ora                                            # trainable copy of the drop-in's head
    ref_new = pipe.only_predict_logprobs(mb_out["sigmas"].float(), mb_out["hidden_states_combineds"].float(), mb_out["tembs"].float())["logprobs"]
    ref_ratio = torch.exp(ref_new.sum(1) - old[idx].sum(1))
    torch.max(-adv * ref_ratio, -adv * torch.clamp(ref_ratio, 0.8, 1.2)).mean().backward()
    for name, p in ora.named_parameters():
        tol = 2e-2 if name.startswith("conv1") else 1e-2
        assert rel(grads[name], p.grad) < tol, (name, rel(grads[name], p.grad))
    # gradient accumulation over a second micro-batch adds up (grads are deposited, not overwritten)
    new2 = wrapper.logprobs(None, wrapper.subset_outputs(outputs, torch.tensor([1], device="cuda")))
    new2.sum().backward()
    assert rel(tp.fc2.bias.grad, grads["fc2.bias"]) > 1e-3
    # torch.optim on the module parameters moves the native path (the packed copies are refreshed from the module)
    opt = torch.optim.AdamW(tp.parameters(), lr=1e-2)
    opt.step()
    new3 = wrapper.logprobs(None, mb_out)
    assert float((new3.detach() - new.detach()).abs().max()) > 1e-4
    # a stale backward is refused instead of using overwritten activations
    a = wrapper.logprobs(None, mb_out)
    wrapper.logprobs(None, mb_out)
    with pytest.raises(RuntimeError):
        a.sum().backward()


def test_rloo_update_end_to_end_tiny():
    """BASELINE config 4 in miniature: 2 prompts x rloo_k 2 rollouts with device-side Beta draws, synthetic reward, PPO epochs;
    the TimePredictor must change, the frozen MMDiT must not, and sampling must keep working with the updated head."""
    from tpdm_b200.modeling_sd3_pnt import SD3PredictNextTimeStepModelRLOOWrapper
    from tpdm_b200.rloo import rloo_update
    from tpdm_b200.tpm_training import TimePredictorTrainer

    torch.manual_seed(3)
    tcfg = dict(sample_size=32, num_layers=2, attention_head_dim=96, num_attention_heads=4, caption_projection_dim=384, pos_embed_max_size=96)
    w = SD3PredictNextTimeStepModelRLOOWrapper(transformer_config=tcfg, torch_dtype=torch.float32, device="cuda", min_sigma=0.01,
                                               max_inference_steps=6)
    assert all(p.requires_grad for p in w.agent_model.time_predictor.parameters())
    assert not any(p.requires_grad for p in w.agent_model.transformer.parameters())
    g = torch.Generator().manual_seed(0)
    data = dict(prompt=["a", "b"], prompt_embeds=torch.randn(2, 333, 4096, generator=g).cuda(),
                negative_prompt_embeds=torch.randn(2, 333, 4096, generator=g).cuda(),
                pooled_prompt_embeds=torch.randn(2, 2048, generator=g).cuda(),
                negative_pooled_prompt_embeds=torch.randn(2, 2048, generator=g).cuda())
    trainer = TimePredictorTrainer(w.agent_model.time_predictor, grid=16, max_samples=4 * 6, lr=1e-3)
    before = {k: v.clone() for k, v in w.agent_model.time_predictor.state_dict().items()}
    t_before = w.agent_model.transformer.proj_out.weight.clone()
    res = rloo_update(w, trainer, data, reward_fn=lambda lat, out: -(lat.float() ** 2).mean(dim=(1, 2, 3)), rloo_k=2, num_ppo_epochs=2,
                      micro_batch_size=2)
    assert len(res["logs"]) == 4 and all(torch.isfinite(torch.tensor(l["loss"])) for l in res["logs"])
    assert abs(float(res["advantages"].reshape(2, -1).sum(0).abs().max())) < 1e-5        # RLOO advantages cancel per prompt (k = 2)
    assert abs(res["logs"][0]["ratio"] - 1.0) < 5e-3                                      # first replay reproduces the rollout log-probs
    after = w.agent_model.time_predictor.state_dict()
    assert any(not torch.equal(before[k], after[k]) for k in before)
    assert torch.equal(t_before, w.agent_model.transformer.proj_out.weight)
    out = w.sample(dict(data, predict=True))
    assert torch.isfinite(out["latents"]).all()


def test_nonfinite_loss_rides_the_gradient_buffer_and_skips_the_step():
    """rloo_trainer.py:497-500 / 516-523: a NaN / Inf loss must not reach the optimizer.  The flag is the element right behind the
    gradients in the ONE buffer that is all-reduced, and the fused AdamW reads it on the device."""
    ora, tp, tr, x, temb = _pair(256, 16, 6)
    mb, T = 2, 3
    xn = x.permute(0, 2, 3, 1).contiguous().bfloat16().reshape(mb, T, 16, 16, 256)
    tm = temb.reshape(mb, T, 128)
    sig = torch.tensor([[0.7, 0.5, 0.3], [0.8, 0.6, 0.2]], device="cuda")
    old = torch.zeros(mb, T, device="cuda")
    assert tr.reduce_buf.data_ptr() == tr.grads.data_ptr() and tr.tail.data_ptr() == tr.grads.data_ptr() + 4 * tr.grads.numel()
    before = tr.params.clone()
    good = tr.ppo_update(sig, old, xn, tm, torch.tensor([0.5, -0.5], device="cuda"), min_sigma=0.01)
    assert float(good["nonfinite"]) == 0.0 and torch.isfinite(good["loss"]) and not torch.equal(tr.params, before)
    assert float(good["loss"]) == float(good["local_loss"])                    # world size 1: the tail carries this rank's loss
    mid = tr.params.clone()
    bad = tr.ppo_update(sig, old, xn, tm, torch.tensor([float("nan"), 1.0], device="cuda"), min_sigma=0.01)
    assert float(bad["nonfinite"]) == 1.0 and float(bad["loss"]) == 0.0       # the NaN itself is kept out of the sum
    assert torch.equal(tr.params, mid)                                         # update skipped on the device
    again = tr.ppo_update(sig, old, xn, tm, torch.tensor([0.5, -0.5], device="cuda"), min_sigma=0.01)
    assert float(again["nonfinite"]) == 0.0 and not torch.equal(tr.params, mid)


def test_successive_rollouts_draw_independent_schedules():
    """the Beta draws of predict=False must consume generator state like beta_dist.sample() (:569) does."""
    from tpdm_b200.modeling_sd3_pnt import SD3PredictNextTimeStepModel

    tcfg = dict(sample_size=32, num_layers=2, attention_head_dim=96, num_attention_heads=4, caption_projection_dim=384, pos_embed_max_size=96)
    torch.manual_seed(5)
    model = SD3PredictNextTimeStepModel(transformer_config=tcfg, torch_dtype=torch.float32, device="cuda")
    g = torch.Generator().manual_seed(0)
    kw = dict(prompt_embeds=torch.randn(2, 333, 4096, generator=g).cuda(), negative_prompt_embeds=torch.randn(2, 333, 4096, generator=g).cuda(),
              pooled_prompt_embeds=torch.randn(2, 2048, generator=g).cuda(), negative_pooled_prompt_embeds=torch.randn(2, 2048, generator=g).cuda(),
              latents=torch.randn(2, 16, 32, 32, generator=g).cuda(), max_inference_steps=5, predict=False)
    gen = torch.Generator().manual_seed(42)
    a = model(**kw, generator=gen)
    b = model(**kw, generator=gen)                       # same generator, second call: new draws
    assert not torch.equal(a.sigmas, b.sigmas)
    c = model(**kw, generator=torch.Generator().manual_seed(42))
    assert torch.equal(a.sigmas, c.sigmas)               # same generator state: same draws


@pytest.mark.parametrize("relative,mean_kl", [(True, False), (False, True)])
def test_rollout_shaping_matches_reference_math(relative, mean_kl):
    """Next-row (f)2: KL to the reference schedule, discounted score, rlhf reward and RLOO advantage on the device vs the
    reference's own get_ref_beta (extracted, golden-pinned in test_oracle) + torch.distributions + the restated loops."""
    from oracle import sd3_oracle as O
    from tpdm_b200.rloo import shape_rollout

    g = torch.Generator().manual_seed(11)
    k, prompts, T = 4, 3, 9
    B = k * prompts
    alphas = 1.0 + 6.0 * torch.rand(B, T, generator=g)
    betas = 1.0 + 6.0 * torch.rand(B, T, generator=g)
    ratios = 0.6 + 0.35 * torch.rand(B, T, generator=g)      # keeps sigma above the range where the reference Beta degenerates
    sigmas = torch.cumprod(ratios, 1)                       # sigma_next after each step
    lengths = torch.randint(1, T + 1, (B,), generator=g)
    masks = torch.arange(T)[None, :] >= lengths[:, None]    # True = step not executed
    last = torch.randn(B, generator=g)
    gamma, kl_coef = 0.97, 0.05
    out = shape_rollout(dict(alphas=alphas.cuda(), betas=betas.cuda(), sigmas=sigmas.cuda(), prob_masks=masks.cuda()), last.cuda(),
                        relative=relative, gamma=gamma, kl_coef=kl_coef, mean_kl=mean_kl, rloo_k=k)
    sig_in = torch.nn.functional.pad(sigmas[:, :-1], (1, 0), value=1.0)
    if relative:
        ra, rb = O.get_ref_beta(sig_in)
    else:
        ra, rb = torch.full_like(alphas, 1.4), torch.full_like(alphas, 11.2)
    kl = torch.distributions.kl_divergence(torch.distributions.Beta(alphas.double(), betas.double()),
                                           torch.distributions.Beta(ra.double(), rb.double())).float()
    kl = torch.where(masks, torch.zeros_like(kl), kl)
    scores = torch.tensor([O.discounted_reward(float(last[i]), int(lengths[i]) - 1, gamma) for i in range(B)])
    non_score = (-kl_coef * kl).mean(1) if mean_kl else (-kl_coef * kl).sum(1)
    rlhf = scores + non_score
    adv = O.rloo_advantage(rlhf, k)
    assert torch.allclose(out["kl"].cpu(), kl, rtol=2e-4, atol=2e-5)
    assert torch.allclose(out["scores"].cpu(), scores, rtol=1e-5, atol=1e-6)
    assert torch.allclose(out["rlhf_reward"].cpu(), rlhf, rtol=1e-4, atol=1e-5)
    assert torch.allclose(out["advantages"].cpu(), adv, rtol=1e-4, atol=2e-5)
    # the same KL through the reference's own closed form (train_utilis.py:6-20, argument order as the reference passes it)
    kl2 = O.get_kl_beta(betas.double(), alphas.double(), rb.double(), ra.double()).float()
    assert torch.allclose(torch.where(masks, torch.zeros_like(kl2), kl2), kl, rtol=1e-4, atol=1e-5)
