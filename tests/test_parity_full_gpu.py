"""Full-depth parity (run with -m gpu on a B200): every BASELINE.json config at its stated size -- 24 joint blocks, hidden
1536, 24 heads -- against the fp32 oracle executed on the same GPU with TF32 off, plus the schedule branches the reference
has besides alpha_beta / relative (modeling_sd3_pnt.py:559-576).

Two kinds of comparison, both against the oracle's own trajectory:

TEACHER FORCED (the kernel-parity statement, BASELINE.json's tolerances): at every step the CUDA path is evaluated at the
oracle's own state -- MMDiT forward at (latents, sigma) of that step through the drop-in transformer, CFG combine, the drop-in
TimePredictor on the combined hidden states -- and must give the velocity within 2e-2 rel-L2 and the next sigma within 1e-3.

CLOSED LOOP (each side follows its own latents, hidden states and predicted times): step count, masks, final latent
<= 3e-2 and the sigma sequence.  With the reference's TimePredictor init (bias dominated; what bench.py times) the closed-loop
sigma sequence agrees to ~1e-5 and is held to 1e-3.  With the head made INPUT-SENSITIVE (the scaling bench.py --workload
config3 uses to get 6-28 step trajectories: fc2 x4, fc1 x4, conv2 x2, i.e. a 32x gain on the input-dependent part of
log(alpha - 1), log(beta - 1)) the loop amplifies rounding-level differences: three attention kernels whose outputs agree to
1.8e-3 per row with each other and with fp32 (tools/attn_rowcheck.py) gave closed-loop max |dsigma| of 3.7e-4, 2.6e-3 and
5.9e-3 on the same seed, while the one-step (teacher-forced) sigma error stays below 1e-3.  The closed-loop sigma drift of the
sensitive head is therefore reported and held to 1e-2 (an accumulation bound, not a kernel tolerance), and the closed-loop
velocities -- which compare the network at two DIFFERENT sigmas -- to the matching 1e-1."""
import pytest
import torch

pytestmark = pytest.mark.gpu

VEL_TOL, SIGMA_TOL, LATENT_TOL = 2e-2, 1e-3, 3e-2
SENSITIVE_CLOSED_LOOP_SIGMA_TOL, SENSITIVE_CLOSED_LOOP_VEL_TOL = 1e-2, 1e-1      # see the module docstring


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def _no_tf32():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.set_float32_matmul_precision("highest")


def sensitive_tpm_(tp, fc2=4.0, fc1=4.0, conv2=2.0):
    """bench.py --workload config3: (alpha, beta) depend on the hidden states instead of the fc2 bias alone."""
    with torch.no_grad():
        tp.fc2.weight.mul_(fc2)
        tp.fc1.weight.mul_(fc1)
        tp.conv2.weight.mul_(conv2)


def sd3m_pipeline_pair(sample_size=128, seed=1234, sensitive=True, min_sigma=0.001, num_layers=24, **pipe_kw):
    """(oracle pipeline in fp32 on the GPU, drop-in model in bf16) holding the same bf16-representable weights."""
    from oracle import sd3_oracle as O
    from tpdm_b200.modeling_sd3_pnt import SD3_MEDIUM_TRANSFORMER_CONFIG, SD3PredictNextTimeStepModel

    _no_tf32()
    cfg = O.sd3_medium_config(sample_size=sample_size)
    cfg.num_layers = num_layers
    torch.manual_seed(seed)
    pipe = O.OraclePipeline(cfg, min_sigma=min_sigma, **pipe_kw).to("cuda")
    if sensitive:
        sensitive_tpm_(pipe.time_predictor)
    tcfg = dict(SD3_MEDIUM_TRANSFORMER_CONFIG, sample_size=sample_size, num_layers=num_layers)
    model = SD3PredictNextTimeStepModel(transformer_config=tcfg, torch_dtype=torch.bfloat16, device="cuda", min_sigma=min_sigma, **pipe_kw)
    model.transformer.load_state_dict(pipe.transformer.state_dict())
    model.time_predictor.load_state_dict(pipe.time_predictor.state_dict())
    pipe.transformer.load_state_dict(model.transformer.state_dict())            # both sides: bf16-representable values
    pipe.time_predictor.load_state_dict(model.time_predictor.state_dict())
    return pipe, model


def inputs(batch, latent, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    mk = lambda *s: torch.randn(*s, device="cuda", generator=g)
    return dict(prompt_embeds=mk(batch, 333, 4096), negative_prompt_embeds=mk(batch, 333, 4096), pooled_prompt_embeds=mk(batch, 2048),
                negative_pooled_prompt_embeds=mk(batch, 2048), latents=mk(batch, 16, latent, latent))


def compare_trajectories(out, ref, tag, sensitive_head=False):
    """Closed loop.  Per-step report first (printed with -s / on failure), then the tolerances of the module docstring."""
    sigma_tol = SENSITIVE_CLOSED_LOOP_SIGMA_TOL if sensitive_head else SIGMA_TOL
    T_ref, T = ref["sigmas"].shape[1], out.sigmas.shape[1]
    n = min(T, T_ref)
    dsig = (out.sigmas[:, :n].float().cpu() - ref["sigmas"][:, :n].float().cpu()).abs()
    vel = [rel(out["velocities"][:, t], ref["velocities"][:, t]) for t in range(n)]
    print(f"[{tag}] steps ours={T} oracle={T_ref}  max|dsigma|={float(dsig.max()):.2e}  max velocity rel-L2={max(vel):.2e}")
    print(f"[{tag}] sigma(oracle) = {[round(float(s), 4) for s in ref['sigmas'][0]]}")
    print(f"[{tag}] |dsigma| per step = {[f'{float(d):.1e}' for d in dsig.max(0).values]}")
    print(f"[{tag}] velocity rel-L2 per step = {[f'{v:.1e}' for v in vel]}")
    assert T == T_ref, f"{tag}: step count {T} != oracle {T_ref}"
    assert torch.equal(out.prob_masks.cpu(), ref["prob_masks"].cpu()), f"{tag}: prob_masks differ"
    assert float(dsig.max()) <= sigma_tol, f"{tag}: sigma drift {float(dsig.max()):.2e}"
    assert max(vel) <= (SENSITIVE_CLOSED_LOOP_VEL_TOL if sensitive_head else VEL_TOL), f"{tag}: velocity {max(vel):.2e}"
    dl = rel(out.latents, ref["final_latents"])
    print(f"[{tag}] final latent rel-L2 = {dl:.2e}")
    assert dl <= LATENT_TOL, f"{tag}: final latent {dl:.2e}"
    assert [int(i) for i in out.last_valid_indices] == ref["last_valid_indices"].tolist()
    return float(dsig.max()), max(vel), dl


def teacher_forced_steps(model, ref, kw, guidance=7.0, predict=True):
    """The kernel-parity statement: at the oracle's own inputs of every step -- latents before the step, timestep = sigma_in * 1000
    -- the drop-in CustomSD3Transformer2DModel.forward (transformer_sd3.py:299-409), the CFG combines (modeling_sd3_pnt.py:536-548) and
    the drop-in TimePredictor.forward (:100-115) must give the oracle's velocity (<= 2e-2 rel-L2) and, through the Beta mode
    (:559-576), the oracle's next sigma (<= 1e-3)."""
    from tpdm_b200.modeling_sd3_pnt import reshape_hidden_states_to_2d

    T = ref["sigmas"].shape[1]
    enc = torch.cat([kw["negative_prompt_embeds"], kw["prompt_embeds"]]).cuda()
    pooled = torch.cat([kw["negative_pooled_prompt_embeds"], kw["pooled_prompt_embeds"]]).cuda()
    cfg = lambda t: t.chunk(2)[0] + guidance * (t.chunk(2)[1] - t.chunk(2)[0])
    verr, serr = [], []
    for t in range(T):
        lat = (ref["init_noise_latents"] if t == 0 else ref["history_latents"][:, t - 1]).cuda().float()
        sig = torch.ones(lat.shape[0], device="cuda") if t == 0 else ref["sigmas"][:, t - 1].cuda().float()
        if bool((ref["prob_masks"][:, t]).all()):
            break
        v, temb, h1, h2 = (x.float() for x in model.transformer(torch.cat([lat] * 2), enc, pooled, sig.repeat(2) * 1000, return_dict=False))
        verr.append(rel(cfg(v), ref["velocities"][:, t]))
        if predict:
            g = lat.shape[-1] // 2
            hc = torch.cat([reshape_hidden_states_to_2d(cfg(h1), g, g), reshape_hidden_states_to_2d(cfg(h2), g, g)], dim=1)
            ab = model.time_predictor(hc, cfg(temb)).float()
            p1, p2 = ab[:, 0], ab[:, 1]
            if model.prediction_type == "mode_concentration":
                p1, p2 = p1 * (p2 - 2) + 1, (1 - p1) * (p2 - 2) + 1
            ratio = (p1 - 1) / (p1 + p2 - 2)
            if model.relative:
                nxt = sig * ratio.clamp(model.epsilon, 1 - model.epsilon)
            else:
                nxt = sig - torch.minimum(ratio.clamp(min=model.epsilon), sig).clamp(0, 1 - model.epsilon)
            live = ~ref["prob_masks"][:, t].cuda()
            serr.append(float(((nxt - ref["sigmas"][:, t].cuda().float()).abs() * live).max()))
    print(f"[teacher forced] velocity rel-L2 per step = {[f'{e:.1e}' for e in verr]}")
    if serr:
        print(f"[teacher forced] |sigma_next - oracle| per step = {[f'{e:.1e}' for e in serr]}")
    assert max(verr) <= VEL_TOL, f"teacher-forced velocity {max(verr):.2e}"
    assert not serr or max(serr) <= SIGMA_TOL, f"teacher-forced sigma {max(serr):.2e}"
    return max(verr), (max(serr) if serr else 0.0)


# ---------------------------------------------------------------------------------------------------------------
# config 2: SD3-medium 1024^2, batch 1, predict=True -- where sigma parity can actually fail
# ---------------------------------------------------------------------------------------------------------------
def assert_beta_parameters_close(out, ref, tag, rtol=3e-2, lp_tol=0.3):
    """alpha / beta / log-prob are not bounded by BASELINE.json; they follow the hidden states (<= 2e-2 rel-L2) through the
    input-sensitive head, so a few per cent on the Beta parameters (and the matching change of the log-density) is the bf16
    signature, not a schedule error -- the schedule itself is held to 1e-3 by compare_trajectories."""
    da = float(((out.alphas.cpu() - ref["alphas"].cpu()).abs() / ref["alphas"].cpu().abs()).max())
    db = float(((out.betas.cpu() - ref["betas"].cpu()).abs() / ref["betas"].cpu().abs()).max())
    dl = float((out.logprobs.cpu() - ref["logprobs"].cpu()).abs().max())
    print(f"[{tag}] max rel |dalpha| {da:.2e}  |dbeta| {db:.2e}  max |dlogprob| {dl:.2e}")
    assert da < rtol and db < rtol and dl < lp_tol, (da, db, dl)


@pytest.mark.parametrize("seed,init", [(0, (1.5, 0.5)), (7, (1.0, 1.2))])
def test_sd3_medium_1024_full_trajectory_sensitive_tpm_vs_oracle(seed, init):
    """(1.5, 0.5) is the reference's default head bias (slow schedule: runs into the 28-step cap); (1.0, 1.2) gives ratios near
    0.45, so the trajectory ends by itself and the step count / masks are a real comparison."""
    pipe, model = sd3m_pipeline_pair(128, init_alpha=init[0], init_beta=init[1])
    kw = inputs(1, 128, seed)
    ref = pipe(**kw, max_inference_steps=28, guidance_scale=7.0, predict=True, record_velocity=True)
    out = model(**kw, max_inference_steps=28, guidance_scale=7.0, predict=True, return_velocities=True)
    # the stress is real: alpha / beta move along the trajectory (not the bias-only constants of the reference init)
    assert float(ref["alphas"].std()) > 1e-2 or float(ref["betas"].std()) > 1e-2
    teacher_forced_steps(model, ref, kw)
    compare_trajectories(out, ref, f"cfg2 seed {seed} init {init}", sensitive_head=True)


def test_sd3_medium_1024_reference_init_trajectory_vs_oracle():
    """The same with the reference's own TimePredictor init (what bench.py times)."""
    pipe, model = sd3m_pipeline_pair(128, sensitive=False)
    kw = inputs(1, 128, 3)
    ref = pipe(**kw, max_inference_steps=28, predict=True, record_velocity=True)
    out = model(**kw, max_inference_steps=28, predict=True, return_velocities=True)
    teacher_forced_steps(model, ref, kw)
    compare_trajectories(out, ref, "cfg2 reference init")
    assert_beta_parameters_close(out, ref, "cfg2 reference init", rtol=2e-3, lp_tol=5e-3)


# ---------------------------------------------------------------------------------------------------------------
# config 4: 512^2 RLOO rollout, 4 prompts x 4 samples = batch 16 (transformer batch 32), injected Beta draws
# ---------------------------------------------------------------------------------------------------------------
def test_sd3_medium_512_rloo_rollout_batch16_vs_oracle():
    pipe, model = sd3m_pipeline_pair(64, seed=4321, min_sigma=0.01)
    P, k, T = 4, 4, 14
    base = inputs(P, 64, 11)
    # rloo_repeat tiles the prompt list k times (modeling_sd3_pnt.py:776); every rollout starts from its own noise
    kw = {n: t.repeat(k, *([1] * (t.dim() - 1))) for n, t in base.items() if n != "latents"}
    kw["latents"] = torch.randn(P * k, 16, 64, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(12))
    torch.manual_seed(13)     # the Beta draws the reference makes on the host (:569), injected into both sides
    ratios = torch.distributions.Beta(torch.tensor(5.0), torch.tensor(2.5)).sample((P * k, T)).clamp(0.05, 0.95).cuda()
    ref = pipe(**kw, max_inference_steps=T, predict=False, ratios=ratios, record_velocity=True)
    out = model(**kw, max_inference_steps=T, predict=False, ratios=ratios, return_velocities=True)
    teacher_forced_steps(model, ref, kw, predict=False)
    compare_trajectories(out, ref, "cfg4 512^2 x16")          # injected ratios: sigma cannot drift, the BASELINE tolerances apply
    assert_beta_parameters_close(out, ref, "cfg4 512^2 x16")
    # the recorded TimePredictor inputs replay to the rollout's own log-probs (what the PPO update starts from)
    lp = model.only_predict_logprobs(out.sigmas, out.hidden_states_combineds, out.tembs)["logprobs"]
    assert float((lp.detach() - out.logprobs).abs().max()) < 5e-3


# ---------------------------------------------------------------------------------------------------------------
# config 5: one 24-layer forward at 2048^2 (S = 16 717)
# ---------------------------------------------------------------------------------------------------------------
def test_sd3_medium_2048_full_depth_forward_vs_oracle():
    from oracle import sd3_oracle as O
    from tpdm_b200.transformer_sd3 import CustomSD3Transformer2DModel

    _no_tf32()
    cfg = O.sd3_medium_config(sample_size=256)
    torch.manual_seed(4321)
    ora = O.OracleSD3Transformer(cfg).requires_grad_(False).eval().to("cuda")
    model = CustomSD3Transformer2DModel(sample_size=256, num_layers=24, attention_head_dim=64, num_attention_heads=24,
                                        caption_projection_dim=1536, pos_embed_max_size=192, device="cuda", dtype=torch.bfloat16)
    model.load_state_dict(ora.state_dict())
    ora.load_state_dict(model.state_dict())
    g = torch.Generator(device="cuda").manual_seed(2)
    lat = torch.randn(1, 16, 256, 256, device="cuda", generator=g).repeat(2, 1, 1, 1)
    enc = torch.randn(2, 333, 4096, device="cuda", generator=g)
    pooled = torch.randn(2, 2048, device="cuda", generator=g)
    ts = torch.tensor([400.0, 400.0], device="cuda")
    with torch.no_grad():
        rv, rt, rh1, rh2 = ora(lat, enc, pooled, ts)
    v, temb, h1, h2 = model(lat, enc, pooled, ts, return_dict=False)
    print(f"[cfg5 2048^2 x24] velocity rel-L2 {rel(v, rv):.2e}  h2 {rel(h2, rh2):.2e}  h1 {rel(h1, rh1):.2e}")
    assert rel(temb, rt) < 2e-3 and rel(h1, rh1) < 1e-3
    assert rel(h2, rh2) < VEL_TOL and rel(v, rv) < VEL_TOL


# ---------------------------------------------------------------------------------------------------------------
# schedule branches (modeling_sd3_pnt.py:559-576) on the tiny config: mode_concentration, relative=False
# ---------------------------------------------------------------------------------------------------------------
def _tiny_pair(init_alpha, init_beta, tpm_epsilon=1.0, **kw):
    from oracle import sd3_oracle as O
    from tpdm_b200.modeling_sd3_pnt import SD3PredictNextTimeStepModel

    cfg = O.tiny_config()
    pipe = O.build_pipeline(cfg, init_alpha=init_alpha, init_beta=init_beta, **kw)
    pipe.time_predictor.epsilon = tpm_epsilon
    sensitive_tpm_(pipe.time_predictor, 1.5, 1.5, 1.2)
    tcfg = dict(sample_size=32, num_layers=2, attention_head_dim=96, num_attention_heads=4, caption_projection_dim=384,
                pos_embed_max_size=96)
    model = SD3PredictNextTimeStepModel(transformer_config=tcfg, torch_dtype=torch.float32, device="cuda", init_alpha=init_alpha,
                                        init_beta=init_beta, **kw)
    model.time_predictor.epsilon = tpm_epsilon
    model.transformer.load_state_dict(pipe.transformer.state_dict())
    model.time_predictor.load_state_dict(pipe.time_predictor.state_dict())
    inp = O.synthetic_inputs(cfg, batch=2)
    return pipe, model, inp


@pytest.mark.parametrize("predict", [True, False])
def test_schedule_mode_concentration_vs_oracle(predict):
    """prediction_type='mode_concentration' (:559-563): the head's outputs are (mode, concentration) and
    alpha = p1 (p2 - 2) + 1, beta = (1 - p1)(p2 - 2) + 1.  A valid Beta needs a mode inside (0, 1), i.e. a head without the
    '+ 1' offset: TimePredictor.epsilon = 0 with fc2 biases (ln 0.7, ln 20)."""
    import math

    pipe, model, inp = _tiny_pair(math.log(0.7), math.log(20.0), tpm_epsilon=0.0, prediction_type="mode_concentration")
    cu = {k: v.cuda() for k, v in inp.items()}
    T = 8
    ratios = None
    if not predict:
        ratios = torch.distributions.Beta(torch.tensor(6.0), torch.tensor(3.0)).sample((2, T))
    ref = pipe(**inp, max_inference_steps=T, predict=predict, ratios=ratios, record_velocity=True)
    out = model(**cu, max_inference_steps=T, predict=predict, ratios=None if ratios is None else ratios.cuda(), return_velocities=True)
    # precondition of the branch: the oracle's Beta parameters are valid (mode inside (0, 1), concentration > 2)
    assert bool((ref["alphas"] > 1).all()) and bool((ref["betas"] > 1).all()), (ref["alphas"], ref["betas"])
    teacher_forced_steps(model, ref, inp, predict=predict)
    compare_trajectories(out, ref, f"mode_concentration predict={predict}", sensitive_head=predict)
    assert_beta_parameters_close(out, ref, f"mode_concentration predict={predict}")
    # alpha/beta are the TRANSFORMED parameters: their mode is the head's first output
    mode = (out.alphas - 1) / (out.alphas + out.betas - 2)
    assert bool(((mode > 0) & (mode < 1)).all())
    assert float(ref["alphas"].std()) > 1e-3          # and they move with the hidden states
    if not predict:   # the replay applies the same transform (:697-698 has the raw outputs; the rollout's are what PPO compares with)
        lp = model.only_predict_logprobs(out.sigmas, out.hidden_states_combineds, out.tembs)["logprobs"]
        assert float((lp.detach() - out.logprobs).abs().max()) < 5e-3


@pytest.mark.parametrize("predict", [True, False])
def test_schedule_absolute_step_vs_oracle(predict):
    """relative=False (:573-576): ratio is clamped to [eps, sigma] then [0, 1 - eps] and SUBTRACTED, sigma_next = sigma - ratio."""
    import math

    pipe, model, inp = _tiny_pair(0.0, math.log(6.0), relative=False)          # Beta(2, 7): mode 1/7 per step
    cu = {k: v.cuda() for k, v in inp.items()}
    T = 10
    ratios = None
    if not predict:
        ratios = torch.distributions.Beta(torch.tensor(2.0), torch.tensor(7.0)).sample((2, T))
    ref = pipe(**inp, max_inference_steps=T, predict=predict, ratios=ratios, record_velocity=True)
    out = model(**cu, max_inference_steps=T, predict=predict, ratios=None if ratios is None else ratios.cuda(), return_velocities=True)
    assert ref["sigmas"].shape[1] >= 5, ref["sigmas"]           # several subtractive steps before sigma reaches zero
    teacher_forced_steps(model, ref, inp, predict=predict)
    compare_trajectories(out, ref, f"relative=False predict={predict}", sensitive_head=predict)
    assert_beta_parameters_close(out, ref, f"relative=False predict={predict}")
    sig = torch.cat([torch.ones(2, 1), out.sigmas.cpu()], dim=1)
    assert bool((sig[:, 1:] <= sig[:, :-1] + 1e-6).all()) and bool((sig >= -1e-6).all())    # steps subtract and never cross zero
    if not predict:
        lp = model.only_predict_logprobs(out.sigmas, out.hidden_states_combineds, out.tembs)["logprobs"]
        assert float((lp.detach() - out.logprobs).abs().max()) < 5e-3
