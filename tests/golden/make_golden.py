"""Generates the golden fixtures in this directory by running the REFERENCE's own code.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It loads ``/root/reference/vall_e/vall_e/{base,nar,ar_discrete}.py`` by file path behind a
shim package (the reference package itself cannot be imported: ``vall_e/config.py:96`` needs
omegaconf, and ``ar_discrete.py:14,16`` import diffusers/timm, which are stubbed), then

* drives the reference D3PM methods (``ar_discrete.py``: cosine_beta_schedule,
  _get_absorbing_transition_mat, create_transition_matrix, _at, _at_onehot, q_probs,
  q_posterior_logits, and the bodies of q_sample/p_sample with ``torch.rand`` monkeypatched so
  the uniforms are the ones stored) unbound on a namespace object, because the reference
  constructor hard-codes ``.to("cuda:0")`` (``ar_discrete.py:269,275,277``);
* runs the reference ``Base`` stack (``base.py``) in NAR configuration on a small model.

Outputs (small, committed): d3pm_absorbing_k1025.npz, d3pm_uniform_k1025.npz,
d3pm_small_*.npz, denoiser_nar_small.npz, denoiser_diffusion_small.npz, ar_discrete_dit.npz
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import detrand  # noqa: E402

REF = Path("/root/reference/vall_e/vall_e")
OUT = Path(__file__).resolve().parent


class TimmLikeMlp(torch.nn.Module):
    """Stand-in for ``timm.models.vision_transformer.Mlp`` (timm is not installed here): the same
    sub-module names and order — fc1, act, drop1, norm (Identity), fc2, drop2 — and forward."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=torch.nn.GELU, drop=0.0):
        super().__init__()
        self.fc1 = torch.nn.Linear(in_features, hidden_features or in_features)
        self.act = act_layer()
        self.drop1 = torch.nn.Dropout(drop)
        self.norm = torch.nn.Identity()
        self.fc2 = torch.nn.Linear(hidden_features or in_features, out_features or in_features)
        self.drop2 = torch.nn.Dropout(drop)

    def forward(self, x):
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


def load_reference():
    diffusers = types.ModuleType("diffusers")
    for n in ("UNet3DConditionModel", "UNet2DConditionModel", "DDPMScheduler",
              "CosineDPMSolverMultistepScheduler", "DDIMScheduler"):
        setattr(diffusers, n, object)
    tv = types.ModuleType("timm.models.vision_transformer")
    for n in ("PatchEmbed", "Attention"):
        setattr(tv, n, object)
    tv.Mlp = TimmLikeMlp        # the reference instantiates timm's Mlp (ar_discrete.py:124,227,235)
    sys.modules.update({"diffusers": diffusers, "timm": types.ModuleType("timm"),
                        "timm.models": types.ModuleType("timm.models"),
                        "timm.models.vision_transformer": tv})
    pkg = types.ModuleType("refpkg")
    pkg.__path__ = [str(REF)]
    sys.modules["refpkg"] = pkg
    mods = {}
    for name in ("base", "nar", "ar_discrete"):
        spec = importlib.util.spec_from_file_location(f"refpkg.{name}", REF / f"{name}.py")
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refpkg.{name}"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


class PatchedRand:
    """Makes the reference's internal ``torch.rand(size=...)`` return the supplied uniforms."""

    def __init__(self, noise):
        self.noise = noise

    def __enter__(self):
        self._orig = torch.rand
        noise = self.noise

        def fake(*a, size=None, **k):
            shape = tuple(size) if size is not None else tuple(a)
            assert tuple(noise.shape) == shape, (noise.shape, shape)
            return noise.clone()
        torch.rand = fake

    def __exit__(self, *exc):
        torch.rand = self._orig


def build_ref_d3pm(AR, S, transition):
    """Reference tensors for the hard-coded K=1025 (ar_discrete.py:257-277), on CPU."""
    s = types.SimpleNamespace()
    s.timesteps = S
    s.eps = 1.0e-6
    s.num_pixel_vals = 1025
    s.betas = AR.cosine_beta_schedule(s, S + 1).to(torch.float16)
    if transition == "absorbing":
        one = [AR._get_absorbing_transition_mat(s, t) for t in range(S)]
    else:
        one = [AR.create_transition_matrix(s, s.betas[t]) for t in range(S)]
    s.q_onestep_mats = torch.stack(one, dim=0).to(torch.float16)
    q = s.q_onestep_mats[0]
    qs = [q]
    for t in range(1, S):
        q = torch.tensordot(q, s.q_onestep_mats[t], dims=[[1], [0]])
        qs.append(q)
    s.q_mats = torch.stack(qs, dim=0).to(torch.float16)
    s.transpose_q_onestep_mats = torch.transpose(s.q_onestep_mats, 1, 2).to(torch.float16)
    for name in ("_at", "_at_onehot", "q_probs", "q_posterior_logits", "p_sample", "q_sample"):
        setattr(s, name, types.MethodType(getattr(AR, name), s))
    return s


def structured(mats, m):
    """Is every (S,K,K) table 'structured' (all diagonals equal, all off-diagonals equal, ...)?"""
    S, K, _ = mats.shape
    a, b = (0 if m != 0 else 1), (1 if m != 1 else 2)
    ok = True
    for t in range(S):
        M = mats[t].float()
        diag = torch.diagonal(M)
        notm = torch.ones(K, dtype=torch.bool)
        notm[m] = False
        ok &= bool((diag[notm] == M[a, a]).all())
        off = M.clone()
        off[torch.arange(K), torch.arange(K)] = M[a, b]
        off[:, m] = M[a, b]
        ok &= bool((off == M[a, b]).all())
        col = M[:, m][notm]
        ok &= bool((col == M[a, m]).all())
    return ok


def gen_d3pm(AR, transition, S=100, W=40, seed=1234):
    K, m = 1025, 512
    s = build_ref_d3pm(AR, S, transition)
    a, b = 0, 1
    one, cum = s.q_onestep_mats.float(), s.q_mats.float()
    out = dict(S=S, K=K, betas=s.betas.numpy(),
               one_aa=one[:, a, a].numpy(), one_ab=one[:, a, b].numpy(), one_am=one[:, a, m].numpy(),
               one_mm=one[:, m, m].numpy(),
               cum_aa=cum[:, a, a].numpy(), cum_ab=cum[:, a, b].numpy(), cum_am=cum[:, a, m].numpy(),
               cum_mm=cum[:, m, m].numpy(), cum_ma=cum[:, m, a].numpy(),
               structured_one=structured(s.q_onestep_mats, m), structured_cum=structured(s.q_mats, m))
    ts = [0, 1, 2, 10, 50, 90, 98, 99]
    # inputs come from detrand (seed-reproducible in the tests); only reference OUTPUTS are stored
    x0 = torch.from_numpy(detrand.integers(seed, 0, 1024, (len(ts), W)))
    x0[:, 0] = m  # a token that already equals the absorbing class
    x0[:, 1] = 1024
    t = torch.tensor(ts, dtype=torch.int64)
    mask = torch.ones(W, dtype=torch.int64)
    mask[-3:] = 0
    noise_q = torch.from_numpy(detrand.uniform(seed + 1, (len(ts), W, K)))
    with PatchedRand(noise_q):
        xt = s.q_sample(x0, t, mask)
    out.update(seed=seed, W=W, q_t=t.numpy(), q_xt=xt.numpy().astype(np.int32))
    # p_sample on fp16 logits: x_t = the forward-noised tokens above (mix of masked / unmasked)
    logits = torch.from_numpy(detrand.normal(seed + 2, (len(ts), W, K)) * 2.0).to(torch.float16)
    noise_p = torch.from_numpy(detrand.uniform(seed + 3, (len(ts), W, K)))
    x_in = xt.to(torch.int32)
    with PatchedRand(noise_p):
        samp, p0 = s.p_sample(logits, t, x_in)
    post = s.q_posterior_logits(logits, x_in, t, x_start_logits=True)
    greedy = torch.argmax(post, dim=-1)
    out.update(p_sample=samp.numpy().astype(np.int32), p_greedy=greedy.numpy().astype(np.int32),
               p_post_head=post[:, :6].numpy())   # fp16 posterior logits of the first 6 tokens per t
    np.savez_compressed(OUT / f"d3pm_{transition}_k1025.npz", **out)
    print(transition, "structured one/cum:", out["structured_one"], out["structured_cum"],
          "q_sample leaked:", int((xt != x0 * mask).sum()), "of", xt.numel())


def gen_denoiser(mods, seed=7):
    base, nar = mods["base"], mods["nar"]
    torch.manual_seed(seed)
    K, d, h, L = 64, 64, 1, 2   # head_dim 64, the only head size the model factory produces
    model = nar.NAR(K, d_model=d, n_heads=h, n_layers=L, p_dropout=0.1).eval()
    for n, p in model.named_parameters():
        if "norm.emb" in n:
            torch.nn.init.normal_(p, std=0.05)
        p.data = p.data.bfloat16().float()   # bf16-representable: the CUDA path sees identical weights
    g = torch.Generator().manual_seed(seed)
    lens = [(5, 9, 11), (3, 4, 20)]
    text = [torch.randint(1, K, (a,), generator=g) for a, _, _ in lens]
    proms = [torch.randint(0, K, (b, 8), generator=g) for _, b, _ in lens]
    resps = [torch.randint(0, K, (c, 3), generator=g) for _, _, c in lens]   # levels 0..2 known
    ql = torch.tensor([2, 2])
    # --- replicate Base.forward up to the un-padded logits with the reference's own modules
    with torch.no_grad():
        x_list = model._samplewise_merge_tensors(model.text_emb(text), model.proms_emb(proms),
                                                 model.resps_emb(resps), sep=model.sep)
        x, m = base.list_to_tensor(x_list)
        x = model.sin_emb.add_pe(x)
        hid = []
        for blk in model.blocks:
            x = blk(x, m, ql)
            hid.append(x.clone())
        hfull = model.classifier(x) * m
    sd = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    out = {f"sd.{k}": v for k, v in sd.items()}
    for i in range(2):
        out[f"text{i}"], out[f"proms{i}"], out[f"resps{i}"] = text[i].numpy(), proms[i].numpy(), resps[i].numpy()
        T = len(x_list[i])
        out[f"logits{i}"] = hfull[i, :T].numpy()
        out[f"hidden_last{i}"] = hid[-1][i, :T].numpy()
        out[f"hidden_first{i}"] = hid[0][i, :T].numpy()
    out["levels"] = ql.numpy()
    out["n_heads"], out["n_layers"] = h, L
    np.savez_compressed(OUT / "denoiser_nar_small.npz", **out)

    # --- D3PM glue assembled from reference modules: 8 response levels,
    #     AdaLN table with S+1 rows indexed by t, time_emb added to the response rows, 8K head.
    S = 20
    torch.manual_seed(seed + 1)

    class Glue(base.Base):
        casual = False
        n_resp_levels = 8
        use_stop_token = False
        norm_type = "adaln"
        resp_loss_only = True

    gm = Glue(K, d_model=d, n_heads=h, n_layers=L, p_dropout=0.1).eval()
    for blk in gm.blocks:
        for sub in (blk.attn, blk.ffn):
            sub.norm = base.AdaLN(d, S + 1)
            torch.nn.init.normal_(sub.norm.emb.weight, std=0.05)
    gm.classifier = torch.nn.Linear(d, 8 * K)
    gm.time_emb = torch.nn.Embedding(S + 1, d)
    for p in gm.parameters():
        p.data = p.data.bfloat16().float()
    xt = [torch.randint(0, K, (c, 8), generator=g) for _, _, c in lens]
    t = torch.tensor([13, 4])
    with torch.no_grad():
        re_ = [r + gm.time_emb(t[i])[None] for i, r in enumerate(gm.resps_emb(xt))]
        x_list = gm._samplewise_merge_tensors(gm.text_emb(text), gm.proms_emb(proms), re_, sep=gm.sep)
        x, m = base.list_to_tensor(x_list)
        x = gm.sin_emb.add_pe(x)
        for blk in gm.blocks:
            x = blk(x, m, t)
        hfull = gm.classifier(x) * m
    out = {f"sd.{k}": v.detach().numpy() for k, v in gm.state_dict().items()}
    for i in range(2):
        out[f"text{i}"], out[f"proms{i}"], out[f"xt{i}"] = text[i].numpy(), proms[i].numpy(), xt[i].numpy()
        T = len(x_list[i])
        out[f"logits{i}"] = hfull[i, :T].numpy()
    out["t"] = t.numpy()
    out["n_heads"], out["n_layers"], out["S"] = h, L, S
    np.savez_compressed(OUT / "denoiser_diffusion_small.npz", **out)
    print("denoiser fixtures written; state-dict keys:", len(sd))


def det_state_dict(shapes, seed=4242):
    """Deterministic weights for the compat DiT, regenerated identically by the test (detrand)."""
    sd = {}
    for i, (k, shape) in enumerate(sorted(shapes.items())):
        v = detrand.normal(seed + 17 * i, tuple(shape)) * np.float32(0.1)
        if k.endswith("weight") and len(shape) == 1:          # LayerNorm gains around 1
            v = v + np.float32(1.0)
        sd[k] = torch.from_numpy(v.astype(np.float32))
    return sd


def gen_ar_discrete_dit(mods, seed=99):
    """The reference's own D3PM class (ar_discrete.py:205-256): denoiser logits for one timestep and
    the two conditioning encodings, computed by the reference's modules with the call sequence of
    its generate_audio (:735-776; the method itself hard-codes cuda:0).  The constructor moves its
    tables to cuda:0 and chains 99 (1025, 1025) fp16 products — both neutralised while it runs."""
    ar = mods["ar_discrete"]
    real_to, real_td = torch.Tensor.to, torch.tensordot

    def cpu_to(self, *a, **k):
        a = tuple(x for x in a if not (isinstance(x, str) and x.startswith("cuda")))
        return real_to(self, *a, **k) if (a or k) else self
    torch.Tensor.to, torch.tensordot = cpu_to, (lambda a, b, dims=None: a)
    try:
        m = ar.AR(32, 100, 1025, 8, 16, 8)
    finally:
        torch.Tensor.to, torch.tensordot = real_to, real_td
    m.eval()
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(det_state_dict(shapes))
    text = torch.from_numpy(detrand.integers(seed, 1, 1025, (37,)))
    proms = torch.from_numpy(detrand.integers(seed + 1, 0, 1025, (180, 8)))
    x_t = torch.zeros(1, 448, dtype=torch.int32)
    x_t[0, :350] = torch.from_numpy(detrand.integers(seed + 2, 1, 1025, (350,))).int()
    t = torch.tensor([40])
    F = torch.nn.functional
    with torch.no_grad():
        mask = x_t[0] != 0                                                   # :712
        text_p = torch.stack([F.pad(text, (0, 50 - text.shape[0]))])        # :715-721
        proms_p = torch.stack([F.pad(proms, (0, 0, 0, 398 - proms.shape[0]))])   # :723-735
        cond1 = m.proms_emb(torch.stack([p[:398, :] for p in proms_p]))[0]  # :736-737
        cond2 = m.text_emb(text_p)                                           # :739
        cond2 = m.sin_emb.add_pe(cond2)[0]
        cond2 = m.encodertext(cond2).unsqueeze(0)
        cond1 = m.sin_emb.add_pe(cond1)[0]
        cond1 = m.encoder2(cond1).unsqueeze(0)
        t_emb = m.time_emb(t)                                                # :752
        x = m.resps_emb(x_t)[0].unsqueeze(0)                                 # :753,760
        for block in m.blocks:
            x = block(x, cond1, cond2, t_emb, mask)
        x = x[:448, :] * mask.unsqueeze(1)                                   # :772
        logits = m.final(x)
    rows = np.array([0, 1, 2, 100, 200, 349, 350, 447])
    np.savez_compressed(OUT / "ar_discrete_dit.npz", keys=np.array(sorted(shapes)),
                        shapes=np.array([str(shapes[k]) for k in sorted(shapes)]),
                        rows=rows, logits_rows=logits[0, rows].numpy(), logits_absmax=float(logits.abs().max()),
                        logits_sum=float(logits.double().sum()), cond1_head=cond1[0, :4].numpy(),
                        cond2_head=cond2[0, :4].numpy(), seed=seed, t=40)
    print("ar_discrete DiT fixture written:", len(shapes), "state-dict entries, logits", tuple(logits.shape),
          "absmax", float(logits.abs().max()))


if __name__ == "__main__":
    mods = load_reference()
    AR = mods["ar_discrete"].AR
    gen_d3pm(AR, "absorbing")
    gen_d3pm(AR, "uniform")
    gen_denoiser(mods)
    gen_ar_discrete_dit(mods)
