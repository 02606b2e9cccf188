"""GPU parity of the PRODUCTION reverse step — ``head_sample_kernel`` (csrc/head_sample_tcgen05.cu), the
classifier GEMM with the D3PM posterior + sampler as its epilogue, the kernel ``Session.run`` issues on
every denoise step — directly against the ORACLE (``oracle/d3pm.py``: the reference's dense fp16-table
``q_posterior_logits`` / ``p_sample``, ar_discrete.py:347-375,401-420), at K = 1024 and 8 levels, both
transitions.  Every test asserts that the fused path is the one taken (``L.head_fused``; kernel-level
calls pass no logits scratch, so the two-kernel path could not even run).

  (a) greedy: fused codes == oracle.p_sample(greedy=True) wherever the oracle's top-2 margin is clear —
      kernel level, through ``generate_audio`` on a teacher-forced multi-step loop, and on the C2 shape
      at the full model size;
  (b) Philox: chi-square of >= 40 k fused draws against softmax(oracle.q_posterior_logits) — the
      oracle's law — for masked / unmasked x_t at t in {1, mid, S-1};
  (c) t = 0 rows: arg max of the raw logits, as ar_discrete.py:407,413.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

DEV = "cuda"
K, LEVELS = 1024, 8


@pytest.fixture(scope="module")
def L():
    from vall_e.b200 import lib
    lib.load()
    return lib


@pytest.fixture(scope="module", params=["absorbing", "uniform"])
def pair(request):
    from oracle.d3pm import D3PM
    from vall_e.vall_e import d3pm as pd
    S = 50
    return request.param, D3PM(S, K, request.param), pd.scalar_table(S, K, request.param).to(DEV), S


def _oracle_step(orc, logits16, t_rows, x_t, chunk=48):
    """oracle.p_sample(greedy) in row chunks (its _at materialises a (rows, K, K) fp16 gather)."""
    codes, posts = [], []
    for r0 in range(0, logits16.shape[0], chunk):
        c, p = orc.p_sample(logits16[r0:r0 + chunk], t_rows[r0:r0 + chunk], x_t[r0:r0 + chunk].to(torch.int32),
                            greedy=True)
        codes.append(c)
        posts.append(p.float())
    return torch.cat(codes), torch.cat(posts)


def _clear(post, margin):
    top2 = post.topk(2, dim=-1).values
    return (top2[..., 0] - top2[..., 1]) > margin


@pytest.mark.parametrize("act", [torch.float16, torch.bfloat16], ids=["f16", "bf16"])
def test_fused_head_greedy_codes_vs_oracle(L, pair, act):
    """Kernel level: 600 rows x 8 levels (a ragged last 256-row block), six utterances at
    t = 0, 1, 2, mid, S-2, S-1, masked and unmasked x_t."""
    tr, orc, table, S = pair
    code = L.ABSORBING if tr == "absorbing" else L.UNIFORM
    rows, d, B = 600, 128, 6
    g = torch.Generator().manual_seed(5)
    head_in = torch.randn(rows, d, generator=g).to(act)
    W = (torch.randn(LEVELS * K, d, generator=g) * 0.25).to(act)
    bias = torch.randn(LEVELS * K, generator=g)
    x_t = torch.randint(0, K, (rows, LEVELS), generator=g, dtype=torch.int32)
    x_t[::2] = K // 2
    row_utt = (torch.arange(rows, dtype=torch.int32) % B).sort().values
    t_utt = torch.tensor([0, 1, 2, S // 2, S - 2, S - 1], dtype=torch.int32)
    utt = torch.zeros(B, L.U_STRIDE, dtype=torch.int32)
    utt[:, L.U_GID] = torch.arange(B, dtype=torch.int32) + 40
    logits = (head_in.float() @ W.float().t() + bias).view(rows, LEVELS, K)          # fp32, the kernel's accumulators
    t_rows = t_utt[row_utt.long()].long()
    ref, post = _oracle_step(orc, logits.to(torch.float16), t_rows, x_t)

    assert L.head_fused(d, K, L.NOISE_GREEDY)
    out = torch.full((rows, LEVELS), -1, dtype=torch.int32, device=DEV)
    L.head_posterior_sample(out, None, head_in.to(DEV), W.to(DEV), bias.to(DEV), x_t.to(DEV), row_utt.to(DEV),
                            t_utt.to(DEV), utt.to(DEV), table, LEVELS, K, code, L.NOISE_GREEDY)
    got = out.cpu().long()
    assert int(got.min()) >= 0 and int(got.max()) < K
    clear = _clear(post, 0.05)                       # fp16 resolution of the oracle's posterior logits
    assert clear.float().mean().item() > 0.7
    assert torch.equal(got[clear], ref[clear]), f"{int((got[clear] != ref[clear]).sum())} greedy codes differ"
    # (c) t == 0 rows: arg max of the raw logits (ar_discrete.py:407,413), with Philox noise too
    t0 = t_rows == 0
    out2 = torch.empty_like(out)
    L.head_posterior_sample(out2, None, head_in.to(DEV), W.to(DEV), bias.to(DEV), x_t.to(DEV), row_utt.to(DEV),
                            t_utt.to(DEV), utt.to(DEV), table, LEVELS, K, code, L.NOISE_PHILOX, seed=9)
    raw_clear = _clear(logits[t0], 2e-2)
    assert torch.equal(out2.cpu().long()[t0][raw_clear], logits[t0].argmax(-1)[raw_clear])
    assert torch.equal(ref[t0][raw_clear], logits[t0].argmax(-1)[raw_clear])       # and that is what the oracle does


def test_fused_head_philox_law_vs_oracle_posterior(L, pair):
    """(b): 40 960 draws per case of one token (same logits row) against the ORACLE's posterior law."""
    tr, orc, table, S = pair
    code = L.ABSORBING if tr == "absorbing" else L.UNIFORM
    n, d, levels = 40960, 128, 2
    g = torch.Generator().manual_seed(23)
    row = torch.randn(1, d, generator=g).to(torch.float16)
    W = (torch.randn(levels * K, d, generator=g) * 0.22).to(torch.float16)
    bias = torch.randn(levels * K, generator=g)
    logits16 = (row.float() @ W.float().t() + bias).view(1, levels, K).to(torch.float16)
    head_in = row.repeat(n, 1).to(DEV)
    ru = torch.zeros(n, dtype=torch.int32, device=DEV)
    u1 = torch.zeros(1, L.U_STRIDE, dtype=torch.int32, device=DEV)
    assert L.head_fused(d, K, L.NOISE_PHILOX)
    for t in (1, S // 2, S - 1):
        for xt_val in (K // 2, 3):
            x1 = torch.full((1, levels), xt_val, dtype=torch.int32)
            law = torch.softmax(orc.q_posterior_logits(logits16, x1, torch.tensor([t])).double(), -1)[0]   # (levels, K)
            xt = torch.full((n, levels), xt_val, dtype=torch.int32, device=DEV)
            out = torch.empty(n, levels, dtype=torch.int32, device=DEV)
            L.head_posterior_sample(out, None, head_in, W.to(DEV), bias.to(DEV), xt, ru,
                                    torch.tensor([t], dtype=torch.int32, device=DEV), u1, table, levels, K, code,
                                    L.NOISE_PHILOX, seed=1000 + t)
            draws = out.cpu().long()
            for lv in range(levels):
                counts = torch.bincount(draws[:, lv], minlength=K).double()
                expected = law[lv] * n
                keep = expected > 5
                chi2 = (((counts - expected) ** 2) / expected)[keep].sum().item()
                rest_obs, rest_exp = counts[~keep].sum().item(), expected[~keep].sum().item()
                cells = int(keep.sum().item())
                if rest_exp > 5:
                    chi2 += (rest_obs - rest_exp) ** 2 / rest_exp
                    cells += 1
                dof = max(cells - 1, 1)
                assert chi2 < dof + 6 * math.sqrt(2 * dof), (tr, t, xt_val, lv, chi2, dof)


def _make(d, h, nl, S, transition, seed):
    from oracle import denoiser as on
    from vall_e.vall_e.diffusion import Diffusion
    sd = on.random_state_dict(K, d, nl, S + 1, n_resp_levels=8, n_out=8 * K, seed=seed, time_rows=S + 1)
    m = Diffusion(K, d_model=d, n_heads=h, n_layers=nl, n_steps=S, transition=transition)
    m.load_state_dict(sd)
    return m.to(DEV), sd


def _teacher_forced(m, sd, h, nl, S, transition, text, proms, x_T, lens, margin):
    """generate_audio (greedy, fused head asserted) with a trace; every step is re-done by the oracle
    (fp32 denoiser -> fp16 logits -> dense-table p_sample) from the CUDA trajectory's previous state."""
    from oracle import denoiser as on
    from oracle.d3pm import D3PM
    from vall_e.b200 import lib as L
    assert L.head_fused(m.engine().w.d, K, L.NOISE_GREEDY)
    orc = D3PM(S, K, transition)
    trace = []
    out = m.generate_audio([x.to(DEV) for x in text], [x.to(DEV) for x in proms], [r.to(DEV) for r in x_T.split(lens)],
                           greedy=True, trace=trace, use_graph=False)
    assert len(trace) == S - 1
    prev, checked, agree = x_T, 0, 0
    for step, t in enumerate(range(S - 1, 0, -1)):
        tt = torch.full((len(lens),), t)
        lg = torch.cat(on.diffusion_logits(sd, text, proms, list(prev.split(lens)), tt, h, nl)).to(torch.float16)
        ref, post = _oracle_step(orc, lg, torch.full((lg.shape[0],), t), prev)
        clear = _clear(post, margin)
        got = trace[step].cpu().long()
        checked += int(clear.sum())
        agree += int((got[clear] == ref[clear]).sum())
        prev = got
    assert torch.equal(torch.cat(out).cpu(), prev)
    return checked, agree, prev.numel() * (S - 1)


@pytest.mark.parametrize("transition", ["absorbing", "uniform"])
def test_fused_head_reverse_loop_teacher_forced_vs_oracle(transition):
    """Multi-step loop through Diffusion.generate_audio / Session.run at K = 1024 x 8 levels."""
    d, h, nl, S = 128, 2, 2, 8
    m, sd = _make(d, h, nl, S, transition, seed=11)
    g = torch.Generator().manual_seed(13)
    lens = [70, 190]
    text = [torch.randint(1, K, (n,), generator=g) for n in (5, 9)]
    proms = [torch.randint(0, K, (n, 8), generator=g) for n in (12, 30)]
    x_T = (torch.full((sum(lens), 8), K // 2, dtype=torch.long) if transition == "absorbing"
           else torch.randint(0, K, (sum(lens), 8), generator=g))
    checked, agree, total = _teacher_forced(m, sd, h, nl, S, transition, text, proms, x_T, lens, margin=0.08)
    assert checked > 0.5 * total and agree == checked, (agree, checked, total)


def test_fused_head_full_size_c2_shape_vs_oracle():
    """BASELINE configs[1] (C2): the full denoiser (d = 1024, 16 heads, 12 layers), one utterance of
    50 phones + 225 prompt frames + 750 frames, three greedy denoise steps through generate_audio."""
    d, h, nl, S = 1024, 16, 12, 4
    m, sd = _make(d, h, nl, S, "absorbing", seed=2)
    g = torch.Generator().manual_seed(17)
    text = [torch.randint(1, K, (50,), generator=g)]
    proms = [torch.randint(0, K, (225, 8), generator=g)]
    x_T = torch.full((750, 8), K // 2, dtype=torch.long)
    torch.set_num_threads(max(1, min(16, torch.get_num_threads())))
    # margin: the 2e-2 logits bar moves a posterior logit by up to ~2e-2 on top of the fp16 grid
    checked, agree, total = _teacher_forced(m, sd, h, nl, S, "absorbing", text, proms, x_T, [750], margin=0.12)
    assert checked > 0.5 * total and agree == checked, (agree, checked, total)
