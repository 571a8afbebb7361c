"""GPU tests of the sequence front-end's BACKWARD kernels (Tier B) against torch autograd over the SELF-ORACLE
oracle/seq_oracle.py.

PARITY UNPINNED BY THE REFERENCE (the reference has no sequence attention / LayerNorm); the checker is
autograd over the plain-PyTorch restatement, evaluated in fp32 on the same bf16-rounded inputs the kernels read.
Tolerance (BASELINE.json north_star, bf16): rel-err <= 2e-2, written at each assert. All calls go through the C ABI
(include/fnd_seq_b200.h) via ultrafnd_git_b200/seq_ops.py.
"""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import seq_oracle as O

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2


def _rel(a, b):
    return O.rel_err(a.detach().float().cpu(), b.detach().float().cpu())


def _attn_case(B, H, Lq, Lk, seed, lens=None, scatter=False):
    from ultrafnd_git_b200 import seq_ops as S
    d = H * 64
    g = torch.Generator().manual_seed(seed)
    q = torch.randn(B, Lq, d, generator=g).bfloat16()
    k = torch.randn(B, Lk, d, generator=g).bfloat16()
    v = torch.randn(B, Lk, d, generator=g).bfloat16()
    d_o = torch.randn(B, Lq, d, generator=g).bfloat16()
    mask = torch.ones(B, Lk, dtype=torch.bool)
    if lens is not None:
        mask = torch.arange(Lk)[None, :] < torch.tensor(lens)[:, None]
    if scatter:
        mask[-1] &= torch.rand(Lk, generator=g) < 0.6
        mask[-1, 0] = True
    # checker: autograd over the restated attention, fp32, same bf16-rounded operands
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    o_ref, _ = O.attention_only(qf, kf, vf, mask, H)
    o_ref.backward(d_o.float())
    dev = torch.device("cuda")
    err = S.new_err_flag(dev)
    qc, kc, vc, doc = (t.cuda().view(-1, d) for t in (q, k, v, d_o))
    mask_u8 = mask.to(torch.uint8).cuda().contiguous()
    pos = torch.arange(1, Lk + 1, dtype=torch.int32)
    kv_len = (mask.to(torch.int32) * pos).amax(dim=1).to(torch.int32).cuda()
    lse = torch.empty(B, H, Lq, dtype=torch.float32, device=dev)
    o = S.coattn_forward(qc, kc, vc, B, H, Lq, Lk, kv_len=kv_len, kv_mask=mask_u8, lse=lse, err=err)
    dq = torch.full((B * Lq, d), float("nan"), dtype=torch.bfloat16, device=dev)
    dk = torch.full((B * Lk, d), float("nan"), dtype=torch.bfloat16, device=dev)
    dv = torch.full((B * Lk, d), float("nan"), dtype=torch.bfloat16, device=dev)
    S.coattn_backward(qc, kc, vc, o, doc, lse, B, H, Lq, Lk, dq, dk, dv, kv_len=kv_len, kv_mask=mask_u8, err=err)
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    return (_rel(o.view(B, Lq, d), o_ref), _rel(dq.view(B, Lq, d), qf.grad), _rel(dk.view(B, Lk, d), kf.grad),
            _rel(dv.view(B, Lk, d), vf.grad), dk.view(B, Lk, d), dv.view(B, Lk, d), mask)


@pytest.mark.parametrize("B,H,Lq,Lk,lens,scatter", [
    (2, 2, 100, 83, None, False),                  # single ragged tile on both sides
    (3, 4, 300, 260, [260, 17, 0], True),          # several tiles, prefix lengths, an EMPTY key sequence, a scattered mask
    (2, 3, 513, 700, [700, 129], False),           # tile / block remainders of one row
    (2, 16, 1024, 512, None, False),               # the stress shape's tile counts
])
def test_coattn_backward_matches_autograd(B, H, Lq, Lk, lens, scatter):
    eo, eq, ek, ev, dk, dv, mask = _attn_case(B, H, Lq, Lk, seed=B * 1000 + Lq + Lk, lens=lens, scatter=scatter)
    print(f"attn bwd B={B} H={H} Lq={Lq} Lk={Lk}: O {eo:.2e} dQ {eq:.2e} dK {ek:.2e} dV {ev:.2e}")
    assert eo < BF16_TOL and eq < BF16_TOL and ek < BF16_TOL and ev < BF16_TOL      # north_star bf16 tolerance 2e-2
    # masked keys receive EXACT zeros (never NaN, never a stale buffer value)
    dead = ~mask.cuda()
    assert torch.all(dk[dead] == 0) and torch.all(dv[dead] == 0)
    assert torch.isfinite(dk.float()).all() and torch.isfinite(dv.float()).all()


def test_layernorm_backward_matches_autograd():
    from ultrafnd_git_b200 import seq_ops as S
    g = torch.Generator().manual_seed(3)
    for M, d in ((50, 256), (1000, 1024), (7, 2048), (33, 64), (4099, 512)):
        t = (torch.randn(M, d, generator=g) * 2 + 0.5).bfloat16()
        dy = torch.randn(M, d, generator=g).bfloat16()
        w = (1 + 0.1 * torch.randn(d, generator=g))
        b = 0.1 * torch.randn(d, generator=g)
        tf = t.float().requires_grad_(True)
        wf, bf = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
        F.layer_norm(tf, (d,), wf, bf, 1e-5).backward(dy.float())
        dt, dg, db = S.layernorm_backward(t.cuda(), dy.cuda(), w.cuda(), 1e-5)
        e = (_rel(dt, tf.grad), _rel(dg, wf.grad), _rel(db, bf.grad))
        print(f"layernorm bwd M={M} d={d}: dt {e[0]:.2e} dgamma {e[1]:.2e} dbeta {e[2]:.2e}")
        assert e[0] < 5e-3 and e[1] < 1e-4 and e[2] < 1e-4      # dt: one bf16 rounding; dgamma / dbeta: fp32 sums


def test_pool_backward_colsum_wgrad():
    from ultrafnd_git_b200 import seq_ops as S
    g = torch.Generator().manual_seed(9)
    B, L, d = 5, 83, 320
    x = torch.randn(B, L, d, generator=g)
    m = torch.rand(B, L, generator=g) < 0.7
    m[2] = False                                    # a sample without any valid token: gradient 0 (clamp_min path)
    dp = torch.randn(B, d, generator=g)
    xf = x.clone().requires_grad_(True)
    O.masked_mean(xf, m).backward(dp)
    dx = S.masked_mean_pool_backward(dp.cuda(), B, L, mask=m.to(torch.uint8).cuda().contiguous())
    e = _rel(dx.view(B, L, d), xf.grad)
    print(f"pool bwd: {e:.2e}")
    assert e < 5e-3                                 # bf16 output rounding
    for M, N in ((1000, 1024), (37, 64), (20000, 3072)):
        y = torch.randn(M, N, generator=g).bfloat16()
        s = S.colsum(y.cuda())
        e = _rel(s, y.float().sum(0))
        print(f"colsum M={M} N={N}: {e:.2e}")
        assert e < 1e-5
    err = S.new_err_flag(torch.device("cuda"))
    for M, N, K in ((4096, 1024, 1024), (512, 64, 128), (8192, 3072, 1024), (664, 512, 4096)):
        dy = torch.randn(M, N, generator=g).bfloat16()
        xx = torch.randn(M, K, generator=g).bfloat16()
        dw = S.wgrad(dy.cuda(), xx.cuda(), err=err)
        ref = dy.float().t() @ xx.float()
        torch.cuda.synchronize()
        assert int(err.item()) == 0
        e = _rel(dw, ref)
        print(f"wgrad M={M} N={N} K={K}: {e:.2e}")
        assert e < 1e-4                             # same bf16 operands, fp32 accumulation


def _frontend_grads(d_model, heads, lengths, batch, seed, full=False, streams=None, blocks=None):
    """(kernel grads, oracle grads, forward errors) of sum_s <y_s, w_s> for fixed random w_s."""
    from ultrafnd_git_b200.seqfront import SequenceFrontEnd
    streams = O.FAKESV_STREAMS if streams is None else streams
    blocks = O.FAKESV_BLOCKS if blocks is None else blocks
    p = O.init_params(streams, blocks, d_model, seed=seed)
    data = O.make_batch(streams, lengths, batch, seed=seed + 1, full=full)
    # the kernels read bf16 features and bf16 GEMM weights: hand the checker the same rounded values
    for n in streams:
        data[n] = data[n].bfloat16().float()
    pr = {k: (v.bfloat16().float() if (k.endswith("weight") and v.dim() == 2) else v.clone()) for k, v in p.items()}
    pr = {k: v.requires_grad_(True) for k, v in pr.items()}
    g = torch.Generator().manual_seed(seed + 2)
    w = {n: torch.randn(batch, streams[n][1], generator=g) for n in streams}
    ref = O.forward(pr, data, streams, blocks, heads)
    sum((ref[n] * w[n]).sum() for n in streams).backward()
    fe = SequenceFrontEnd(d_model, heads, streams, blocks).cuda()
    fe.load_state_dict({k: v for k, v in p.items()})
    out = fe({k: v.cuda() for k, v in data.items()})
    loss = sum((out[n] * w[n].cuda()).sum() for n in streams)
    loss.backward()
    torch.cuda.synchronize()
    fe.check_error()
    fwd = {n: _rel(out[n], ref[n]) for n in streams}
    got = {k: v.grad for k, v in fe.named_parameters()}
    return got, {k: v.grad for k, v in pr.items()}, fwd


@pytest.mark.parametrize("d_model,heads,lengths,batch,full", [
    (256, 4, {"text": 40, "frames": 83, "audio": 50, "c3d": 83}, 3, False),        # FakeSV-shaped, ragged, one scattered mask
    (512, 8, {"text": 300, "frames": 83, "audio": 50, "c3d": 83}, 4, False),       # several query / key tiles
    (128, 2, {"text": 130, "frames": 260, "audio": 64, "c3d": 16}, 2, True),       # full-length sequences
])
def test_frontend_backward_matches_oracle_autograd(d_model, heads, lengths, batch, full):
    got, ref, fwd = _frontend_grads(d_model, heads, lengths, batch, seed=17 + d_model, full=full)
    assert max(fwd.values()) < BF16_TOL
    worst = ("", 0.0)
    for k, gref in ref.items():
        assert got[k] is not None, k
        e = _rel(got[k], gref)
        if e > worst[1]:
            worst = (k, e)
        assert torch.isfinite(got[k]).all(), k
    errs = {k: _rel(got[k], ref[k]) for k in ref}
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:4]
    print(f"front-end backward d={d_model}: worst parameter-gradient rel-err {worst[1]:.2e} ({worst[0]}); top: "
          + ", ".join(f"{k} {v:.1e}" for k, v in top))
    assert worst[1] < BF16_TOL                    # north_star bf16 tolerance 2e-2, every parameter gradient


def test_frontend_backward_at_the_stress_shape():
    """BASELINE.json configs[4] shape (1024 text tokens x 512 frames, hidden 1024, 16 heads), batch 2, full-length
    sequences: every parameter gradient of the front-end against autograd over the self-oracle."""
    streams = {"text": (768, 768, "text_features"), "frames": (4096, 512, "visual_features")}
    got, ref, fwd = _frontend_grads(1024, 16, {"text": 1024, "frames": 512}, 2, seed=99, full=True, streams=streams,
                                    blocks=(("text", "frames"),))
    errs = {k: _rel(got[k], ref[k]) for k in ref}
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    print(f"front-end backward, stress shape: forward {max(fwd.values()):.2e}; parameter gradients top: " + ", ".join(f"{k} {v:.1e}" for k, v in top))
    assert max(fwd.values()) < BF16_TOL and max(errs.values()) < BF16_TOL        # north_star bf16 tolerance 2e-2


def test_end_to_end_training_gradients_through_fusion_and_classifier():
    """Sequences -> SequenceFrontEnd -> CrossModalTransformer -> DeepTruthClassifier -> cross-entropy, ONE backward.
    The reference's fusion forward is not smooth (|a - b| interactions, evidence gates: a 1e-3 relative change of its
    inputs moves its input gradients by 2-3 %, measured on the CPU oracle), so the chain is checked link by link at the
    SAME evaluation point instead of end to end at two slightly different ones:
      (1) loss against the all-oracle chain (bf16 tolerance);
      (2) dL/d(pooled vectors) that the fusion backward hands to the front-end, against autograd over the restated
          reference evaluated AT the front-end's own outputs;
      (3) the front-end's parameter gradients against autograd over the self-oracle driven by that same upstream gradient."""
    from oracle import fnd_oracle as FO
    from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier
    from ultrafnd_git_b200.seqfront import SequenceFrontEnd
    streams, blocks = O.FAKESV_STREAMS, O.FAKESV_BLOCKS
    d_model, heads, B = 128, 2, 6
    lengths = {"text": 40, "frames": 83, "audio": 50, "c3d": 83}
    p = O.init_params(streams, blocks, d_model, seed=5)
    data = O.make_batch(streams, lengths, B, seed=6)
    for n in streams:
        data[n] = data[n].bfloat16().float()
    fus, clf = FO.init_params(42)
    FO.perturb_node_head(clf)
    tier_a = FO.make_batch(B, seed=8)
    label, gnn, aux = tier_a["label"], tier_a["gnn_feat"], tier_a["aux"]
    # ---- kernels: one forward, one backward through all three modules
    f = CrossModalTransformer(precision="fp32"); c = DeepTruthClassifier(precision="fp32")
    f.load_state_dict(fus); c.load_state_dict(clf)
    for m in list(f.modules()) + list(c.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    f.train(); c.train(); f._sync_dropout(); c._sync_dropout()
    fe = SequenceFrontEnd(d_model, heads, streams, blocks).cuda()
    fe.load_state_dict(p)
    seq_feats = {streams[n][2]: data[n].cuda() for n in streams}
    for n in streams:
        seq_feats[streams[n][2] + "_mask"] = data[n + "_mask"].cuda()
    seq_feats["gnn_feat"] = gnn.cuda()
    pooled = fe.forward_features(seq_feats)                 # what CrossModalTransformer.forward does for 3-D inputs
    for n in streams:
        pooled[streams[n][2]].retain_grad()
    out = f(pooled)
    loss = F.cross_entropy(c(out["fused"], aux.cuda())["logits"], label.cuda())
    loss.backward()
    torch.cuda.synchronize()
    fe.check_error()
    # ---- (1) + (2): restated reference at the front-end's own outputs
    y_in = {streams[n][2]: pooled[streams[n][2]].detach().cpu().clone().requires_grad_(True) for n in streams}
    feats = dict(y_in); feats["gnn_feat"] = gnn
    fo = FO.fusion_forward({k: v.clone() for k, v in fus.items()}, feats, dropout=0.0)
    co = FO.classifier_forward({k: v.clone() for k, v in clf.items()}, fo["fused"], aux, dropout=0.0)
    loss_ref = F.cross_entropy(co["logits"], label)
    loss_ref.backward()
    e_loss = abs(float(loss.detach()) - float(loss_ref.detach())) / float(loss_ref.detach())
    e_dy = {k: _rel(pooled[k].grad, y_in[k].grad) for k in y_in}
    # ---- (3): self-oracle front-end driven by the kernels' upstream gradient
    pr = {k: (v.bfloat16().float() if (k.endswith("weight") and v.dim() == 2) else v.clone()).requires_grad_(True) for k, v in p.items()}
    ys = O.forward(pr, data, streams, blocks, heads)
    torch.autograd.backward([ys[n] for n in streams], [pooled[streams[n][2]].grad.cpu() for n in streams])
    errs = {k: _rel(v.grad, pr[k].grad) for k, v in fe.named_parameters()}
    top = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    print(f"end-to-end: loss {float(loss.detach()):.6f} vs {float(loss_ref.detach()):.6f} ({e_loss:.1e}); dL/d pooled "
          + ", ".join(f"{k.split('_')[0]} {v:.1e}" for k, v in e_dy.items()) + "; front-end gradients top: "
          + ", ".join(f"{k} {v:.1e}" for k, v in top))
    assert e_loss < 1e-3                                   # Tier A in fp32 mode at identical inputs: north_star fp32 tolerance
    assert max(e_dy.values()) < 5e-3                       # fp32-mode gradient bound used throughout (5x the logits tolerance)
    assert max(errs.values()) < BF16_TOL                   # the front-end is bf16: north_star 2e-2


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 2e-2)])
def test_fusion_input_gradients_match_oracle(precision, tol):
    """dL/d(text, audio, visual, temporal vectors) returned by CrossModalTransformer's backward (dX_m = dP_m W_m) against
    autograd over the restated reference forward (cross_modal_transformer.py:141-198), same inputs on both sides."""
    from oracle import fnd_oracle as FO
    from ultrafnd_git_b200.modules import CrossModalTransformer, DeepTruthClassifier
    B = 24
    fus, clf = FO.init_params(42)
    FO.perturb_node_head(clf)
    batch = FO.make_batch(B, seed=31)
    keys = ("text_features", "audio_features", "visual_features", "temporal_features")
    ref_in = {k: batch[k].clone().requires_grad_(True) for k in keys}
    feats = dict(ref_in); feats["gnn_feat"] = batch["gnn_feat"]
    fo = FO.fusion_forward({k: v.clone() for k, v in fus.items()}, feats, dropout=0.0)
    co = FO.classifier_forward({k: v.clone() for k, v in clf.items()}, fo["fused"], batch["aux"], dropout=0.0)
    F.cross_entropy(co["logits"], batch["label"]).backward()
    f = CrossModalTransformer(precision=precision); c = DeepTruthClassifier(precision=precision)
    f.load_state_dict(fus); c.load_state_dict(clf)
    for m in list(f.modules()) + list(c.modules()):
        if isinstance(m, torch.nn.Dropout):
            m.p = 0.0
    f.train(); c.train(); f._sync_dropout(); c._sync_dropout()
    got_in = {k: batch[k].cuda().requires_grad_(True) for k in keys}
    gf = dict(got_in); gf["gnn_feat"] = batch["gnn_feat"].cuda()
    out = f(gf)
    F.cross_entropy(c(out["fused"], batch["aux"].cuda())["logits"], batch["label"].cuda()).backward()
    torch.cuda.synchronize()
    errs = {k: _rel(got_in[k].grad, ref_in[k].grad) for k in keys}
    print(f"fusion input gradients [{precision}]: " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert max(errs.values()) < 5 * tol          # same bound the per-parameter gradient tests use (5x the logits tolerance)


def test_flat_adamw_matches_torch_clip_and_adamw():
    """fnd_seq_grad_sumsq + fnd_seq_adamw_step against torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW on the SAME gradients
    (three steps, clipping active in the first and inactive in the last), incl. the 1 / world gradient scale."""
    from ultrafnd_git_b200 import seq_ops as S
    g = torch.Generator().manual_seed(12)
    n = 4 * 25_003
    w0 = torch.randn(n, generator=g)
    grads = [torch.randn(n, generator=g) * s for s in (3.0, 0.5, 0.001)]
    lr, betas, eps, wd, max_norm, world = 3e-3, (0.9, 0.999), 1e-8, 1e-2, 5.0, 4
    p = torch.nn.Parameter(w0.clone().double())
    opt = torch.optim.AdamW([p], lr=lr, betas=betas, eps=eps, weight_decay=wd)
    w = w0.clone().cuda()
    m, v = torch.zeros_like(w), torch.zeros_like(w)
    norms = []
    for t, gr in enumerate(grads, 1):
        p.grad = (gr.double() / world).clone()                 # the all-reduced SUM is gr; the mean is what the reference steps on
        norms.append(float(torch.nn.utils.clip_grad_norm_([p], max_norm)))
        opt.step()
        gc = gr.cuda()
        ss = S.grad_sumsq(gc)
        S.adamw_step(w, gc, m, v, ss, t, lr, betas, eps, wd, max_norm, 1.0 / world)
        torch.cuda.synchronize()
        e_n = abs(float(ss[0].sqrt()) / world - norms[-1]) / norms[-1]
        e_w = _rel(w, p.detach())
        print(f"flat AdamW step {t}: grad norm {norms[-1]:.4f} (rel-err {e_n:.1e}), parameters rel-err {e_w:.1e}")
        assert e_n < 1e-5 and e_w < 1e-6                        # fp32 arithmetic against an fp64 torch trajectory
    assert norms[0] > max_norm > norms[-1]                      # both branches of the clip were exercised


def test_sequence_trainer_flat_buffers_and_steps():
    """SequenceTrainer: gradients land in the flat buffer bit-identically to the plain autograd path, the parameters are
    views of the flat parameter buffer, and a few optimizer steps reduce a regression loss."""
    from ultrafnd_git_b200.seqfront import SequenceFrontEnd, SequenceTrainer
    streams, blocks = O.FAKESV_STREAMS, O.FAKESV_BLOCKS
    d_model, heads, B = 128, 2, 4
    lengths = {"text": 40, "frames": 83, "audio": 50, "c3d": 83}
    p = O.init_params(streams, blocks, d_model, seed=3)
    data = {k: v.cuda() for k, v in O.make_batch(streams, lengths, B, seed=4).items()}
    g = torch.Generator().manual_seed(5)
    target = {n: torch.randn(B, streams[n][1], generator=g).cuda() for n in streams}

    def loss_of(fe):
        out = fe(data)
        return sum(((out[n] - target[n]) ** 2).mean() for n in streams)

    fe_plain = SequenceFrontEnd(d_model, heads, streams, blocks).cuda(); fe_plain.load_state_dict(p)
    loss_of(fe_plain).backward()
    ref = {k: v.grad.clone() for k, v in fe_plain.named_parameters()}
    fe = SequenceFrontEnd(d_model, heads, streams, blocks).cuda(); fe.load_state_dict(p)
    tr = SequenceTrainer(fe, lr=2e-3, max_norm=5.0)
    assert sum(len(b) for b in fe.grad_order()) == len(ref) and tr.n == sum(v.numel() for v in ref.values())
    for k, q in fe.named_parameters():
        o, n = tr.offset[k]
        assert q.data_ptr() == tr.flat_w[o:o + n].data_ptr()
    losses = []
    for it in range(6):
        tr.zero_grad()
        loss = loss_of(fe)
        loss.backward()
        if it == 0:
            for k in ref:
                o, n = tr.offset[k]
                assert torch.equal(tr.flat_g[o:o + n].view(ref[k].shape), ref[k]), k      # same kernels, same order: bit-identical
        tr.step()
        losses.append(float(loss.detach()))
    torch.cuda.synchronize()
    fe.check_error()
    print("SequenceTrainer losses:", ", ".join(f"{x:.4f}" for x in losses), f"| grad norm {tr.grad_norm():.3f}")
    assert losses[-1] < 0.9 * losses[0] and all(math.isfinite(x) for x in losses)
