"""GPU tests of the sequence front-end (Tier B) against the SELF-ORACLE oracle/seq_oracle.py.

PARITY UNPINNED BY THE REFERENCE: the reference has no sequence attention / LayerNorm / token projection;
these tests pin the CUDA kernels to a plain-PyTorch restatement of the textbook operators, nothing more.
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


@pytest.mark.parametrize("M,N,K,extras", [(100, 64, 128, False), (300, 384, 4096, True), (4096, 1024, 1024, True),
                                           (37, 768, 256, True), (20000, 512, 512, False), (129, 1536, 512, True),
                                           # CTA-pair kernel (fnd_seq_gemm2.cuh: >= 74 tiles of 256 x 256): ragged M / N / K with
                                           # every epilogue extra, and the stress shape's two projections
                                           (19001, 320, 200, True), (18500, 576, 328, False), (32768, 1024, 1024, True), (16384, 3072, 1024, False)])
def test_seq_linear_matches_torch(M, N, K, extras):
    from ultrafnd_git_b200 import seq_ops as S
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).cuda().bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).cuda().bfloat16()
    bias = torch.randn(N, generator=g).cuda() if extras else None
    resid = torch.randn(M, N, generator=g).cuda().bfloat16() if extras else None
    err = S.new_err_flag(a.device)
    out = S.linear(a, w, bias, resid=resid, act=1 if extras else 0, err=err)
    of = torch.empty(M, N, device="cuda")
    S.linear(a, w, bias, resid=resid, act=1 if extras else 0, out_f32=of, want_bf16=False, err=err)
    ref = a.float() @ w.float().t()
    if extras:
        ref = F.gelu(ref + bias + resid.float())
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    e32, e16 = _rel(of, ref), _rel(out, ref)
    from ultrafnd_git_b200 import _lib
    print(f"pair clusters {_lib.load().fnd_seq_pair_clusters()}", end=" ")
    print(f"seq_linear M={M} N={N} K={K}: fp32-out rel-err {e32:.2e}, bf16-out rel-err {e16:.2e}")
    assert e32 < 1e-4            # same bf16 operands, fp32 accumulation: only summation order differs
    assert e16 < 5e-3            # + one bf16 rounding of the output


def test_seq_layernorm_and_pool_and_cast():
    from ultrafnd_git_b200 import seq_ops as S
    g = torch.Generator().manual_seed(5)
    for M, d in ((50, 256), (1000, 1024), (7, 2048), (33, 64)):
        x = (torch.randn(M, d, generator=g) * 3 + 1).cuda().bfloat16()
        w, b = (1 + 0.1 * torch.randn(d, generator=g)).cuda(), (0.1 * torch.randn(d, generator=g)).cuda()
        y = S.layernorm(x, w, b, 1e-5)
        ref = F.layer_norm(x.float(), (d,), w, b, 1e-5)
        e = _rel(y, ref)
        print(f"layernorm M={M} d={d}: rel-err {e:.2e}")
        assert e < 5e-3          # bf16 output rounding
    B, L, d = 5, 83, 320
    x = torch.randn(B, L, d, generator=g).cuda().bfloat16()
    m = (torch.rand(B, L, generator=g) < 0.6)
    m[0] = True; m[1] = False
    out = S.masked_mean_pool(x.view(B * L, d), B, L, mask=m.to(torch.uint8).cuda())
    ref = O.masked_mean(x.float().cpu(), m)
    assert torch.isfinite(out).all() and float(out[1].abs().max()) == 0.0     # no valid token -> 0 / 1e-6 = 0
    e = _rel(out, ref)
    print(f"masked_mean_pool rel-err {e:.2e}")
    assert e < 1e-5
    length = torch.tensor([83, 0, 40, 1, 64], dtype=torch.int32).cuda()
    out2 = S.masked_mean_pool(x.view(B * L, d), B, L, length=length)
    ref2 = O.masked_mean(x.float().cpu(), torch.arange(L)[None, :] < length.cpu()[:, None])
    assert _rel(out2, ref2) < 1e-5
    xf = torch.randn(1000, 768, generator=g).cuda()
    assert torch.equal(S.cast_bf16(xf), xf.bfloat16())


@pytest.mark.parametrize("B,H,Lq,Lk,mode", [(2, 2, 100, 70, "ragged"), (1, 4, 256, 192, "full"), (3, 1, 64, 300, "holes"),
                                             (2, 16, 130, 83, "ragged"), (2, 3, 83, 512, "ragged"), (1, 2, 1024, 512, "full")])
def test_coattn_forward_matches_self_oracle(B, H, Lq, Lk, mode):
    from ultrafnd_git_b200 import seq_ops as S
    d = 64 * H
    g = torch.Generator().manual_seed(B * 1000 + Lq + Lk)
    # fused [Q|K|V] layout on both sides, as the front-end produces it
    qkv_q = torch.randn(B * Lq, 3 * d, generator=g).cuda().bfloat16()
    qkv_k = torch.randn(B * Lk, 3 * d, generator=g).cuda().bfloat16()
    if mode == "full":
        mask = torch.ones(B, Lk, dtype=torch.bool)
    elif mode == "ragged":
        n = torch.randint(1, Lk + 1, (B,), generator=g)
        n[0] = Lk
        mask = torch.arange(Lk)[None, :] < n[:, None]
    else:
        mask = torch.rand(B, Lk, generator=g) < 0.5
        mask[:, 0] = True
        mask[B - 1] = False                      # one sample with NO valid key: output must be zeros
    length = (mask.int() * torch.arange(1, Lk + 1)[None, :]).amax(1).int().cuda()
    lse = torch.empty(B, H, Lq, device="cuda")
    err = S.new_err_flag(qkv_q.device)
    out = S.coattn_forward(qkv_q, qkv_k, qkv_k, B, H, Lq, Lk, q_col0=0, k_col0=d, v_col0=2 * d, kv_len=length,
                           kv_mask=mask.to(torch.uint8).cuda(), lse=lse, err=err)
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    q = qkv_q[:, :d].float().cpu().view(B, Lq, d)
    k = qkv_k[:, d:2 * d].float().cpu().view(B, Lk, d)
    v = qkv_k[:, 2 * d:].float().cpu().view(B, Lk, d)
    ref, ref_lse = O.attention_only(q, k, v, mask, H)
    o = out.float().cpu().view(B, Lq, d)
    e = _rel(o, ref)
    fin = torch.isfinite(ref_lse)
    e_lse = float((lse.cpu()[fin] - ref_lse[fin]).abs().max()) if fin.any() else 0.0
    print(f"coattn B={B} H={H} Lq={Lq} Lk={Lk} {mode}: out rel-err {e:.2e}, max |lse err| {e_lse:.2e}")
    assert torch.isfinite(o).all()
    assert e < BF16_TOL                          # bf16 P and bf16 output against an fp32 softmax
    assert e_lse < 2e-3
    if mode == "holes":
        assert float(o[B - 1].abs().max()) == 0.0 and bool(torch.isinf(lse[B - 1]).all())
    # prefix lengths alone (no byte mask) must give the same answer when the mask is a prefix mask
    if mode != "holes":
        out2 = S.coattn_forward(qkv_q, qkv_k, qkv_k, B, H, Lq, Lk, q_col0=0, k_col0=d, v_col0=2 * d, kv_len=length, err=err)
        assert torch.equal(out2, out)


def _frontend_pair(streams, blocks, d_model, heads, seed=7):
    from ultrafnd_git_b200.seqfront import SequenceFrontEnd
    p = O.init_params(streams, blocks, d_model, seed=seed)
    fe = SequenceFrontEnd(d_model, heads, streams, blocks).cuda()
    missing = fe.load_state_dict(p, strict=True)
    return p, fe


@pytest.mark.parametrize("case", ["fakesv", "fakesv_full", "stress_small"])
def test_sequence_frontend_matches_self_oracle(case):
    if case.startswith("fakesv"):
        streams, blocks = O.FAKESV_STREAMS, O.FAKESV_BLOCKS
        lengths, B, d_model, heads = {"text": 200, "frames": 83, "audio": 50, "c3d": 83}, 4, 256, 4
    else:
        streams, blocks = {"text": (768, 768, "text_features"), "frames": (4096, 512, "visual_features")}, (("text", "frames"),)
        lengths, B, d_model, heads = {"text": 1024, "frames": 512}, 2, 1024, 16
    p, fe = _frontend_pair(streams, blocks, d_model, heads)
    batch = O.make_batch(streams, lengths, B, seed=11, full=(case == "fakesv_full"))
    ref = O.forward(p, batch, streams, blocks, heads, return_states=True)
    got = fe({k: v.cuda() for k, v in batch.items()}, return_states=True)
    torch.cuda.synchronize()
    fe.check_error()
    worst = 0.0
    for name in streams:
        m = batch[name + "_mask"]
        st_ref = ref["state." + name] * m[..., None]
        st_got = got["state." + name].float().cpu() * m[..., None]
        e_state, e_pool, e_out = _rel(st_got, st_ref), _rel(got["pooled." + name], ref["pooled." + name]), _rel(got[name], ref[name])
        print(f"[{case}] {name}: state rel-err {e_state:.2e}, pooled {e_pool:.2e}, head output {e_out:.2e}")
        worst = max(worst, e_state, e_pool, e_out)
        assert torch.isfinite(got[name]).all()
    assert worst < BF16_TOL, worst               # north_star: rel-err <= 2e-2 in bf16


def test_cross_modal_transformer_accepts_sequences():
    """when feats[...] is 3-D the attached front-end pools it first; 2-D behaviour is untouched."""
    from ultrafnd_git_b200.modules import CrossModalTransformer
    from ultrafnd_git_b200.seqfront import SequenceFrontEnd
    torch.manual_seed(0)
    fusion = CrossModalTransformer(precision="bf16").eval()
    fe = SequenceFrontEnd(256, 4).cuda()
    fusion.attach_sequence_frontend(fe)
    B = 6
    g = torch.Generator().manual_seed(2)
    feats = {"text_features": torch.randn(B, 40, 768, generator=g).cuda(), "visual_features": torch.randn(B, 83, 4096, generator=g).cuda(),
             "audio_features": torch.randn(B, 50, 128, generator=g).cuda(), "temporal_features": torch.randn(B, 83, 4096, generator=g).cuda(),
             "gnn_feat": torch.randn(B, 128, generator=g).cuda(),
             "text_features_mask": (torch.arange(40)[None, :] < torch.randint(1, 41, (B, 1), generator=g)).cuda()}
    with torch.no_grad():
        out = fusion(feats)
        pooled = fe.forward_features(feats)
        out2 = fusion(pooled)
    assert out["fused"].shape == (B, 512) and out["logits"].shape == (B, 2)
    assert torch.equal(out["fused"], out2["fused"])
