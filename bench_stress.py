"""bench_stress.py — `bench.py --workload stress`: BASELINE.json configs[4], the sequence front-end (Tier B) at the stress
shape: 1024 text tokens x 512 frames, hidden 1024, 16 heads (d_k 64), per-GPU batch 32 (the 8-GPU sweep 32..512 is 4..64 per
GPU). SELF-ORACLE SCOPE: the reference has no sequence co-attention; nothing here is a reference-parity claim.

One "step" = one forward pass of the front-end over one batch: fp32 -> bf16 cast, per-modality Linear + LayerNorm, one
bidirectional co-attention block (fused [Q|K|V] projections, flash-style tcgen05 attention in both directions, output
projections with the residual in the epilogue, LayerNorm), masked mean-pool and the pooled heads. Full-length sequences
(no padding: padded keys would be SKIPPED by the kernel and inflate samples/s against the fixed FLOP count).

JSON line (same contract as the train workload): value = samples/s with inputs resident in HBM (CUDA-graph replay, L2
flushed between steps, CUDA events, max over ranks); e2e = the same through SequenceFrontEnd.forward with pinned-host
inputs (H2D of the fp32 features and D2H of the pooled outputs inside the timed region); roofline = the dominant kernel
(the [Q|K|V] projection GEMM) against the measured bf16 peak; coattn = the co-attention block alone (north_star's
"co-attn % of BF16 tensor peak": FLOPs of SURVEY.md §8d, 17.18 GFLOP per sample, over the block's time).
Data parallel: forward-only replicas, batch sharded by rank, no collective ("weak").
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

LT, LF, D_MODEL, HEADS = 1024, 512, 1024, 16
STREAMS = {"text": (768, 768, "text_features"), "frames": (4096, 512, "visual_features")}
BLOCKS = (("text", "frames"),)


def _median_ms(fn, reps=10, warm=2, flush=None):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def run_reference_arm(args):
    """CPU arm: the self-oracle (plain PyTorch fp32, all host threads) on a bounded sample of the same workload."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import seq_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    B = 2
    p = O.init_params(STREAMS, BLOCKS, D_MODEL)
    batch = O.make_batch(STREAMS, {"text": LT, "frames": LF}, B, full=True)
    steps = max(1, min(args.steps, 8))
    for _ in range(min(args.warmup, 1)):
        O.forward(p, batch, STREAMS, BLOCKS, HEADS)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.forward(p, batch, STREAMS, BLOCKS, HEADS)
    dt = time.perf_counter() - t0
    value = steps * B / dt
    print(json.dumps({
        "impl": "reference", "metric": "forward samples/sec (sequence front-end, stress shape)", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"stress: sequence front-end forward, {LT} text tokens x {LF} frames, hidden {D_MODEL}, {HEADS} heads, batch {B}",
                   "device": "host CPU", "note": "self-oracle (oracle/seq_oracle.py): the reference has no sequence co-attention"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{steps} forward passes at batch {B} (oracle/seq_oracle.py, torch CPU ops)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}), flush=True)


def run_b200_arm(args):
    import torch.distributed as dist
    import bench
    from ultrafnd_git_b200 import seq_ops as S
    from ultrafnd_git_b200.seqfront import SequenceFrontEnd, SequenceTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    d, H = D_MODEL, HEADS

    torch.manual_seed(42)
    fe = SequenceFrontEnd(d, H, STREAMS, BLOCKS).to(dev)
    # flat-buffer optimizer + bucketed gradient all-reduce (re-points the parameters at its flat buffer: before any capture)
    trainer = SequenceTrainer(fe, lr=1e-4, max_norm=5.0)
    g = torch.Generator().manual_seed(1234 + rank)
    pool = 2
    host = [{"text": torch.randn(B, LT, 768, generator=g).pin_memory(), "frames": torch.randn(B, LF, 4096, generator=g).pin_memory()}
            for _ in range(pool)]
    resident = [{k: v.to(dev) for k, v in hb.items()} for hb in host]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- timed region 1: forward with inputs resident in HBM, one CUDA graph per input set ----
    graphs, outs = [], []
    for i in range(pool):
        with torch.no_grad():                         # forward-only legs run without autograd (two-stream blocks)
            fe(resident[i])                           # warm-up (bf16 weight shadows, kernel attributes)
        torch.cuda.synchronize()
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph), torch.no_grad():
            o = fe(resident[i])
        graphs.append(gph); outs.append(o)
    for i in range(W):
        graphs[i % pool].replay()
    barrier()
    fe.check_error()
    sampler = bench.ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    for i in range(K):
        flush_buf.zero_()
        evs[i][0].record()
        graphs[i % pool].replay()
        evs[i][1].record()
    barrier()
    total_ms = sum(a.elapsed_time(b) for a, b in evs)
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = world * B * K / (total_ms / 1e3)

    # ---- timed region 2: end to end through the public API (pinned host -> H2D -> forward -> pooled outputs D2H) ----
    dev_in = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]
    out_host = {n: torch.zeros(B, s[1]).pin_memory() for n, s in STREAMS.items()}
    h2d_bytes = sum(v.numel() * 4 for v in host[0].values())
    d2h_bytes = sum(v.numel() * 4 for v in out_host.values())
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)
    h2d_done = [torch.cuda.Event() for _ in range(2)]
    in_free = [torch.cuda.Event() for _ in range(2)]
    for e in in_free:
        e.record(main)

    def e2e_step(i):
        s = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(in_free[s])
            for k in dev_in[s]:
                dev_in[s][k].copy_(host[i % pool][k], non_blocking=True)
            h2d_done[s].record(copy_stream)
        main.wait_event(h2d_done[s])
        with torch.no_grad():
            o = fe(dev_in[s])
        in_free[s].record(main)
        for n in out_host:
            out_host[n].copy_(o[n], non_blocking=True)

    KE = max(4, min(K, 20))
    for i in range(2):
        e2e_step(i)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    dbg_t = []
    for i in range(KE):
        t0 = time.perf_counter()
        e2e_step(2 + i)
        dbg_t.append(time.perf_counter() - t0)
    e1.record()
    if os.environ.get("FND_STRESS_DEBUG"):
        print("e2e host enqueue ms per step:", [round(1e3 * x, 2) for x in dbg_t], file=sys.stderr)
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_ms = float(t.item())
    e2e_value = world * B * KE / (e2e_ms / 1e3)
    # ---- timed region 3: TRAINING steps on every rank (forward with activations kept, fused backward, bucketed NCCL
    #      all-reduce launched as each gradient bucket completes, gradient norm, clip + AdamW over the flat buffer) ----
    gw = torch.Generator(device="cuda").manual_seed(7 + rank)
    wts = {n: torch.randn(B, s_[1], device=dev, generator=gw) for n, s_ in STREAMS.items()}

    def train_step(i):
        trainer.zero_grad()
        o = fe(resident[i % pool])
        sum((o[n] * wts[n]).sum() for n in STREAMS).backward()
        trainer.step()
    KT = max(3, min(K, 10))
    for i in range(3):
        train_step(i)
    barrier()
    tev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(KT)]
    for i in range(KT):
        flush_buf.zero_()
        tev[i][0].record()
        train_step(i)
        tev[i][1].record()
    barrier()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in tev)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    train_ms = float(t.item()) / KT
    clocks = sampler.stop() if rank == 0 else None
    fe.check_error()
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the co-attention block alone (resident bf16 states), as one graph, and kernel by kernel ----
    Wt = fe._weights()
    blk = fe.blocks[0]
    err = S.new_err_flag(dev)
    gg = torch.Generator(device="cuda").manual_seed(0)
    xt = torch.randn(B * LT, d, device=dev, generator=gg).bfloat16()
    xf = torch.randn(B * LF, d, device=dev, generator=gg).bfloat16()
    qkv_t = torch.empty(B * LT, 3 * d, device=dev, dtype=torch.bfloat16)
    qkv_f = torch.empty(B * LF, 3 * d, device=dev, dtype=torch.bfloat16)
    at, yt, zt = (torch.empty(B * LT, d, device=dev, dtype=torch.bfloat16) for _ in range(3))
    af, yf, zf = (torch.empty(B * LF, d, device=dev, dtype=torch.bfloat16) for _ in range(3))
    ops = {
        "in_proj_text": (lambda: S.linear(xt, Wt["blocks.0.a.in_proj.weight"], blk.a.in_proj.bias, out=qkv_t, err=err), 2.0 * B * LT * 3 * d * d),
        "in_proj_frames": (lambda: S.linear(xf, Wt["blocks.0.b.in_proj.weight"], blk.b.in_proj.bias, out=qkv_f, err=err), 2.0 * B * LF * 3 * d * d),
        "attn_text_from_frames": (lambda: S.coattn_forward(qkv_t, qkv_f, qkv_f, B, H, LT, LF, 0, d, 2 * d, out=at, err=err), 4.0 * B * LT * LF * d),
        "attn_frames_from_text": (lambda: S.coattn_forward(qkv_f, qkv_t, qkv_t, B, H, LF, LT, 0, d, 2 * d, out=af, err=err), 4.0 * B * LT * LF * d),
        "out_proj_text": (lambda: S.linear(at, Wt["blocks.0.a.out_proj.weight"], blk.a.out_proj.bias, resid=xt, out=yt, err=err), 2.0 * B * LT * d * d),
        "out_proj_frames": (lambda: S.linear(af, Wt["blocks.0.b.out_proj.weight"], blk.b.out_proj.bias, resid=xf, out=yf, err=err), 2.0 * B * LF * d * d),
        "layernorm_text": (lambda: S.layernorm(yt, blk.a.ln.weight, blk.a.ln.bias, out=zt), 0.0),
        "layernorm_frames": (lambda: S.layernorm(yf, blk.b.ln.weight, blk.b.ln.bias, out=zf), 0.0),
    }

    def layer():
        for fn, _ in ops.values():
            fn()
    side = torch.cuda.Stream(dev)

    def layer_two_streams():
        # the two directions of the block are independent except that each attention needs BOTH [Q|K|V] projections: text side
        # on the current stream, frames side on `side` (forked / joined by events, so the whole block is still ONE graph)
        cur = torch.cuda.current_stream(dev)
        f = {n: fn for n, (fn, _) in ops.items()}
        side.wait_stream(cur)
        f["in_proj_text"]()
        ev_t = torch.cuda.Event(); ev_t.record(cur)
        with torch.cuda.stream(side):
            f["in_proj_frames"]()
            ev_f = torch.cuda.Event(); ev_f.record(side)
            side.wait_event(ev_t)
        cur.wait_event(ev_f)
        f["attn_text_from_frames"]()
        f["out_proj_text"]()
        f["layernorm_text"]()
        with torch.cuda.stream(side):
            f["attn_frames_from_text"]()
            f["out_proj_frames"]()
            f["layernorm_frames"]()
        cur.wait_stream(side)

    layer()
    torch.cuda.synchronize()
    lg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(lg):
        layer()
    layer_ms_one = _median_ms(lg.replay, reps=20, warm=3, flush=flush_buf)
    lg2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(lg2):
        layer_two_streams()
    layer_ms_two = _median_ms(lg2.replay, reps=20, warm=3, flush=flush_buf)
    layer_ms = min(layer_ms_one, layer_ms_two)
    kern = {n: _median_ms(fn, reps=10, warm=2) for n, (fn, _) in ops.items()}
    torch.cuda.synchronize()
    assert int(err.item()) == 0

    # ---- fused attention backward alone, and one TRAINING step of the front-end (forward with activations kept +
    #      fused backward through SequenceFrontEnd's autograd node; eager launches: the backward allocates its buffers) ----
    d_at = torch.randn(B * LT, d, device=dev, generator=gg).bfloat16()
    lse_t = torch.empty(B, H, LT, device=dev, dtype=torch.float32)
    S.coattn_forward(qkv_t, qkv_f, qkv_f, B, H, LT, LF, 0, d, 2 * d, out=at, lse=lse_t, err=err)
    dqkv_t, dqkv_f = torch.empty_like(qkv_t), torch.empty_like(qkv_f)
    bwd_ms = _median_ms(lambda: S.coattn_backward(qkv_t, qkv_f, qkv_f, at, d_at, lse_t, B, H, LT, LF, dqkv_t, dqkv_f, dqkv_f,
                                                  q_col0=0, k_col0=d, v_col0=2 * d, dq_col0=0, dk_col0=d, dv_col0=2 * d, err=err),
                        reps=10, warm=2)
    torch.cuda.synchronize()
    fe.check_error()
    assert int(err.item()) == 0

    peaks = bench.measured_peaks()
    layer_flops = B * 2.0 * (2.0 * (2.0 * LT * d * d + 2.0 * LF * d * d + 2.0 * LT * LF * d))
    peak = peaks["bf16_tflops"] * 1e12
    kernels = {n: {"ms": round(ms, 5), "tflops": (ops[n][1] / (ms / 1e3) / 1e12) if ops[n][1] else None,
                   "frac_of_measured_bf16_peak": (ops[n][1] / (ms / 1e3) / peak) if ops[n][1] else None} for n, ms in kern.items()}
    dom = max((n for n in kern if ops[n][1]), key=lambda n: kern[n])
    ach = ops[dom][1] / (kern[dom] / 1e3) / 1e12
    traffic = None
    tr = os.path.join(ROOT, "profiles", "r02_seq_traffic.json")
    if os.path.exists(tr):
        with open(tr) as f:
            traffic = json.load(f).get(dom)
    roofline = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"], "traffic": traffic,
                "traffic_source": "committed ncu --set full capture (profiles/r02_seq_traffic.json)" if traffic else None,
                "peak_source": (peaks["source"] + " (MEASURED_PEAKS.json, burst: kernel timed alone)") if peaks["source"] == "measured" else "fallback",
                "algorithmic_flops": ops[dom][1], "kernel_ms": kern[dom]}
    att_fl = ops["attn_text_from_frames"][1] + ops["attn_frames_from_text"][1]
    att_ms = kern["attn_text_from_frames"] + kern["attn_frames_from_text"]
    coattn = {"layer_ms": layer_ms, "layer_flops": layer_flops, "layer_tflops": layer_flops / (layer_ms / 1e3) / 1e12,
              "frac_of_measured_bf16_peak": layer_flops / (layer_ms / 1e3) / peak,
              "attention_kernels_frac": att_fl / (att_ms / 1e3) / peak, "projection_gemms_frac":
              (layer_flops - att_fl) / ((sum(kern[n] for n in kern if n.startswith(("in_proj", "out_proj")))) / 1e3) / peak,
              "layer_ms_one_stream": layer_ms_one, "layer_ms_two_streams": layer_ms_two,
              "target": 0.60, "note": "one bidirectional co-attention block (in/out projections + scores + PV + LayerNorm), one CUDA graph, "
                                      "L2 flushed between replays; FLOPs per SURVEY.md §8d (17.18 GFLOP / sample)"}

    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import seq_oracle as O
        torch.set_num_threads(os.cpu_count() or 1)
        p = O.init_params(STREAMS, BLOCKS, D_MODEL)
        cb = O.make_batch(STREAMS, {"text": LT, "frames": LF}, 2, full=True)
        O.forward(p, cb, STREAMS, BLOCKS, HEADS)
        t0 = time.perf_counter()
        n = 0
        while n < 8 and time.perf_counter() - t0 < 12.0:
            O.forward(p, cb, STREAMS, BLOCKS, HEADS)
            n += 1
        dt = time.perf_counter() - t0
        cpu_baseline = {"value": 2 * n / dt, "unit": "samples/s", "cores": torch.get_num_threads(), "kind": "port",
                        "sample": f"{n} forward passes at batch 2 in {dt:.1f} s (self-oracle oracle/seq_oracle.py on the host CPU)"}

    attn_fwd_flops = ops["attn_text_from_frames"][1]
    kernels["attn_bwd_text_from_frames"] = {"ms": round(bwd_ms, 5), "tflops": 2.5 * attn_fwd_flops / (bwd_ms / 1e3) / 1e12,
                                            "frac_of_measured_bf16_peak": 2.5 * attn_fwd_flops / (bwd_ms / 1e3) / peak,
                                            "note": "row prologue + dQ kernel + dK/dV kernel; algorithmic FLOPs = 2.5x the forward's "
                                                    "(the two-kernel split executes 3.5x: S and dP are recomputed in both)"}
    fl_fwd = fe.flops({"text": LT, "frames": LF}, B)
    train = {"ms_per_step": train_ms, "value": world * B / (train_ms / 1e3), "unit": "samples/s", "n_gpus": world, "steps": KT,
             "tflops_per_gpu": 3.0 * fl_fwd / (train_ms / 1e3) / 1e12, "frac_of_measured_bf16_peak": 3.0 * fl_fwd / (train_ms / 1e3) / peak,
             "grad_allreduce": ("NCCL, one all-reduce per gradient bucket (heads | block | embeddings) issued as the backward completes it"
                                if world > 1 else None),
             "note": "whole-job training throughput, max over ranks: forward (activations kept) + fused backward of every front-end "
                     "parameter (attention backward, LayerNorm / pool backward, dgrad + token-major wgrad GEMMs, bias column sums) + gradient "
                     "norm + clip + AdamW over the flat buffer (SequenceTrainer), eager launches, L2 flushed; FLOPs counted as 3x forward"}
    nlaunch = 2 * 3 + 8 + 2 * 2                       # cast+embed+LN per stream, block, pool+head per stream
    fl_total = fe.flops({"text": LT, "frames": LF}, B)
    line = {
        "metric": "forward samples/sec (sequence front-end, stress shape); co-attn % of BF16 tensor peak", "value": value,
        "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": total_ms / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"stress: sequence front-end forward, {LT} text tokens x 768, {LF} frames x 4096, hidden {d}, {H} heads, "
                               f"batch {B}/GPU, full-length sequences, random-init weights", "batch_per_gpu": B, "global_batch": B * world,
                   "parallelism": f"dp{world} (forward-only replicas, no collective)", "cuda_graph": True,
                   "l2": "flushed between timed steps (256 MiB write)", "parity": "self-oracle (oracle/seq_oracle.py); unpinned by the reference"},
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                "ms_per_step": e2e_ms / KE, "note": "bound by the H2D copy of the fp32 features (PCIe)"},
        "gpu_launches": nlaunch * K, "clocks": clocks, "roofline": roofline, "coattn": coattn, "coattn_tensor_frac": coattn["frac_of_measured_bf16_peak"],
        "cpu_baseline": cpu_baseline, "kernels": kernels, "train": train,
        "step_tensor_frac": fl_total / (total_ms / K / 1e3) / peak,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
