// fnd_api.cu — C-ABI of libfnd_b200.so (see include/fnd_b200.h for the contract of every entry point).
#include "fnd_engine.h"
#include <new>

using namespace fnd;

#define FND_CUDA_OK(expr)                                        \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return -1000 - static_cast<int>(_e);  \
  } while (0)
#define FND_OK(expr)          \
  do {                        \
    int _r = (expr);          \
    if (_r) return _r;        \
  } while (0)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// "slots" workspace buffer (16384 floats): [0, 8192) fused-step norm slots (wgrad CTAs, then finalize CTAs);
// [8192, 12288) sumsq_kernel slots; [12288, 16384) scratch for the module-level finalize launches.
static const int kSlotSumsq = 8192;
static const int kSlotScratch = 12288;

// ---------------------------------------------------------------------------------------------------
// table construction (runs once per plan at bind time)
// ---------------------------------------------------------------------------------------------------
static int build_tables(Plan& P) {
  TableBuilder tb(P);
  const int B = P.B, H = P.H;
  const fnd_dims& d = P.d;
  const char* pn[5] = {"text_proj", "audio_proj", "visual_proj", "temporal_proj", "gnn_proj"};
  // Q layout: input column (in P) and stacked weight of each q/k/v group
  struct QGroup { int in_col; int q_col; int n; const char* first; };
  const QGroup qg[4] = {{0, 0, 2 * H, "fusion.attn_tv.q"}, {2 * H, 2 * H, 3 * H, "fusion.attn_tv.k"},
                        {H, 5 * H, 2 * H, "fusion.attn_ta.k"}, {3 * H, 7 * H, 2 * H, "fusion.attn_vu.k"}};
  const int catw = P.nslots * H;
  const int p0w = H + d.aux_dim;

  // ---------------- forward ----------------
  P.fwd_proj.kind = 0;
  for (int i = 0; i < P.nmod; ++i) {
    EpiParams e = epi_zero();
    const std::string key = std::string("fusion.") + pn[i];
    e.bias = P.W(key + ".bias");
    e.out_f32 = P.buf<float>("P") + i * H; e.f32_pitch = 5 * H;
    if (i < 4) {
      e.out_hi = P.buf<__nv_bfloat16>("pbf_hi") + i * H;
      e.out_lo = P.buf<__nv_bfloat16>("pbf_lo") ? P.buf<__nv_bfloat16>("pbf_lo") + i * H : nullptr;
      e.bf_pitch = 5 * H;
    }
    FND_OK(add_problem(P, P.fwd_proj, tb.act("xbf", P.xoff[i], P.dsum, false), tb.weight(key + ".weight", P.xdim[i], false),
                       B, H, P.xdim[i], P.cfg_proj.bn, 1, e, "", 1));
  }
  P.fwd_qkv.kind = 0;
  for (int g = 0; g < 4; ++g) {
    EpiParams e = epi_zero();
    e.bias = P.W(std::string(qg[g].first) + ".bias");
    e.out_f32 = P.buf<float>("Q") + qg[g].q_col; e.f32_pitch = 9 * H;
    FND_OK(add_problem(P, P.fwd_qkv, tb.act("pbf", qg[g].in_col, 5 * H, false),
                       tb.weight(std::string(qg[g].first) + ".weight", H, false), B, qg[g].n, H, P.cfg_qkv.bn, 1, e, "", 1));
  }
  {
    P.fwd_f0.kind = 0;
    EpiParams e = epi_zero();
    e.bias = P.W("fusion.fuse_mlp.0.bias");
    e.out_pre = P.buf<float>("z_f0"); e.pre_pitch = 2 * H;
    e.act = 1; e.drop_p = d.fusion_dropout; e.drop_stream = kStreamFuse0;
    e.out_hi = P.buf<__nv_bfloat16>("h1_hi"); e.out_lo = P.buf<__nv_bfloat16>("h1_lo"); e.bf_pitch = 2 * H;
    FND_OK(add_problem(P, P.fwd_f0, tb.act("fused_cat", 0, catw, false), tb.weight("fusion.fuse_mlp.0.weight", catw, false),
                       B, 2 * H, catw, P.cfg_f0.bn, P.cfg_f0.splits, e, "f0", 1));
  }
  {
    P.fwd_f1.kind = 0;
    EpiParams e = epi_zero();
    e.bias = P.W("fusion.fuse_mlp.3.bias");
    e.out_pre = P.buf<float>("z_f1"); e.pre_pitch = H;
    e.act = 1; e.drop_p = d.fusion_dropout; e.drop_stream = kStreamFuse1;
    e.out_f32 = P.buf<float>("fused"); e.f32_pitch = H;
    e.out_hi = P.buf<__nv_bfloat16>("fusedbf_hi"); e.out_lo = P.buf<__nv_bfloat16>("fusedbf_lo"); e.bf_pitch = H;
    FND_OK(add_problem(P, P.fwd_f1, tb.act("h1", 0, 2 * H, false), tb.weight("fusion.fuse_mlp.3.weight", 2 * H, false),
                       B, H, 2 * H, P.cfg_f1.bn, P.cfg_f1.splits, e, "f1", 1));
  }
  {
    P.fwd_p0.kind = 0;
    EpiParams e = epi_zero();
    e.bias = P.W("clf.pre.0.bias");
    if (d.aux_dim) {
      e.aux = P.buf<float>("aux");
      e.aux_w = P.W("clf.pre.0.weight") + H; e.aux_w_pitch = p0w;
    }
    e.out_pre = P.buf<float>("z_p0"); e.pre_pitch = H;
    e.act = 1; e.drop_p = d.clf_dropout; e.drop_stream = kStreamPre0;
    e.out_hi = P.buf<__nv_bfloat16>("xp1_hi"); e.out_lo = P.buf<__nv_bfloat16>("xp1_lo"); e.bf_pitch = H;
    FND_OK(add_problem(P, P.fwd_p0, tb.act("fusedbf", 0, H, false), tb.weight_rp(false), B, H, H, P.cfg_pre.bn, P.cfg_pre.splits, e, "p0", 1));
  }
  {
    P.fwd_p1.kind = 0;
    EpiParams e = epi_zero();
    e.bias = P.W("clf.pre.3.bias");
    e.out_pre = P.buf<float>("z_p1"); e.pre_pitch = H;
    e.act = 1; e.drop_p = d.clf_dropout; e.drop_stream = kStreamPre1;
    e.out_f32 = P.buf<float>("h"); e.f32_pitch = H;
    e.out_hi = P.buf<__nv_bfloat16>("hbf_hi"); e.out_lo = P.buf<__nv_bfloat16>("hbf_lo"); e.bf_pitch = H;
    FND_OK(add_problem(P, P.fwd_p1, tb.act("xp1", 0, H, false), tb.weight("clf.pre.3.weight", H, false), B, H, H, P.cfg_pre.bn, P.cfg_pre.splits, e, "p1", 1));
  }

  // ---------------- dgrad (A = dY K-major, B = W viewed MN-major) ----------------
  {
    P.dg_p1.kind = 1;
    EpiParams e = epi_zero();
    e.gate_z = P.buf<float>("z_p0"); e.gate_pitch = H; e.gate_p = d.clf_dropout; e.gate_stream = kStreamPre0;
    e.out_hi = P.buf<__nv_bfloat16>("dz_p0_hi"); e.out_lo = P.buf<__nv_bfloat16>("dz_p0_lo"); e.bf_pitch = H;
    FND_OK(add_problem(P, P.dg_p1, tb.act("dz_p1", 0, H, false), tb.weight("clf.pre.3.weight", H, true), B, H, H, P.cfg_dg_pre.bn,
                       P.cfg_dg_pre.splits, e, "dgp1", 1));
  }
  for (int fusedpath = 0; fusedpath < 2; ++fusedpath) {
    GemmTable& T = fusedpath ? P.dg_p0_fused : P.dg_p0_split;
    T.kind = 1;
    EpiParams e = epi_zero();
    if (fusedpath) {
      e.gate_z = P.buf<float>("z_f1"); e.gate_pitch = H; e.gate_p = d.fusion_dropout; e.gate_stream = kStreamFuse1;
      e.out_hi = P.buf<__nv_bfloat16>("dz_f1_hi"); e.out_lo = P.buf<__nv_bfloat16>("dz_f1_lo"); e.bf_pitch = H;
    } else {
      e.out_f32 = P.buf<float>("dfused"); e.f32_pitch = H;
    }
    FND_OK(add_problem(P, T, tb.act("dz_p0", 0, H, false), tb.weight_rp(true), B, H, H, P.cfg_dg_pre.bn, P.cfg_dg_pre.splits, e,
                       fusedpath ? "dgp0f" : "dgp0s", 1));
  }
  {
    P.dg_f1.kind = 1;
    EpiParams e = epi_zero();
    e.gate_z = P.buf<float>("z_f0"); e.gate_pitch = 2 * H; e.gate_p = d.fusion_dropout; e.gate_stream = kStreamFuse0;
    e.out_hi = P.buf<__nv_bfloat16>("dz_f0_hi"); e.out_lo = P.buf<__nv_bfloat16>("dz_f0_lo"); e.bf_pitch = 2 * H;
    FND_OK(add_problem(P, P.dg_f1, tb.act("dz_f1", 0, H, false), tb.weight("fusion.fuse_mlp.3.weight", 2 * H, true),
                       B, 2 * H, H, P.cfg_dg_f1.bn, P.cfg_dg_f1.splits, e, "dgf1", 1));
  }
  {
    P.dg_f0.kind = 1;
    EpiParams e = epi_zero();
    e.out_f32 = P.buf<float>("dcat"); e.f32_pitch = catw;
    FND_OK(add_problem(P, P.dg_f0, tb.act("dz_f0", 0, 2 * H, false), tb.weight("fusion.fuse_mlp.0.weight", catw, true),
                       B, catw, 2 * H, P.cfg_dg_f0.bn, P.cfg_dg_f0.splits, e, "dgf0", 1));
  }
  P.dg_qkv.kind = 1;
  for (int g = 0; g < 4; ++g) {
    EpiParams e = epi_zero();
    e.add_in = P.buf<float>("dPdirect") + qg[g].in_col; e.add_pitch = 5 * H;
    e.out_hi = P.buf<__nv_bfloat16>("dP_hi") + qg[g].in_col;
    e.out_lo = P.buf<__nv_bfloat16>("dP_lo") ? P.buf<__nv_bfloat16>("dP_lo") + qg[g].in_col : nullptr;
    e.bf_pitch = 5 * H;
    FND_OK(add_problem(P, P.dg_qkv, tb.act("dQ", qg[g].q_col, 9 * H, false),
                       tb.weight(std::string(qg[g].first) + ".weight", H, true), B, H, qg[g].n, P.cfg_dg_qkv.bn,
                       P.cfg_dg_qkv.splits, e, "dgqkv" + std::to_string(g), 1));
  }

  // ---------------- wgrad (dW[N_out, K_in] = dY^T X; both operands MN-major; contraction = batch) -------------
  auto wg = [&](GemmTable& T, const Operand& dY, const Operand& X, int n_out, int k_in, float* dst, int pitch) -> int {
    EpiParams e = epi_zero();
    e.out_f32 = dst; e.f32_pitch = pitch;
    // dAraw feeds finalize jobs that may run as trailing CTAs of the same launch: its tiles announce completion
    if (dst == P.buf<float>("dAraw")) e.done_ctr = &P.state()->wg_done;
    // bf16 mirror of the weight gradients (same element offsets as the gradient arena); pre.0.weight's rows (pitch 514)
    // are not 16-byte aligned in bf16 and stay fp32-only
    if (P.grads_bf && dst >= P.grads && dst < P.grads + P.L.n_hot && (pitch & 7) == 0) e.out_bf = P.grads_bf + (dst - P.grads);
    // eligible for the data-parallel fused push (routed store-only epilogue): a GEMM weight in the arena with 16-byte rows
    if (dst >= P.grads && dst < P.grads + P.L.n_hot && (pitch & 63) == 0 && (k_in & 63) == 0) e.routed = 1;
    return add_problem(P, T, dY, X, n_out, k_in, B, 128, 1, e, "", 0);
  };
  const int dA_tiles = ceil_div(P.TD + 2, kGemmBM) * ceil_div(H, 128);
  auto build_wg = [&](GemmTable& T, bool clf, bool fus) -> int {
    T.kind = 2;
    if (clf) {
      FND_OK(wg(T, tb.act("dz_p1", 0, H, true), tb.act("xp1", 0, H, true), H, H, P.G("clf.pre.3.weight"), H));
      FND_OK(wg(T, tb.act("dz_p0", 0, H, true), tb.act("fusedbf", 0, H, true), H, H, P.G("clf.pre.0.weight"), p0w));
      FND_OK(wg(T, tb.act("dFbf", 0, kDFCols, true), tb.act("hbf", 0, H, true), P.TD + 2, H, P.buf<float>("dAraw"), H));
    }
    if (fus) {
      FND_OK(wg(T, tb.act("dz_f1", 0, H, true), tb.act("h1", 0, 2 * H, true), H, 2 * H, P.G("fusion.fuse_mlp.3.weight"), 2 * H));
      FND_OK(wg(T, tb.act("dz_f0", 0, 2 * H, true), tb.act("fused_cat", 0, catw, true), 2 * H, catw,
                P.G("fusion.fuse_mlp.0.weight"), catw));
      for (int g = 0; g < 4; ++g)
        FND_OK(wg(T, tb.act("dQ", qg[g].q_col, 9 * H, true), tb.act("pbf", qg[g].in_col, 5 * H, true), qg[g].n, H,
                  P.G(std::string(qg[g].first) + ".weight"), H));
      for (int i = 0; i < P.nmod; ++i)
        FND_OK(wg(T, tb.act("dP", i * H, 5 * H, true), tb.act("xbf", P.xoff[i], P.dsum, true), H, P.xdim[i],
                  P.G(std::string("fusion.") + pn[i] + ".weight"), P.xdim[i]));
    }
    return 0;
  };
  {
    // data-parallel overlap: the early range (fuse_mlp.3, fuse_mlp.0), and everything else
    P.wg_early.kind = 2;
    FND_OK(wg(P.wg_early, tb.act("dz_f1", 0, H, true), tb.act("h1", 0, 2 * H, true), H, 2 * H, P.G("fusion.fuse_mlp.3.weight"), 2 * H));
    FND_OK(wg(P.wg_early, tb.act("dz_f0", 0, 2 * H, true), tb.act("fused_cat", 0, catw, true), 2 * H, catw,
              P.G("fusion.fuse_mlp.0.weight"), catw));
    GemmTable& T = P.wg_rest;
    T.kind = 2;
    FND_OK(wg(T, tb.act("dz_p1", 0, H, true), tb.act("xp1", 0, H, true), H, H, P.G("clf.pre.3.weight"), H));
    FND_OK(wg(T, tb.act("dz_p0", 0, H, true), tb.act("fusedbf", 0, H, true), H, H, P.G("clf.pre.0.weight"), p0w));
    FND_OK(wg(T, tb.act("dFbf", 0, kDFCols, true), tb.act("hbf", 0, H, true), P.TD + 2, H, P.buf<float>("dAraw"), H));
    for (int g = 0; g < 4; ++g)
      FND_OK(wg(T, tb.act("dQ", qg[g].q_col, 9 * H, true), tb.act("pbf", qg[g].in_col, 5 * H, true), qg[g].n, H,
                P.G(std::string(qg[g].first) + ".weight"), H));
    for (int i = 0; i < P.nmod; ++i)
      FND_OK(wg(T, tb.act("dP", i * H, 5 * H, true), tb.act("xbf", P.xoff[i], P.dsum, true), H, P.xdim[i],
                P.G(std::string("fusion.") + pn[i] + ".weight"), P.xdim[i]));
  }
  FND_OK(build_wg(P.wg_all, true, true));
  FND_OK(build_wg(P.wg_clf, true, false));
  FND_OK(build_wg(P.wg_fus, false, true));

  // ---------------- finalize job tables ----------------
  auto job = [&](FinTable& T, int type, int rows, int cols, int src_pitch, const float* f32, const __nv_bfloat16* hi,
                 const __nv_bfloat16* lo, const float* aux, float* dst, int dst_pitch, float scale) {
    FinJob j;
    memset(&j, 0, sizeof(j));
    j.type = type; j.rows = rows; j.cols = cols; j.src_pitch = src_pitch;
    j.src_f32 = f32; j.src_hi = hi; j.src_lo = lo; j.aux = aux; j.dst = dst; j.dst_pitch = dst_pitch; j.scale = scale;
    j.cta_count = (type == kJobSoftmaxBwd) ? rows : ceil_div(cols, 64);
    j.want_norm = 1;
    T.host.push_back(j);
  };
  auto bfj = [&](FinTable& T, const std::string& name, int col, int cols, int pitch, float* dst) {
    const __nv_bfloat16* lo = P.buf<__nv_bfloat16>(name + "_lo");
    job(T, kJobColsumBF16, B, cols, pitch, nullptr, P.buf<__nv_bfloat16>(name + "_hi") + col, lo ? lo + col : nullptr,
        nullptr, dst, 0, 1.0f);
  };
  auto build_fin = [&](FinTable& T, bool clf, bool fus) {
    if (clf) {
      bfj(T, "dz_p1", 0, H, H, P.G("clf.pre.3.bias"));
      bfj(T, "dz_p0", 0, H, H, P.G("clf.pre.0.bias"));
      if (d.aux_dim) {
        const __nv_bfloat16* lo = P.buf<__nv_bfloat16>("dz_p0_lo");
        job(T, kJobColsumAux, B, H, H, nullptr, P.buf<__nv_bfloat16>("dz_p0_hi"), lo, P.buf<float>("aux"),
            P.G("clf.pre.0.weight") + H, p0w, 1.0f);
      }
      job(T, kJobColsumF32, B, P.TD, kDFCols, P.buf<float>("dF"), nullptr, nullptr, nullptr, P.G("clf.node.trees.0.thresh.0"), 0, -1.0f);
      job(T, kJobColsumF32, B, 2, kDFCols, P.buf<float>("dF") + P.TD, nullptr, nullptr, nullptr, P.G("clf.bypass.bias"), 0, 1.0f);
      job(T, kJobColsumF32, B, d.trees * P.leaves * 2, d.trees * P.leaves * 2, P.buf<float>("leafc"), nullptr, nullptr, nullptr,
          P.G("clf.node.trees.0.leaf_logits"), 0, 1.0f);
      job(T, kJobColsumF32, 1, 2 * H, 2 * H, P.buf<float>("dAraw") + static_cast<size_t>(P.TD) * H, nullptr, nullptr, nullptr,
          P.G("clf.bypass.weight"), 0, 1.0f);
      T.host.back().wait_ctr = &P.state()->wg_done; T.host.back().wait_count = dA_tiles;
      job(T, kJobSoftmaxBwd, P.TD, H, H, P.buf<float>("dAraw"), nullptr, nullptr, P.buf<float>("alpha"),
          P.G("clf.node.trees.0.gates.0"), H, 1.0f);
      T.host.back().wait_ctr = &P.state()->wg_done; T.host.back().wait_count = dA_tiles;
    }
    if (fus) {
      bfj(T, "dP", 0, P.nmod * H, 5 * H, P.G("fusion.text_proj.bias"));
      bfj(T, "dQ", 0, 9 * H, 9 * H, P.G("fusion.attn_tv.q.bias"));
      bfj(T, "dz_f0", 0, 2 * H, 2 * H, P.G("fusion.fuse_mlp.0.bias"));
      bfj(T, "dz_f1", 0, H, H, P.G("fusion.fuse_mlp.3.bias"));
      job(T, kJobColsumF32, P.n_asm_ctas, 3 * P.L.evstride, 3 * P.L.evstride, P.buf<float>("ev_partial"), nullptr, nullptr,
          nullptr, P.G("fusion.attn_tv.evidence_proj.0.weight"), 0, 1.0f);
    }
    if (clf && fus) {   // fused step only: mean loss (the stand-alone finalize launches compute it in their last CTA)
      job(T, kJobLossMean, B, 1, 1, P.buf<float>("loss_row"), nullptr, nullptr, nullptr, P.loss_mirror, P.loss_mirror ? P.loss_ring : 0, 1.0f);
      T.host.back().cta_count = 1;
      T.host.back().want_norm = 0;
    }
    int c = 0;
    for (auto& j : T.host) { j.cta_begin = c; c += j.cta_count; }
    T.grid = c;
  };
  build_fin(P.fin_all, true, true);
  build_fin(P.fin_clf, true, false);
  build_fin(P.fin_fus, false, true);

  // ---------------- finish: CTA prefixes, slots, upload ----------------
  GemmTable* all[] = {&P.fwd_proj, &P.fwd_qkv, &P.fwd_f0, &P.fwd_f1, &P.fwd_p0, &P.fwd_p1, &P.dg_p1, &P.dg_p0_fused,
                      &P.dg_p0_split, &P.dg_f1, &P.dg_f0, &P.dg_qkv, &P.wg_all, &P.wg_clf, &P.wg_fus, &P.wg_early, &P.wg_rest};
  for (GemmTable* T : all) T->grid = finish_table(T->host.data(), static_cast<int>(T->host.size()));
  // gradient-norm slots: fused-step wgrad CTAs first (the dAraw problem writes none: its slots stay 0)
  float* slots = P.buf<float>("slots");
  for (auto& g : P.wg_all.host)
    if (g.epi.out_f32 != P.buf<float>("dAraw")) g.epi.sumsq_slots = slots + g.cta_begin;
  // the early / rest split of the same problems (fnd_train_step_overlap, data-parallel overlap) shares wg_all's slot
  // layout: a problem's tiles write the slots its twin in wg_all owns
  for (GemmTable* T : {&P.wg_early, &P.wg_rest})
    for (auto& g : T->host) {
      if (g.epi.out_f32 == P.buf<float>("dAraw")) continue;
      for (const auto& a : P.wg_all.host)
        if (a.epi.out_f32 == g.epi.out_f32 && a.cta_count == g.cta_count) g.epi.sumsq_slots = slots + a.cta_begin;
    }
  P.total_slots = P.wg_all.grid + P.fin_all.grid;
  if (P.total_slots > kSlotSumsq || P.fin_all.grid > 16384 - kSlotScratch) return -31;
  return 0;
}

static int upload_tables(Plan& P, cudaStream_t st) {
  uint8_t* base = P.buf<uint8_t>("tables");
  size_t off = 0;
  const size_t cap = static_cast<size_t>(P.bufs["tables"].bytes);
  FinTable* fins[] = {&P.fin_all, &P.fin_clf, &P.fin_fus};
  for (FinTable* T : fins) {
    const size_t bytes = T->host.size() * sizeof(FinJob);
    if (off + bytes > cap) return -32;
    T->dev = reinterpret_cast<FinJob*>(base + off);
    FND_CUDA_OK(cudaMemcpyAsync(T->dev, T->host.data(), bytes, cudaMemcpyHostToDevice, st));
    off = align_up(off + bytes, 256);
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------
// launch helpers
// ---------------------------------------------------------------------------------------------------
static inline RunCtx make_ctx(const Plan& P, int training) {
  DevState* S = P.state();
  RunCtx c;
  memset(&c, 0, sizeof(c));
  c.err = &S->err;
  c.rng = S->rng;
  c.training = training;
  long long* dbg = P.buf<long long>("dbg");
  Plan& Pm = const_cast<Plan&>(P);
  c.dbg = dbg ? dbg + static_cast<size_t>((Pm.dbg_launch++) % 40) * 1024 * 8 : nullptr;
  return c;
}
// PDL chaining: the first launch of an entry point is an ordinary launch (full dependency on whatever precedes it in
// the stream); every later launch of the same entry point carries the programmatic-serialization attribute.
static bool take_pdl(const Plan& Pc) {
  Plan& P = const_cast<Plan&>(Pc);
  const bool v = P.pdl_next;
  P.pdl_next = true;
  return v;
}
// Debug aid (tools/prefix_probe.py): when a launch limit is set, only the first `limit` launches of an entry point are
// issued, so the marginal in-graph cost of every kernel can be measured by timing prefixes of the step.
static bool launch_skipped(const Plan& Pc) {
  Plan& P = const_cast<Plan&>(Pc);
  if (P.launch_limit < 0) return false;
  return P.launch_seq++ >= P.launch_limit;
}
#define FND_SKIP(P) do { if (launch_skipped(P)) return 0; } while (0)
static void mark(const Plan& Pc, const char* name, cudaStream_t st) {
  Plan& P = const_cast<Plan&>(Pc);
  if (!P.profiling) return;
  P.pdl_next = false;          // an event record sits between the two kernels
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, st);
  P.marks.emplace_back(name, ev);
}
static int run_gemm(const Plan& P, const GemmTable& T, int training, cudaStream_t st, const char* name,
                    const FinParams* fin = nullptr, int fin_ctas = 0, const DpRoute* route = nullptr) {
  FND_SKIP(P);
  const bool merged = fin && gemm_launch_is_light(T.kind, T.host.data(), static_cast<int>(T.host.size()));
  RunCtx ctx = make_ctx(P, training);
  if (route) ctx.route = *route;
  FND_CUDA_OK(launch_gemm(T.kind, T.host.data(), static_cast<int>(T.host.size()), T.grid, ctx, st,
                          take_pdl(P), merged ? fin : nullptr, merged ? fin_ctas : 0));
  mark(P, name, st);
  if (fin && !merged) {
    // large batches (deep wgrad ring, one CTA per SM): the finalize jobs run as a kernel of their own
    FND_CUDA_OK(launch_k(finalize_kernel, fin_ctas, 256, 0, st, take_pdl(P), *fin));
    mark(P, "finalize", st);
  }
  return 0;
}

static int run_prep(Plan& P, const fnd_inputs* in, int training, bool bump_clf, cudaStream_t st) {
  FND_SKIP(P);
  PrepParams pp;
  memset(&pp, 0, sizeof(pp));
  for (int i = 0; i < P.nmod; ++i) {
    if (!in->x[i]) return -40;
    if (in->pitch[i] % 4 || (reinterpret_cast<uintptr_t>(in->x[i]) & 15)) return -41;
    pp.x[i] = in->x[i]; pp.xpitch[i] = in->pitch[i]; pp.xdim[i] = P.xdim[i]; pp.xoff[i] = P.xoff[i];
  }
  pp.nmod = P.nmod; pp.dsum = P.dsum; pp.gather = in->gather;
  pp.aux_src = P.d.aux_dim ? in->aux : nullptr; pp.aux_pitch = in->aux_pitch;
  pp.label_src = in->labels;
  pp.aux_dst = P.buf<float>("aux"); pp.label_dst = P.buf<long long>("labels");
  pp.out_hi = P.buf<__nv_bfloat16>("xbf_hi"); pp.out_lo = P.buf<__nv_bfloat16>("xbf_lo");
  pp.B = P.B;
  pp.gates = P.W("clf.node.trees.0.gates.0"); pp.alpha = P.buf<float>("alpha");
  pp.TD = P.TD; pp.H = P.H;
  pp.rng = P.state()->rng; pp.bump_fusion = training ? 1 : 0; pp.bump_clf = (training && bump_clf) ? 1 : 0;
  pp.wg_done = &P.state()->wg_done;
  FND_CUDA_OK(launch_k(prep_kernel, P.B + P.TD, kRowThreads, 0, st, take_pdl(P), pp));
  mark(P, "prep", st);
  return 0;
}

static void fill_ev(const Plan& P, EvidenceParams (&ev)[3]) {
  const char* an[3] = {"attn_tv", "attn_ta", "attn_vu"};
  for (int k = 0; k < 3; ++k) {
    const std::string b = std::string("fusion.") + an[k] + ".evidence_proj.";
    ev[k].w1 = P.W(b + "0.weight"); ev[k].b1 = P.W(b + "0.bias");
    ev[k].w2 = P.W(b + "2.weight"); ev[k].b2 = P.W(b + "2.bias");
  }
}

static int run_assemble_fwd(Plan& P, cudaStream_t st) {
  FND_SKIP(P);
  AssembleParams a;
  memset(&a, 0, sizeof(a));
  a.P = P.buf<float>("P"); a.Q = P.buf<float>("Q");
  fill_ev(P, a.ev);
  a.cat_hi = P.buf<__nv_bfloat16>("fused_cat_hi"); a.cat_lo = P.buf<__nv_bfloat16>("fused_cat_lo");
  a.rowstat = P.buf<float>("rowstat");
  a.B = P.B; a.H = P.H; a.use_gnn = P.d.use_gnn;
  const int grid = P.B < 1184 ? P.B : 1184;
  if (P.H == 512) FND_CUDA_OK(launch_k(assemble_fwd_kernel<1>, grid, kRowThreads, 0, st, take_pdl(P), a));
  else FND_CUDA_OK(launch_k(assemble_fwd_kernel<2>, grid, kRowThreads, 0, st, take_pdl(P), a));
  mark(P, "assemble_fwd", st);
  return 0;
}

static int run_assemble_bwd(Plan& P, cudaStream_t st) {
  FND_SKIP(P);
  AssembleBwdParams a;
  memset(&a, 0, sizeof(a));
  a.P = P.buf<float>("P"); a.Q = P.buf<float>("Q"); a.dcat = P.buf<float>("dcat"); a.rowstat = P.buf<float>("rowstat");
  fill_ev(P, a.ev);
  a.dPdirect = P.buf<float>("dPdirect");
  a.dQ_hi = P.buf<__nv_bfloat16>("dQ_hi"); a.dQ_lo = P.buf<__nv_bfloat16>("dQ_lo");
  a.dP_hi = P.buf<__nv_bfloat16>("dP_hi"); a.dP_lo = P.buf<__nv_bfloat16>("dP_lo");
  a.ev_partial = P.buf<float>("ev_partial"); a.evstride = P.L.evstride;
  a.B = P.B; a.H = P.H; a.use_gnn = P.d.use_gnn;
  if (P.H == 512) FND_CUDA_OK(launch_k(assemble_bwd_kernel<1>, P.n_asm_ctas, kRowThreads, 0, st, take_pdl(P), a));
  else FND_CUDA_OK(launch_k(assemble_bwd_kernel<2>, P.n_asm_ctas, kRowThreads, 0, st, take_pdl(P), a));
  mark(P, "assemble_bwd", st);
  return 0;
}

static HeadParams head_params(const Plan& P, int training) {
  HeadParams h;
  memset(&h, 0, sizeof(h));
  h.h = P.buf<float>("h"); h.alpha = P.buf<float>("alpha");
  h.thresh = P.W("clf.node.trees.0.thresh.0"); h.leaf = P.W("clf.node.trees.0.leaf_logits");
  h.wb = P.W("clf.bypass.weight"); h.bb = P.W("clf.bypass.bias"); h.temperature = P.W("clf.temperature");
  h.labels = P.buf<long long>("labels");
  h.tau = P.d.node_tau; h.tree_drop_p = P.d.tree_dropout; h.training = training;
  h.logits = P.buf<float>("logits"); h.probs = P.buf<float>("probs"); h.svals = P.buf<float>("svals");
  h.loss_row = P.buf<float>("loss_row"); h.dlogits_out = P.buf<float>("dlogits");
  h.dF = P.buf<float>("dF"); h.dF_hi = P.buf<__nv_bfloat16>("dFbf_hi"); h.dF_lo = P.buf<__nv_bfloat16>("dFbf_lo");
  h.leafc = P.buf<float>("leafc");
  h.z_pre1 = P.buf<float>("z_p1"); h.pre_drop_p = P.d.clf_dropout;
  h.dz_hi = P.buf<__nv_bfloat16>("dz_p1_hi"); h.dz_lo = P.buf<__nv_bfloat16>("dz_p1_lo");
  h.state = P.state();
  h.B = P.B; h.H = P.H; h.T = P.d.trees; h.D = P.d.depth;
  long long* dbg = P.buf<long long>("dbg");
  Plan& Pm = const_cast<Plan&>(P);
  h.dbg = dbg ? dbg + static_cast<size_t>((Pm.dbg_launch++) % 40) * 1024 * 8 : nullptr;
  return h;
}
template <bool FWD, bool CE, bool BWD>
static int run_head(const Plan& P, const HeadParams& h, cudaStream_t st) {
  FND_SKIP(P);
  if (P.d.trees != 6 || P.d.depth != 4) return -50;     // only the reference's NODE shape is instantiated
  const int grid = ceil_div(P.B, 8) < 296 ? ceil_div(P.B, 8) : 296;
  const size_t smem = static_cast<size_t>(P.TD + 2) * P.H * sizeof(float) + 8 * (P.TD + 2) * 33 * sizeof(float);
  if (P.H == 512) {
    FND_CUDA_OK(cudaFuncSetAttribute(head_kernel<FWD, CE, BWD, 4, 6, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    FND_CUDA_OK(launch_k(head_kernel<FWD, CE, BWD, 4, 6, 4>, grid, 256, smem, st, take_pdl(P), h));
  } else {
    FND_CUDA_OK(cudaFuncSetAttribute(head_kernel<FWD, CE, BWD, 8, 6, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    FND_CUDA_OK(launch_k(head_kernel<FWD, CE, BWD, 8, 6, 4>, grid, 256, smem, st, take_pdl(P), h));
  }
  FND_CUDA_OK(cudaGetLastError());
  mark(P, "head", st);
  return 0;
}

static FinParams fin_params(const Plan& P, const FinTable& T, int slot_base, int total_slots, bool with_loss, int update_step,
                            int elect_last) {
  FinParams f;
  memset(&f, 0, sizeof(f));
  f.jobs = T.dev; f.njobs = static_cast<int>(T.host.size());
  f.slots = P.buf<float>("slots"); f.slot_base = slot_base; f.total_slots = total_slots;
  f.loss_row = with_loss ? P.buf<float>("loss_row") : nullptr; f.B = P.B;
  f.state = P.state(); f.update_step = update_step; f.elect_last = elect_last;
  return f;
}
static int run_finalize(Plan& P, const FinTable& T, int slot_base, int total_slots, bool with_loss, int update_step,
                        cudaStream_t st) {
  FND_SKIP(P);
  FinParams f = fin_params(P, T, slot_base, total_slots, with_loss, update_step, 1);
  FND_CUDA_OK(launch_k(finalize_kernel, T.grid, 256, 0, st, take_pdl(P), f));
  mark(P, "finalize", st);
  return 0;
}

// one persistent wave: as many CTAs as are resident at once
static int adamw_grid() {
  static int g = 0;
  if (!g) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adamw_kernel, 256, 0) != cudaSuccess || per_sm < 1) per_sm = 4;
    g = 148 * per_sm;
  }
  return g;
}
static AdamWParams adamw_params(const Plan& P) {
  AdamWParams a;
  memset(&a, 0, sizeof(a));
  a.p = P.params; a.g = P.grads; a.m = P.m; a.v = P.v;
  a.n = static_cast<size_t>(P.L.n_hot);
  a.sh_hi = P.sh_hi; a.sh_lo = P.sh_lo; a.n_shadow = static_cast<size_t>(P.L.n_shadow);
  if (P.d.aux_dim) {
    a.rp_begin = static_cast<size_t>(P.L.at("clf.pre.0.weight"));
    a.rp_end = a.rp_begin + static_cast<size_t>(P.H) * (P.H + P.d.aux_dim);
    a.rp_cols = P.H + P.d.aux_dim; a.rp_pitch = P.L.rp_pitch;
    a.rp_hi = P.sh_hi + P.L.n_shadow; a.rp_lo = P.sh_lo ? P.sh_lo + P.L.n_shadow : nullptr;
  }
  a.state = P.state();
  a.slots = nullptr; a.nslots = 0;
  return a;
}

static int fusion_forward_impl(Plan& P, const fnd_inputs* in, int training, bool bump_clf, cudaStream_t st,
                               bool join_deferred = false) {
  P.last_training = training;
  P.dbg_launch = 0;
  FND_OK(run_prep(P, in, training, bump_clf, st));
  FND_OK(run_gemm(P, P.fwd_proj, training, st, "gemm_proj"));
  FND_OK(run_gemm(P, P.fwd_qkv, training, st, "gemm_qkv"));
  FND_OK(run_assemble_fwd(P, st));
  if (join_deferred) {
    // the deferred all-gather of the fuse_mlp shadows (side stream) must have landed before gemm_fuse0 reads them —
    // including its PDL-early weight loads, hence an ordinary launch
    FND_CUDA_OK(cudaStreamWaitEvent(st, P.ev_join2, 0));
    P.pdl_next = false;
  }
  FND_OK(run_gemm(P, P.fwd_f0, training, st, "gemm_fuse0"));
  FND_OK(run_gemm(P, P.fwd_f1, training, st, "gemm_fuse1"));
  return 0;
}
static int fusion_head_impl(Plan& P, cudaStream_t st) {   // fusion.classifier: returned for API parity only
  FND_CUDA_OK(launch_k(rowlinear2_fwd_kernel, ceil_div(P.B, 8), 256, 0, st, take_pdl(P), P.buf<float>("fused"),
                       P.W("fusion.classifier.weight"), P.W("fusion.classifier.bias"), P.buf<float>("fusion_logits"), P.B, P.H));
  mark(P, "fusion_head", st);
  return 0;
}
static int classifier_gemms_impl(Plan& P, int training, cudaStream_t st) {
  FND_OK(run_gemm(P, P.fwd_p0, training, st, "gemm_pre0"));
  FND_OK(run_gemm(P, P.fwd_p1, training, st, "gemm_pre1"));
  return 0;
}

static Plan* as_plan(void* p) { return static_cast<Plan*>(p); }
#define FND_PLAN(p)                      \
  Plan* PP = as_plan(p);                 \
  if (!PP) return -1;                    \
  if (!PP->bound) return -5;             \
  Plan& P = *PP;                         \
  P.pdl_next = false;                    \
  P.launch_seq = 0;                      \
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream)

extern "C" {

int fnd_version(void) { return 100; }
const char* fnd_build_arch(void) { return "sm_100a"; }

// ------------------------------- arena -------------------------------
int fnd_param_count(const fnd_dims* dims) {
  if (!dims || !dims_ok(*dims)) return -1;
  return static_cast<int>(make_layout(*dims).entries.size());
}
int fnd_param_info(const fnd_dims* dims, int index, char* name, int name_cap, long long* offset, int* ndim, int* rows,
                   int* cols, int* hot) {
  if (!dims || !dims_ok(*dims)) return -1;
  const ArenaLayout L = make_layout(*dims);
  if (index < 0 || index >= static_cast<int>(L.entries.size())) return -2;
  const ParamEntry& e = L.entries[index];
  if (name && name_cap > 0) {
    strncpy(name, e.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (offset) *offset = e.offset;
  if (ndim) *ndim = e.ndim;
  if (rows) *rows = e.rows;
  if (cols) *cols = e.cols;
  if (hot) *hot = e.hot;
  return 0;
}
long long fnd_arena_total_elems(const fnd_dims* dims) { return (dims && dims_ok(*dims)) ? make_layout(*dims).n_total : -1; }
long long fnd_arena_hot_elems(const fnd_dims* dims) { return (dims && dims_ok(*dims)) ? make_layout(*dims).n_hot : -1; }
long long fnd_arena_shadow_elems(const fnd_dims* dims) { return (dims && dims_ok(*dims)) ? make_layout(*dims).n_shadow : -1; }
long long fnd_arena_shadow_buffer_elems(const fnd_dims* dims) {
  if (!dims || !dims_ok(*dims)) return -1;
  const ArenaLayout L = make_layout(*dims);
  return L.n_shadow + align64(L.rp_elems);
}

// ------------------------------- plan -------------------------------
int fnd_plan_create(const fnd_dims* dims, int batch, int mode, void** plan_out) {
  if (!dims || !plan_out || batch < 1 || (mode != 0 && mode != 1)) return -1;
  if (!dims_ok(*dims)) return -2;
  Plan* P = new (std::nothrow) Plan();
  if (!P) return -3;
  P->d = *dims;
  P->L = make_layout(*dims);
  P->B = batch; P->mode = mode; P->ncombo = mode ? 3 : 1;
  P->H = dims->hidden; P->nmod = dims->use_gnn ? 5 : 4; P->nslots = dims->use_gnn ? 16 : 15;
  P->TD = dims->trees * dims->depth; P->leaves = 1 << dims->depth;
  const int pd[5] = {dims->d_text, dims->d_audio, dims->d_visual, dims->d_temporal, dims->d_gnn};
  int o = 0;
  for (int i = 0; i < P->nmod; ++i) { P->xoff[i] = o; P->xdim[i] = pd[i]; o += pd[i]; }
  P->dsum = o;
  carve(*P);
  *plan_out = P;
  return 0;
}
void fnd_plan_destroy(void* plan) { delete as_plan(plan); }
size_t fnd_plan_workspace_bytes(const void* plan) { return plan ? static_cast<size_t>(static_cast<const Plan*>(plan)->ws_bytes) : 0; }
long long fnd_plan_buffer_offset(const void* plan, const char* name) {
  if (!plan || !name) return -1;
  const Plan* P = static_cast<const Plan*>(plan);
  auto it = P->bufs.find(name);
  return it == P->bufs.end() ? -1 : it->second.off;
}
long long fnd_plan_buffer_bytes(const void* plan, const char* name) {
  if (!plan || !name) return -1;
  const Plan* P = static_cast<const Plan*>(plan);
  auto it = P->bufs.find(name);
  return it == P->bufs.end() ? -1 : it->second.bytes;
}

int fnd_plan_set_grad_mirror(void* plan, void* grads_bf16) {
  Plan* PP = as_plan(plan);
  if (!PP) return -1;
  if (reinterpret_cast<uintptr_t>(grads_bf16) & 255) return -3;
  PP->grads_bf = static_cast<__nv_bfloat16*>(grads_bf16);
  return 0;
}

int fnd_plan_bind(void* plan, void* workspace, float* params, float* grads, float* adam_m, float* adam_v, void* shadow_hi,
                  void* shadow_lo, void* stream) {
  Plan* PP = as_plan(plan);
  if (!PP || !workspace || !params || !grads || !shadow_hi) return -1;
  Plan& P = *PP;
  if (P.ncombo == 3 && !shadow_lo) return -2;
  if ((reinterpret_cast<uintptr_t>(workspace) & 255) || (reinterpret_cast<uintptr_t>(params) & 255) ||
      (reinterpret_cast<uintptr_t>(grads) & 255) || (reinterpret_cast<uintptr_t>(shadow_hi) & 255))
    return -3;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  FND_CUDA_OK(init_gemm_attrs());
  P.ws = static_cast<uint8_t*>(workspace);
  P.params = params; P.grads = grads; P.m = adam_m; P.v = adam_v;
  P.sh_hi = static_cast<__nv_bfloat16*>(shadow_hi);
  P.sh_lo = P.ncombo == 3 ? static_cast<__nv_bfloat16*>(shadow_lo) : nullptr;
  FND_CUDA_OK(cudaMemsetAsync(P.ws, 0, static_cast<size_t>(P.ws_bytes), st));
  GemmTable* all[] = {&P.fwd_proj, &P.fwd_qkv, &P.fwd_f0, &P.fwd_f1, &P.fwd_p0, &P.fwd_p1, &P.dg_p1, &P.dg_p0_fused,
                      &P.dg_p0_split, &P.dg_f1, &P.dg_f0, &P.dg_qkv, &P.wg_all, &P.wg_clf, &P.wg_fus, &P.wg_early, &P.wg_rest};
  for (GemmTable* T : all) T->host.clear();
  P.fin_all.host.clear(); P.fin_clf.host.clear(); P.fin_fus.host.clear();
  FND_OK(build_tables(P));
  FND_OK(upload_tables(P, st));
  DevState hs;
  memset(&hs, 0, sizeof(hs));
  hs.rng[0] = 0x9E3779B9u; hs.rng[1] = 0x7F4A7C15u;
  hs.lr = 2e-4f; hs.beta1 = 0.9f; hs.beta2 = 0.999f; hs.eps = 1e-8f; hs.weight_decay = 1e-4f; hs.max_norm = 5.0f;
  hs.bc1 = 1.0f; hs.bc2 = 1.0f; hs.clip_coef = 1.0f;
  hs.loss_scale = 1.0f / static_cast<float>(P.B);
  FND_CUDA_OK(cudaMemcpyAsync(P.state(), &hs, sizeof(hs), cudaMemcpyHostToDevice, st));
  FND_CUDA_OK(cudaStreamSynchronize(st));   // hs / host tables are stack or plan-owned pageable memory
  if (!P.ev_fork) {                         // side-stream fork / join events (never created inside a capture)
    FND_CUDA_OK(cudaEventCreateWithFlags(&P.ev_fork, cudaEventDisableTiming));
    FND_CUDA_OK(cudaEventCreateWithFlags(&P.ev_join, cudaEventDisableTiming));
    FND_CUDA_OK(cudaEventCreateWithFlags(&P.ev_fork2, cudaEventDisableTiming));
    FND_CUDA_OK(cudaEventCreateWithFlags(&P.ev_join2, cudaEventDisableTiming));
  }
  P.bound = true;
  return 0;
}

// ------------------------------- state -------------------------------
static int write_state(Plan& P, size_t field_off, const void* src, size_t bytes, cudaStream_t st) {
  FND_CUDA_OK(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(P.state()) + field_off, src, bytes, cudaMemcpyHostToDevice, st));
  FND_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}
int fnd_set_hyper(void* plan, float lr, float beta1, float beta2, float eps, float weight_decay, float max_norm, void* stream) {
  FND_PLAN(plan);
  const float h[6] = {lr, beta1, beta2, eps, weight_decay, max_norm};
  return write_state(P, offsetof(DevState, lr), h, sizeof(h), st);
}
int fnd_set_lr(void* plan, float lr, void* stream) {
  FND_PLAN(plan);
  return write_state(P, offsetof(DevState, lr), &lr, sizeof(lr), st);
}
int fnd_set_seed(void* plan, unsigned long long seed, void* stream) {
  FND_PLAN(plan);
  const uint32_t r[2] = {static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32)};
  return write_state(P, offsetof(DevState, rng), r, sizeof(r), st);
}
int fnd_set_loss_scale(void* plan, float scale, void* stream) {
  FND_PLAN(plan);
  return write_state(P, offsetof(DevState, loss_scale), &scale, sizeof(scale), st);
}
int fnd_set_loss_mirror(void* plan, float* host_mapped, int ring, void* stream) {
  FND_PLAN(plan);
  if (host_mapped && ring <= 0) return -1;
  P.loss_mirror = host_mapped;
  P.loss_ring = host_mapped ? ring : 0;
  bool found = false;
  for (auto& j : P.fin_all.host)
    if (j.type == kJobLossMean) {
      j.dst = host_mapped;
      j.dst_pitch = host_mapped ? ring : 0;
      found = true;
    }
  if (!found) return -2;
  // the job table lives in device memory and is read at run time: already-captured graphs see the new pointer too
  FND_CUDA_OK(cudaMemcpyAsync(P.fin_all.dev, P.fin_all.host.data(), P.fin_all.host.size() * sizeof(FinJob), cudaMemcpyHostToDevice, st));
  FND_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}
int fnd_collect_rows(void* plan, int k, long long offset, long long batch_no, float* loss_rows, float* p1, long long* ys,
                     long long* bid, float* forensic3, void* stream) {
  FND_PLAN(plan);
  if (k <= 0) return 0;
  if (k > P.B || offset < 0) return -1;
  P.pdl_next = false;
  collect_rows_kernel<<<ceil_div(k, 256), 256, 0, st>>>(P.buf<float>("loss_row"), P.buf<float>("probs"), P.buf<long long>("labels"),
                                                       P.buf<float>("rowstat"), k, offset, batch_no, loss_rows, p1, ys, bid, forensic3);
  FND_CUDA_OK(cudaGetLastError());
  return 0;
}
int fnd_refresh_shadows(void* plan, void* stream) {
  FND_PLAN(plan);
  AdamWParams a = adamw_params(P);
  shadow_refresh_kernel<<<148 * 8, 256, 0, st>>>(a);
  FND_CUDA_OK(cudaGetLastError());
  return 0;
}
int fnd_check_error(void* plan, void* stream) {
  FND_PLAN(plan);
  int herr = 0;
  FND_CUDA_OK(cudaMemcpyAsync(&herr, &P.state()->err, sizeof(int), cudaMemcpyDeviceToHost, st));
  FND_CUDA_OK(cudaStreamSynchronize(st));
  if (herr) {
    FND_CUDA_OK(cudaMemsetAsync(&P.state()->err, 0, sizeof(int), st));
    FND_CUDA_OK(cudaStreamSynchronize(st));
  }
  return herr;
}
int fnd_export_dropout_mask(void* plan, int layer, float* out, long long n, void* stream) {
  FND_PLAN(plan);
  if (!out || n <= 0) return -1;
  float p = 0.f;
  switch (layer) {
    case kStreamFuse0: case kStreamFuse1: p = P.d.fusion_dropout; break;
    case kStreamPre0: case kStreamPre1: p = P.d.clf_dropout; break;
    case kStreamTree: p = P.d.tree_dropout; break;
    default: return -2;
  }
  dropout_mask_kernel<<<256, 256, 0, st>>>(out, static_cast<size_t>(n), p, layer, P.state());
  FND_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------- module-level entry points -------------------------------
int fnd_fusion_forward(void* plan, const fnd_inputs* in, int training, void* stream) {
  FND_PLAN(plan);
  if (!in) return -1;
  FND_OK(fusion_forward_impl(P, in, training, false, st));
  return fusion_head_impl(P, st);
}

int fnd_classifier_forward(void* plan, const float* fused, const float* aux, int aux_pitch, int training, void* stream) {
  FND_PLAN(plan);
  P.last_training = training;
  if (fused || aux) {
    // stage external inputs: cast `fused` to the bf16 operand planes and copy aux rows
    PrepParams pp;
    memset(&pp, 0, sizeof(pp));
    pp.nmod = fused ? 1 : 0;
    pp.x[0] = fused; pp.xpitch[0] = P.H; pp.xdim[0] = P.H; pp.xoff[0] = 0; pp.dsum = P.H;
    pp.aux_src = (P.d.aux_dim && aux) ? aux : nullptr; pp.aux_pitch = aux_pitch; pp.aux_dst = P.buf<float>("aux");
    pp.out_hi = P.buf<__nv_bfloat16>("fusedbf_hi"); pp.out_lo = P.buf<__nv_bfloat16>("fusedbf_lo");
    pp.B = P.B; pp.gates = P.W("clf.node.trees.0.gates.0"); pp.alpha = P.buf<float>("alpha"); pp.TD = P.TD; pp.H = P.H;
    pp.rng = P.state()->rng; pp.bump_clf = training ? 1 : 0;
    pp.wg_done = &P.state()->wg_done;
    FND_CUDA_OK(launch_k(prep_kernel, P.B + P.TD, kRowThreads, 0, st, take_pdl(P), pp));
  }
  FND_OK(classifier_gemms_impl(P, training, st));
  HeadParams h = head_params(P, training);
  return run_head<true, false, false>(P, h, st);
}

int fnd_ce_loss_fwd_bwd(void* plan, const long long* labels, void* stream) {
  FND_PLAN(plan);
  HeadParams h = head_params(P, P.last_training);
  if (labels) h.labels = labels;
  return run_head<false, true, false>(P, h, st);
}

int fnd_classifier_backward(void* plan, const float* dlogits, void* stream) {
  FND_PLAN(plan);
  const int tr = P.last_training;
  HeadParams h = head_params(P, tr);
  h.dlogits_in = dlogits ? dlogits : P.buf<float>("dlogits");
  FND_OK((run_head<false, false, true>(P, h, st)));
  FND_OK(run_gemm(P, P.dg_p1, tr, st, "dgrad_pre1"));
  FND_OK(run_gemm(P, P.dg_p0_split, tr, st, "dgrad_pre0"));
  FND_OK(run_gemm(P, P.wg_clf, tr, st, "wgrad_clf"));
  return run_finalize(P, P.fin_clf, kSlotScratch, 0, false, 0, st);
}

int fnd_fusion_backward(void* plan, const float* dfused, const float* dfusion_logits, void* stream) {
  FND_PLAN(plan);
  const int tr = P.last_training;
  const float* df = dfused ? dfused : P.buf<float>("dfused");
  if (dfusion_logits) {
    // d fused += dlogits . W_classifier   (fusion.classifier, cross_modal_transformer.py:198)
    float* acc = P.buf<float>("dfused");
    if (df != acc) FND_CUDA_OK(cudaMemcpyAsync(acc, df, static_cast<size_t>(P.B) * P.H * 4, cudaMemcpyDeviceToDevice, st));
    P.pdl_next = false;     // a memcpy node may precede this launch
    FND_CUDA_OK(launch_k(rowlinear2_dgrad_kernel, P.B, 256, 0, st, take_pdl(P), dfusion_logits,
                         P.W("fusion.classifier.weight"), acc, P.B, P.H, 1));
    df = acc;
  }
  GateParams g;
  memset(&g, 0, sizeof(g));
  g.dy = df; g.z = P.buf<float>("z_f1");
  g.out_hi = P.buf<__nv_bfloat16>("dz_f1_hi"); g.out_lo = P.buf<__nv_bfloat16>("dz_f1_lo");
  g.drop_p = P.d.fusion_dropout; g.stream = kStreamFuse1; g.training = tr; g.state = P.state();
  g.n = static_cast<size_t>(P.B) * P.H;
  FND_CUDA_OK(launch_k(gate_kernel, ceil_div(static_cast<int>(g.n / 8), 256), 256, 0, st, take_pdl(P), g));
  FND_OK(run_gemm(P, P.dg_f1, tr, st, "dgrad_fuse1"));
  FND_OK(run_gemm(P, P.dg_f0, tr, st, "dgrad_fuse0"));
  FND_OK(run_assemble_bwd(P, st));
  FND_OK(run_gemm(P, P.dg_qkv, tr, st, "dgrad_qkv"));
  FND_OK(run_gemm(P, P.wg_fus, tr, st, "wgrad_fusion"));
  return run_finalize(P, P.fin_fus, kSlotScratch, 0, false, 0, st);
}

int fnd_clip_adamw_step(void* plan, int norm_from_slots, void* stream) {
  FND_PLAN(plan);
  if (!P.m || !P.v) return -6;
  if (!norm_from_slots) {
    // norm over the (possibly all-reduced) gradient arena + optimizer-step bookkeeping
    float* slots = P.buf<float>("slots");
    const int nb = 148 * 4;
    FND_CUDA_OK(launch_k(sumsq_kernel, nb, 256, 0, st, take_pdl(P), P.grads, static_cast<size_t>(P.L.n_hot), slots + kSlotSumsq));
    FND_CUDA_OK(launch_k(norm_finish_kernel, 1, 32, 0, st, take_pdl(P), slots + kSlotSumsq, nb, P.state(), 1));
    mark(P, "grad_norm", st);
  } else {
    FND_CUDA_OK(launch_k(step_kernel, 1, 32, 0, st, take_pdl(P), P.state()));
    mark(P, "step_bookkeeping", st);
  }
  AdamWParams a = adamw_params(P);
  FND_CUDA_OK(launch_k(adamw_kernel, adamw_grid(), 256, 0, st, take_pdl(P), a));
  mark(P, "adamw", st);
  return 0;
}

// ------------------------------- fused trainer step -------------------------------
// data-parallel tail (defined with the other fnd_dp_* entry points below). Its kernels are ordinary (non-PDL) launches:
// they spin on remote flags and must not become resident early.
static const int kDpGrid = 148 * 4;
static int dp_tail(Plan& P, bool early_done, bool defer, cudaStream_t st);
static int dp_deferred(Plan& P, int grid, cudaStream_t st);
static int dp_push(Plan& P, int s0, int s1, int bank, int ctr, int grid, int block, cudaStream_t st);
static void dp_segments(const Plan& P, int rank, int world, size_t (&lo)[kDpMaxSeg], size_t (&hi)[kDpMaxSeg]);

// dp_flags (data-parallel step only): bit 0 = push the early gradient range from the side stream under the backward,
// bit 1 = defer the optimizer update / all-gather of that range to the next step (see fnd_dp.cuh).
static int train_fwd_bwd_impl(void* plan, const fnd_inputs* in, int fused_optimizer, void* stream, void* side_stream = nullptr,
                              bool skip_norm = false, int dp_flags = 0) {
  FND_PLAN(plan);
  if (!in || !in->labels) return -1;
  cudaStream_t side = reinterpret_cast<cudaStream_t>(side_stream);
  const bool overlap = side_stream != nullptr && P.dp_bound && (dp_flags & 1);
  // single-GPU fused step with the fuse_mlp weight gradients on the side stream (fnd_train_step_overlap)
  const bool early1 = side_stream != nullptr && fused_optimizer && dp_flags == 0 && !skip_norm && P.launch_limit < 0 &&
                      gemm_launch_is_light(P.wg_rest.kind, P.wg_rest.host.data(), static_cast<int>(P.wg_rest.host.size()));
  const bool defer = side_stream != nullptr && P.dp_bound && (dp_flags & 2);
  // bit 2: fused push — the weight-gradient launch's epilogue stores each tile into its owner's staging slot (bf16 wire
  // format, merged-finalize light launch only; not combined with the early push)
  bool fused_push = P.dp_bound && (dp_flags & 4) && !overlap && P.dp.stage_bf16 && P.dp.world > 1 && !P.dp.mc_grads &&
                    gemm_launch_is_light(P.wg_all.kind, P.wg_all.host.data(), static_cast<int>(P.wg_all.host.size()));
  // every routed problem must take the staged store-only epilogue (its dead operand ring holds the 36 KB of staging;
  // not the case for very small batches, whose k-loop is a single partial block)
  for (const auto& g : P.wg_all.host)
    if (g.epi.routed && !(g.epi.plain_f32 == 2 && g.bn >= 64)) fused_push = false;
  P.dp_fused_now = fused_push;
  if (P.dp_bound && (dp_flags || skip_norm)) {
    // A deferred update of the previous step: run it on the side stream under the first four kernels of this forward
    // pass (one CTA per SM, so their GEMM CTAs still fit), or right here when there is no side stream.
    if (side_stream) {
      FND_CUDA_OK(cudaEventRecord(P.ev_fork2, st));
      FND_CUDA_OK(cudaStreamWaitEvent(side, P.ev_fork2, 0));
      FND_OK(dp_deferred(P, 148, side));
      FND_CUDA_OK(cudaEventRecord(P.ev_join2, side));
    } else {
      FND_OK(dp_deferred(P, kDpGrid, st));
    }
  }
  FND_OK(fusion_forward_impl(P, in, 1, true, st, P.dp_bound && (dp_flags || skip_norm) && side_stream != nullptr));
  FND_OK(classifier_gemms_impl(P, 1, st));
  HeadParams h = head_params(P, 1);
  FND_OK((run_head<true, true, true>(P, h, st)));
  FND_OK(run_gemm(P, P.dg_p1, 1, st, "dgrad_pre1"));
  FND_OK(run_gemm(P, P.dg_p0_fused, 1, st, "dgrad_pre0"));
  FND_OK(run_gemm(P, P.dg_f1, 1, st, "dgrad_fuse1"));
  if (overlap) {
    // Data-parallel step with overlap: the fuse_mlp.0 / fuse_mlp.3 weight gradients (70 % of all gradient bytes) need
    // only dz_f0, dz_f1, h1 and fused_cat, so they are produced NOW and pushed to their owners from the side stream
    // under the rest of the backward pass (fork / join through two events; capturable). The push uses ONE 256-thread
    // CTA per SM (64 registers, no shared memory), so a GEMM CTA (320 threads x 152 registers, 213 KB smem) still fits
    // beside it.
    FND_OK(run_gemm(P, P.wg_early, 1, st, "wgrad_early"));
    FND_CUDA_OK(cudaEventRecord(P.ev_fork, st));
    FND_CUDA_OK(cudaStreamWaitEvent(side, P.ev_fork, 0));
    P.dp.a = adamw_params(P);
    FND_OK(dp_push(P, 0, 1, kPadReadyEarly, kPadCounterEarly, 148, 256, side));
    FND_CUDA_OK(cudaEventRecord(P.ev_join, side));
    P.pdl_next = false;
  }
  if (early1) {
    // fork: dz_f0 / dz_f1 / h1 / fused_cat are complete; the side stream's launch is an ordinary (non-PDL) one
    FND_CUDA_OK(cudaEventRecord(P.ev_fork, st));
    FND_CUDA_OK(cudaStreamWaitEvent(side, P.ev_fork, 0));
    const bool keep = P.pdl_next;
    P.pdl_next = false;
    FND_OK(run_gemm(P, P.wg_early, 1, side, "wgrad_early"));
    FND_CUDA_OK(cudaEventRecord(P.ev_join, side));
    P.pdl_next = keep;
  }
  FND_OK(run_gemm(P, P.dg_f0, 1, st, "dgrad_fuse0"));
  FND_OK(run_assemble_bwd(P, st));
  FND_OK(run_gemm(P, P.dg_qkv, 1, st, "dgrad_qkv"));
  if (early1) {
    // the remaining weight gradients + finalize CTAs (slot layout of wg_all), then join before the optimizer
    const FinParams f = fin_params(P, P.fin_all, P.wg_all.grid, P.total_slots, false, 0, 0);
    FND_OK(run_gemm(P, P.wg_rest, 1, st, "wgrad_rest", &f, P.fin_all.grid));
    FND_CUDA_OK(cudaStreamWaitEvent(st, P.ev_join, 0));
    P.joined_side = true;
    return 0;
  }
  if (overlap) {
    const FinParams f = fin_params(P, P.fin_all, P.wg_rest.grid, P.total_slots, false, 0, 0);
    FND_OK(run_gemm(P, P.wg_rest, 1, st, "wgrad_rest", &f, P.fin_all.grid));
    FND_CUDA_OK(cudaStreamWaitEvent(st, P.ev_join, 0));
    return dp_tail(P, true, defer, st);
  }
  // Weight gradients of every GEMM; the trailing CTAs of the same launch run the finalize jobs (bias / threshold /
  // leaf / evidence reductions, mean loss). Every CTA leaves its sum of squares in "slots"; whoever consumes the
  // gradients next reduces the slots to the global norm: adamw_kernel in the fused step (no election, fence or atomic
  // on the tile CTAs' critical path — measured 10 us), norm_finish_kernel otherwise.
  const FinParams f = fin_params(P, P.fin_all, P.wg_all.grid, P.total_slots, false, 0, 0);
  FND_OK(run_gemm(P, P.wg_all, 1, st, "wgrad_all", &f, P.fin_all.grid, fused_push ? &P.dp_route : nullptr));
  if (!fused_optimizer && !skip_norm) {
    FND_SKIP(P);
    FND_CUDA_OK(launch_k(norm_finish_kernel, 1, 32, 0, st, take_pdl(P), P.buf<float>("slots"), P.total_slots, P.state(), 0));
    mark(P, "grad_norm", st);
  }
  return 0;
}

int fnd_train_fwd_bwd(void* plan, const fnd_inputs* in, void* stream) { return train_fwd_bwd_impl(plan, in, 0, stream); }

int fnd_train_step_dp(void* plan, const fnd_inputs* in, void* stream, void* side_stream, int flags) {
  Plan* PP = as_plan(plan);
  if (!PP || !PP->bound) return -5;
  if (!PP->dp_bound) return -7;
  if (!side_stream) flags &= 4;      // overlap / deferral need the side stream; the fused push (bit 2) does not
  if (flags & 1) return train_fwd_bwd_impl(plan, in, 0, stream, side_stream, true, flags);
  // (skip_norm: the reduced gradient's norm is what counts)
  FND_OK(train_fwd_bwd_impl(plan, in, 0, stream, side_stream, true, flags));
  Plan& P = *PP;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  P.pdl_next = false;
  return dp_tail(P, false, (flags & 2) != 0, st);
}

int fnd_train_step(void* plan, const fnd_inputs* in, void* stream) { return fnd_train_step_overlap(plan, in, stream, nullptr); }

int fnd_train_step_overlap(void* plan, const fnd_inputs* in, void* stream, void* side_stream) {
  {
    Plan* PP = as_plan(plan);
    if (PP && PP->bound && (!PP->m || !PP->v)) return -6;
    if (PP) PP->joined_side = false;
  }
  FND_OK(train_fwd_bwd_impl(plan, in, 1, stream, side_stream));
  FND_PLAN(plan);
  P.pdl_next = !P.joined_side;   // continues the chain started by train_fwd_bwd_impl (an event wait breaks it)
  P.launch_seq = 16;
  FND_SKIP(P);
  // AdamW reduces the norm slots itself (identical in every CTA), clips, steps and publishes the bookkeeping.
  AdamWParams a = adamw_params(P);
  a.slots = P.buf<float>("slots"); a.nslots = P.total_slots;
  FND_CUDA_OK(launch_k(adamw_kernel, adamw_grid(), 256, 0, st, take_pdl(P), a));
  mark(P, "adamw", st);
  return 0;
}

// ------------------------------- data-parallel optimizer step over peer memory -------------------------------
// The hot arena [0, n_hot) is cut into three ranges — the "early" range fuse_mlp.0.weight + fuse_mlp.3.weight (adjacent
// in the arena; their gradients are complete after dgrad_fuse1), the arena before it and the arena after it — and every
// rank owns an equal share (rounded to 1024 elements) of EACH range, so that the early push is balanced over ranks.
static void dp_segments(const Plan& P, int rank, int world, size_t (&lo)[kDpMaxSeg], size_t (&hi)[kDpMaxSeg]) {
  const size_t a0 = static_cast<size_t>(P.L.at("fusion.fuse_mlp.0.weight"));
  const size_t a1 = static_cast<size_t>(P.L.at("clf.pre.0.weight"));
  const size_t rb[kDpMaxSeg][2] = {{a0, a1}, {0, a0}, {a1, static_cast<size_t>(P.L.n_hot)}};
  for (int s = 0; s < kDpMaxSeg; ++s) {
    const size_t n = rb[s][1] - rb[s][0];
    size_t per = (n + world - 1) / world;
    per = (per + 1023) / 1024 * 1024;
    const size_t l = per * rank < n ? per * rank : n;
    lo[s] = rb[s][0] + l;
    hi[s] = rb[s][0] + (l + per < n ? l + per : n);
  }
}
// elements of the largest slice (rank 0's: shares are rounded up to 1024)
static size_t dp_slot_cap(const Plan& P, int world) {
  size_t lo[kDpMaxSeg], hi[kDpMaxSeg], n = 0;
  dp_segments(P, 0, world, lo, hi);
  for (int s = 0; s < kDpMaxSeg; ++s) n += hi[s] - lo[s];
  return n;
}

long long fnd_dp_stage_bytes(const void* plan, int world, int bf16) {
  if (!plan || world < 1 || world > kDpMaxWorld) return -1;
  const Plan* P = static_cast<const Plan*>(plan);
  return static_cast<long long>(dp_slot_cap(*P, world)) * world * (bf16 ? 2 : 4);
}

int fnd_dp_bind(void* plan, int rank, int world, const unsigned long long* peer_bases, long long off_params,
                long long off_grads, long long off_shadow_hi, long long off_shadow_lo, long long off_pad,
                long long off_stage, int stage_bf16, unsigned long long multicast_base, long long off_grads_bf16,
                float* gred, long long gred_elems, float* slots, long long slots_elems) {
  Plan* PP = as_plan(plan);
  if (!PP || !PP->bound) return -5;
  Plan& P = *PP;
  if (!peer_bases || !gred || !slots || world < 1 || world > kDpMaxWorld || rank < 0 || rank >= world) return -1;
  if (!P.m || !P.v) return -6;
  DpParams d;
  memset(&d, 0, sizeof(d));
  d.rank = rank; d.world = world;
  for (int p = 0; p < world; ++p) {
    uint8_t* base = reinterpret_cast<uint8_t*>(static_cast<uintptr_t>(peer_bases[p]));
    if (!base) return -2;
    d.params[p] = reinterpret_cast<float*>(base + off_params);
    d.sh_hi[p] = reinterpret_cast<__nv_bfloat16*>(base + off_shadow_hi);
    d.sh_lo[p] = P.sh_lo ? reinterpret_cast<__nv_bfloat16*>(base + off_shadow_lo) : nullptr;
    d.pad[p] = reinterpret_cast<unsigned int*>(base + off_pad);
    d.stage[p] = base + off_stage;
    dp_segments(P, p, world, d.seg_lo[p], d.seg_hi[p]);
    size_t off = 0;
    for (int s = 0; s < kDpMaxSeg; ++s) { d.seg_goff[p][s] = off; off += d.seg_hi[p][s] - d.seg_lo[p][s]; }
  }
  // the plan must already be bound to THIS rank's slices of the symmetric buffer
  const float* my_grads = reinterpret_cast<const float*>(reinterpret_cast<uint8_t*>(static_cast<uintptr_t>(peer_bases[rank])) + off_grads);
  if (d.params[rank] != P.params || my_grads != P.grads || d.sh_hi[rank] != P.sh_hi || (P.sh_lo && d.sh_lo[rank] != P.sh_lo))
    return -3;
  d.grads = P.grads;
  if (multicast_base) {
    uint8_t* mc = reinterpret_cast<uint8_t*>(static_cast<uintptr_t>(multicast_base));
    d.mc_params = reinterpret_cast<float*>(mc + off_params);
    d.mc_sh_hi = reinterpret_cast<__nv_bfloat16*>(mc + off_shadow_hi);
    d.mc_sh_lo = P.sh_lo ? reinterpret_cast<__nv_bfloat16*>(mc + off_shadow_lo) : nullptr;
    // pull mode (off_grads_bf16 >= -1): the reduce-scatter is one multimem.ld_reduce kernel over the gradient arenas; a
    // non-negative offset names the bf16 gradient mirror every rank keeps in the symmetric allocation
    if (off_grads_bf16 >= -1) {
      d.mc_grads = reinterpret_cast<const float*>(mc + off_grads);
      if (off_grads_bf16 >= 0) {
        uint8_t* mine = reinterpret_cast<uint8_t*>(static_cast<uintptr_t>(peer_bases[rank]));
        if (reinterpret_cast<__nv_bfloat16*>(mine + off_grads_bf16) != P.grads_bf) return -8;   // plan not bound with this mirror
        d.mc_grads_bf = reinterpret_cast<const __nv_bfloat16*>(mc + off_grads_bf16);
      }
    }
  }
  d.nseg = kDpMaxSeg;
  d.slot_cap = dp_slot_cap(P, world);
  d.stage_bf16 = stage_bf16 ? 1 : 0;
  if (gred_elems < static_cast<long long>(d.slot_cap) || slots_elems < 1024) return -4;
  d.gred = gred; d.slots = slots;
  d.a = adamw_params(P);
  P.dp = d;
  P.dp_bound = true;
  {
    // routing table of the fused push (fnd_gemm.cuh: DpRoute); same geometry as dp_segments
    DpRoute r;
    memset(&r, 0, sizeof(r));
    r.rank = rank; r.world = world;
    r.grads = P.grads;
    r.slot_cap = static_cast<uint32_t>(d.slot_cap);
    const size_t a0 = static_cast<size_t>(P.L.at("fusion.fuse_mlp.0.weight")), a1 = static_cast<size_t>(P.L.at("clf.pre.0.weight"));
    r.a0 = static_cast<uint32_t>(a0); r.a1 = static_cast<uint32_t>(a1);
    const size_t rb[kDpMaxSeg][2] = {{a0, a1}, {0, a0}, {a1, static_cast<size_t>(P.L.n_hot)}};
    for (int s = 0; s < kDpMaxSeg; ++s) {
      const size_t n = rb[s][1] - rb[s][0];
      size_t per = (n + world - 1) / world;
      per = (per + 1023) / 1024 * 1024;
      r.r_lo[s] = static_cast<uint32_t>(rb[s][0]);
      r.per[s] = static_cast<uint32_t>(per);
    }
    for (int p = 0; p < world; ++p) {
      r.stage[p] = static_cast<__nv_bfloat16*>(d.stage[p]);
      for (int s = 0; s < kDpMaxSeg; ++s) r.goff[p][s] = static_cast<uint32_t>(d.seg_goff[p][s]);
    }
    P.dp_route = r;
  }
  if (!P.ev_fork) {
    FND_CUDA_OK(cudaEventCreateWithFlags(&P.ev_fork, cudaEventDisableTiming));
    FND_CUDA_OK(cudaEventCreateWithFlags(&P.ev_join, cudaEventDisableTiming));
    FND_CUDA_OK(cudaEventCreateWithFlags(&P.ev_fork2, cudaEventDisableTiming));
    FND_CUDA_OK(cudaEventCreateWithFlags(&P.ev_join2, cudaEventDisableTiming));
  }
  return 0;
}

int fnd_dp_shard_ranges(const void* plan, int rank, int world, long long* lo3, long long* hi3) {
  if (!plan || !lo3 || !hi3 || world < 1 || rank < 0 || rank >= world) return -1;
  const Plan* P = static_cast<const Plan*>(plan);
  size_t lo[kDpMaxSeg], hi[kDpMaxSeg];
  dp_segments(*P, rank, world, lo, hi);
  for (int s = 0; s < kDpMaxSeg; ++s) { lo3[s] = static_cast<long long>(lo[s]); hi3[s] = static_cast<long long>(hi[s]); }
  return kDpMaxSeg;
}

static int dp_push(Plan& P, int s0, int s1, int bank, int ctr, int grid, int block, cudaStream_t st) {
  if (P.dp.stage_bf16) FND_CUDA_OK(launch_k(dp_push_kernel<true>, grid, block, 0, st, false, P.dp, s0, s1, bank, ctr));
  else FND_CUDA_OK(launch_k(dp_push_kernel<false>, grid, block, 0, st, false, P.dp, s0, s1, bank, ctr));
  return 0;
}

// The deferred (range 0) optimizer launch + the wait for every peer's deferred shadows. No-op on the device when nothing
// is pending.
static int dp_deferred(Plan& P, int grid, cudaStream_t st) {
  P.dp.a = adamw_params(P);
  FND_CUDA_OK(launch_k(dp_adamw_kernel, grid, 256, 0, st, false, P.dp, 0, 1, 1, static_cast<int>(kPadCounterEarly)));
  FND_CUDA_OK(launch_k(dp_wait_kernel, 1, 32, 0, st, false, P.dp, 1));
  return 0;
}

static int dp_tail(Plan& P, bool early_done, bool defer, cudaStream_t st) {
  P.dp.a = adamw_params(P);
  if (P.dp.mc_grads && !early_done) {
    // NVSwitch multicast available: the reduce-scatter is one in-switch-reduction kernel
    FND_CUDA_OK(launch_k(dp_pull_kernel, kDpGrid, 256, 0, st, false, P.dp));
    mark(P, "dp_pull", st);
  } else if (P.dp_fused_now) {
    // fused push: the wgrad epilogue already delivered the GEMM-weight tiles; push the two un-routed arena intervals
    // (pre.0.weight and the non-GEMM parameters), raise the flags, reduce all N staged pieces
    const size_t r0 = static_cast<size_t>(P.L.at("clf.pre.0.weight"));
    // (with aux_dim == 0 pre.0.weight has 16-byte rows and is routed like every other GEMM weight: empty interval)
    const size_t r0e = P.d.aux_dim ? r0 + static_cast<size_t>(P.H) * (P.H + P.d.aux_dim) : r0;
    FND_CUDA_OK(launch_k(dp_push_residual_kernel, 32, 256, 0, st, false, P.dp, r0, (r0e + 63) / 64 * 64,
                         static_cast<size_t>(P.L.n_shadow), static_cast<size_t>(P.L.n_hot)));
    mark(P, "dp_push_residual", st);
    FND_CUDA_OK(launch_k(dp_reduce_kernel<true>, kDpGrid, 256, 0, st, false, P.dp, 0, 1));
    mark(P, "dp_reduce", st);
  } else {
    // without an early push, ONE launch moves all three ranges (it raises the late flags; the early bank is not used)
    FND_OK(dp_push(P, early_done ? 1 : 0, kDpMaxSeg, kPadReadyLate, kPadCounter, kDpGrid, 256, st));
    mark(P, "dp_push", st);
    if (P.dp.stage_bf16) FND_CUDA_OK(launch_k(dp_reduce_kernel<true>, kDpGrid, 256, 0, st, false, P.dp, early_done ? 1 : 0, 0));
    else FND_CUDA_OK(launch_k(dp_reduce_kernel<false>, kDpGrid, 256, 0, st, false, P.dp, early_done ? 1 : 0, 0));
    mark(P, "dp_reduce", st);
  }
  FND_CUDA_OK(launch_k(dp_adamw_kernel, kDpGrid, 256, 0, st, false, P.dp, defer ? 1 : 0, static_cast<int>(kDpMaxSeg), 0,
                       static_cast<int>(kPadCounter)));
  mark(P, "dp_adamw", st);
  FND_CUDA_OK(launch_k(dp_wait_kernel, 1, 32, 0, st, false, P.dp, 0));
  mark(P, "dp_wait", st);
  return 0;
}

int fnd_dp_optimizer_step(void* plan, void* stream) {
  FND_PLAN(plan);
  if (!P.dp_bound) return -7;
  FND_OK(dp_deferred(P, kDpGrid, st));          // a deferred update of an earlier step must land first
  return dp_tail(P, false, false, st);
}

int fnd_dp_flush(void* plan, void* stream) {
  FND_PLAN(plan);
  if (!P.dp_bound) return -7;
  return dp_deferred(P, kDpGrid, st);
}

int fnd_eval_step(void* plan, const fnd_inputs* in, void* stream) {
  FND_PLAN(plan);
  if (!in) return -1;
  FND_OK(fusion_forward_impl(P, in, 0, false, st));
  FND_OK(fusion_head_impl(P, st));
  FND_OK(classifier_gemms_impl(P, 0, st));
  HeadParams h = head_params(P, 0);
  if (in->labels) return run_head<true, true, false>(P, h, st);
  return run_head<true, false, false>(P, h, st);
}

int fnd_debug_set_cluster_splitk(int on) {
  const int old = cluster_splitk_flag();
  cluster_splitk_flag() = on ? 1 : 0;
  return old;
}

int fnd_debug_set_launch_limit(void* plan, int limit) {
  Plan* PP = as_plan(plan);
  if (!PP) return -1;
  PP->launch_limit = limit;
  return 0;
}

int fnd_profile_begin(void* plan, void* stream) {
  FND_PLAN(plan);
  for (auto& m : P.marks) cudaEventDestroy(m.second);
  P.marks.clear();
  P.profiling = true;
  mark(P, "begin", st);
  return 0;
}

int fnd_profile_end(void* plan, void* stream, char* names, float* ms, int cap, int* count) {
  FND_PLAN(plan);
  if (!names || !ms || !count || cap < 1) return -1;
  P.profiling = false;
  FND_CUDA_OK(cudaStreamSynchronize(st));
  // aggregate by name, in order of first appearance: ms[i] = total time attributed to kernel `names[i]`
  int n = 0;
  for (size_t i = 1; i < P.marks.size(); ++i) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, P.marks[i - 1].second, P.marks[i].second) != cudaSuccess) t = 0.f;
    int slot = -1;
    for (int j = 0; j < n; ++j)
      if (strncmp(names + 64 * j, P.marks[i].first, 63) == 0) { slot = j; break; }
    if (slot < 0) {
      if (n >= cap) continue;
      slot = n++;
      strncpy(names + 64 * slot, P.marks[i].first, 63);
      names[64 * slot + 63] = 0;
      ms[slot] = 0.f;
    }
    ms[slot] += t;
  }
  *count = n;
  for (auto& m : P.marks) cudaEventDestroy(m.second);
  P.marks.clear();
  return 0;
}

int fnd_launch_count(const void* plan, const char* entry) {
  if (!plan || !entry) return -1;
  const std::string e(entry);
  const int fusion_fwd = 6;   // prep, proj, qkv, assemble, f0, f1
  if (e == "fusion_forward") return fusion_fwd + 1;
  if (e == "classifier_forward") return 3;
  if (e == "eval_step") return fusion_fwd + 1 + 2 + 1;
  if (e == "train_fwd_bwd") return fusion_fwd + 2 + 1 + 4 + 1 + 1 + 1 + 1;   // ... wgrad(+finalize CTAs), grad_norm
  if (e == "clip_adamw_step") return 3;
  if (e == "train_step") return fusion_fwd + 2 + 1 + 4 + 1 + 1 + 1 + 1;       // ... wgrad(+finalize CTAs), adamw
  return -2;
}

// ------------------------------- raw GEMM utility -------------------------------
size_t fnd_gemm_scratch_bytes(int M, int N, int bn, int splits) {
  size_t b = align_up(sizeof(GemmProblem), 256);
  b += 256;                                                                   // error flag
  b += align_up(sizeof(int) * 2 * ceil_div(M, kGemmBM) * ceil_div(N, bn), 256);   // split-K counters (arrive, depart)
  if (splits > 1) b += splitk_ws_floats(M, N, bn, splits) * sizeof(float);
  return b + 256;
}

static int gemm_bf16_impl(const void* a_hi, const void* a_lo, int a_pitch, int a_mn, const void* b_hi, const void* b_lo,
                          int b_pitch, int b_mn, float* c, int c_pitch, int M, int N, int K, int bn, int splits, int ncombo,
                          void* scratch, size_t scratch_bytes, void* stream, long long* stamps, int reps, bool sync = true,
                          int* user_err = nullptr);

int fnd_gemm_bf16_async(const void* a_hi, int a_pitch, int a_mn, const void* b_hi, int b_pitch, int b_mn, float* c,
                        int c_pitch, int M, int N, int K, int bn, int splits, void* scratch, size_t scratch_bytes,
                        int* err_flag, void* stream) {
  return gemm_bf16_impl(a_hi, nullptr, a_pitch, a_mn, b_hi, nullptr, b_pitch, b_mn, c, c_pitch, M, N, K, bn, splits, 1,
                        scratch, scratch_bytes, stream, nullptr, 1, false, err_flag);
}

int fnd_gemm_bf16(const void* a_hi, const void* a_lo, int a_pitch, int a_mn, const void* b_hi, const void* b_lo,
                  int b_pitch, int b_mn, float* c, int c_pitch, int M, int N, int K, int bn, int splits, int ncombo,
                  void* scratch, size_t scratch_bytes, void* stream) {
  return gemm_bf16_impl(a_hi, a_lo, a_pitch, a_mn, b_hi, b_lo, b_pitch, b_mn, c, c_pitch, M, N, K, bn, splits, ncombo,
                        scratch, scratch_bytes, stream, nullptr, 1);
}

int fnd_gemm_bf16_probe(const void* a_hi, const void* a_lo, int a_pitch, int a_mn, const void* b_hi, const void* b_lo,
                        int b_pitch, int b_mn, float* c, int c_pitch, int M, int N, int K, int bn, int splits, int ncombo,
                        void* scratch, size_t scratch_bytes, void* stream, long long* stamps, int reps) {
  return gemm_bf16_impl(a_hi, a_lo, a_pitch, a_mn, b_hi, b_lo, b_pitch, b_mn, c, c_pitch, M, N, K, bn, splits, ncombo,
                        scratch, scratch_bytes, stream, stamps, reps);
}
}  // extern "C"

static int gemm_bf16_impl(const void* a_hi, const void* a_lo, int a_pitch, int a_mn, const void* b_hi, const void* b_lo,
                          int b_pitch, int b_mn, float* c, int c_pitch, int M, int N, int K, int bn, int splits, int ncombo,
                          void* scratch, size_t scratch_bytes, void* stream, long long* stamps, int reps, bool sync,
                          int* user_err) {
  if (!a_hi || !b_hi || !c || !scratch) return -1;
  if (scratch_bytes < fnd_gemm_scratch_bytes(M, N, bn, splits)) return -2;
  if ((reinterpret_cast<uintptr_t>(scratch) & 255) != 0) return -3;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  uint8_t* base = static_cast<uint8_t*>(scratch);
  GemmProblem* dtab = reinterpret_cast<GemmProblem*>(base);
  size_t off = align_up(sizeof(GemmProblem), 256);
  int* derr = reinterpret_cast<int*>(base + off);
  off += 256;
  int* dctr = reinterpret_cast<int*>(base + off);
  const size_t ctr_bytes = align_up(sizeof(int) * 2 * ceil_div(M, kGemmBM) * ceil_div(N, bn), 256);
  off += ctr_bytes;
  float* dws = reinterpret_cast<float*>(base + off);

  EpiParams epi = epi_zero();
  epi.out_f32 = c;
  epi.f32_pitch = c_pitch;
  Operand A{static_cast<const __nv_bfloat16*>(a_hi), static_cast<const __nv_bfloat16*>(a_lo), a_pitch, a_mn != 0};
  Operand B{static_cast<const __nv_bfloat16*>(b_hi), static_cast<const __nv_bfloat16*>(b_lo), b_pitch, b_mn != 0};
  GemmProblem hp;
  int r = fill_problem(hp, A, B, M, N, K, bn, splits, ncombo, kEvictNormal, kEvictNormal, dws, dctr, epi);
  if (r) return r;
  const int grid = finish_table(&hp, 1);
  FND_CUDA_OK(init_gemm_attrs());
  FND_CUDA_OK(cudaMemsetAsync(derr, 0, 256 + ctr_bytes, st));
  RunCtx ctx{user_err ? user_err : derr, nullptr, 0, stamps};
  const int kind = (a_mn ? (b_mn ? 2 : 3) : (b_mn ? 1 : 0));
  (void)dtab;
  for (int r = 0; r < (reps < 1 ? 1 : reps); ++r) FND_CUDA_OK(launch_gemm(kind, &hp, 1, grid, ctx, st));
  if (!sync) return 0;
  int herr = 0;
  FND_CUDA_OK(cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, st));
  FND_CUDA_OK(cudaStreamSynchronize(st));
  return herr;
}
