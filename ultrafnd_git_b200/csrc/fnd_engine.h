// fnd_engine.h — host side of the fusion hot path: parameter arena layout, per-batch plan (workspace carve-up,
// TMA descriptors, GEMM problem tables, finalize job tables) and the launch sequences of every entry point.
//
// Data layout in HBM (all row-major):
//   arena (fp32)  : [GEMM weights | bypass.weight | NODE gates | biases | thresholds | leaf tables | evidence MLPs | cold]
//   shadows (bf16): same element offsets as the arena for the GEMM weights (+ a re-pitched pre.0.weight)
//   workspace     : DevState | packed inputs Xbf [B,dsum] | P [B,5H] | Q [B,9H] | fused_cat [B,16H] | ... (see carve())
#pragma once
#include "../../include/fnd_b200.h"
#include "fnd_gemm_host.h"
#include "fnd_optim.cuh"
#include "fnd_dp.cuh"
#include <cstdlib>
#include <map>
#include <string>
#include <vector>

namespace fnd {

inline long long align64(long long x) { return (x + 63) / 64 * 64; }

// -------------------------------------------------------------------------------------------------
// Arena layout
// -------------------------------------------------------------------------------------------------
struct ParamEntry {
  std::string name;
  long long offset;
  int ndim, rows, cols;
  int hot;
  long long numel() const { return ndim == 0 ? 1 : (ndim == 1 ? rows : static_cast<long long>(rows) * cols); }
};

struct ArenaLayout {
  std::vector<ParamEntry> entries;
  std::map<std::string, long long> off;
  long long n_shadow = 0, n_hot = 0, n_total = 0;
  long long rp_elems = 0;    // re-pitched pre.0.weight shadow (elements)
  int rp_pitch = 0;
  int evstride = 0;
  long long at(const std::string& k) const { return off.at(k); }
};

inline int dims_ok(const fnd_dims& d) {
  if (d.hidden != 512 && d.hidden != 1024) return 0;
  const int ds[5] = {d.d_text, d.d_audio, d.d_visual, d.d_temporal, d.d_gnn};
  for (int i = 0; i < 5; ++i)
    if (ds[i] <= 0 || ds[i] % 64) return 0;
  if (d.aux_dim != 0 && d.aux_dim != 2) return 0;
  if (d.trees < 1 || d.depth < 1 || d.depth > 4 || d.trees * d.depth > kMaxTD) return 0;
  return 1;
}

inline ArenaLayout make_layout(const fnd_dims& d) {
  ArenaLayout L;
  const int H = d.hidden;
  long long cur = 0;
  auto add = [&](const std::string& name, int ndim, int rows, int cols, int hot) {
    ParamEntry e{name, cur, ndim, rows, cols, hot};
    L.entries.push_back(e);
    L.off[name] = cur;
    cur += e.numel();
  };
  auto pad = [&]() { cur = align64(cur); };
  // ---- GEMM weights (shadowed) ----
  const char* pn[5] = {"text_proj", "audio_proj", "visual_proj", "temporal_proj", "gnn_proj"};
  const int pd[5] = {d.d_text, d.d_audio, d.d_visual, d.d_temporal, d.d_gnn};
  const int nmod = d.use_gnn ? 5 : 4;
  for (int i = 0; i < nmod; ++i) { add(std::string("fusion.") + pn[i] + ".weight", 2, H, pd[i], 1); pad(); }
  // q/k/v weights stacked by the projection's INPUT so one GEMM serves each input (Q layout, fnd_rows.cuh)
  const char* qn[9] = {"attn_tv.q", "attn_ta.q", "attn_tv.k", "attn_tv.v", "attn_vu.q", "attn_ta.k", "attn_ta.v",
                       "attn_vu.k", "attn_vu.v"};
  for (int i = 0; i < 9; ++i) add(std::string("fusion.") + qn[i] + ".weight", 2, H, H, 1);
  pad();
  const int nslots = d.use_gnn ? 16 : 15;
  add("fusion.fuse_mlp.0.weight", 2, 2 * H, nslots * H, 1); pad();
  add("fusion.fuse_mlp.3.weight", 2, H, 2 * H, 1); pad();
  add("clf.pre.0.weight", 2, H, H + d.aux_dim, 1); pad();
  add("clf.pre.3.weight", 2, H, H, 1); pad();
  L.n_shadow = cur;
  // ---- everything else that trains ----
  add("clf.bypass.weight", 2, 2, H, 1); pad();
  for (int t = 0; t < d.trees; ++t)
    for (int k = 0; k < d.depth; ++k) add("clf.node.trees." + std::to_string(t) + ".gates." + std::to_string(k), 1, H, 0, 1);
  pad();
  for (int i = 0; i < nmod; ++i) add(std::string("fusion.") + pn[i] + ".bias", 1, H, 0, 1);
  pad();
  for (int i = 0; i < 9; ++i) add(std::string("fusion.") + qn[i] + ".bias", 1, H, 0, 1);
  pad();
  add("fusion.fuse_mlp.0.bias", 1, 2 * H, 0, 1); pad();
  add("fusion.fuse_mlp.3.bias", 1, H, 0, 1); pad();
  add("clf.pre.0.bias", 1, H, 0, 1); pad();
  add("clf.pre.3.bias", 1, H, 0, 1); pad();
  add("clf.bypass.bias", 1, 2, 0, 1); pad();
  for (int t = 0; t < d.trees; ++t)
    for (int k = 0; k < d.depth; ++k) add("clf.node.trees." + std::to_string(t) + ".thresh." + std::to_string(k), 1, 1, 0, 1);
  pad();
  for (int t = 0; t < d.trees; ++t) add("clf.node.trees." + std::to_string(t) + ".leaf_logits", 2, 1 << d.depth, 2, 1);
  pad();
  const char* an[3] = {"attn_tv", "attn_ta", "attn_vu"};
  L.evstride = static_cast<int>(align64(5 * H + 1));
  for (int k = 0; k < 3; ++k) {
    const std::string b = std::string("fusion.") + an[k] + ".evidence_proj.";
    add(b + "0.weight", 2, H, 3, 1);
    add(b + "0.bias", 1, H, 0, 1);
    add(b + "2.weight", 2, 1, H, 1);
    add(b + "2.bias", 1, 1, 0, 1);
    pad();
  }
  L.n_hot = cur;
  // ---- cold: present in state_dict / parameters() but never updated by the reference step ----
  add("fusion.classifier.weight", 2, 2, H, 0); pad();
  add("fusion.classifier.bias", 1, 2, 0, 0); pad();
  add("clf.temperature", 0, 0, 0, 0);
  for (int t = 0; t < d.trees; ++t) add("clf.node.trees." + std::to_string(t) + ".tau", 0, 0, 0, 0);
  pad();
  add("fusion.semantic.text_proj.0.weight", 2, 512, 512, 0);
  add("fusion.semantic.text_proj.0.bias", 1, 512, 0, 0);
  add("fusion.semantic.vision_proj.0.weight", 2, 512, 512, 0);
  add("fusion.semantic.vision_proj.0.bias", 1, 512, 0, 0);
  pad();
  L.n_total = cur;
  L.rp_pitch = d.aux_dim ? H + 8 : H;
  L.rp_elems = d.aux_dim ? static_cast<long long>(H) * L.rp_pitch : 0;
  return L;
}

// -------------------------------------------------------------------------------------------------
// Plan
// -------------------------------------------------------------------------------------------------
struct WsBuf { long long off; long long bytes; };

struct GemmTable {
  std::vector<GemmProblem> host;
  GemmProblem* dev = nullptr;
  int kind = 0;     // 0 fwd (K,K), 1 dgrad (K,MN), 2 wgrad (MN,MN)
  int grid = 0;
};
struct FinTable {
  std::vector<FinJob> host;
  FinJob* dev = nullptr;
  int grid = 0;
};

// Tile width and split-K factor of one GEMM launch (all problems of a launch share them).
struct GemmCfg { int bn = 64; int splits = 1; };

// At the reference's batch sizes every GEMM is a latency chain: prefer the NARROWEST tile that still fits the launch
// in one wave of 148 CTAs (more SMs streaming weights, shorter epilogue per CTA). Long-K problems (K >= 2048) take
// 128-wide tiles instead so the activation panel is re-read by few CTAs, and are split along K (one CTA per SM, see
// fnd_gemm.cuh) when a CTA would otherwise stream more than ~600 KB by itself.
inline GemmCfg pick_cfg(int B, const std::vector<std::pair<int, int>>& nk, bool b_mn, int ncombo, bool allow_split,
                        int cluster_split = 0) {
  const int tm = ceil_div(B, kGemmBM);
  int kmax = 0;
  for (auto& p : nk) kmax = p.second > kmax ? p.second : kmax;
  const int kb = ceil_div(kmax, kGemmBK);
  auto ctas_at = [&](int bn) {
    int c = 0;
    for (auto& p : nk) c += tm * ceil_div(p.first, bn);
    return c;
  };
  GemmCfg c;
  c.bn = 128;
  if (kb < 32) {
    const int cands[4] = {16, 32, 64, 128};
    for (int i = b_mn ? 2 : 0; i < 4; ++i)
      if (ctas_at(cands[i]) <= 148) { c.bn = cands[i]; break; }
  }
  const int ctas = ctas_at(c.bn);
  const long long per_cta = static_cast<long long>(kb) * ncombo * (kGemmStageBytesA + c.bn * kGemmBK * 2);
  // (a CTA ingests ~55 GB/s through TMA in 128-byte rows; the split-K exchange costs ~2.5 us = ~140 KB of streaming)
  if (allow_split && per_cta > 400 * 1024) {
    int sp = static_cast<int>(per_cta / (128 * 1024));
    if (sp > 148 / ctas) sp = 148 / ctas;
    if (sp > 16) sp = 16;
    if (sp > c.bn / 2) sp = c.bn / 2;
    if (sp > kb) sp = kb;
    c.splits = sp < 1 ? 1 : sp;
  } else if (allow_split && ctas <= 16 && kb >= 8) {
    // a handful of wide (MN-major B: >= 64 columns) tiles: four splits shorten both the operand stream and the
    // per-CTA epilogue (the fix-up is spread over the splits) by more than the exchange costs (measured; with the cluster
    // exchange too: 2 / 4 / 8 splits give a 230.9 / 223.5 / 225.1 us step)
    c.splits = 4;
  }
  if (c.splits == 1 && cluster_split > 1 && kb >= 8) {
    // Cluster split-K (fnd_gemm.cuh: partial tiles exchanged through distributed shared memory, ~1 us instead of the ~2.5 us
    // L2 rendezvous): a latency-chain GEMM whose 32 CTAs each stream the whole 128 x K activation panel becomes 128 CTAs
    // that stream a quarter of it. Narrowest tile whose split grid still fits one wave.
    const int cands[4] = {16, 32, 64, 128};
    for (int i = b_mn ? 2 : 0; i < 4; ++i)
      if (ctas_at(cands[i]) * cluster_split <= 148) { c.bn = cands[i]; c.splits = cluster_split; break; }
  }
  return c;
}

struct Plan {
  fnd_dims d;
  ArenaLayout L;
  int B = 0, mode = 0, ncombo = 1;
  int H = 0, nmod = 0, nslots = 0, dsum = 0, TD = 0, leaves = 0;
  int xoff[5] = {0, 0, 0, 0, 0}, xdim[5] = {0, 0, 0, 0, 0};
  int n_asm_ctas = 0;
  GemmCfg cfg_proj, cfg_qkv, cfg_f0, cfg_f1, cfg_pre, cfg_dg_pre, cfg_dg_f1, cfg_dg_f0, cfg_dg_qkv;
  bool bound = false;
  int last_training = 0;
  // workspace
  std::map<std::string, WsBuf> bufs;
  std::vector<std::string> order;
  long long ws_bytes = 0;
  uint8_t* ws = nullptr;
  float *params = nullptr, *grads = nullptr, *m = nullptr, *v = nullptr;
  __nv_bfloat16 *sh_hi = nullptr, *sh_lo = nullptr;
  __nv_bfloat16* grads_bf = nullptr;   // optional bf16 mirror of the GEMM-weight gradients (fnd_plan_set_grad_mirror)
  // tables
  GemmTable fwd_proj, fwd_qkv, fwd_f0, fwd_f1, fwd_p0, fwd_p1;
  GemmTable dg_p1, dg_p0_fused, dg_p0_split, dg_f1, dg_f0, dg_qkv;
  GemmTable wg_all, wg_clf, wg_fus, wg_early, wg_rest;
  DpRoute dp_route;                 // fused-push routing table (filled by fnd_dp_bind)
  float* loss_mirror = nullptr;     // fnd_set_loss_mirror: pinned device-mapped host ring the loss job also writes (kept over re-binds)
  int loss_ring = 0;
  bool joined_side = false;         // fnd_train_step_overlap forked / joined the side stream in the step being enqueued
  bool dp_fused_now = false;        // the step being enqueued uses the fused push (set by train_fwd_bwd_impl, read by dp_tail)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;   // data-parallel overlap: side-stream fork / join (early push)
  cudaEvent_t ev_fork2 = nullptr, ev_join2 = nullptr; // ... and for the deferred update at the start of a step
  FinTable fin_all, fin_clf, fin_fus;
  int total_slots = 0;
  // optional per-kernel timing (bench/profiling only): an event is recorded after every launch
  int dbg_launch = 0;            // probe builds: index of the next launch's stamp region
  bool pdl_next = false;         // the next launch may carry the programmatic-serialization attribute
  int launch_limit = -1, launch_seq = 0;   // debug: issue only the first `launch_limit` launches of an entry point
  bool dp_bound = false;         // fnd_dp_bind succeeded: dp holds the peer-mapped buffers of every rank
  DpParams dp;
  bool profiling = false;
  std::vector<std::pair<const char*, cudaEvent_t>> marks;

  template <class T> T* buf(const std::string& name) const {
    auto it = bufs.find(name);
    return it == bufs.end() ? nullptr : reinterpret_cast<T*>(ws + it->second.off);
  }
  DevState* state() const { return buf<DevState>("state"); }
  float* W(const std::string& k) const { return params + L.at(k); }
  float* G(const std::string& k) const { return grads + L.at(k); }
  const __nv_bfloat16* Sh(const std::string& k) const { return sh_hi + L.at(k); }
  const __nv_bfloat16* Sl(const std::string& k) const { return sh_lo ? sh_lo + L.at(k) : nullptr; }
};

inline void plan_add(Plan& P, const std::string& name, long long bytes) {
  WsBuf b{P.ws_bytes, bytes};
  P.bufs[name] = b;
  P.order.push_back(name);
  P.ws_bytes += (bytes + 255) / 256 * 256;
}
// bf16 activation plane(s): "<name>_hi" always, "<name>_lo" in fp32x3 mode
inline void plan_add_bf(Plan& P, const std::string& name, long long elems) {
  plan_add(P, name + "_hi", elems * 2);
  if (P.ncombo == 3) plan_add(P, name + "_lo", elems * 2);
}

inline void carve(Plan& P) {
  const long long B = P.B, H = P.H;
  const int TD = P.TD;
  plan_add(P, "state", sizeof(DevState));
  plan_add(P, "aux", B * 2 * 4);
  plan_add(P, "labels", B * 8);
  plan_add_bf(P, "xbf", B * P.dsum);
  plan_add(P, "P", B * 5 * H * 4);
  plan_add_bf(P, "pbf", B * 5 * H);
  plan_add(P, "Q", B * 9 * H * 4);
  plan_add(P, "rowstat", B * 16 * 4);
  plan_add_bf(P, "fused_cat", B * P.nslots * H);
  plan_add(P, "z_f0", B * 2 * H * 4);
  plan_add_bf(P, "h1", B * 2 * H);
  plan_add(P, "z_f1", B * H * 4);
  plan_add(P, "fused", B * H * 4);
  plan_add_bf(P, "fusedbf", B * H);
  plan_add(P, "fusion_logits", B * 2 * 4);
  plan_add(P, "z_p0", B * H * 4);
  plan_add_bf(P, "xp1", B * H);
  plan_add(P, "z_p1", B * H * 4);
  plan_add(P, "h", B * H * 4);
  plan_add_bf(P, "hbf", B * H);
  plan_add(P, "alpha", static_cast<long long>(kMaxTD) * H * 4);
  plan_add(P, "svals", B * 32 * 4);
  plan_add(P, "logits", B * 2 * 4);
  plan_add(P, "probs", B * 2 * 4);
  plan_add(P, "loss_row", B * 4);
  plan_add(P, "dlogits", B * 2 * 4);
  // backward
  plan_add(P, "dF", B * kDFCols * 4);
  plan_add_bf(P, "dFbf", B * kDFCols);
  plan_add(P, "leafc", B * P.d.trees * P.leaves * 2 * 4);
  plan_add(P, "dAraw", static_cast<long long>(64) * H * 4);
  plan_add_bf(P, "dz_p1", B * H);
  plan_add_bf(P, "dz_p0", B * H);
  plan_add(P, "dfused", B * H * 4);
  plan_add_bf(P, "dz_f1", B * H);
  plan_add_bf(P, "dz_f0", B * 2 * H);
  plan_add(P, "dcat", B * P.nslots * H * 4);
  plan_add(P, "dPdirect", B * 5 * H * 4);
  plan_add_bf(P, "dQ", B * 9 * H);
  plan_add_bf(P, "dP", B * 5 * H);
  P.n_asm_ctas = static_cast<int>(B < 296 ? B : 296);
  plan_add(P, "ev_partial", static_cast<long long>(P.n_asm_ctas) * 3 * P.L.evstride * 4);
  (void)TD;
  // per-launch tile / split-K configuration; split-K workspaces + (arrive, depart) counters where a launch splits
  {
    const int Hh = P.H, cat = P.nslots * P.H, nc = P.ncombo;
    std::vector<std::pair<int, int>> proj, qkv = {{2 * Hh, Hh}, {3 * Hh, Hh}, {2 * Hh, Hh}, {2 * Hh, Hh}};
    for (int i = 0; i < P.nmod; ++i) proj.push_back({Hh, P.xdim[i]});
    // (measured and not kept: cluster split-K x2 for the projections, 19.6 -> 21.0 us, and x4 with 128-column tiles for the
    //  q/k/v launch, 14.5 -> 25.4 us — their 80 / 144 narrow-tile CTAs already spread the panel over the chip)
    P.cfg_proj = pick_cfg(P.B, proj, false, nc, false);
    P.cfg_qkv = pick_cfg(P.B, qkv, false, nc, false);
    P.cfg_f0 = pick_cfg(P.B, {{2 * Hh, cat}}, false, nc, true);
    const int csk = cluster_splitk_enabled() ? 4 : 0;
    P.cfg_f1 = pick_cfg(P.B, {{Hh, 2 * Hh}}, false, nc, true, csk);
    P.cfg_pre = pick_cfg(P.B, {{Hh, Hh}}, false, nc, false, csk);
    P.cfg_dg_pre = pick_cfg(P.B, {{Hh, Hh}}, true, nc, true);
    P.cfg_dg_f1 = pick_cfg(P.B, {{2 * Hh, Hh}}, true, nc, true);
    P.cfg_dg_f0 = pick_cfg(P.B, {{cat, 2 * Hh}}, true, nc, true);
    P.cfg_dg_qkv = pick_cfg(P.B, {{Hh, 2 * Hh}, {Hh, 3 * Hh}, {Hh, 2 * Hh}, {Hh, 2 * Hh}}, true, nc, true);
    auto split_bufs = [&](const char* tag, const GemmCfg& c, int N) {
      if (c.splits <= 1) return;
      plan_add(P, std::string("splitws_") + tag, static_cast<long long>(splitk_ws_floats(P.B, N, c.bn, c.splits)) * 4);
      plan_add(P, std::string("splitctr_") + tag, static_cast<long long>(ceil_div(P.B, kGemmBM)) * ceil_div(N, c.bn) * 2 * 4);
    };
    split_bufs("f0", P.cfg_f0, 2 * Hh);
    split_bufs("f1", P.cfg_f1, Hh);
    split_bufs("p0", P.cfg_pre, Hh);
    split_bufs("p1", P.cfg_pre, Hh);
    split_bufs("dgf1", P.cfg_dg_f1, 2 * Hh);
    split_bufs("dgf0", P.cfg_dg_f0, cat);
    split_bufs("dgp1", P.cfg_dg_pre, Hh);
    split_bufs("dgp0f", P.cfg_dg_pre, Hh);
    split_bufs("dgp0s", P.cfg_dg_pre, Hh);
    for (int g = 0; g < 4; ++g) split_bufs(("dgqkv" + std::to_string(g)).c_str(), P.cfg_dg_qkv, Hh);
  }
  if (getenv("FND_DEBUG_STAMPS")) plan_add(P, "dbg", 8LL * 8 * 1024 * 40);     // clock64 stamps of the row kernels (probes)
  // per-CTA sum-of-squares slots (wgrad CTAs + finalize CTAs); generous upper bound, zero-initialised at bind
  plan_add(P, "slots", 16384 * 4);
  // device copies of the kernel tables
  plan_add(P, "tables", 64 * static_cast<long long>(sizeof(GemmProblem)) + 96 * static_cast<long long>(sizeof(FinJob)) + 4096);
}

// ---- small helpers for table construction ----
struct TableBuilder {
  Plan& P;
  explicit TableBuilder(Plan& p) : P(p) {}
  Operand act(const std::string& name, long long col, int pitch, bool mn) const {
    Operand o;
    o.hi = P.buf<__nv_bfloat16>(name + "_hi") + col;
    const __nv_bfloat16* lo = P.buf<__nv_bfloat16>(name + "_lo");
    o.lo = lo ? lo + col : nullptr;
    o.pitch = pitch;
    o.mn_major = mn;
    return o;
  }
  Operand weight(const std::string& key, int pitch, bool mn) const {
    Operand o;
    o.hi = P.Sh(key);
    o.lo = P.Sl(key);
    o.pitch = pitch;
    o.mn_major = mn;
    return o;
  }
  Operand weight_rp(bool mn) const {      // re-pitched pre.0.weight shadow
    Operand o;
    if (P.d.aux_dim) {
      o.hi = P.sh_hi + P.L.n_shadow;
      o.lo = P.sh_lo ? P.sh_lo + P.L.n_shadow : nullptr;
      o.pitch = P.L.rp_pitch;
    } else {
      o.hi = P.Sh("clf.pre.0.weight");
      o.lo = P.Sl("clf.pre.0.weight");
      o.pitch = P.H;
    }
    o.mn_major = mn;
    return o;
  }
};

inline EpiParams epi_zero() {
  EpiParams e;
  memset(&e, 0, sizeof(e));
  return e;
}

struct SplitAlloc { std::string ws_name, ctr_name; };

inline int add_problem(Plan& P, GemmTable& T, const Operand& A, const Operand& B, int M, int N, int K, int bn, int splits,
                       const EpiParams& epi, const std::string& tag, int b_static) {
  GemmProblem g;
  float* ws = nullptr;
  int* ctr = nullptr;
  // clamp the split factor exactly as fill_problem will
  const int kb_total = ceil_div(K, kGemmBK);
  if (splits > kb_total) splits = kb_total;
  if (splits > 1) {
    const int kps = ceil_div(kb_total, splits);
    splits = ceil_div(kb_total, kps);
  }
  if (splits > 1) {
    ws = P.buf<float>("splitws_" + tag);
    ctr = P.buf<int>("splitctr_" + tag);
    if (!ws || !ctr) return -30;
  }
  int r = fill_problem(g, A, B, M, N, K, bn, splits, P.ncombo, kEvictNormal, kEvictNormal, ws, ctr, epi, b_static);
  if (r) return r;
  T.host.push_back(g);
  return 0;
}

}  // namespace fnd
