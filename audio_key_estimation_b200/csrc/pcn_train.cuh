// pcn_train.cuh -- training step of PitchClassNet (BASELINE config 5): a train-mode forward
// that keeps every activation (bn_mode = 2) and the backward pass over them.  Included by pcn.cu after Fwd.
//
// Graph at the train_model.py defaults (num_layers 2), names used below:
//   L0  mel -> s0 [pool_semi conv, BN, LReLU] -> octave max -> q0 -> e0[i] [equivariant conv, BN, LReLU] -> pc0
//   L1  pc0 -> up [ConvTranspose, BN, LReLU] -> tile x8, cat with mel -> p[i] [7x7 circular conv, BN, LReLU]
//       -> s1 [pool_semi] -> octave max -> cat1 = [pc0 | .] -> e1[i] -> MaxPool (1,2) -> pcp
//   heads: tonic / key [equivariant 16->32 BN LReLU, 32->1], genre [1x7 16->32 BN LReLU, 2x7 32->1] -> masked mean (-> sigmoid)
// The backward walks this list in reverse.  Gradients land in a flat buffer with the layout of the parameter buffer.
#pragma once

namespace ake {

struct ConvSite {
  int id = -1;
  View in0, in1;   // in1.p == nullptr: single input
  ConvGeom g{};
  View z;          // raw convolution output (bias included)
  View a;          // after BatchNorm + LeakyReLU (== z for the last head convolutions)
  bool has_bn = false;
};

struct TrainTape {
  bool valid = false;
  int B = 0, T = 0;
  const void* ws = nullptr;
  size_t ws_off = 0;  // workspace bytes the forward used; the backward carves behind them
  const int* seq_len = nullptr;
  const float* key_out = nullptr;
  View mel, q0, up_z, up_a, cat1, pcp;
  ConvSite s0, s1;
  std::vector<ConvSite> e0, p, e1;
  ConvSite ht[2], hk[2], hg[2];
  bool genre = false;
  float* d_ss = nullptr;  // [scale | shift] of every BN site of this forward (batch statistics), n_ss each
  float* d_mi = nullptr;  // [mean | invstd], n_ss each
  // scratch of the tensor-core 7x7 convolutions (pcn_train_tc.cuh), carved by the forward and reused by the backward
  bool tc_ready = false;
  __half* tc_hi = nullptr;
  __half* tc_lo = nullptr;
  float* tc_raw = nullptr;
  bool eq_ready = false;
  __half* eq_hi = nullptr;
  __half* eq_lo = nullptr;
  __half* tcw_block = nullptr;
  std::map<int, long long> tcw_fwd, tcw_bwd, tcw_head;
  bool heads_tc = false;  // the heads' first convs ran on the tensor cores (their data gradients do too)
};

__global__ void fill_kernel(float* __restrict__ x, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) x[i] = v;
}

template <int KH, int KW, int SR, int RB, int RS>
static void launch_wgrad_t(WgradArgs a, int B, cudaStream_t st) {
  constexpr int RIN = (RB - 1) * SR + KH;
  const size_t smem = sizeof(float) * ((size_t)kWgCI * RIN * wg_xp(KW) + (size_t)RB * kWgGP);
  auto kern = conv_wgrad_kernel<KH, KW, SR, RB, RS>;
  ensure_dyn_smem(kern, smem);
  a.n_cob = cdiv(a.Cout, kWgCO);
  const int row_tiles = cdiv(a.rows_out, RB), y = cdiv(a.Cin, kWgCI) * a.n_cob;
  // enough blocks for two per SM, at least one 16-frame tile each
  const long long base = (long long)row_tiles * y * B;
  a.t_splits = (int)std::max<long long>(1, std::min<long long>(cdiv(a.T_out, kWgTB), cdiv64(2LL * sm_count(), base)));
  kern<<<dim3(row_tiles * a.t_splits, y, B), kWgCI * KH * RS, smem, st>>>(a);
  AKE_LAUNCHED();
}

static void launch_wgrad(const WgradArgs& a, const ConvGeom& g, int B, cudaStream_t st) {
  if (g.KH == 12 && g.KW == 7 && g.SR == 1) return launch_wgrad_t<12, 7, 1, 12, 2>(a, B, st);
  if (g.KH == 7 && g.KW == 7 && g.SR == 1) return launch_wgrad_t<7, 7, 1, 8, 4>(a, B, st);
  if (g.KH == 3 && g.KW == 3 && g.SR == 3) return launch_wgrad_t<3, 3, 3, 8, 8>(a, B, st);
  if (g.KH == 1 && g.KW == 7 && g.SR == 1) return launch_wgrad_t<1, 7, 1, 12, 4>(a, B, st);
  if (g.KH == 2 && g.KW == 7 && g.SR == 1) return launch_wgrad_t<2, 7, 1, 12, 4>(a, B, st);
  fail(AKE_ERR_UNSUPPORTED, "no weight-gradient kernel for KH=%d KW=%d SR=%d", g.KH, g.KW, g.SR);
}

// ------------------------------------------------------------------------------------------------ forward (kept)
void Fwd::run_keep(const float* mel, float* key_out, float* tonic_out, float* genre_out, TrainTape& tp) {
  const ake_pcn_config& cfg = p->cfg;
  if (cfg.num_layers != 2 || cfg.head_layers != 2)
    fail(AKE_ERR_UNSUPPORTED, "the training step is built for num_layers = 2, head_layers = 2 (train_model.py defaults)");
  if (cfg.max_pool) fail(AKE_ERR_UNSUPPORTED, "the training step does not cover opt.max_pool");
  if (cfg.resblock || cfg.denseblock || cfg.stay_sixth || cfg.p2pc_conv || cfg.pc2p_mem || cfg.local)
    fail(AKE_ERR_UNSUPPORTED, "the training step (backward pass) is built for the default architecture; the non-default switches run forward only");
  const int P = cfg.pitches, S = P / 3, k = cfg.kernel_size;
  if (!dry) p->taps.clear();
  d_stats = arena.take<double>(2 * (size_t)p->n_ss);
  d_ss_train = arena.take<float>(2 * (size_t)p->n_ss);
  float* d_mi = arena.take<float>(2 * (size_t)p->n_ss);
  if (!dry) AKE_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(double) * 2 * p->n_ss, st));
  tp = TrainTape();
  tp.B = B, tp.T = T, tp.seq_len = seq_len, tp.key_out = key_out, tp.genre = cfg.genre != 0;
  tp.d_ss = d_ss_train, tp.d_mi = d_mi;
  tp.mel.p = const_cast<float*>(mel), tp.mel.C = 1, tp.mel.R = P, tp.mel.T = T;

  d_mi_train = d_mi;  // bn_finalize_kernel also writes the mean / inverse standard deviation the backward pass needs
  auto stats_of = [&](const Conv& c, const View& z) { train_bn(c, z, 0); };
  auto bn_act = [&](const Conv& c, const View& z) {
    View a = alloc(z.C, z.R, z.T);
    if (!dry) {
      bn_act_out_kernel<<<ew_blocks(z.numel(B)), 256, 0, st>>>(z.p, a.p, B, z.C, z.R * z.T, d_ss_train + c.ss_off,
                                                               d_ss_train + p->n_ss + c.ss_off);
      AKE_LAUNCHED();
    }
    return a;
  };
  auto tail_conv_ok = [&](const Conv& c, const ConvGeom& g) {
    return c.Cout == 1 && g.KW == 7 && g.KH == c.KH && (c.KH == 12 || c.KH == 2) && g.SR == 1 && g.pad_t == 0 && !g.time_circ && g.row_off == 0 && c.Cin * c.KH * 7 <= 8192;
  };
  // conv (+ batch statistics) (+ BN + LReLU into a separate buffer)
  auto site = [&](int id, const View& in0, const View* in1, const ConvGeom& g, bool act_now) {
    const Conv& c = p->convs[id];
    ConvSite s;
    s.id = id, s.in0 = in0, s.g = g, s.has_bn = c.bn >= 0;
    if (in1) s.in1 = *in1;
    s.z = alloc(c.Cout, g.rows_out, g.T_out);
    if (tc_conv_ok(c, g, in0.T)) {
      // 7x7 circular conv on the tensor cores; the BatchNorm sums of z come out of the same pass
      tc_conv(in0, in1, c, false, nullptr, s.z, s.has_bn ? d_stats + 2 * c.ss_off : nullptr);
      s.a = s.z;
      if (s.has_bn) {
        train_bn_finalize(c, s.z.R * s.z.T);
        if (act_now) s.a = bn_act(c, s.z);
      }
      return s;
    }
    if (!in1 && tail_conv_ok(c, g) && !s.has_bn) {
      // last conv of a classifier head (Cin -> 1 channel): a dedicated latency-sized kernel
      if (!dry) {
        const size_t smem = sizeof(float) * ((size_t)c.Cin * c.KH * 7 + 256);
        auto kern = c.KH == 12 ? tail_conv_fwd_kernel<12> : tail_conv_fwd_kernel<2>;
        kern<<<dim3(g.rows_out, B), 256, smem, st>>>(in0.p, p->d_params + c.w_off, c.has_bias ? p->d_params + c.b_off : nullptr, s.z.p, c.Cin,
                                                     in0.R, g.rows_out, in0.T, g.T_out, g.row_circ);
        AKE_LAUNCHED();
      }
      s.a = s.z;
      return s;
    }
    if (!in1 && eq_conv_ok(c, g, in0.T)) {
      // equivariant 12 x 7 conv on the tensor cores (raw planar output), then its batch statistics
      eq_conv(in0, c, false, nullptr, nullptr, nullptr, s.z);
      s.a = s.z;
      if (s.has_bn) {
        stats_of(c, s.z);
        if (act_now) s.a = bn_act(c, s.z);
      }
      return s;
    }
    conv(id, in0, in1, g, s.z, 0, false, s.has_bn ? d_stats + 2 * c.ss_off : nullptr);  // the BatchNorm sums come out of the conv's epilogue
    s.a = s.z;
    if (s.has_bn) {
      train_bn_finalize(c, s.z.R * s.z.T);
      if (act_now) s.a = bn_act(c, s.z);
    }
    return s;
  };
  auto octmax = [&](const ConvSite& s, View& dst, int coff) {
    const Conv& c = p->convs[s.id];
    if (dry) return;
    octmax_kernel<<<ew_blocks((long long)B * c.Cout * 12 * s.z.T), 256, 0, st>>>(s.z.p, B, c.Cout, s.z.R, s.z.T, d_ss_train + c.ss_off,
                                                                               d_ss_train + p->n_ss + c.ss_off, 1, dst.p, dst.C, coff);
    AKE_LAUNCHED();
  };
  const ConvGeom g_sem{3, 3, 3, P, 0, 0, 1, 1, S, T};
  auto g_equiv = [&](int Tn, bool same) { return ConvGeom{12, k, 1, 12, 1, 0, same ? k / 2 : 0, 0, 12, same ? Tn : Tn - k + 1}; };
  const LayerPlan& l0 = p->layers[0];
  const LayerPlan& l1 = p->layers[1];

  {
    // operand images of the step's tensor-core convolutions (7x7 stack, both equivariant stacks), one launch
    std::vector<TcSite> cand;
    const ConvGeom gp0{k, k, 1, P, 1, -(k / 2), k / 2, 1, P, T};
    for (int id : l0.pc2pc) cand.push_back(TcSite{id, g_equiv(T, true), T, false});
    for (int id : l1.p2p) cand.push_back(TcSite{id, gp0, T, false});
    for (int id : l1.pc2pc) cand.push_back(TcSite{id, g_equiv(T, true), T, false});
    std::vector<int> head_first;
    {
      const Conv& ct = p->convs[p->tonic_head[0]];
      const Conv& ck = p->convs[p->key_head[0]];
      if (heads_tc_ok(ct, ck, T / 2) && ct.bn >= 0 && ck.bn >= 0 && eq_conv_ok(p->convs[l1.pc2pc.back()], g_equiv(T, true), T))
        head_first = {p->tonic_head[0], p->key_head[0]};
    }
    tp.heads_tc = !head_first.empty();
    tc_pack_images(cand, head_first);
  }
  // ---- layer 0 (models.py:359-369)
  tp.s0 = site(l0.sem, tp.mel, nullptr, g_sem, false);
  tp.q0 = alloc(1, 12, T);
  octmax(tp.s0, tp.q0, 0);
  {
    View src = tp.q0;
    for (int id : l0.pc2pc) {
      tp.e0.push_back(site(id, src, nullptr, g_equiv(T, true), true));
      src = tp.e0.back().a;
    }
  }
  const View pc0 = tp.e0.back().a;
  // ---- layer 1 (models.py:370-396)
  {
    const Conv& cu = p->convs[l1.up];
    tp.up_z = alloc(l1.prev_pc, 36, T);
    if (!dry) {
      upsixth_kernel<<<ew_blocks(tp.up_z.numel(B)), 256, 0, st>>>(pc0.p, p->d_params + cu.w_off, scale_of(cu, true), shift_of(cu, true), 0,
                                                                tp.up_z.p, B, l1.prev_pc, T);
      AKE_LAUNCHED();
    }
    stats_of(cu, tp.up_z);
    tp.up_a = bn_act(cu, tp.up_z);
  }
  {
    const ConvGeom gp{k, k, 1, P, 1, -(k / 2), k / 2, 1, P, T};
    for (size_t i = 0; i < l1.p2p.size(); ++i) {
      if (i == 0) tp.p.push_back(site(l1.p2p[i], tp.mel, &tp.up_a, gp, true));
      else tp.p.push_back(site(l1.p2p[i], tp.p.back().a, nullptr, gp, true));
    }
  }
  tp.s1 = site(l1.sem, tp.p.back().a, nullptr, g_sem, false);
  tp.cat1 = alloc(l1.prev_pc + l1.out_p, 12, T);
  if (!dry) {
    copy_channels_kernel<<<ew_blocks(pc0.numel(B)), 256, 0, st>>>(pc0.p, pc0.C, 0, tp.cat1.p, tp.cat1.C, 0, B, pc0.C, 12 * T, 0);
    AKE_LAUNCHED();
  }
  octmax(tp.s1, tp.cat1, l1.prev_pc);
  {
    View src = tp.cat1;
    for (size_t i = 0; i < l1.pc2pc.size(); ++i) {
      const bool last = i + 1 == l1.pc2pc.size();
      tp.e1.push_back(site(l1.pc2pc[i], src, nullptr, g_equiv(T, true), !last));
      src = tp.e1.back().a;
    }
  }
  const int T2 = T / 2;
  tp.pcp = alloc(l1.out_pc, 12, T2);
  {
    const ConvSite& s = tp.e1.back();
    const Conv& c = p->convs[s.id];
    if (!dry) {
      timepool_kernel<<<ew_blocks(tp.pcp.numel(B)), 256, 0, st>>>(s.z.p, B, c.Cout, 12, T, d_ss_train + c.ss_off, d_ss_train + p->n_ss + c.ss_off,
                                                                 1, tp.pcp.p);
      AKE_LAUNCHED();
    }
  }
  // ---- heads (models.py:713-742, 750-753)
  const int T1 = T2 - (k - 1), Tf = T1 - (k - 1);
  if (Tf <= 0) fail(AKE_ERR_INVALID, "T=%d is too short: the heads need more than %d frames after pooling", T, (k - 1) * cfg.head_layers);
  {
    const Conv& ct = p->convs[p->tonic_head[0]];
    const Conv& ck = p->convs[p->key_head[0]];
    if (heads_tc_ok(ct, ck, T2) && ct.bn >= 0 && ck.bn >= 0) {
      // first conv of both heads in ONE tensor-core pass (they read the same input), then each head's BatchNorm + LeakyReLU
      const Conv* hc[2] = {&ct, &ck};
      ConvSite* hs[2] = {&tp.ht[0], &tp.hk[0]};
      const int ids[2] = {p->tonic_head[0], p->key_head[0]};
      for (int h = 0; h < 2; ++h) {
        ConvSite s;
        s.id = ids[h], s.in0 = tp.pcp, s.g = g_equiv(T2, false), s.has_bn = true;
        s.z = alloc(hc[h]->Cout, 12, T1);
        *hs[h] = s;
      }
      heads_tc(tp.pcp, ct, ck, tp.ht[0].z, tp.hk[0].z);
      for (int h = 0; h < 2; ++h) {
        stats_of(*hc[h], hs[h]->z);
        hs[h]->a = bn_act(*hc[h], hs[h]->z);
      }
    } else {
      tp.ht[0] = site(p->tonic_head[0], tp.pcp, nullptr, g_equiv(T2, false), true);
      tp.hk[0] = site(p->key_head[0], tp.pcp, nullptr, g_equiv(T2, false), true);
    }
  }
  tp.ht[1] = site(p->tonic_head[1], tp.ht[0].a, nullptr, g_equiv(T1, false), false);
  tp.hk[1] = site(p->key_head[1], tp.hk[0].a, nullptr, g_equiv(T1, false), false);
  if (cfg.genre) {
    tp.hg[0] = site(p->genre_head[0], tp.pcp, nullptr, ConvGeom{1, k, 1, 12, 0, 0, 0, 0, 12, T1}, true);
    tp.hg[1] = site(p->genre_head[1], tp.hg[0].a, nullptr, ConvGeom{2, k, 1, 12, 0, 0, 0, 0, 11, Tf}, false);
  }
  if (!dry) {
    head_reduce_kernel<<<cdiv(B * (cfg.genre ? 35 : 24) * 32, 256), 256, 0, st>>>(tp.hk[1].z.p, tp.ht[1].z.p, cfg.genre ? tp.hg[1].z.p : nullptr, B, Tf,
                                                                               seq_len, cfg.time_pool_size, (k - 1) * cfg.head_layers, 0, key_out,
                                                                               tonic_out, genre_out);
    AKE_LAUNCHED();
  }
  tp.tc_ready = tc_ready, tp.tc_hi = tc_hi, tp.tc_lo = tc_lo, tp.tc_raw = tc_raw;
  tp.eq_ready = eq_ready, tp.eq_hi = eq_hi, tp.eq_lo = eq_lo;
  tp.tcw_block = tcw_block, tp.tcw_fwd = tcw_fwd, tp.tcw_bwd = tcw_bwd, tp.tcw_head = tcw_head;
  tp.ws_off = arena.off;
  tp.valid = !dry;
}

// ------------------------------------------------------------------------------------------------ backward
// One side stream per host thread and device for the weight-gradient kernels of the backward pass.
struct SideStream {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  bool used = false;
};
static SideStream& side_stream() {
  thread_local SideStream s;
  const int dev = current_device();
  if (s.device != dev) {
    s = SideStream();
    s.device = dev;
    AKE_CUDA(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    AKE_CUDA(cudaEventCreateWithFlags(&s.fork, cudaEventDisableTiming));
    AKE_CUDA(cudaEventCreateWithFlags(&s.join, cudaEventDisableTiming));
  }
  s.used = false;
  return s;
}

void Fwd::backward_keep(const TrainTape& tp, const float* d_key, const float* d_tonic, const float* d_genre, float* grads) {
  const ake_pcn_config& cfg = p->cfg;
  const int P = cfg.pitches, k = cfg.kernel_size;
  const int n_ss = p->n_ss;
  tc_ready = tp.tc_ready, tc_hi = tp.tc_hi, tc_lo = tp.tc_lo, tc_raw = tp.tc_raw;
  eq_ready = tp.eq_ready, eq_hi = tp.eq_hi, eq_lo = tp.eq_lo;
  tcw_block = tp.tcw_block, tcw_fwd = tp.tcw_fwd, tcw_bwd = tp.tcw_bwd, tcw_head = tp.tcw_head;
  float* ones = arena.take<float>(64);
  float* zeros = arena.take<float>(64);
  double* bsums = arena.take<double>(2 * (size_t)n_ss);
  unsigned* d_maxbits = arena.take<unsigned>(64);  // largest |dz| per tensor-core data gradient (its operand scale)
  int n_maxbits = 0;
  std::map<int, const float*> wd_of;  // conv id -> flipped / transposed weights of its FFMA data gradient
  if (!dry) {
    AKE_CUDA(cudaMemsetAsync(d_maxbits, 0, sizeof(unsigned) * 64, st));
    AKE_CUDA(cudaMemsetAsync(grads, 0, sizeof(float) * p->n_params, st));
    AKE_CUDA(cudaMemsetAsync(zeros, 0, sizeof(float) * 64, st));
    AKE_CUDA(cudaMemsetAsync(bsums, 0, sizeof(double) * 2 * n_ss, st));
    fill_kernel<<<1, 64, 0, st>>>(ones, 64, 1.f);
    AKE_LAUNCHED();
  }
  auto zalloc = [&](const View& like) { return alloc(like.C, like.R, like.T); };

  // BatchNorm + LeakyReLU backward of conv `id`: da (grad w.r.t. the activation) -> dz; writes dgamma / dbeta
  auto bn_bwd = [&](int id, const View& z, const View& da, unsigned* maxbits = nullptr) {
    const Conv& c = p->convs[id];
    const BnSite& bn = p->bns[c.bn];
    View dz = zalloc(z);
    if (dry) return dz;
    const int RT = z.R * z.T;
    const float* sc = tp.d_ss + c.ss_off;
    const float* sh = tp.d_ss + n_ss + c.ss_off;
    const float* mu = tp.d_mi + c.ss_off;
    const float* is = tp.d_mi + n_ss + c.ss_off;
    dim3 grid(std::max(1, std::min(64, (int)cdiv64((long long)B * RT, 1024))), c.Cout);
    bn_bwd_reduce_kernel<<<grid, 256, 0, st>>>(z.p, da.p, B, c.Cout, RT, sc, sh, mu, is, bsums + 2 * c.ss_off);
    AKE_LAUNCHED();
    bn_bwd_apply_kernel<<<ew_blocks(z.numel(B)), 256, 0, st>>>(z.p, da.p, dz.p, B, c.Cout, RT, sc, sh, mu, is, bsums + 2 * c.ss_off,
                                                              (double)B * RT, grads + bn.gamma, grads + bn.beta, maxbits);
    AKE_LAUNCHED();
    return dz;
  };
  auto bias_grad = [&](const Conv& c, const View& dz) {
    if (dry) return;
    dim3 grid(std::max(1, std::min(32, (int)cdiv64((long long)B * dz.R * dz.T, 1024))), c.Cout);
    channel_sum_kernel<<<grid, 256, 0, st>>>(dz.p, B, c.Cout, dz.R * dz.T, grads + c.b_off);
    AKE_LAUNCHED();
  };
  // Weight and bias gradients only feed the gradient buffer: they run on a side stream, forked behind the kernel that produced dz
  // and joined at the end of the backward pass, so that they overlap the data-gradient chain (at batch 8 neither fills the GPU).
  SideStream& side = side_stream();
  auto wgrad = [&](const ConvSite& s, const View& dz, const unsigned* maxbits = nullptr) {
    const Conv& c = p->convs[s.id];
    const bool tc = maxbits != nullptr && tc_wgrad_ok(c, s.g, s.in0.T);
    const bool eqw = maxbits != nullptr && !s.in1.p && eq_wgrad_ok(c, s.g, s.in0.T);
    if (dry) {
      if (tc) tc_wgrad(s.in0, s.in1.p ? &s.in1 : nullptr, c, dz, maxbits, nullptr, st);  // (carves the scratch)
      if (eqw) eq_wgrad(s.in0, c, s.g, dz, maxbits, nullptr, st);
      return;
    }
    const cudaStream_t main_st = st;
    static const bool use_side = [] { const char* e = getenv("AKE_WGRAD_SIDE"); return e ? atoi(e) != 0 : true; }();  // 5.40 -> 4.63 ms per step
    if (use_side) {
      AKE_CUDA(cudaEventRecord(side.fork, main_st));
      AKE_CUDA(cudaStreamWaitEvent(side.stream, side.fork, 0));
      st = side.stream;  // bias_grad and launch_wgrad below launch on `st`
      side.used = true;
    }
    struct Restore {
      cudaStream_t& s;
      cudaStream_t v;
      ~Restore() { s = v; }
    } restore{st, main_st};
    // a conv bias in front of a train-mode BatchNorm has an exactly-zero gradient (dz sums to zero per channel; the reference
    // leaves ~1e-15 of rounding there): the zeroed gradient buffer already holds it
    if (!s.has_bn) bias_grad(c, dz);
    if (tc) {  // 7x7 conv: one tensor-core GEMM over positions, deterministic reduction (pcn_train_tc.cuh)
      tc_wgrad(s.in0, s.in1.p ? &s.in1 : nullptr, c, dz, maxbits, grads + c.w_off, st);
      return;
    }
    if (eqw) {  // 16-channel equivariant conv: likewise (eq_wgrad_umma_kernel)
      eq_wgrad(s.in0, c, s.g, dz, maxbits, grads + c.w_off, st);
      return;
    }
    WgradArgs a{};
    a.in0 = s.in0.p, a.c0 = s.in0.C, a.rows0 = s.in0.R, a.bs0 = s.in0.bstride();
    if (s.in1.p) a.in1 = s.in1.p, a.c1 = s.in1.C, a.rows1 = s.in1.R, a.bs1 = s.in1.bstride();
    else a.in1 = s.in0.p, a.c1 = 0, a.rows1 = 1, a.bs1 = 0;
    a.T_in = s.in0.T;
    a.rows_v = s.g.rows_v, a.row_circ = s.g.row_circ, a.row_off = s.g.row_off, a.pad_t = s.g.pad_t, a.time_circ = s.g.time_circ;
    a.rows_out = s.g.rows_out, a.T_out = s.g.T_out, a.Cin = c.Cin, a.Cout = c.Cout;
    a.dz = dz.p, a.dw = grads + c.w_off;
    launch_wgrad(a, s.g, B, st);
  };
  // data gradient of a stride-1 convolution: the forward kernel on dz with flipped / transposed weights
  auto dgrad = [&](const ConvSite& s, const View& dz, View& dx, bool accumulate, const unsigned* maxbits = nullptr) {
    const Conv& c = p->convs[s.id];
    if (maxbits) {  // tensor-core data gradient (the caller checked tc_conv_ok / eq_conv_ok and had bn_bwd record the largest |dz|)
      if (accumulate || dx.C != c.Cin) fail(AKE_ERR_INVALID, "internal: tensor-core dgrad overwrites a (B, Cin, rows, T) tensor");
      if (s.g.KH == 12) eq_conv(dz, c, true, maxbits, ones, zeros, dx);
      else tc_conv(dz, nullptr, c, true, maxbits, dx, nullptr);
      return;
    }
    const int tile = co_tile_for(c.Cin), cin_pad = cdiv(c.Cin, tile) * tile;
    if (dry) return;
    if (dx.C != c.Cin) fail(AKE_ERR_INVALID, "internal: dgrad channel mismatch");
    const auto wit = wd_of.find(s.id);
    if (wit == wd_of.end()) fail(AKE_ERR_INVALID, "internal: no data-gradient weights were packed for conv %d", s.id);
    const float* wd = wit->second;
    ConvArgs a{};
    a.in0 = dz.p, a.c0 = dz.C, a.rows0 = dz.R, a.bs0 = dz.bstride();
    a.in1 = dz.p, a.c1 = 0, a.rows1 = 1, a.bs1 = 0;
    a.T_in = dz.T;
    a.rows_v = dz.R, a.row_circ = s.g.row_circ, a.row_off = -(s.g.row_off + c.KH - 1), a.pad_t = c.KW - 1 - s.g.pad_t, a.time_circ = s.g.time_circ;
    a.rows_out = dx.R, a.T_out = dx.T;
    a.Cin = c.Cout, a.Cout = c.Cin;
    a.w = wd, a.cout_pad = cin_pad;
    a.scale = ones, a.shift = zeros, a.act = 0;
    a.out = dx.p, a.obs = dx.bstride(), a.ocs = (long long)dx.R * dx.T, a.out_coff = 0;
    a.pool_t = 0, a.T_store = dx.T, a.accum = accumulate ? 1 : 0;
    ConvGeom gd = s.g;
    const int co_tile = cin_pad % 8 == 0 ? 8 : (cin_pad % 4 == 0 ? 4 : 1);
    launch_conv(a, gd, co_tile, B, st);
  };
  // a stack of conv + BN + LReLU sites, last to first: d_a is the gradient w.r.t. the last activation; returns d(input of site 0)
  auto stack_bwd = [&](const std::vector<ConvSite>& sites, View d_a, int first) {
    for (int i = (int)sites.size() - 1; i >= first; --i) {
      const ConvSite& s = sites[i];
      const Conv& c = p->convs[s.id];
      unsigned* mb = nullptr;
      if ((tc_conv_ok(c, s.g, s.in0.T) || (!s.in1.p && eq_conv_ok(c, s.g, s.in0.T))) && n_maxbits < 64) mb = d_maxbits + n_maxbits++;
      View dz = bn_bwd(s.id, s.z, d_a, mb);
      wgrad(s, dz, mb);
      View dx = alloc(c.Cin, s.in0.R, s.in0.T);
      dgrad(s, dz, dx, false, mb);
      d_a = dx;
    }
    return d_a;
  };

  // ---- flipped / transposed weights of every FFMA data-gradient convolution of this pass, packed by one launch
  {
    std::vector<const ConvSite*> sites = {&tp.ht[1], &tp.ht[0], &tp.hk[1], &tp.hk[0]};
    if (tp.genre) sites.push_back(&tp.hg[1]), sites.push_back(&tp.hg[0]);
    for (const ConvSite& s : tp.e1)
      if (!eq_conv_ok(p->convs[s.id], s.g, s.in0.T)) sites.push_back(&s);
    for (const ConvSite& s : tp.p)
      if (!tc_conv_ok(p->convs[s.id], s.g, s.in0.T)) sites.push_back(&s);
    for (const ConvSite& s : tp.e0)
      if (!eq_conv_ok(p->convs[s.id], s.g, s.in0.T)) sites.push_back(&s);
    std::vector<DgradPackEntry> ents;
    size_t total = 0;
    int max_n = 1;
    for (const ConvSite* s : sites) {
      const Conv& c = p->convs[s->id];
      const int tile = co_tile_for(c.Cin), cin_pad = cdiv(c.Cin, tile) * tile;
      const size_t n = (size_t)c.Cout * c.KH * c.KW * cin_pad;
      ents.push_back(DgradPackEntry{c.w_off, (long long)total, c.Cout, c.Cin, c.KH, c.KW, cin_pad});
      total += align_up(n, 64);
      max_n = std::max(max_n, (int)n);
    }
    float* block = arena.take<float>(total);
    for (size_t i = 0; i < sites.size(); ++i) wd_of[sites[i]->id] = block + ents[i].dst_off;
    for (size_t i0 = 0; !dry && i0 < ents.size(); i0 += kDgradPackMax) {
      DgradPackTable t{};
      t.n = (int)std::min<size_t>(kDgradPackMax, ents.size() - i0);
      for (int i = 0; i < t.n; ++i) t.e[i] = ents[i0 + i];
      pack_dgrad_all_kernel<<<dim3(std::min(32, cdiv(max_n, 256)), t.n), 256, 0, st>>>(t, p->d_params, block);
      AKE_LAUNCHED();
    }
  }
  // ---- masked means (+ sigmoid)
  const View& tf = tp.ht[1].z;
  View d_tf = zalloc(tf), d_kf = zalloc(tp.hk[1].z), d_gf;
  if (tp.genre) d_gf = zalloc(tp.hg[1].z);
  if (!dry) {
    const long long n = (long long)B * (tp.genre ? 35 : 24) * tf.T;
    head_reduce_bwd_kernel<<<ew_blocks(n), 256, 0, st>>>(d_key, d_tonic, tp.genre ? d_genre : nullptr, tp.key_out, B, tf.T, tp.seq_len,
                                                        cfg.time_pool_size, (k - 1) * cfg.head_layers, d_kf.p, d_tf.p,
                                                        tp.genre ? d_gf.p : nullptr);
    AKE_LAUNCHED();
  }
  // ---- heads
  View d_pcp = zalloc(tp.pcp);
  auto head_bwd = [&](const ConvSite (&h)[2], const View& d_frames, bool accumulate) {
    wgrad(h[1], d_frames);
    View d_a0 = zalloc(h[0].a);
    dgrad(h[1], d_frames, d_a0, false);
    const bool tc = tp.heads_tc && &h != &tp.hg && n_maxbits < 64 && eq_ready;
    unsigned* mb = tc ? d_maxbits + n_maxbits++ : nullptr;
    View dz0 = bn_bwd(h[0].id, h[0].z, d_a0, mb);
    wgrad(h[0], dz0, mb);
    if (tc) heads_dgrad_tc(h[0].id, dz0, mb, ones, zeros, d_pcp, accumulate);
    else dgrad(h[0], dz0, d_pcp, accumulate);
  };
  head_bwd(tp.ht, d_tf, false);
  head_bwd(tp.hk, d_kf, true);
  if (tp.genre) head_bwd(tp.hg, d_gf, true);
  // ---- time pool + PitchClass2PitchClass stack of layer 1
  View d_cat1;
  {
    const ConvSite& s = tp.e1.back();
    const Conv& c = p->convs[s.id];
    View da = zalloc(s.z);
    if (!dry) {
      timepool_bwd_kernel<<<ew_blocks(tp.pcp.numel(B)), 256, 0, st>>>(s.z.p, B, c.Cout, 12, s.z.T, tp.d_ss + c.ss_off, tp.d_ss + n_ss + c.ss_off,
                                                                     d_pcp.p, da.p);
      AKE_LAUNCHED();
    }
    d_cat1 = stack_bwd(tp.e1, da, 0);
  }
  // ---- concat: [pc0 | octave pool of pool_semi(p)]
  const View& pc0 = tp.e0.back().a;
  View d_pc0 = zalloc(pc0);
  View d_pf;
  {
    const ConvSite& s = tp.s1;
    const Conv& c = p->convs[s.id];
    View da = zalloc(s.z);
    if (!dry) {
      copy_channels_kernel<<<ew_blocks(pc0.numel(B)), 256, 0, st>>>(d_cat1.p, d_cat1.C, 0, d_pc0.p, d_pc0.C, 0, B, pc0.C, 12 * pc0.T, 0);
      AKE_LAUNCHED();
      octmax_bwd_kernel<<<ew_blocks((long long)B * c.Cout * 12 * s.z.T), 256, 0, st>>>(s.z.p, B, c.Cout, s.z.R, s.z.T, tp.d_ss + c.ss_off,
                                                                                     tp.d_ss + n_ss + c.ss_off, d_cat1.p, d_cat1.C, pc0.C, da.p);
      AKE_LAUNCHED();
    }
    View dz = bn_bwd(s.id, s.z, da);
    wgrad(s, dz);
    d_pf = alloc(c.Cin, P, s.in0.T);
    if (!dry) {
      semitone_dgrad_kernel<<<ew_blocks(d_pf.numel(B)), 256, 0, st>>>(dz.p, p->d_params + c.w_off, B, c.Cin, c.Cout, P, s.in0.T, d_pf.p);
      AKE_LAUNCHED();
    }
  }
  // ---- Pitch2Pitch stack; its first conv reads cat[mel, tile(up)]
  {
    View dX0 = stack_bwd(tp.p, d_pf, 0);  // (B, 1 + prev_pc, P, T); channel 0 is the gradient w.r.t. the log-CQT (unused)
    const Conv& cu = p->convs[p->layers[1].up];
    View d_up = zalloc(tp.up_a);
    if (!dry) {
      tile_sum_kernel<<<ew_blocks(d_up.numel(B)), 256, 0, st>>>(dX0.p, B, dX0.C, 1, d_up.C, P, dX0.T, d_up.p);
      AKE_LAUNCHED();
    }
    View dz_up = bn_bwd(p->layers[1].up, tp.up_z, d_up);  // (the ConvTranspose bias sits in front of a BatchNorm: zero gradient)
    if (!dry) {
      if (cu.Cin > 8) fail(AKE_ERR_UNSUPPORTED, "up_sixth backward: more than 8 channels");
      const long long n = (long long)B * 12 * pc0.T;
      upsixth_bwd_kernel<<<(unsigned)cdiv64(n, 128), 128, 0, st>>>(pc0.p, p->d_params + cu.w_off, dz_up.p, B, cu.Cin, pc0.T, d_pc0.p,
                                                                 grads + cu.w_off);
      AKE_LAUNCHED();
    }
  }
  // ---- layer 0
  {
    View d_q0 = stack_bwd(tp.e0, d_pc0, 0);
    const ConvSite& s = tp.s0;
    const Conv& c = p->convs[s.id];
    View da = zalloc(s.z);
    if (!dry) {
      octmax_bwd_kernel<<<ew_blocks((long long)B * c.Cout * 12 * s.z.T), 256, 0, st>>>(s.z.p, B, c.Cout, s.z.R, s.z.T, tp.d_ss + c.ss_off,
                                                                                     tp.d_ss + n_ss + c.ss_off, d_q0.p, d_q0.C, 0, da.p);
      AKE_LAUNCHED();
    }
    View dz = bn_bwd(s.id, s.z, da);
    wgrad(s, dz);
  }
  if (!dry && side.used) {  // join: the gradient buffer is complete only once the side stream's kernels are done
    AKE_CUDA(cudaEventRecord(side.join, side.stream));
    AKE_CUDA(cudaStreamWaitEvent(st, side.join, 0));
  }
}

}  // namespace ake
