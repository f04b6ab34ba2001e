// pcn.cu -- PitchClassNet plan + forward orchestration + C ABI (see include/ake_b200.h).
//
// Architecture facts restated from the reference (cited per item):
//   channel plan            models.py:266-308, 694-710
//   layer 0 / layer >= 1    models.py:359-369 / 370-396
//   heads                   models.py:713-742
//   masked mean + sigmoid   models.py:754-804
#include <cmath>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <atomic>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "pcn_kernels.cuh"
#include "pcn_train_kernels.cuh"
#include "pcn_umma.cuh"
#include "pcn_p2p1.cuh"
#include "pcn_train_tc.cuh"

namespace ake {

thread_local std::string g_last_error;
thread_local int64_t g_launches = 0;
void set_last_error(const std::string& msg) { g_last_error = msg; }
int64_t& launch_counter() { return g_launches; }

// ---- section profiler ---------------------------------------------------------------------------
namespace {
struct ProfEntry {
  std::string tag;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int launches = 0;
};
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;
std::vector<ProfEntry> g_prof_entries;
std::vector<cudaEvent_t> g_prof_pool;
thread_local std::vector<size_t> g_prof_open;  // indices of sections opened by this thread

cudaEvent_t prof_event() {
  if (!g_prof_pool.empty()) {
    cudaEvent_t e = g_prof_pool.back();
    g_prof_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
  return e;
}
}  // namespace

bool profile_enabled() { return g_prof_on.load(std::memory_order_relaxed); }

void profile_record(const char* tag, cudaStream_t st, bool begin, int n_launches) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (begin) {
    ProfEntry e;
    e.tag = tag;
    e.ev0 = prof_event(), e.ev1 = prof_event();
    if (e.ev0) cudaEventRecord(e.ev0, st);
    g_prof_entries.push_back(e);
    g_prof_open.push_back(g_prof_entries.size() - 1);
  } else if (!g_prof_open.empty()) {
    const size_t i = g_prof_open.back();
    g_prof_open.pop_back();
    if (i < g_prof_entries.size()) {
      if (g_prof_entries[i].ev1) cudaEventRecord(g_prof_entries[i].ev1, st);
      g_prof_entries[i].launches = n_launches;
    }
  }
}

struct TensorInfo {
  std::string name;
  int ndim;
  int64_t shape[4];
  int64_t off, numel;
};

struct BnSite {
  int C;
  int64_t gamma, beta, mean, var;  // offsets into the flat parameter buffer
  int stat_off;                    // channel offset into the concatenated BN-statistics output
};

struct Conv {
  int Cout, Cin, KH, KW;
  bool transposed = false;  // ConvTranspose2d weight layout (Cin, Cout, KH, KW)
  bool has_bias = true;     // dense layers: nn.Conv2d(..., bias=False) (models.py:463, 467)
  bool norm_only = false;   // a BatchNorm with no convolution in front (dense layers' norm1): KH = KW = 0, no weights
  int64_t w_off = 0, b_off = 0;
  int bn = -1;              // index into bn sites, -1: none
  int cout_pad = 0;
  int64_t packed_off = 0;   // into packed weights
  int ss_off = 0;           // into the scale/shift tables
};

struct TrainTape;

struct ResPair {  // ResBlock / ResBlockEquivariant (models.py:402-454): conv1 (C -> 2C) + BN + act, conv2 (2C -> C) + BN, + x, act
  int c1, c2;
};

struct DenseLayer {  // _DenseLayer / _DenseLayerEquivariant (models.py:456-580): norm1 + LeakyReLU + conv1 (1 wide) + norm2 + ReLU + conv2
  int norm1, conv1, conv2;
};

struct LayerPlan {
  int prev_p = 0, prev_pc = 0, out_p = 0, out_pc = 0;
  int sem = -1, up = -1;
  int pool_conv = -1;                  // opt.p2pc_conv: Pitch2PitchClassConv instead of the octave max pool
  std::vector<int> p2p, pc2pc;         // the plain conv stacks; with opt.resblock only their first conv ...
  std::vector<ResPair> p2p_res, pc2pc_res;  // ... followed by conv_layers residual blocks
  std::vector<DenseLayer> p2p_dense, pc2pc_dense;  // opt.denseblock: DenseBlock / DenseBlockEquivariant instead of the stacks
};

}  // namespace ake

using namespace ake;

struct ake_pcn {
  ake_pcn_config cfg;
  std::vector<TensorInfo> tensors;
  std::vector<BnSite> bns;
  std::vector<Conv> convs;
  std::vector<LayerPlan> layers;
  std::vector<int> tonic_head, key_head, genre_head;
  int64_t n_params = 0, n_packed = 0;
  int n_ss = 0, n_bn_ch = 0;
  // device state owned by the plan (small: weights only)
  float* d_params = nullptr;   // flat fp32 copy of the state_dict
  float* d_packed = nullptr;   // repacked conv weights
  float* d_ss_eval = nullptr;  // [scale | shift] eval-mode epilogues, n_ss each
  float* d_ss_raw = nullptr;   // [1 | bias] raw epilogues
  bool has_params = false;
  // tensor-core path (eval mode, default channel plan): fp16 hi/lo operand images of the 7x7 convolutions
  bool umma = false;
  __half* d_wimg = nullptr;         // one kP2PWBytes image per Pitch2Pitch conv, in conv-id order of `umma_convs`
  std::vector<int> umma_convs;
  __half* d_wimg_f1 = nullptr;      // first Pitch2Pitch conv split by input channel (pcn_p2p1.cuh): [mel image, 4 KB | periodic-part image, kP2PWBytes]
  __half* d_wimg_pc = nullptr;      // equivariant convs of the layer-1 PitchClass2PitchClass stack (kPcWBytes each)
  __half* d_wimg_l0 = nullptr;      // equivariant convs of the layer-0 PitchClass2PitchClass stack (kPc8WBytes each)
  __half* d_wimg_semi = nullptr;    // pool_semi conv of layer 1 (kSemiWBytes)
  __half* d_wimg_heads = nullptr;   // first conv of the tonic and key heads, fused along N (344,064 B)
  float* d_ss_heads = nullptr;      // [scale 64 | shift 64] of that fused conv (tonic channels first)
  __half* d_wimg_genre = nullptr;   // first conv of the genre head (1 x 7, 16 -> 32): one 16 KB stage
  __half* d_wimg_tail = nullptr;    // last conv of the tonic / key / genre heads (head_tail_umma_kernel), 12 KB each
  bool umma_heads = false;
  bool umma_dirty = false;  // parameters uploaded since the operand images were last built (built lazily by the next eval-mode forward)
  std::map<std::string, std::pair<const float*, int64_t>> taps;
  // activations kept by the bn_mode = 2 forwards that have not been back-propagated yet, keyed by their workspace
  // (pcn_train.cuh): several kept forwards may be outstanding (summed losses, siamese use), each with its own workspace
  std::map<const void*, struct ake::TrainTape*> tapes;
  int device = -1;  // device the weights / operand images live on (set by the first upload)
  std::vector<int64_t> bn_count;  // per BatchNorm site: elements per channel it normalised over in the last train-mode forward
};

namespace ake {

static int64_t add_tensor(ake_pcn* p, const std::string& name, std::initializer_list<int64_t> shape) {
  TensorInfo t;
  t.name = name;
  t.ndim = (int)shape.size();
  t.numel = 1;
  int i = 0;
  for (auto s : shape) t.shape[i++] = s, t.numel *= s;
  for (; i < 4; ++i) t.shape[i] = 1;
  t.off = p->n_params;
  p->n_params += t.numel;
  p->tensors.push_back(t);
  return t.off;
}

static int add_bn(ake_pcn* p, const std::string& prefix, int C) {
  BnSite b;
  b.C = C;
  b.gamma = add_tensor(p, prefix + ".weight", {C});
  b.beta = add_tensor(p, prefix + ".bias", {C});
  b.mean = add_tensor(p, prefix + ".running_mean", {C});
  b.var = add_tensor(p, prefix + ".running_var", {C});
  b.stat_off = p->n_bn_ch;
  p->n_bn_ch += C;
  p->bns.push_back(b);
  return (int)p->bns.size() - 1;
}

static int add_conv(ake_pcn* p, const std::string& wname, int Cout, int Cin, int KH, int KW, int co_tile,
                    bool transposed = false, bool has_bias = true) {
  Conv c;
  c.Cout = Cout, c.Cin = Cin, c.KH = KH, c.KW = KW, c.transposed = transposed, c.has_bias = has_bias;
  if (transposed)
    c.w_off = add_tensor(p, wname + ".weight", {Cin, Cout, KH, KW});
  else
    c.w_off = add_tensor(p, wname + ".weight", {Cout, Cin, KH, KW});
  if (has_bias) c.b_off = add_tensor(p, wname + ".bias", {Cout});
  c.cout_pad = cdiv(Cout, co_tile) * co_tile;
  c.packed_off = p->n_packed;
  if (!transposed) p->n_packed += (int64_t)Cin * KH * KW * c.cout_pad;
  c.ss_off = p->n_ss;
  p->n_ss += Cout;
  p->convs.push_back(c);
  return (int)p->convs.size() - 1;
}

// A BatchNorm site with no convolution in front of it (its scale / shift tables still live in a Conv slot).
static int add_norm(ake_pcn* p, const std::string& prefix, int C) {
  Conv c;
  c.Cout = C, c.Cin = C, c.KH = 0, c.KW = 0, c.norm_only = true, c.has_bias = false;
  c.ss_off = p->n_ss;
  p->n_ss += C;
  c.bn = add_bn(p, prefix, C);
  p->convs.push_back(c);
  return (int)p->convs.size() - 1;
}

static int co_tile_for(int Cout) { return Cout >= 8 ? 8 : (Cout >= 4 ? 4 : 1); }

static void build_plan(ake_pcn* p) {
  const ake_pcn_config& c = p->cfg;
  if (c.denseblock && (c.resblock || c.pc2p_mem || c.stay_sixth))
    fail(AKE_ERR_UNSUPPORTED, "denseblock together with resblock / pc2p_mem / stay_sixth is not built (the reference's own channel plan "
         "does not cover those combinations, models.py:266-283, 331-335)");
  if (c.only_semitones) fail(AKE_ERR_UNSUPPORTED, "only_semitones is not built (the reference itself cannot run it: models.py:366-367)");
  if (c.stay_sixth && c.pc2p_mem) fail(AKE_ERR_UNSUPPORTED, "stay_sixth together with pc2p_mem is not built");
  if (c.p2pc_conv && (c.pitches / 3) % 12) fail(AKE_ERR_UNSUPPORTED, "p2pc_conv needs pitches / 3 to be a multiple of 12 (the reference pads with -inf otherwise)");
  const bool variant = c.resblock || c.stay_sixth || c.p2pc_conv || c.pc2p_mem || c.local || c.denseblock;
  if (c.local && (c.frames <= 0 || c.loc_window_size <= 0)) fail(AKE_ERR_INVALID, "opt.local needs opt.frames > 0 and opt.loc_window_size > 0");
  if (c.pitch_classes != 12) fail(AKE_ERR_INVALID, "pitch_classes must be 12 (models.py:171)");
  if (c.pitches <= 0 || c.pitches % 36) fail(AKE_ERR_INVALID, "pitches must be a positive multiple of 36");
  if (c.kernel_size != 7) fail(AKE_ERR_UNSUPPORTED, "kernel_size %d: only 7 is built", c.kernel_size);
  if (c.time_pool_size != 2) fail(AKE_ERR_UNSUPPORTED, "time_pool_size %d: only 2 is built", c.time_pool_size);
  if (c.num_layers < 1 || c.num_layers > 3) fail(AKE_ERR_UNSUPPORTED, "num_layers %d: 1..3 are built", c.num_layers);
  if (c.conv_layers < 1 || c.n_filters < 1 || c.head_layers < 1) fail(AKE_ERR_INVALID, "bad layer counts");
  const int k = c.kernel_size, nf = c.n_filters;

  // Pitch2Pitch / PitchClass2PitchClass stack (models.py:168-243): conv_layers x (conv + BN + LeakyReLU), or with opt.resblock
  // conv + BN + LeakyReLU followed by conv_layers residual blocks ("<stack>.layer.{3 + j}.conv1 / b1 / conv2 / b2")
  auto build_stack = [&](const std::string& stack, bool equivariant, int Cin, int Cout, std::vector<int>& ids, std::vector<ResPair>& res) {
    const int KH = equivariant ? 12 : k;
    const std::string leaf = equivariant ? ".conv2d" : "";
    for (int i = 0; i < (c.resblock ? 1 : c.conv_layers); ++i) {
      int id = add_conv(p, stack + ".layer." + std::to_string(3 * i) + leaf, Cout, i == 0 ? Cin : Cout, KH, k, co_tile_for(Cout));
      p->convs[id].bn = add_bn(p, stack + ".layer." + std::to_string(3 * i + 1), Cout);
      ids.push_back(id);
    }
    if (c.resblock)
      for (int j = 0; j < c.conv_layers; ++j) {
        const std::string rb = stack + ".layer." + std::to_string(3 + j);
        ResPair rp;
        rp.c1 = add_conv(p, rb + ".conv1" + leaf, 2 * Cout, Cout, KH, k, co_tile_for(2 * Cout));
        p->convs[rp.c1].bn = add_bn(p, rb + ".b1", 2 * Cout);
        rp.c2 = add_conv(p, rb + ".conv2" + leaf, Cout, 2 * Cout, KH, k, co_tile_for(Cout));
        p->convs[rp.c2].bn = add_bn(p, rb + ".b2", Cout);
        res.push_back(rp);
      }
  };
  // DenseBlock / DenseBlockEquivariant(num_layers = conv_layers, num_input_features = Cin, bn_size = Cin // 2 (1 for Cin = 1),
  // growth_rate = n_filters, multi_path off): "<stack>.layer.0.denselayer{i+1}.norm1 / conv1 / norm2 / conv2" (models.py:186, 224, 582-648)
  auto build_dense = [&](const std::string& stack, bool equivariant, int Cin, std::vector<DenseLayer>& dl) {
    const int bn_size = Cin > 1 ? Cin / 2 : 1, g = nf;
    const std::string leaf = equivariant ? ".conv2d" : "";
    for (int i = 0; i < c.conv_layers; ++i) {
      const std::string d = stack + ".layer.0.denselayer" + std::to_string(i + 1);
      const int Cn = Cin + i * g;
      DenseLayer L;
      L.norm1 = add_norm(p, d + ".norm1", Cn);
      L.conv1 = add_conv(p, d + ".conv1" + leaf, bn_size * g, Cn, equivariant ? 12 : 1, 1, co_tile_for(bn_size * g), false, equivariant);
      p->convs[L.conv1].bn = add_bn(p, d + ".norm2", bn_size * g);
      L.conv2 = add_conv(p, d + ".conv2" + leaf, g, bn_size * g, equivariant ? 12 : k, k, co_tile_for(g), false, equivariant);
      dl.push_back(L);
    }
    return Cin + c.conv_layers * g;
  };
  auto build_pool = [&](const std::string& pre, int C, LayerPlan& lp) {
    if (!c.p2pc_conv) return;
    const int KS = cdiv(c.pitches / 3, 12);  // Pitch2PitchClassConv.kernel_size (models.py:115)
    lp.pool_conv = add_conv(p, pre + "pool.conv", C, C, KS, 1, 1);
    p->convs[lp.pool_conv].bn = add_bn(p, pre + "pool.bn", C);
  };
  for (int L = 0; L < c.num_layers; ++L) {
    LayerPlan lp;
    const std::string pre = "model." + std::to_string(L) + ".";
    if (L == 0) {
      lp.out_p = 1, lp.out_pc = nf;  // models.py:298-300 (pc2pc built with num_filters, :320)
      lp.sem = add_conv(p, pre + "pool_semi", 1, 1, 3, 3, 1);
      p->convs[lp.sem].bn = add_bn(p, pre + "pool_semi_b", 1);
      build_pool(pre, 1, lp);
      if (c.denseblock) lp.out_pc = build_dense(pre + "pc2pc", true, 1, lp.pc2pc_dense);
      else build_stack(pre + "pc2pc", true, 1, nf, lp.pc2pc, lp.pc2pc_res);
    } else if (c.denseblock) {
      // models.py:266-283: every layer passes all of its input features on (DenseNet concatenation)
      lp.prev_p = p->layers[L - 1].out_p, lp.prev_pc = p->layers[L - 1].out_pc;
      lp.up = add_conv(p, pre + "up_sixth", lp.prev_pc, lp.prev_pc, 3, 1, 1, /*transposed=*/true);
      p->convs[lp.up].bn = add_bn(p, pre + "up_sixth_b", lp.prev_pc);
      lp.out_p = build_dense(pre + "p2p", false, lp.prev_p + lp.prev_pc, lp.p2p_dense);
      lp.sem = add_conv(p, pre + "pool_semi", lp.out_p, lp.out_p, 3, 3, 8);
      p->convs[lp.sem].bn = add_bn(p, pre + "pool_semi_b", lp.out_p);
      build_pool(pre, lp.out_p, lp);
      lp.out_pc = build_dense(pre + "pc2pc", true, lp.out_p + lp.prev_pc, lp.pc2pc_dense);
    } else {
      // models.py:285-308
      if (L == 1) lp.prev_p = 1, lp.prev_pc = nf, lp.out_p = 2 * nf, lp.out_pc = 2 * lp.out_p;
      else {
        lp.prev_p = (2 * nf) * (int)std::pow(4.0, L - 2);
        lp.prev_pc = 2 * lp.prev_p;
        lp.out_p = 4 * lp.prev_p, lp.out_pc = 4 * lp.prev_pc;
      }
      if (!c.stay_sixth) {  // opt.stay_sixth: the pitch-class rows are tiled straight over the semitones (models.py:322-323)
        lp.up = add_conv(p, pre + "up_sixth", lp.prev_pc, lp.prev_pc, 3, 1, 1, /*transposed=*/true);
        p->convs[lp.up].bn = add_bn(p, pre + "up_sixth_b", lp.prev_pc);
      }
      // opt.pc2p_mem: the up-sampled features are added to p instead of concatenated (models.py:335)
      build_stack(pre + "p2p", false, c.pc2p_mem ? lp.prev_p : lp.prev_p + lp.prev_pc, lp.out_p, lp.p2p, lp.p2p_res);
      if (!c.stay_sixth) {
        lp.sem = add_conv(p, pre + "pool_semi", lp.out_p, lp.out_p, 3, 3, 8);
        p->convs[lp.sem].bn = add_bn(p, pre + "pool_semi_b", lp.out_p);
      }
      build_pool(pre, lp.out_p, lp);
      build_stack(pre + "pc2pc", true, lp.out_p + lp.prev_pc, lp.out_pc, lp.pc2pc, lp.pc2pc_res);
    }
    p->layers.push_back(lp);
  }
  // heads (models.py:713-742); registration order tonic, key, genre (models.py:739-742)
  const int final_ch = p->layers.back().out_pc;
  auto build_head = [&](const std::string& name, bool equivariant, std::vector<int>& ids) {
    int fc = final_ch;
    for (int i = 0; i < c.head_layers; ++i) {
      const std::string s = name + "." + std::to_string(3 * i);
      if (i == c.head_layers - 1) {
        ids.push_back(add_conv(p, equivariant ? s + ".conv2d" : s, 1, fc, equivariant ? 12 : 2, k, 1));
      } else {
        const int oc = i == 0 ? 2 * fc : fc;
        int id = add_conv(p, equivariant ? s + ".conv2d" : s, oc, fc, equivariant ? 12 : 1, k, 8);
        p->convs[id].bn = add_bn(p, name + "." + std::to_string(3 * i + 1), oc);
        ids.push_back(id);
        fc = oc;
      }
    }
  };
  {
    // Tensor-core path: the Pitch2Pitch stack of layer 1 at the train_model.py channel plan (1 + 4 -> 8 -> 8 channels).
    const char* off = getenv("AKE_DISABLE_UMMA");
    p->umma = !(off && off[0] == '1') && c.num_layers == 2 && nf == 4 && k == 7 && !variant;
    if (p->umma) p->umma_convs = p->layers[1].p2p;
  }
  build_head("tonic_classifier", true, p->tonic_head);
  build_head("key_classifier", true, p->key_head);
  p->umma_heads = p->umma && c.head_layers == 2;  // 16 -> 32 (BN, act) -> 1: the first conv runs on tensor cores
  if (c.genre) build_head("genre_classifier", false, p->genre_head);
}

// ------------------------------------------------------------------------------ kernel launch helpers
struct View {  // (B, C, R, T) contiguous fp32
  float* p = nullptr;
  int C = 0, R = 0, T = 0;
  long long bstride() const { return (long long)C * R * T; }
  long long numel(int B) const { return bstride() * B; }
};

struct ConvGeom {
  int KH, KW, SR;
  int rows_v, row_circ, row_off, pad_t, time_circ;
  int rows_out, T_out;
  int pool_t = 0;
};

template <int KH, int KW, int SR, int RB, int CO_T, int RT>
static void launch_conv_t(ConvArgs a, int B, int max_tg, cudaStream_t st) {
  constexpr int RIN = (RB - 1) * SR + KH;
  const int groups = cdiv(a.T_out, RT);
  // (Narrower time tiles for small batches were measured and dropped: 5.40 -> 7.92 ms for the 8-clip training step -- every block
  // re-stages its weight slice and input halo.)
  const int n_tiles = cdiv(groups, max_tg);
  a.tgroups = cdiv(groups, n_tiles);
  const int TBW = a.tgroups * RT + KW - 1;
  if (TBW > 96) fail(AKE_ERR_INVALID, "internal: time tile too wide");
  int xp = (TBW + 3) / 4 * 4;
  while (xp % 32 != 12 && xp % 32 != 4 && xp % 32 != 20 && xp % 32 != 28) xp += 4;  // odd multiple of 4: conflict-free float4 rows
  a.xp = xp;
  a.n_row_tiles = cdiv(a.rows_out, RB);
  int threads = cdiv(RB * a.tgroups, 32) * 32;
  const size_t smem = sizeof(float) * ((size_t)kConvCI * RIN * xp + (size_t)kConvCI * KH * KW * CO_T);
  auto kern = conv_rows_kernel<KH, KW, SR, RB, CO_T, RT>;
  ensure_dyn_smem(kern, smem);
  dim3 grid(n_tiles, a.n_row_tiles * (a.cout_pad / CO_T), B);
  // few blocks (the 8-clip training step): a block is alone on its SM and its staging phase is latency-bound with 3-4 warps;
  // extra warps only stage (the kernel's `active` guard keeps them out of the arithmetic)
  if ((long long)grid.x * grid.y * grid.z <= 4LL * sm_count()) threads = 256;
  kern<<<grid, threads, smem, st>>>(a);
  AKE_LAUNCHED();
}

static void launch_conv(const ConvArgs& a, const ConvGeom& g, int co_tile, int B, cudaStream_t st) {
  // (KH, KW, SR) families: equivariant 12x7, pitch 7x7, semitone 3x3/3, genre 1x7 and 2x7
  if (g.KH == 12 && g.KW == 7 && g.SR == 1) {
    // few blocks (small batch: the 8-clip training step): four output channels per thread instead of eight doubles the blocks of
    // the equivariant convs (5.40 -> 5.21 ms per step); the accumulation order of an output does not depend on the tiling
    if (co_tile == 8 && (long long)cdiv(cdiv(a.T_out, 4), 8) * (a.cout_pad / 8) * B < 2LL * sm_count()) co_tile = 4;
    // ... and two per thread once more when even that leaves the GPU with fewer than two blocks per SM (a thread's serial FMA chain
    // -- Cin x 84 taps x co_tile x 4 frames -- is what the launch waits for; one channel per thread was measured slower: 2.60 -> 2.81 ms)
    static const int min_tile = [] { const char* e = getenv("AKE_EQUIV_MIN_TILE"); return e ? atoi(e) : 2; }();
    if (co_tile == 4 && min_tile <= 2 && (long long)cdiv(cdiv(a.T_out, 4), 8) * (a.cout_pad / 4) * B < 2LL * sm_count()) co_tile = 2;
    if (co_tile == 8) return launch_conv_t<12, 7, 1, 12, 8, 4>(a, B, 8, st);
    if (co_tile == 4) return launch_conv_t<12, 7, 1, 12, 4, 4>(a, B, 8, st);
    if (co_tile == 2) return launch_conv_t<12, 7, 1, 12, 2, 4>(a, B, 8, st);
    if (co_tile == 1) return launch_conv_t<12, 7, 1, 12, 1, 4>(a, B, 8, st);
  } else if (g.KH == 7 && g.KW == 7 && g.SR == 1) {
    if (co_tile == 8) return launch_conv_t<7, 7, 1, 32, 8, 8>(a, B, 4, st);
    if (co_tile == 4) return launch_conv_t<7, 7, 1, 32, 4, 8>(a, B, 4, st);
    if (co_tile == 1) return launch_conv_t<7, 7, 1, 32, 1, 8>(a, B, 4, st);
  } else if (g.KH == 3 && g.KW == 3 && g.SR == 3) {
    if (co_tile == 8) return launch_conv_t<3, 3, 3, 32, 8, 8>(a, B, 4, st);
    if (co_tile == 1) return launch_conv_t<3, 3, 3, 32, 1, 8>(a, B, 4, st);
  } else if (g.KH == 1 && g.KW == 7 && g.SR == 1) {
    if (co_tile == 8) return launch_conv_t<1, 7, 1, 12, 8, 4>(a, B, 8, st);
    if (co_tile == 4) return launch_conv_t<1, 7, 1, 12, 4, 4>(a, B, 8, st);
    if (co_tile == 1) return launch_conv_t<1, 7, 1, 12, 1, 4>(a, B, 8, st);
  } else if (g.KH == 1 && g.KW == 1 && g.SR == 1) {  // dense layers: 1 x 1 bottleneck on the pitch rows
    if (co_tile == 8) return launch_conv_t<1, 1, 1, 32, 8, 8>(a, B, 4, st);
    if (co_tile == 4) return launch_conv_t<1, 1, 1, 32, 4, 8>(a, B, 4, st);
    if (co_tile == 1) return launch_conv_t<1, 1, 1, 32, 1, 8>(a, B, 4, st);
  } else if (g.KH == 12 && g.KW == 1 && g.SR == 1) {  // dense layers: equivariant bottleneck (kernel_depth 1)
    if (co_tile == 8) return launch_conv_t<12, 1, 1, 12, 8, 4>(a, B, 8, st);
    if (co_tile == 4) return launch_conv_t<12, 1, 1, 12, 4, 4>(a, B, 8, st);
    if (co_tile == 1) return launch_conv_t<12, 1, 1, 12, 1, 4>(a, B, 8, st);
  } else if (g.KH == 2 && g.KW == 7 && g.SR == 1) {
    if (co_tile == 1) return launch_conv_t<2, 7, 1, 12, 1, 4>(a, B, 8, st);
    if (co_tile == 8) return launch_conv_t<2, 7, 1, 12, 8, 4>(a, B, 8, st);
  }
  fail(AKE_ERR_UNSUPPORTED, "no conv kernel for KH=%d KW=%d SR=%d co_tile=%d", g.KH, g.KW, g.SR, co_tile);
}

// The persistent kernels decode work-item indices with multiply-high (x / d == umulhi(x, 2^32 / d + 1)), which is exact while
// x * d < 2^32: refuse launches beyond that instead of decoding wrongly.
static void check_decode_range(long long n_items, long long divisor, const char* what) {
  if (n_items * divisor >= (1LL << 32))
    fail(AKE_ERR_UNSUPPORTED, "%s: %lld work items x %lld exceed the index decode range; split the batch", what, n_items, divisor);
}

static int ew_blocks(long long n) { return (int)std::min<long long>(cdiv64(n, 256), 148LL * 16); }

struct Fwd {
  ake_pcn* p;
  int B, T;
  bool train, dry;
  cudaStream_t st;
  Arena arena;
  const int* seq_len;
  float* bn_stats_out;
  __half* umma_pc_hi = nullptr;  // tensor-core path: final pitch-class features as chunk planes [B][2][23][T/2][8]
  __half* umma_pc_lo = nullptr;
  bool umma_pc_ready = false;
  bool l0_fast = false;
  __half* l0_planes[3][2] = {};
  int os_key = 12, os_tonic = 12, os_genre = 11;  // floats between consecutive clips' outputs (35: (B, 35) result rows)
  double* d_stats = nullptr;  // train: per conv channel (sum, sumsq)
  float* d_ss_train = nullptr;
  float* d_mi_train = nullptr;  // kept forward: [mean | invstd] per conv channel for the BatchNorm backward
  // kept forward / backward: 7x7 convolutions on the tensor cores (pcn_train_tc.cuh); scratch shared by every site of the pass
  bool tc_ready = false;
  __half* tc_hi = nullptr;
  __half* tc_lo = nullptr;
  float* tc_raw = nullptr;
  __half* tcw_block = nullptr;                 // every operand image of the step (tc_pack_all_weights_kernel)
  std::map<int, long long> tcw_fwd, tcw_bwd;  // conv id -> offset (halves) of its forward / data-gradient image in the block
  std::map<int, long long> tcw_head;          // 2 * conv id + part -> data-gradient image of gradient channels [16 part, 16 part + 16) of a head's first conv
  // ... and the equivariant 12 x 7 convolutions of the PitchClass2PitchClass stacks (pc2pc_umma_kernel<3>, pc8_umma_kernel<2> raw)
  bool eq_ready = false;
  __half* eq_hi = nullptr;
  __half* eq_lo = nullptr;
  // backward: scratch of the tensor-core weight gradient (side stream: its own planes)
  bool wg_ready = false;
  __half* wg_x[2] = {};
  __half* wg_g[2] = {};
  float* wg_partial = nullptr;

  Fwd(ake_pcn* p_, int B_, int T_, bool train_, void* ws, size_t ws_bytes, cudaStream_t st_)
      : p(p_), B(B_), T(T_), train(train_), dry(ws == nullptr), st(st_), arena(ws, ws_bytes) {}

  View alloc(int C, int R, int Tn) {
    View v;
    v.C = C, v.R = R, v.T = Tn;
    v.p = arena.take<float>((size_t)B * C * R * Tn);
    return v;
  }
  void tap(const std::string& name, const View& v, int C = -1) {
    if (dry) return;
    // ping-pong buffers are reused: a newer tap on the same storage invalidates the older one
    for (auto it = p->taps.begin(); it != p->taps.end();)
      it = (it->second.first == v.p) ? p->taps.erase(it) : std::next(it);
    p->taps[name] = {v.p, (int64_t)B * (C < 0 ? v.C : C) * v.R * v.T};
  }
  const float* scale_of(const Conv& c, bool raw) const { return (raw ? p->d_ss_raw : p->d_ss_eval) + c.ss_off; }
  const float* shift_of(const Conv& c, bool raw) const {
    return (raw ? p->d_ss_raw : p->d_ss_eval) + p->n_ss + c.ss_off;
  }

  // Batch statistics of channels [coff, coff+C) of `v`, then scale/shift for the train-mode epilogue.
  void train_bn(const Conv& c, const View& v, int coff) {
    if (dry) return;
    const int rt = v.R * v.T;
    dim3 grid(std::max(1, std::min(64, (int)cdiv64((long long)B * rt, 1024))), c.Cout);
    bn_stats_kernel<<<grid, 256, 0, st>>>(v.p, B, v.C, coff, rt, d_stats + 2 * c.ss_off);
    AKE_LAUNCHED();
    train_bn_finalize(c, rt);
  }
  // ... the second half: (sum, sumsq) of conv `c` over B * rt elements per channel -> scale / shift (+ mean / invstd, running-stat inputs)
  void train_bn_finalize(const Conv& c, int rt) {
    const BnSite& bn = p->bns[c.bn];
    double* stats = d_stats + 2 * c.ss_off;
    if (dry) return;
    if (p->bn_count.size() != p->bns.size()) p->bn_count.assign(p->bns.size(), 0);
    p->bn_count[c.bn] = (int64_t)B * rt;
    bn_finalize_kernel<<<cdiv(c.Cout, 64), 64, 0, st>>>(stats, (double)B * rt, p->d_params + bn.gamma,
                                                         p->d_params + bn.beta, c.Cout, d_ss_train + c.ss_off,
                                                         d_ss_train + p->n_ss + c.ss_off,
                                                         bn_stats_out ? bn_stats_out + 2 * bn.stat_off : nullptr,
                                                         d_mi_train ? d_mi_train + c.ss_off : nullptr,
                                                         d_mi_train ? d_mi_train + p->n_ss + c.ss_off : nullptr);
    AKE_LAUNCHED();
  }

  // ---- train mode: operand images of every tensor-core convolution of the step, packed by ONE launch at the start of the kept forward
  // (the parameters changed since the last step); `sites`: (conv id, geometry, input frames) of the candidate convolutions
  struct TcSite {
    int id;
    ConvGeom g;
    int Tn;
    bool two_inputs;
  };
  void tc_pack_images(const std::vector<TcSite>& sites, const std::vector<int>& head_first = {}) {
    std::vector<TcWeightEntry> ents;
    long long total = 0;
    for (const TcSite& s : sites) {
      const Conv& c = p->convs[s.id];
      int kind = -1;
      if (tc_conv_ok(c, s.g, s.Tn)) kind = 0;
      else if (!s.two_inputs && eq_conv_ok(c, s.g, s.Tn)) kind = (c.Cin > 8 || c.Cout > 8) ? 2 : 1;
      if (kind < 0) continue;
      const long long halves = (kind == 0 ? kP2PWBytes : (kind == 1 ? kPc8WBytes : kPcWBytes)) / 2;
      for (int flip = 0; flip < 2; ++flip) {
        ents.push_back(TcWeightEntry{c.w_off, total, c.Cout, c.Cin, kind, flip, 0});
        (flip ? tcw_bwd : tcw_fwd)[s.id] = total;
        total += (halves + 127) / 128 * 128;
      }
    }
    // data gradient of the heads' first convs (32 gradient channels each -> 16): two 16-channel images per head
    for (int id : head_first) {
      const Conv& c = p->convs[id];
      for (int part = 0; part < 2; ++part) {
        ents.push_back(TcWeightEntry{c.w_off, total, c.Cout, c.Cin, 2, 1, 16 * part});
        tcw_head[id * 2 + part] = total;
        total += (kPcWBytes / 2 + 127) / 128 * 128;
      }
    }
    tcw_block = arena.take<__half>((size_t)std::max<long long>(total, 128));
    for (size_t i0 = 0; !dry && i0 < ents.size(); i0 += kTcWeightMax) {
      TcWeightTable t{};
      t.n = (int)std::min<size_t>(kTcWeightMax, ents.size() - i0);
      for (int i = 0; i < t.n; ++i) t.e[i] = ents[i0 + i];
      tc_pack_all_weights_kernel<<<dim3(21, t.n), 256, 0, st>>>(t, p->d_params, tcw_block);
      AKE_LAUNCHED();
    }
  }
  const __half* tc_image(const Conv& c, bool dgrad) const {
    const int id = (int)(&c - p->convs.data());
    const auto& m = dgrad ? tcw_bwd : tcw_fwd;
    const auto it = m.find(id);
    if (it == m.end()) fail(AKE_ERR_INVALID, "internal: no tensor-core operand image was packed for conv %d", id);
    return tcw_block ? tcw_block + it->second : nullptr;
  }

  // ---- train mode: a 7x7 circular convolution (or its data gradient) on the tensor cores (pcn_train_tc.cuh)
  bool tc_conv_ok(const Conv& c, const ConvGeom& g, int Tn) const {
    static const bool on = [] { const char* e = getenv("AKE_TRAIN_TC"); return e ? atoi(e) != 0 : true; }();
    return on && p->umma && g.KH == 7 && g.KW == 7 && g.SR == 1 && g.row_circ && g.time_circ && g.row_off == -3 && g.pad_t == 3 &&
           g.rows_v == g.rows_out && c.Cin <= 8 && c.Cout <= 8 && Tn >= 7;
  }
  // out = conv(cat[in0, tile(in1)], W) + bias   (dgrad = false; `stats` += the BatchNorm sums of the result), or
  // out = the data gradient of that conv for the output gradient in0 (dgrad = true; maxbits: largest |in0| as float bits)
  void tc_conv(const View& in0, const View* in1, const Conv& c, bool dgrad, const unsigned* maxbits, View& out, double* stats) {
    const int P = in0.R, Tn = in0.T, Wd = Tn + 6;
    if (!tc_ready) {
      const size_t halves = (size_t)B * (P + 6) * Wd * 8;
      tc_hi = arena.take<__half>(halves), tc_lo = arena.take<__half>(halves);
      tc_raw = arena.take<float>((size_t)B * P * Tn * 8);
      tc_ready = true;
    }
    if (dry) return;
    ProfScope prof("pcn.p2p", st);
    const __half* wimg = tc_image(c, dgrad);
    TcPackArgs pa{};
    pa.in0 = in0.p, pa.bs0 = in0.bstride(), pa.c0 = in0.C;
    pa.in1 = in1 ? in1->p : in0.p, pa.bs1 = in1 ? in1->bstride() : 0, pa.c1 = in1 ? in1->C : 0, pa.rows1 = in1 ? in1->R : 1;
    pa.B = B, pa.P = P, pa.T = Tn, pa.Wd = Wd, pa.col_shift = 3, pa.wrap_cols = 1, pa.maxbits = maxbits, pa.hi = tc_hi, pa.lo = tc_lo;
    tc_pack_planes_kernel<<<ew_blocks((long long)B * (P + 6) * Wd), 256, 0, st>>>(pa);
    AKE_LAUNCHED();
    const int n_tt = cdiv(Tn, kP2PMaxTB), TB = cdiv(Tn, n_tt), n_rt = cdiv(P, kP2PRows), n_tiles = B * n_rt * n_tt;
    const size_t smem = p2p_smem_bytes(TB + 6);
    ensure_dyn_smem(p2p_umma_kernel<false, true>, smem);
    check_decode_range((long long)n_tiles, (long long)n_rt * n_tt, "Pitch2Pitch (training)");
    P2PArgs a{tc_hi, tc_lo, nullptr, nullptr, wimg, p->d_ss_raw, p->d_ss_raw, P, Tn, Wd, TB, n_tt, n_rt, n_tiles, nullptr, nullptr, tc_raw};
    p2p_umma_kernel<false, true><<<std::min(n_tiles, sm_count()), kP2PThreads, smem, st>>>(a);
    AKE_LAUNCHED();
    TcUnpackArgs ua{};
    ua.raw = tc_raw, ua.out = out.p, ua.maxbits = maxbits, ua.B = B, ua.P = P, ua.T = Tn, ua.stats = stats;
    ua.C = dgrad ? c.Cin : c.Cout;
    ua.bias = (!dgrad && c.has_bias) ? p->d_params + c.b_off : nullptr;
    tc_unpack_kernel<<<std::min<int>((int)cdiv64((long long)B * P * Tn, 256), 4 * sm_count()), 256, 0, st>>>(ua);
    AKE_LAUNCHED();
  }

  // ---- train mode: an equivariant 12 x 7 convolution ("same" zero padding in time) or its data gradient on the tensor cores
  bool eq_conv_ok(const Conv& c, const ConvGeom& g, int Tn) const {
    static const bool on = [] { const char* e = getenv("AKE_TRAIN_TC_EQUIV"); return e ? atoi(e) != 0 : true; }();
    return on && p->umma && g.KH == 12 && g.KW == 7 && g.SR == 1 && g.row_circ && !g.time_circ && g.row_off == 0 && g.pad_t == 3 &&
           g.rows_v == 12 && g.rows_out == 12 && g.T_out == Tn && c.Cin <= 16 && c.Cout <= 16 && Tn >= 7;
  }
  // out = conv(in, W) + bias (dgrad = false), or the data gradient for the output gradient `in` (dgrad = true; maxbits: largest |in|);
  // ones / zeros: 16 floats each on the device (the data gradient's epilogue)
  void eq_conv(const View& in, const Conv& c, bool dgrad, const unsigned* maxbits, const float* ones, const float* zeros, View& out) {
    const int Tn = in.T, Wd = Tn + 6;
    if (!eq_ready) {
      const size_t halves = (size_t)B * 2 * 23 * Wd * 8;
      eq_hi = arena.take<__half>(halves), eq_lo = arena.take<__half>(halves);
      eq_ready = true;
    }
    if (dry) return;
    ProfScope prof("pcn.equiv", st);
    const int Cin = dgrad ? c.Cout : c.Cin, Cout = dgrad ? c.Cin : c.Cout;  // of the convolution that runs
    const bool wide = Cin > 8 || Cout > 8;
    const __half* wimg = tc_image(c, dgrad);
    EqPackArgs pa{in.p, B, in.C, wide ? 2 : 1, Tn, Wd, 3, 0, 0, maxbits, eq_hi, eq_lo};
    eq_pack_planes_kernel<<<ew_blocks((long long)B * pa.G * 23 * Wd), 256, 0, st>>>(pa);
    AKE_LAUNCHED();
    const float* scale = dgrad ? ones : scale_of(c, true);
    const float* shift = dgrad ? zeros : shift_of(c, true);
    if (wide) {
      const int n_tt = cdiv(Tn, 32), TBe = (cdiv(Tn, n_tt) + 1) / 2 * 2;
      const size_t smem_e = pc2pc_smem_bytes(TBe + 6);
      ensure_dyn_smem(pc2pc_umma_kernel<3>, smem_e);
      Pc2PcArgs ea{};
      ea.in_hi = eq_hi, ea.in_lo = eq_lo, ea.Wd_in = Wd, ea.T_out = Tn, ea.TB = TBe, ea.n_ttiles = cdiv(Tn, TBe);
      ea.n_tiles = ea.n_ttiles * B;
      check_decode_range((long long)ea.n_tiles, ea.n_ttiles, "PitchClass2PitchClass (training)");
      ea.wimg = wimg, ea.scale = scale, ea.shift = shift, ea.out_f32 = out.p, ea.Cout_store = Cout, ea.maxbits = maxbits;
      pc2pc_umma_kernel<3><<<std::min(ea.n_tiles, sm_count()), kPcThreads, smem_e, st>>>(ea);
    } else {
      const int n_tt = cdiv(Tn, kPc8MaxTB), TB8 = (cdiv(Tn, n_tt) + 1) / 2 * 2;
      const size_t smem8 = pc8_smem_bytes(TB8 + 6);
      ensure_dyn_smem(pc8_umma_kernel<2>, smem8);
      Pc8Args a8{};
      a8.in_hi = eq_hi, a8.in_lo = eq_lo, a8.Wd_in = Wd, a8.T_out = Tn, a8.TB = TB8, a8.n_ttiles = cdiv(Tn, TB8);
      a8.n_tiles = a8.n_ttiles * B;
      check_decode_range((long long)a8.n_tiles, a8.n_ttiles, "layer-0 PitchClass2PitchClass (training)");
      a8.wimg = wimg, a8.scale = scale, a8.shift = shift, a8.Cout = Cout, a8.out_f32 = out.p, a8.raw = 1, a8.maxbits = maxbits;
      pc8_umma_kernel<2><<<std::min(a8.n_tiles, sm_count()), kPc8Threads, smem8, st>>>(a8);
    }
    AKE_LAUNCHED();
  }

  // ---- train mode: the first conv of the tonic AND the key head (16 -> 32 | 32 channels, valid in time) in one tensor-core pass
  // (equiv_umma_kernel<64, 1, 2> raw: the eval-mode kernel without BatchNorm / activation), raw planar outputs with their biases
  bool heads_tc_ready = false;
  __half* hd_hi = nullptr;
  __half* hd_lo = nullptr;
  __half* hd_wimg = nullptr;
  float* hd_ss = nullptr;
  bool heads_tc_ok(const Conv& ct, const Conv& ck, int T2) const {
    static const bool on = [] { const char* e = getenv("AKE_TRAIN_TC_HEADS"); return e ? atoi(e) != 0 : true; }();
    return on && p->umma && p->umma_heads && ct.Cin == 16 && ck.Cin == 16 && ct.Cout == 32 && ck.Cout == 32 && ct.KH == 12 && ck.KH == 12 && T2 >= 13;
  }
  void heads_tc(const View& pcp, const Conv& ct, const Conv& ck, View& zt, View& zk) {
    const int T2 = pcp.T, T1 = T2 - 6;
    if (!heads_tc_ready) {
      const size_t halves = (size_t)B * 2 * 23 * T2 * 8 + 64 * 8;
      hd_hi = arena.take<__half>(halves), hd_lo = arena.take<__half>(halves);
      hd_wimg = arena.take<__half>((size_t)84 * 4096 / 2);
      hd_ss = arena.take<float>(128);
      heads_tc_ready = true;
    }
    if (dry) return;
    ProfScope prof("pcn.equiv", st);
    EqPackArgs pa{pcp.p, B, pcp.C, 2, T2, T2, 0, 0, 0, nullptr, hd_hi, hd_lo};
    eq_pack_planes_kernel<<<ew_blocks((long long)B * 2 * 23 * T2), 256, 0, st>>>(pa);
    AKE_LAUNCHED();
    equiv_pack_weights_kernel<<<168, 256, 0, st>>>(p->d_params + ct.w_off, p->d_params + ck.w_off, 32, 64, 16, 1, hd_wimg);
    AKE_LAUNCHED();
    heads_raw_ss_kernel<<<1, 64, 0, st>>>(ct.has_bias ? p->d_params + ct.b_off : nullptr, ck.has_bias ? p->d_params + ck.b_off : nullptr, hd_ss);
    AKE_LAUNCHED();
    const int n_tt = cdiv(T1, 32), TBe = (cdiv(T1, n_tt) + 1) / 2 * 2;
    const size_t smem_e = equiv_smem_bytes(TBe + 6);
    ensure_dyn_smem(equiv_umma_kernel<64, 1, 2>, smem_e);
    EquivArgs ea{};
    ea.in_hi = hd_hi, ea.in_lo = hd_lo, ea.Wd_in = T2, ea.T_out = T1, ea.TB = TBe, ea.n_ttiles = cdiv(T1, TBe);
    ea.wimg = hd_wimg, ea.scale = hd_ss, ea.shift = hd_ss + 64, ea.out_f32 = zt.p, ea.out_f32_b = zk.p, ea.raw = 1;
    equiv_umma_kernel<64, 1, 2><<<dim3(ea.n_ttiles, B), 192, smem_e, st>>>(ea);
    AKE_LAUNCHED();
  }

  // data gradient of a head's first conv (16 -> 32 channels, valid in time): d_in (B, 16, 12, T2) (+)= the "full" conv of dz (B, 32, 12, T2 - 6)
  // with the tap-flipped weights = pc2pc_umma_kernel<3> on planes with six zero halo columns, once per 16 gradient channels (the second
  // launch, and a second head, accumulate)
  void heads_dgrad_tc(int id, const View& dz, const unsigned* maxbits, const float* ones, const float* zeros, View& d_in, bool accumulate) {
    const Conv& c = p->convs[id];
    const int T1 = dz.T, T2 = d_in.T, Wd = T2 + 6;
    if (dry) return;
    ProfScope prof("pcn.equiv", st);
    const int n_tt = cdiv(T2, 32), TBe = (cdiv(T2, n_tt) + 1) / 2 * 2;
    const size_t smem_e = pc2pc_smem_bytes(TBe + 6);
    ensure_dyn_smem(pc2pc_umma_kernel<3>, smem_e);
    for (int part = 0; part < 2; ++part) {
      EqPackArgs pa{dz.p, B, dz.C, 2, T1, Wd, 6, 16 * part, 0, maxbits, eq_hi, eq_lo};
      eq_pack_planes_kernel<<<ew_blocks((long long)B * 2 * 23 * Wd), 256, 0, st>>>(pa);
      AKE_LAUNCHED();
      Pc2PcArgs ea{};
      ea.in_hi = eq_hi, ea.in_lo = eq_lo, ea.Wd_in = Wd, ea.T_out = T2, ea.TB = TBe, ea.n_ttiles = cdiv(T2, TBe);
      ea.n_tiles = ea.n_ttiles * B;
      ea.wimg = tcw_block + tcw_head.at(id * 2 + part), ea.scale = ones, ea.shift = zeros, ea.out_f32 = d_in.p, ea.Cout_store = c.Cin;
      ea.maxbits = maxbits, ea.accumulate = (accumulate || part > 0) ? 1 : 0;
      pc2pc_umma_kernel<3><<<std::min(ea.n_tiles, sm_count()), kPcThreads, smem_e, st>>>(ea);
      AKE_LAUNCHED();
    }
  }

  // dW of a 7x7 circular conv (input cat[in0, tile(in1)], output gradient dz with largest |dz| = maxbits) on the tensor cores,
  // on stream `ws` (the backward's side stream); overwrites dw (Cout, Cin, 7, 7).  pcn_train_tc.cuh: p2p_wgrad_umma_kernel.
  void tc_wgrad(const View& in0, const View* in1, const Conv& c, const View& dz, const unsigned* maxbits, float* dw, cudaStream_t ws) {
    const int P = in0.R, Tn = in0.T, Wd = Tn + 6, Wg = cdiv(Tn, 16) * 16;
    if (!wg_ready) {
      for (int i = 0; i < 2; ++i) {
        wg_x[i] = arena.take<__half>((size_t)B * (P + 6) * Wd * 8 + 64 * 8);
        wg_g[i] = arena.take<__half>((size_t)B * (P + 6) * Wg * 8 + 64 * 8);
      }
      wg_partial = arena.take<float>((size_t)sm_count() * 64 * 56);
      wg_ready = true;
    }
    if (dry) return;
    TcPackArgs px{};
    px.in0 = in0.p, px.bs0 = in0.bstride(), px.c0 = in0.C;
    px.in1 = in1 ? in1->p : in0.p, px.bs1 = in1 ? in1->bstride() : 0, px.c1 = in1 ? in1->C : 0, px.rows1 = in1 ? in1->R : 1;
    px.B = B, px.P = P, px.T = Tn, px.Wd = Wd, px.col_shift = 3, px.wrap_cols = 1, px.maxbits = nullptr, px.hi = wg_x[0], px.lo = wg_x[1];
    tc_pack_planes_kernel<<<ew_blocks((long long)B * (P + 6) * Wd), 256, 0, ws>>>(px);
    AKE_LAUNCHED();
    TcPackArgs pg{};
    pg.in0 = dz.p, pg.bs0 = dz.bstride(), pg.c0 = dz.C, pg.in1 = dz.p, pg.bs1 = 0, pg.c1 = 0, pg.rows1 = 1;
    pg.B = B, pg.P = P, pg.T = Tn, pg.Wd = Wg, pg.col_shift = 0, pg.wrap_cols = 0, pg.maxbits = maxbits, pg.hi = wg_g[0], pg.lo = wg_g[1];
    tc_pack_planes_kernel<<<ew_blocks((long long)B * (P + 6) * Wg), 256, 0, ws>>>(pg);
    AKE_LAUNCHED();
    WgradTcArgs wa{wg_x[0], wg_x[1], wg_g[0], wg_g[1], wg_partial, B, P, Tn, Wd, Wg, cdiv(P, kWgR), B * cdiv(P, kWgR)};
    const size_t smem = wgrad_tc_smem_bytes(Wd, Wg);
    ensure_dyn_smem(p2p_wgrad_umma_kernel, smem);
    const int grid = std::min(wa.n_tiles, sm_count());
    p2p_wgrad_umma_kernel<<<grid, kWgThreads, smem, ws>>>(wa);
    AKE_LAUNCHED();
    wgrad_tc_reduce_kernel<<<cdiv(c.Cout * c.Cin * 49, 128), 128, 0, ws>>>(wg_partial, grid, maxbits, c.Cout, c.Cin, dw);
    AKE_LAUNCHED();
  }
  // dW of a 16-channel equivariant conv on the tensor cores, on stream `ws` (eq_wgrad_umma_kernel); overwrites dw (Cout, Cin, 12, 7)
  bool eqw_ready = false;
  size_t eqw_x_halves = 0, eqw_g_halves = 0;
  __half* eqw_x[2] = {};
  __half* eqw_g[2] = {};
  float* eqw_partial = nullptr;
  // equivariant 12 x 7 convs: the "same" convs of both stacks (operand planes with three zero halo columns) and the heads' valid first convs
  bool eq_wgrad_ok(const Conv& c, const ConvGeom& g, int Tn) const {
    static const bool on = [] { const char* e = getenv("AKE_TRAIN_TC_EQ_WGRAD"); return e ? atoi(e) != 0 : true; }();
    const bool same = g.pad_t == 3 && g.T_out == Tn, valid = g.pad_t == 0 && g.T_out == Tn - 6;
    const int n_pairs = cdiv(c.Cin, 8) * cdiv(c.Cout, 8);
    return on && p->umma && g.KH == 12 && c.KH == 12 && g.KW == 7 && g.SR == 1 && g.row_circ && !g.time_circ && g.row_off == 0 && (same || valid) &&
           g.rows_v == 12 && g.rows_out == 12 && Tn >= 13 && c.Cin <= 32 && c.Cout <= 32 && n_pairs <= 16 &&
           eq_wgrad_smem_bytes(Tn + (same ? 6 : 0), cdiv(g.T_out, 16) * 16) <= 227 * 1024 - 256;
  }
  void eq_wgrad(const View& in, const Conv& c, const ConvGeom& g, const View& dz, const unsigned* maxbits, float* dw, cudaStream_t ws) {
    const bool same = g.pad_t == 3;
    const int Tn = in.T, To = dz.T, Wx = Tn + (same ? 6 : 0), Wg = cdiv(To, 16) * 16;
    const int n_gi = cdiv(c.Cin, 8), n_go = cdiv(c.Cout, 8), n_pairs = n_gi * n_go;
    const size_t xh = (size_t)B * n_gi * 23 * Wx * 8 + 64 * 8, gh = (size_t)B * n_go * 23 * Wg * 8 + 64 * 8;
    if (!eqw_ready || xh > eqw_x_halves || gh > eqw_g_halves) {  // the heads come first in a backward pass, the larger stack tensors later: regrow
      eqw_x_halves = std::max(eqw_x_halves, xh), eqw_g_halves = std::max(eqw_g_halves, gh);
      for (int i = 0; i < 2; ++i) eqw_x[i] = arena.take<__half>(eqw_x_halves), eqw_g[i] = arena.take<__half>(eqw_g_halves);
      if (!eqw_ready) eqw_partial = arena.take<float>((size_t)sm_count() * 64 * 96);
      eqw_ready = true;
    }
    if (dry) return;
    EqPackArgs px{in.p, B, in.C, n_gi, Tn, Wx, same ? 3 : 0, 0, 0, nullptr, eqw_x[0], eqw_x[1]};
    eq_pack_planes_kernel<<<ew_blocks((long long)B * n_gi * 23 * Wx), 256, 0, ws>>>(px);
    AKE_LAUNCHED();
    EqPackArgs pg{dz.p, B, dz.C, n_go, To, Wg, 0, 0, 1, maxbits, eqw_g[0], eqw_g[1]};
    eq_pack_planes_kernel<<<ew_blocks((long long)B * n_go * 23 * Wg), 256, 0, ws>>>(pg);
    AKE_LAUNCHED();
    EqWgradArgs wa{eqw_x[0], eqw_x[1], eqw_g[0], eqw_g[1], eqw_partial, B, To, Wx, Wg, n_gi, n_go};
    const size_t smem = eq_wgrad_smem_bytes(Wx, Wg);
    ensure_dyn_smem(eq_wgrad_umma_kernel, smem);
    const int grid = std::min(n_pairs * B, sm_count() / n_pairs * n_pairs);
    eq_wgrad_umma_kernel<<<grid, kEqWgThreads, smem, ws>>>(wa);
    AKE_LAUNCHED();
    eq_wgrad_reduce_kernel<<<cdiv(c.Cout * c.Cin * 84, 128), 128, 0, ws>>>(eqw_partial, grid, maxbits, c.Cout, c.Cin, n_gi, n_go, dw);
    AKE_LAUNCHED();
  }
  bool tc_wgrad_ok(const Conv& c, const ConvGeom& g, int Tn) const {
    static const bool on = [] { const char* e = getenv("AKE_TRAIN_TC_WGRAD"); return e ? atoi(e) != 0 : true; }();
    return on && tc_conv_ok(c, g, Tn) && wgrad_tc_smem_bytes(Tn + 6, cdiv(Tn, 16) * 16) <= 227 * 1024;
  }

  // One row convolution + BatchNorm + LeakyReLU (+ fused time pool).  Returns nothing; writes `out`.
  void conv(int id, const View& in0, const View* in1, const ConvGeom& g, View& out, int out_coff, int act,  // act: 0 none, 1 LeakyReLU, 2 ReLU
            double* stats = nullptr) {  // train mode with a BatchNorm behind: also accumulate the batch sums of the raw outputs
    const Conv& c = p->convs[id];
    const bool has_bn = c.bn >= 0;
    const bool raw = train && has_bn;
    if (dry) return;
    ProfScope prof(g.KH == 7 ? "pcn.p2p" : (g.SR == 3 ? "pcn.semitone" : (g.rows_v == 12 && g.row_circ ? "pcn.equiv" : "pcn.genre")), st);
    ConvArgs a{};
    a.in0 = in0.p, a.c0 = in0.C, a.rows0 = in0.R, a.bs0 = in0.bstride();
    if (in1) a.in1 = in1->p, a.c1 = in1->C, a.rows1 = in1->R, a.bs1 = in1->bstride();
    else a.in1 = in0.p, a.c1 = 0, a.rows1 = 1, a.bs1 = 0;
    a.T_in = in0.T;
    a.rows_v = g.rows_v, a.row_circ = g.row_circ, a.row_off = g.row_off, a.pad_t = g.pad_t, a.time_circ = g.time_circ;
    a.rows_out = g.rows_out, a.T_out = g.T_out;
    a.Cin = c.Cin, a.Cout = c.Cout;
    if (a.c0 + a.c1 != c.Cin) fail(AKE_ERR_INVALID, "internal: conv %d channel mismatch %d+%d vs %d", id, a.c0, a.c1, c.Cin);
    a.w = p->d_packed + c.packed_off, a.cout_pad = c.cout_pad;
    a.scale = scale_of(c, raw), a.shift = shift_of(c, raw);
    a.act = raw ? 0 : act;
    a.out = out.p, a.obs = out.bstride(), a.ocs = (long long)out.R * out.T, a.out_coff = out_coff;
    a.pool_t = raw ? 0 : g.pool_t;
    a.T_store = out.T;
    a.stats = raw ? stats : nullptr;
    const int co_tile = c.cout_pad % 8 == 0 ? 8 : (c.cout_pad % 4 == 0 ? 4 : 1);
    launch_conv(a, g, co_tile, B, st);
  }

  void affine_act(const Conv& c, View& v, int coff, int act = 1) {
    if (dry) return;
    const long long n = (long long)B * c.Cout * v.R * v.T;
    affine_act_kernel<<<ew_blocks(n), 256, 0, st>>>(v.p, B, v.C, coff, c.Cout, v.R * v.T, d_ss_train + c.ss_off,
                                                   d_ss_train + p->n_ss + c.ss_off, act);
    AKE_LAUNCHED();
  }

  // conv + BN + act writing channels [coff, coff+Cout) of `out` (same T), both BN modes.
  void conv_bn_act(int id, const View& in0, const View* in1, const ConvGeom& g, View& out, int coff) {
    conv(id, in0, in1, g, out, coff, true);
    if (train) {
      train_bn(p->convs[id], out, coff);
      affine_act(p->convs[id], out, coff);
    }
  }

  // A Pitch2Pitch / PitchClass2PitchClass stack on the generic path: conv + BN + LeakyReLU repeated, or (opt.resblock) one such
  // conv followed by residual blocks x -> LeakyReLU(x + bn2(conv2(LeakyReLU(bn1(conv1(x)))))) (models.py:402-454).
  View conv_stack(const std::vector<int>& ids, const std::vector<ResPair>& res, const View& in0, const View* in1, const ConvGeom& g,
                  const std::string& tapbase) {
    const int Cout = p->convs[ids[0]].Cout;
    View a = alloc(Cout, g.rows_out, g.T_out), b2 = alloc(Cout, g.rows_out, g.T_out);
    View cur = a;
    conv_bn_act(ids[0], in0, in1, g, a, 0);
    tap(tapbase + "0", a);
    for (size_t i = 1; i < ids.size(); ++i) {
      View dst = cur.p == a.p ? b2 : a;
      conv_bn_act(ids[i], cur, nullptr, g, dst, 0);
      tap(tapbase + std::to_string(i), dst);
      cur = dst;
    }
    if (!res.empty()) {
      View mid = alloc(2 * Cout, g.rows_out, g.T_out), z = alloc(Cout, g.rows_out, g.T_out);
      for (size_t j = 0; j < res.size(); ++j) {
        const Conv& c2 = p->convs[res[j].c2];
        conv_bn_act(res[j].c1, cur, nullptr, g, mid, 0);
        conv(res[j].c2, mid, nullptr, g, z, 0, false);  // eval: BatchNorm folded into the epilogue; train: raw, statistics below
        if (train) train_bn(c2, z, 0);
        View dst = cur.p == a.p ? b2 : a;
        if (!dry) {
          residual_act_kernel<<<ew_blocks(z.numel(B)), 256, 0, st>>>(z.p, cur.p, B, Cout, z.R * z.T, train ? d_ss_train + c2.ss_off : nullptr,
                                                                    train ? d_ss_train + p->n_ss + c2.ss_off : nullptr, dst.p);
          AKE_LAUNCHED();
        }
        tap(tapbase + "res" + std::to_string(j), dst);
        cur = dst;
      }
    }
    return cur;
  }

  // DenseBlock / DenseBlockEquivariant (models.py:582-648): every layer reads the concatenation of the block's input and all
  // earlier layers' outputs -- norm1 + LeakyReLU on that concatenation, a 1-wide bottleneck conv, norm2 + ReLU, a k-wide conv
  // (zero padded on the pitch rows / circular over the pitch classes, zero padded in time) -- and appends `growth` channels.
  // `init`: the block's input, materialised (C_in, R, T).  Returns the (C_in + layers * growth, R, T) concatenation.
  View dense_stack(const std::vector<DenseLayer>& dl, const View& init, bool equivariant, const std::string& tapbase) {
    const int k = p->cfg.kernel_size, g = p->convs[dl[0].conv2].Cout, Tn = init.T, R = init.R;
    View feat = alloc(init.C + (int)dl.size() * g, R, Tn);
    if (!dry) {
      tile_rows_kernel<<<ew_blocks(init.numel(B)), 256, 0, st>>>(init.p, B, init.C, R, R, Tn, feat.p, feat.C, 0);
      AKE_LAUNCHED();
    }
    const ConvGeom g1 = equivariant ? ConvGeom{12, 1, 1, 12, 1, 0, 0, 0, 12, Tn} : ConvGeom{1, 1, 1, R, 0, 0, 0, 0, R, Tn};
    const ConvGeom g2 = equivariant ? ConvGeom{12, k, 1, 12, 1, 0, k / 2, 0, 12, Tn} : ConvGeom{k, k, 1, R, 0, -(k / 2), k / 2, 0, R, Tn};
    for (size_t i = 0; i < dl.size(); ++i) {
      const Conv& n1 = p->convs[dl[i].norm1];
      const Conv& c1 = p->convs[dl[i].conv1];
      const int Cn = n1.Cout;
      View nb = alloc(Cn, R, Tn), mid = alloc(c1.Cout, R, Tn);
      if (train) train_bn(n1, feat, 0);  // statistics of the first Cn channels of the concatenation
      if (!dry) {
        bn_act_copy_kernel<<<ew_blocks(nb.numel(B)), 256, 0, st>>>(feat.p, B, feat.C, Cn, R * Tn, train ? d_ss_train + n1.ss_off : scale_of(n1, false),
                                                                  train ? d_ss_train + p->n_ss + n1.ss_off : shift_of(n1, false), 1, nb.p);
        AKE_LAUNCHED();
      }
      conv(dl[i].conv1, nb, nullptr, g1, mid, 0, 2);  // eval: norm2 folded into the epilogue, ReLU
      if (train) {
        train_bn(c1, mid, 0);
        affine_act(c1, mid, 0, 2);
      }
      conv(dl[i].conv2, mid, nullptr, g2, feat, Cn, 0);  // new features appended behind the concatenation
    }
    tap(tapbase, feat);
    return feat;
  }

  void run(const float* mel, float* key_out, float* tonic_out, float* genre_out);
  // training step (pcn_train.cuh): forward that keeps every activation, and the backward pass over them
  void run_keep(const float* mel, float* key_out, float* tonic_out, float* genre_out, TrainTape& tape);
  void backward_keep(const TrainTape& tape, const float* d_key, const float* d_tonic, const float* d_genre, float* grads);
};

void Fwd::run(const float* mel, float* key_out, float* tonic_out, float* genre_out) {
  const ake_pcn_config& cfg = p->cfg;
  const int P = cfg.pitches, S = P / 3, k = cfg.kernel_size;
  if (!dry) p->taps.clear();
  if (train) {
    d_stats = arena.take<double>(2 * (size_t)p->n_ss);
    d_ss_train = arena.take<float>(2 * (size_t)p->n_ss);
    if (!dry) AKE_CUDA(cudaMemsetAsync(d_stats, 0, sizeof(double) * 2 * p->n_ss, st));
  }
  const ConvGeom g_sem{3, 3, 3, P, 0, 0, 1, 1, S, 0};
  auto g_equiv = [&](int Tn, bool same) { return ConvGeom{12, k, 1, 12, 1, 0, same ? k / 2 : 0, 0, 12, same ? Tn : Tn - k + 1}; };

  View p_in;  // current pitch-wise features
  p_in.p = const_cast<float*>(mel), p_in.C = 1, p_in.R = P, p_in.T = T;
  View pc;    // current pitch-class features
  int Tn = T;

  const bool v_stay = cfg.stay_sixth != 0, v_mem = cfg.pc2p_mem != 0, v_local = cfg.local != 0;
  const bool variant = v_stay || v_mem || v_local || cfg.resblock || cfg.p2pc_conv || cfg.denseblock;
  // Pitch2PitchClass on semitone rows (C, S, Tn): the octave max pool (models.py:82-106) or, with opt.p2pc_conv, the dilated
  // convolution + BN + LeakyReLU of models.py:108-133.  `raw_of`: train mode, the BN + act of that conv are still to be applied.
  auto pitch_class_pool = [&](int L, View& semi, const Conv* raw_of, View& dst_cat, int coff) {
    const LayerPlan& lp = p->layers[L];
    if (lp.pool_conv < 0) {
      if (!dry) {
        octmax_kernel<<<ew_blocks((long long)B * semi.C * 12 * Tn), 256, 0, st>>>(
            semi.p, B, semi.C, semi.R, Tn, raw_of ? d_ss_train + raw_of->ss_off : nullptr,
            raw_of ? d_ss_train + p->n_ss + raw_of->ss_off : nullptr, raw_of ? 1 : 0, dst_cat.p, dst_cat.C, coff);
        AKE_LAUNCHED();
      }
      return;
    }
    if (raw_of) affine_act(*raw_of, semi, 0);  // the convolution needs the activated map
    const Conv& cp = p->convs[lp.pool_conv];
    if (!dry) {
      p2pc_conv_kernel<<<ew_blocks((long long)B * semi.C * 12 * Tn), 256, 0, st>>>(semi.p, p->d_params + cp.w_off, B, semi.C, semi.R, cp.KH, Tn,
                                                                                 scale_of(cp, train), shift_of(cp, train), train ? 0 : 1,
                                                                                 dst_cat.p, dst_cat.C, coff);
      AKE_LAUNCHED();
    }
    if (train) {
      train_bn(cp, dst_cat, coff);
      affine_act(cp, dst_cat, coff);
    }
  };
  // pool_semi conv + BN + act (models.py:361-363 / 386-388), then Pitch2PitchClass (:368 / :389).  Returns the semitone map
  // (activated when `need_act`, e.g. opt.stay_sixth keeps it as the next layer's pitch-wise input).
  auto semitone_pool = [&](int L, const View& src, View& dst_cat, int coff, bool need_act = false) {
    const LayerPlan& lp = p->layers[L];
    const Conv& c = p->convs[lp.sem];
    View semi = alloc(c.Cout, S, Tn);
    ConvGeom g = g_sem;
    g.T_out = Tn;
    conv(lp.sem, src, nullptr, g, semi, 0, true);
    const Conv* raw_of = nullptr;
    if (train) {
      train_bn(c, semi, 0);
      raw_of = &c;
      if (need_act) affine_act(c, semi, 0), raw_of = nullptr;
    }
    pitch_class_pool(L, semi, raw_of, dst_cat, coff);
    tap("l" + std::to_string(L) + (raw_of ? ".semi_raw" : ".semi"), semi);
    return semi;
  };

  for (int L = 0; L < cfg.num_layers; ++L) {
    const LayerPlan& lp = p->layers[L];
    const std::string ln = "l" + std::to_string(L);
    View cat;  // input of the pc2pc stack
    if (L == 0) {
      cat = alloc(1, 12, Tn);
      if (!train && p->convs[lp.sem].Cout == 1 && lp.pool_conv < 0) {
        // eval mode: conv + BN + act + octave pool in one pass over the log-CQT
        const Conv& c = p->convs[lp.sem];
        View semi = alloc(1, S, Tn);
        // tensor-core path of the layer-0 equivariant stack (1 -> 4 -> 4 -> 4 channels): 8-channel chunk planes
        l0_fast = p->umma && Tn >= 7;
        for (int id : lp.pc2pc) l0_fast = l0_fast && p->convs[id].Cin <= 8 && p->convs[id].Cout <= 8;
        if (l0_fast)
          for (int i = 0; i < 3; ++i)
            for (int j = 0; j < 2; ++j) l0_planes[i][j] = arena.take<__half>((size_t)B * 23 * (Tn + 6) * 8);
        if (!dry) {
          ProfScope prof("pcn.semitone", st);
          l0_semitone_pool_kernel<<<dim3(cdiv(12 * Tn, 128), 1, B), 128, 0, st>>>(p_in.p, p->d_params + c.w_off, scale_of(c, false),
                                                                            shift_of(c, false), semi.p, cat.p, P, Tn, 1, 0,
                                                                            l0_fast ? l0_planes[0][0] : nullptr,
                                                                            l0_fast ? l0_planes[0][1] : nullptr);
          AKE_LAUNCHED();
        }
        tap("l0.semi", semi);
        if (v_stay) p_in = semi;  // opt.stay_sixth: the semitone map is the pitch-wise feature from here on (models.py:366-367)
      } else {
        View semi = semitone_pool(0, p_in, cat, 0, v_stay);
        if (v_stay) p_in = semi;
      }
      tap("l0.pool", cat);
    } else {
      const bool fast = p->umma && !train && L == 1 && Tn >= 7;
      View p_feat;
      if (!fast) {
        cat = alloc(lp.prev_pc + lp.out_p, 12, Tn);
        if (!dry)  // concat [pc, pc2] (models.py:392): previous pc is copied in, the pool_semi result is written beside it
          AKE_CUDA(cudaMemcpy2DAsync(cat.p, sizeof(float) * cat.bstride(), pc.p, sizeof(float) * pc.bstride(),
                                     sizeof(float) * pc.bstride(), B, cudaMemcpyDeviceToDevice, st));
      }
      if (fast) {
        // ---- tensor-core path: chunk-plane activations (pcn_umma.cuh), up_sixth fused into the first operand build
        const int Wd = Tn + 6;
        const size_t plane_halves = (size_t)B * (P + 6) * Wd * 8;
        __half* x[2][2];
        for (int i = 0; i < 2; ++i)
          for (int j = 0; j < 2; ++j) x[i][j] = arena.take<__half>(plane_halves);
        const Conv& cu = p->convs[lp.up];
        // up_sixth (models.py:372-374) as a (B, 36, T) x 4-channel table; the first 7x7 conv generates its input tiles
        // cat[p, tile(up)] from it and the log-CQT in shared memory (no operand planes for the 5-channel input at all)
        // first conv split by input channel (pcn_p2p1.cuh): the 4 up-sampled channels have period 36 in pitch
        const Conv& c_first = p->convs[lp.p2p[0]];
        const bool split1 = p->d_wimg_f1 && P % 36 == 0 && Tn >= 64 && c_first.Cin == 5 && c_first.Cout == 8;
        float4* up_tab = nullptr;
        __half *up_hi = nullptr, *up_lo = nullptr;
        float* per_tab = nullptr;
        if (split1) {
          up_hi = arena.take<__half>((size_t)B * 42 * Wd * 8), up_lo = arena.take<__half>((size_t)B * 42 * Wd * 8);
          per_tab = arena.take<float>((size_t)B * 36 * Tn * 8);
        } else {
          up_tab = reinterpret_cast<float4*>(arena.take<float>((size_t)B * 36 * Tn * 4));
        }
        if (!dry) {
          ProfScope prof("pcn.prep", st);
          if (split1)
            upsixth_planes_kernel<<<dim3(cdiv(36 * Tn, 128), 1, B), 128, 0, st>>>(pc.p, p->d_params + cu.w_off, scale_of(cu, false),
                                                                              shift_of(cu, false), up_hi, up_lo, Tn, Wd);
          else
            upsixth_table_kernel<<<dim3(cdiv(Tn, 128), 36, B), 128, 0, st>>>(pc.p, p->d_params + cu.w_off, scale_of(cu, false),
                                                                             shift_of(cu, false), up_tab, Tn);
          AKE_LAUNCHED();
        }
        const int n_tt = cdiv(Tn, kP2PMaxTB), TB = cdiv(Tn, n_tt);
        const size_t smem = p2p_smem_bytes(TB + 6);
        int cur = 0;
        for (size_t i = 0; i < lp.p2p.size(); ++i) {
          const Conv& c = p->convs[lp.p2p[i]];
          if (!dry) {
            ProfScope prof("pcn.p2p", st);
            ensure_dyn_smem(p2p_umma_kernel<false>, smem);
            ensure_dyn_smem(p2p_umma_kernel<true>, smem);
            ensure_dyn_smem(p2p_umma_kernel<false, true>, smem);
            ensure_dyn_smem(p2p1_umma_kernel, f1_smem_bytes(TB));
            const int n_rt = cdiv(P, kP2PRows), n_tiles = B * n_rt * cdiv(Tn, TB);
            check_decode_range((long long)B * n_rt * cdiv(Tn, TB), (long long)n_rt * cdiv(Tn, TB), "Pitch2Pitch");
            P2PArgs a{x[cur][0], x[cur][1], x[cur ^ 1][0], x[cur ^ 1][1],
                      reinterpret_cast<const __half*>(reinterpret_cast<const uint8_t*>(p->d_wimg) + i * kP2PWBytes),
                      scale_of(c, false), shift_of(c, false), P, Tn, Wd, TB, cdiv(Tn, TB), n_rt, n_tiles, p_in.p, up_tab};
            const int grid = std::min(n_tiles, sm_count());  // persistent: one CTA per SM
            if (i == 0 && split1) {
              // periodic part on the 36-row image (raw accumulators), then the mel channel + periodic part + BN + LeakyReLU
              const int n_rt36 = cdiv(36, kP2PRows), n_tiles36 = B * n_rt36 * cdiv(Tn, TB);
              const __half* w_per = reinterpret_cast<const __half*>(reinterpret_cast<const uint8_t*>(p->d_wimg_f1) + 4096);
              P2PArgs ap{up_hi, up_lo, nullptr, nullptr, w_per, scale_of(c, false), shift_of(c, false), 36, Tn, Wd, TB, cdiv(Tn, TB),
                         n_rt36, n_tiles36, nullptr, nullptr, per_tab};
              p2p_umma_kernel<false, true><<<std::min(n_tiles36, sm_count()), kP2PThreads, smem, st>>>(ap);
              AKE_LAUNCHED();
              P2P1Args a1{p_in.p, per_tab, x[cur ^ 1][0], x[cur ^ 1][1], p->d_wimg_f1, scale_of(c, false), shift_of(c, false),
                          P, Tn, Wd, TB, cdiv(Tn, TB), n_rt, n_tiles};
              p2p1_umma_kernel<<<grid, kF1Threads, f1_smem_bytes(TB), st>>>(a1);
            } else if (i == 0) {
              p2p_umma_kernel<true><<<grid, kP2PThreadsGen, smem, st>>>(a);
            } else {
              p2p_umma_kernel<false><<<grid, kP2PThreads, smem, st>>>(a);
            }
            AKE_LAUNCHED();
          }
          cur ^= 1;
        }
        const Conv& cs = p->convs[lp.sem];
        // concat [pc, pool_semi(p)] as 16-channel chunk planes (wrap rows + zero halo columns) for the equivariant convs
        const size_t eq_halves = (size_t)B * 2 * 23 * Wd * 8;
        __half* e[3][2];
        for (int i = 0; i < 3; ++i)
          for (int j = 0; j < 2; ++j) e[i][j] = arena.take<__half>(eq_halves);
        if (!dry) {
          ProfScope prof("pcn.semitone", st);
          const int n_oct = P / 36;
          if (p->d_wimg_semi && P % 36 == 0 && n_oct <= kSemiMaxOct) {
            // tensor cores: all octaves of a (clip, pitch class, time tile) accumulate side by side in TMEM
            // tile width: TB + 2 <= 128 anchors, and kSemiBufs tiles of 3 n_oct rows (hi + lo) must fit in shared memory
            // 227 KB of shared memory: kSemiBufs tile buffers of 3 n_oct (TB + 2) + 136 positions (hi + lo), the weights and the
            // hand-over buffers (semi_smem_bytes), ~1 KB of static shared memory
            const int pos_cap = (int)((232448 - 1280 - (long long)semi_smem_bytes(n_oct, 0) + (long long)kSemiBufs * 2 * 136 * 16) / (kSemiBufs * 32)) - 136;
            const int tb_cap = std::min(kSemiMaxTB, pos_cap / (3 * n_oct) - 2);
            const int n_tt = cdiv(Tn, tb_cap), TBs = cdiv(Tn, n_tt);
            SemiUmmaArgs sa{x[cur][0], x[cur][1], p->d_wimg_semi, scale_of(cs, false), shift_of(cs, false), pc.p, e[0][0], e[0][1],
                            B, P, Tn, Wd, n_oct, TBs, n_tt, B * 12 * n_tt};
            check_decode_range((long long)B * 12 * n_tt, 12LL * n_tt, "pool_semi");
            const size_t smem_s = semi_smem_bytes(n_oct, TBs + 2);
            ensure_dyn_smem(semi_umma_kernel, smem_s);
            semi_umma_kernel<<<std::min(sa.n_items, sm_count()), kSemiThreads, smem_s, st>>>(sa);
          } else {
            SemiArgs sa{x[cur][0], x[cur][1], p->d_params + cs.w_off, scale_of(cs, false), shift_of(cs, false), pc.p,
                        e[0][0], e[0][1], B, P, Tn, Wd};
            semitone_pool_chunks_kernel<<<dim3(cdiv(Wd, 128), 12, B), 128, 0, st>>>(sa);
          }
          AKE_LAUNCHED();
        }
        // PitchClass2PitchClass stack (models.py:393) on tensor cores; MaxPool2d((1,2)) fused into the last conv
        const int Th = Tn / 2;
        View pooled = alloc(lp.out_pc, 12, Th);
        umma_pc_hi = arena.take<__half>((size_t)B * 2 * 23 * Th * 8);
        umma_pc_lo = arena.take<__half>((size_t)B * 2 * 23 * Th * 8);
        if (!dry) {
          ProfScope prof("pcn.equiv", st);
          // the "same" padding of the following conv = zero halo columns of its input planes
          for (int i = 1; i < 3; ++i)
            for (int j = 0; j < 2; ++j) AKE_CUDA(cudaMemsetAsync(e[i][j], 0, sizeof(__half) * eq_halves, st));
          const int n_tt = cdiv(Tn, 32), TBe = (cdiv(Tn, n_tt) + 1) / 2 * 2;
          const size_t smem_e = pc2pc_smem_bytes(TBe + 6);
          ensure_dyn_smem(pc2pc_umma_kernel<0>, smem_e);
          ensure_dyn_smem(pc2pc_umma_kernel<1>, smem_e);
          int ce = 0;
          for (size_t i = 0; i < lp.pc2pc.size(); ++i) {
            const Conv& c = p->convs[lp.pc2pc[i]];
            const bool last = i + 1 == lp.pc2pc.size();
            const int nxt = last ? ce : (ce == 1 ? 2 : 1);
            Pc2PcArgs ea{};
            ea.in_hi = e[ce][0], ea.in_lo = e[ce][1], ea.Wd_in = Wd, ea.T_out = Tn, ea.TB = TBe, ea.n_ttiles = cdiv(Tn, TBe);
            ea.n_tiles = ea.n_ttiles * B;
            check_decode_range((long long)ea.n_ttiles * B, ea.n_ttiles, "PitchClass2PitchClass");
            ea.wimg = reinterpret_cast<const __half*>(reinterpret_cast<const uint8_t*>(p->d_wimg_pc) + i * kPcWBytes);
            ea.scale = scale_of(c, false), ea.shift = shift_of(c, false);
            const int grid = std::min(ea.n_tiles, sm_count());  // persistent: one CTA per SM
            if (!last) {
              ea.out_hi = e[nxt][0], ea.out_lo = e[nxt][1], ea.Wd_out = Wd, ea.col_off = 3;
              pc2pc_umma_kernel<0><<<grid, kPcThreads, smem_e, st>>>(ea);
            } else {
              ea.out_hi = umma_pc_hi, ea.out_lo = umma_pc_lo, ea.Wd_out = Th, ea.col_off = 0, ea.out_f32 = pooled.p;
              pc2pc_umma_kernel<1><<<grid, kPcThreads, smem_e, st>>>(ea);
            }
            AKE_LAUNCHED();
            ce = nxt;
          }
        }
        pc = pooled;
        Tn = Th;
        umma_pc_ready = true;
      } else {
      const int Pl = v_stay ? S : P;  // rows of the pitch-wise features (opt.stay_sixth: semitones)
      View up;                         // second input of the first conv: rows tiled over the pitches (models.py:135-143)
      View p_first = p_in;
      const View* in1 = nullptr;
      if (!v_stay) {
        // up_sixth ConvTranspose + BN + act (models.py:372-374); tiled to all pitches by the conv loader (:378)
        const Conv& cu = p->convs[lp.up];
        up = alloc(lp.prev_pc, 36, Tn);
        if (!dry) {
          const bool raw = train;
          const long long n = up.numel(B);
          upsixth_kernel<<<ew_blocks(n), 256, 0, st>>>(pc.p, p->d_params + cu.w_off, scale_of(cu, raw), shift_of(cu, raw),
                                                      raw ? 0 : 1, up.p, B, lp.prev_pc, Tn);
          AKE_LAUNCHED();
        }
        if (train) {
          train_bn(cu, up, 0);
          affine_act(cu, up, 0);
        }
        tap(ln + ".up", up);
        if (v_mem) {
          // opt.pc2p_mem (models.py:145-166, 376): p += channel-group sums of the up-sampled features; nothing is concatenated
          if (P % 36 || lp.prev_pc % lp.prev_p) fail(AKE_ERR_UNSUPPORTED, "pc2p_mem needs pitches %% 36 == 0 and %d %% %d == 0", lp.prev_pc, lp.prev_p);
          View pm = alloc(lp.prev_p, P, Tn);
          if (!dry) {
            pc2p_mem_add_kernel<<<ew_blocks(pm.numel(B)), 256, 0, st>>>(p_in.p, up.p, B, lp.prev_p, lp.prev_pc, P, Tn, pm.p);
            AKE_LAUNCHED();
          }
          tap(ln + ".mem", pm);
          p_first = pm;
        } else {
          in1 = &up;
        }
      } else {
        in1 = &pc;  // PitchClass2Pitch(pitches // 3): the 12 pitch-class rows tiled over the semitones (models.py:323, 380)
      }
      // Pitch2Pitch stack (models.py:384): circular in pitch and time
      ConvGeom gp{k, k, 1, Pl, 1, -(k / 2), k / 2, 1, Pl, Tn};
      if (!lp.p2p_dense.empty()) {
        // opt.denseblock: the block normalises its whole input, so cat[p, tile(up_sixth(pc))] (models.py:378-383) is materialised
        View p_cat = alloc(lp.prev_p + lp.prev_pc, P, Tn);
        if (!dry) {
          tile_rows_kernel<<<ew_blocks((long long)B * lp.prev_p * P * Tn), 256, 0, st>>>(p_in.p, B, lp.prev_p, P, P, Tn, p_cat.p, p_cat.C, 0);
          AKE_LAUNCHED();
          tile_rows_kernel<<<ew_blocks((long long)B * lp.prev_pc * P * Tn), 256, 0, st>>>(up.p, B, lp.prev_pc, 36, P, Tn, p_cat.p, p_cat.C, lp.prev_p);
          AKE_LAUNCHED();
        }
        p_feat = dense_stack(lp.p2p_dense, p_cat, false, ln + ".p2p_dense");
      } else if (variant) {
        p_feat = conv_stack(lp.p2p, lp.p2p_res, p_first, in1, gp, ln + ".p2p");
      } else {
        View a = alloc(lp.out_p, P, Tn), b2 = alloc(lp.out_p, P, Tn);
        View* src = nullptr;
        View* dst = &a;
        for (size_t i = 0; i < lp.p2p.size(); ++i) {
          if (i == 0) conv_bn_act(lp.p2p[i], p_in, &up, gp, *dst, 0);
          else conv_bn_act(lp.p2p[i], *src, nullptr, gp, *dst, 0);
          tap(ln + ".p2p" + std::to_string(i), *dst);
          src = dst;
          dst = (dst == &a) ? &b2 : &a;
        }
        p_feat = *src;
      }
      if (v_stay) pitch_class_pool(L, p_feat, nullptr, cat, lp.prev_pc);  // no pool_semi: p already sits on semitones (models.py:390-391)
      else semitone_pool(L, p_feat, cat, lp.prev_pc);
      tap(ln + ".cat", cat);
      }
      // time pooling of the pitch-wise features is only needed if another layer follows (models.py:395)
      if (L + 1 < cfg.num_layers) {
        if (v_local) {
          p_in = p_feat;  // opt.local: no time pooling (models.py:349, 394)
        } else {
          View pp = alloc(lp.out_p, p_feat.R, Tn / 2);
          if (!dry) {
            timepool_kernel<<<ew_blocks(pp.numel(B)), 256, 0, st>>>(p_feat.p, B, lp.out_p, p_feat.R, Tn, nullptr, nullptr, 0, pp.p);
            AKE_LAUNCHED();
          }
          p_in = pp;
        }
      }
    }
    if (umma_pc_ready && L == 1) continue;
    if (L == 0 && l0_fast) {
      // PitchClass2PitchClass stack of layer 0 on tensor cores (pc8_umma_kernel): planes -> planes -> planes -> fp32 pc
      View out = alloc(lp.out_pc, 12, Tn);
      if (!dry) {
        ProfScope prof("pcn.l0", st);
        const int Wd = Tn + 6;
        const size_t halves = (size_t)B * 23 * Wd * 8;
        for (int i = 1; i < 3; ++i)  // zero halo columns of the intermediate planes = the "same" padding of the next conv
          for (int j = 0; j < 2; ++j) AKE_CUDA(cudaMemsetAsync(l0_planes[i][j], 0, sizeof(__half) * halves, st));
        const int n_tt = cdiv(Tn, kPc8MaxTB), TB8 = (cdiv(Tn, n_tt) + 1) / 2 * 2;
        const size_t smem8 = pc8_smem_bytes(TB8 + 6);
        ensure_dyn_smem(pc8_umma_kernel<0>, smem8);
        ensure_dyn_smem(pc8_umma_kernel<2>, smem8);
        for (size_t i = 0; i < lp.pc2pc.size(); ++i) {
          const Conv& c = p->convs[lp.pc2pc[i]];
          const bool last = i + 1 == lp.pc2pc.size();
          const int in = (int)(i % 3), nxt = (int)((i + 1) % 3);
          Pc8Args pa{};
          pa.in_hi = l0_planes[in][0], pa.in_lo = l0_planes[in][1], pa.Wd_in = Wd, pa.T_out = Tn, pa.TB = TB8;
          pa.n_ttiles = cdiv(Tn, TB8), pa.n_tiles = pa.n_ttiles * B;
          check_decode_range((long long)pa.n_ttiles * B, pa.n_ttiles, "layer-0 PitchClass2PitchClass");
          pa.wimg = reinterpret_cast<const __half*>(reinterpret_cast<const uint8_t*>(p->d_wimg_l0) + i * kPc8WBytes);
          pa.scale = scale_of(c, false), pa.shift = shift_of(c, false), pa.Cout = c.Cout;
          const int grid = std::min(pa.n_tiles, sm_count());
          if (!last) {
            pa.out_hi = l0_planes[nxt][0], pa.out_lo = l0_planes[nxt][1], pa.Wd_out = Wd;
            pc8_umma_kernel<0><<<grid, kPc8Threads, smem8, st>>>(pa);
          } else {
            pa.out_f32 = out.p;
            pc8_umma_kernel<2><<<grid, kPc8Threads, smem8, st>>>(pa);
          }
          AKE_LAUNCHED();
        }
      }
      pc = out;
      continue;
    }
    if (variant) {
      // PitchClass2PitchClass stack (models.py:369 / 393) of the non-default architectures, then MaxPool2d((1, 2)) (:396)
      View y = !lp.pc2pc_dense.empty() ? dense_stack(lp.pc2pc_dense, cat, true, ln + ".pc2pc_dense")
                                       : conv_stack(lp.pc2pc, lp.pc2pc_res, cat, nullptr, g_equiv(Tn, true), ln + ".pc2pc");
      if (L > 0 && !v_local) {
        View pooled = alloc(lp.out_pc, 12, Tn / 2);
        if (!dry) {
          timepool_kernel<<<ew_blocks(pooled.numel(B)), 256, 0, st>>>(y.p, B, lp.out_pc, 12, Tn, nullptr, nullptr, 0, pooled.p);
          AKE_LAUNCHED();
        }
        y = pooled;
        Tn /= 2;
      }
      pc = y;
      continue;
    }
    // PitchClass2PitchClass stack (models.py:369 / 393), zero padding in time
    View a = alloc(lp.out_pc, 12, Tn), b2 = alloc(lp.out_pc, 12, Tn);
    View* src = &cat;
    View* dst = &a;
    for (size_t i = 0; i < lp.pc2pc.size(); ++i) {
      const bool last = i + 1 == lp.pc2pc.size();
      ConvGeom g = g_equiv(Tn, true);
      if (last && L > 0) {
        // fuse time_pool_pc (models.py:396) into the last conv's epilogue (eval) or the BN apply (train)
        View pooled = alloc(lp.out_pc, 12, Tn / 2);
        const Conv& c = p->convs[lp.pc2pc[i]];
        if (!train) {
          g.pool_t = 1;
          conv(lp.pc2pc[i], *src, nullptr, g, pooled, 0, true);
        } else {
          conv(lp.pc2pc[i], *src, nullptr, g, *dst, 0, true);
          train_bn(c, *dst, 0);
          if (!dry) {
            timepool_kernel<<<ew_blocks(pooled.numel(B)), 256, 0, st>>>(dst->p, B, lp.out_pc, 12, Tn, d_ss_train + c.ss_off,
                                                                       d_ss_train + p->n_ss + c.ss_off, 1, pooled.p);
            AKE_LAUNCHED();
          }
        }
        pc = pooled;
        Tn /= 2;
      } else {
        conv_bn_act(lp.pc2pc[i], *src, nullptr, g, *dst, 0);
        tap(ln + ".pc2pc" + std::to_string(i), *dst);
        pc = *dst;
        src = dst;
        dst = (dst == &a) ? &b2 : &a;
      }
    }
  }
  tap("pc_final", pc);

  // heads (models.py:750-753)
  auto head = [&](const std::vector<int>& ids, bool equivariant, const char* name) -> View {
    View x = pc;
    int Th = Tn;
    for (size_t i = 0; i < ids.size(); ++i) {
      const Conv& c = p->convs[ids[i]];
      const bool last = i + 1 == ids.size();
      ConvGeom g;
      if (equivariant) g = g_equiv(Th, false);
      else g = ConvGeom{c.KH, k, 1, 12, 0, 0, 0, 0, 12 - c.KH + 1, Th - k + 1};
      View y = alloc(c.Cout, g.rows_out, g.T_out);
      if (last) conv(ids[i], x, nullptr, g, y, 0, false);
      else conv_bn_act(ids[i], x, nullptr, g, y, 0);
      x = y;
      Th = g.T_out;
    }
    tap(name, x);
    return x;
  };
  View tonic_f, key_f, genre_f;
  bool heads_done = false, heads_folded = false;
  if (umma_pc_ready && p->umma_heads && Tn >= 13) {
    // first conv of both heads in one tensor-core pass (16 -> 32 | 32, valid in time), then the 32 -> 1 convs
    const int T1 = Tn - (k - 1), Tf = T1 - (k - 1);
    // first conv of the tonic / key heads (N = 64 channels) and of the genre head (1 x 7) -> fp16 chunk planes
    const size_t hk_halves = (size_t)B * 8 * 23 * T1 * 8, g_halves = (size_t)B * 4 * 12 * T1 * 8;
    __half* hk_hi = arena.take<__half>(hk_halves);
    __half* hk_lo = arena.take<__half>(hk_halves);
    __half* g_hi = cfg.genre ? arena.take<__half>(g_halves) : nullptr;
    __half* g_lo = cfg.genre ? arena.take<__half>(g_halves) : nullptr;
    if (!dry) {
      ProfScope prof("pcn.equiv", st);
      const int n_tt = cdiv(T1, 32), TBe = (cdiv(T1, n_tt) + 1) / 2 * 2;
      const size_t smem_e = equiv_smem_bytes(TBe + 6);
      ensure_dyn_smem(equiv_umma_kernel<64, 1, 3>, smem_e);
      EquivArgs ea{};
      ea.in_hi = umma_pc_hi, ea.in_lo = umma_pc_lo, ea.Wd_in = Tn, ea.T_out = T1, ea.TB = TBe, ea.n_ttiles = cdiv(T1, TBe);
      ea.wimg = p->d_wimg_heads, ea.scale = p->d_ss_heads, ea.shift = p->d_ss_heads + 64;
      ea.out_hi = hk_hi, ea.out_lo = hk_lo, ea.out_rows = 23;
      equiv_umma_kernel<64, 1, 3><<<dim3(ea.n_ttiles, B), 192, smem_e, st>>>(ea);
      AKE_LAUNCHED();
      if (cfg.genre) {
        ensure_dyn_smem(equiv_umma_kernel<32, 1, 3, 1>, smem_e);
        const Conv& cg = p->convs[p->genre_head[0]];
        ea.wimg = p->d_wimg_genre, ea.scale = scale_of(cg, false), ea.shift = shift_of(cg, false);
        ea.out_hi = g_hi, ea.out_lo = g_lo, ea.out_rows = 12;
        equiv_umma_kernel<32, 1, 3, 1><<<dim3(ea.n_ttiles, B), 192, smem_e, st>>>(ea);
        AKE_LAUNCHED();
      }
    }
    // last conv of every head (32 -> 1) in one tensor-core launch
    tonic_f = alloc(1, 12, Tf), key_f = alloc(1, 12, Tf);
    if (cfg.genre) genre_f = alloc(1, 11, Tf);
    if (!dry && Tf > 0 && !cfg.max_pool) {
      // the last conv is linear and only its temporal mean is used: fold it into 7 windowed time sums per (channel, row)
      ProfScope prof("pcn.heads", st);
      HeadFoldArgs fa{};
      const int nh = cfg.genre ? 3 : 2;
      const Conv* hc[3] = {&p->convs[p->tonic_head[1]], &p->convs[p->key_head[1]], cfg.genre ? &p->convs[p->genre_head[1]] : nullptr};
      float* ho[3] = {tonic_out, key_out, genre_out};
      for (int h = 0; h < nh; ++h) {
        const bool eq = h < 2;
        fa.in_hi[h] = eq ? hk_hi : g_hi, fa.in_lo[h] = eq ? hk_lo : g_lo;
        fa.G_total[h] = eq ? 8 : 4, fa.g0[h] = h == 1 ? 4 : 0, fa.R[h] = eq ? 23 : 12, fa.KH[h] = hc[h]->KH, fa.rows_out[h] = eq ? 12 : 11;
        fa.wrap[h] = eq, fa.sigmoid[h] = h == 1;
        fa.w[h] = p->d_params + hc[h]->w_off, fa.bias[h] = p->d_params + hc[h]->b_off, fa.out[h] = ho[h];
      }
      int pool_div = 1;
      for (int i = 0; i < cfg.num_layers - 1; ++i) pool_div *= cfg.time_pool_size;
      fa.seq_len = seq_len, fa.T1 = T1, fa.Tf = Tf, fa.pool_div = pool_div, fa.head_shrink = (k - 1) * cfg.head_layers;
      fa.out_stride[0] = os_tonic, fa.out_stride[1] = os_key, fa.out_stride[2] = os_genre;
      head_fold_kernel<<<dim3(B, nh), 512, 0, st>>>(fa);
      AKE_LAUNCHED();
      heads_folded = true;
    } else if (!dry && Tf > 0) {
      ProfScope prof("pcn.heads", st);
      HeadUmmaArgs ha{};
      const int nh = cfg.genre ? 3 : 2;
      const Conv* hc[3] = {&p->convs[p->tonic_head[1]], &p->convs[p->key_head[1]], cfg.genre ? &p->convs[p->genre_head[1]] : nullptr};
      float* ho[3] = {tonic_f.p, key_f.p, genre_f.p};
      for (int h = 0; h < nh; ++h) {
        const bool eq = h < 2;
        ha.in_hi[h] = eq ? hk_hi : g_hi, ha.in_lo[h] = eq ? hk_lo : g_lo;
        ha.G_total[h] = eq ? 8 : 4, ha.g0[h] = h == 1 ? 4 : 0, ha.R[h] = eq ? 23 : 12, ha.KH[h] = hc[h]->KH, ha.rows_out[h] = eq ? 12 : 11;
        ha.wimg[h] = reinterpret_cast<const __half*>(reinterpret_cast<const uint8_t*>(p->d_wimg_tail) + (size_t)h * 24 * 512);
        ha.bias[h] = p->d_params + hc[h]->b_off, ha.out[h] = ho[h];
      }
      const int n_tt = cdiv(Tf, 20);
      ha.T1 = T1, ha.Tf = Tf, ha.TB = cdiv(Tf, n_tt), ha.n_ttiles = cdiv(Tf, ha.TB);
      const size_t smem_h = head_tail_smem_bytes(ha.TB + 6);
      ensure_dyn_smem(head_tail_umma_kernel, smem_h);
      head_tail_umma_kernel<<<dim3(ha.n_ttiles, nh, B), 160, smem_h, st>>>(ha);
      AKE_LAUNCHED();
    }
    if (!heads_folded) tap("tonic_frames", tonic_f), tap("key_frames", key_f);  // folded: no per-frame head outputs exist
    heads_done = true;
  } else {
    tonic_f = head(p->tonic_head, true, "tonic_frames");
    key_f = head(p->key_head, true, "key_frames");
  }
  if (cfg.genre && !heads_done) genre_f = head(p->genre_head, false, "genre_frames");
  if (tonic_f.T <= 0) fail(AKE_ERR_INVALID, "T=%d is too short: the heads need more than %d frames after pooling", T,
                           (k - 1) * cfg.head_layers);
  if (v_local) {
    // opt.local (models.py:720-722, 804-810): MaxPool2d((1, W), stride 1) behind the last conv of the key and tonic heads, no
    // temporal mean; the outputs stay (B, 1, rows, T') and the reference merely REINTERPRETS them as (B, T', rows)
    const int W = cfg.frames * cfg.loc_window_size - cfg.head_layers * (k - 1);
    if (W < 1 || tonic_f.T - W + 1 < 1)
      fail(AKE_ERR_INVALID, "opt.local: pooling window %d does not fit the %d head frames", W, tonic_f.T);
    if (!dry) {
      const int To = tonic_f.T - W + 1;
      slide_max_kernel<<<ew_blocks((long long)B * 12 * To), 256, 0, st>>>(key_f.p, B, 12, key_f.T, W, 1, key_out);
      AKE_LAUNCHED();
      slide_max_kernel<<<ew_blocks((long long)B * 12 * To), 256, 0, st>>>(tonic_f.p, B, 12, tonic_f.T, W, 0, tonic_out);
      AKE_LAUNCHED();
      if (cfg.genre) {
        slide_max_kernel<<<ew_blocks((long long)B * 11 * genre_f.T), 256, 0, st>>>(genre_f.p, B, 11, genre_f.T, 1, 0, genre_out);
        AKE_LAUNCHED();
      }
    }
    return;
  }
  if (!dry && !heads_folded) {
    int pool_div = 1;
    for (int i = 0; i < cfg.num_layers - 1; ++i) pool_div *= cfg.time_pool_size;
    const int rows = cfg.genre ? 35 : 24;
    head_reduce_kernel<<<cdiv(B * rows * 32, 256), 256, 0, st>>>(key_f.p, tonic_f.p, cfg.genre ? genre_f.p : nullptr, B,
                                                                tonic_f.T, seq_len, pool_div, (k - 1) * cfg.head_layers,
                                                                cfg.max_pool, key_out, tonic_out, genre_out, os_key, os_tonic, os_genre);
    AKE_LAUNCHED();
  }
}

}  // namespace ake
#include "pcn_train.cuh"
namespace ake {

// Everything the plan allocated on its device (weights and operand images).
static void free_device_state(ake_pcn* p) {
  void** ptrs[] = {(void**)&p->d_params, (void**)&p->d_packed, (void**)&p->d_ss_eval, (void**)&p->d_ss_raw, (void**)&p->d_wimg,
                   (void**)&p->d_wimg_f1, (void**)&p->d_wimg_pc, (void**)&p->d_wimg_l0, (void**)&p->d_wimg_semi, (void**)&p->d_wimg_heads,
                   (void**)&p->d_ss_heads, (void**)&p->d_wimg_genre, (void**)&p->d_wimg_tail};
  for (void** q : ptrs) {
    if (*q) cudaFree(*q);
    *q = nullptr;
  }
  p->has_params = false;
}

static void pack_umma_images(ake_pcn* p, cudaStream_t st);

static void upload_params(ake_pcn* p, const float* flat_dev, int64_t n, cudaStream_t st) {
  if (n != p->n_params) fail(AKE_ERR_INVALID, "expected %lld parameter floats, got %lld", (long long)p->n_params, (long long)n);
  // The plan's buffers live on the device that is current at upload time; a module moved to another GPU re-uploads there.
  const int dev = current_device();
  if (p->device != dev) {
    free_device_state(p);
    p->device = dev;
  }
  if (!p->d_params) {
    AKE_CUDA(cudaMalloc(&p->d_params, sizeof(float) * p->n_params));
    AKE_CUDA(cudaMalloc(&p->d_packed, sizeof(float) * std::max<int64_t>(p->n_packed, 1)));
    AKE_CUDA(cudaMalloc(&p->d_ss_eval, sizeof(float) * 2 * p->n_ss));
    AKE_CUDA(cudaMalloc(&p->d_ss_raw, sizeof(float) * 2 * p->n_ss));
  }
  AKE_CUDA(cudaMemcpyAsync(p->d_params, flat_dev, sizeof(float) * n, cudaMemcpyDeviceToDevice, st));
  // repacked weights + folded epilogues of every convolution: one launch per kPackTableMax convolutions (pack_all_kernel)
  for (size_t c0 = 0; c0 < p->convs.size(); c0 += kPackTableMax) {
    PackTable t{};
    t.n = (int)std::min<size_t>(kPackTableMax, p->convs.size() - c0), t.n_ss = p->n_ss;
    int max_n = 1;
    for (int i = 0; i < t.n; ++i) {
      const Conv& c = p->convs[c0 + i];
      const BnSite* bn = c.bn >= 0 ? &p->bns[c.bn] : nullptr;
      PackEntry& e = t.e[i];
      e.w_off = c.w_off, e.packed_off = c.packed_off, e.b_off = c.b_off;
      e.gamma = bn ? bn->gamma : 0, e.beta = bn ? bn->beta : 0, e.mean = bn ? bn->mean : 0, e.var = bn ? bn->var : 0;
      e.Cout = c.Cout, e.Cin = c.Cin, e.KHW = c.KH * c.KW, e.cout_pad = c.cout_pad, e.ss_off = c.ss_off;
      e.pack = (!c.transposed && !c.norm_only) ? 1 : 0, e.has_bias = c.has_bias ? 1 : 0, e.has_bn = bn ? 1 : 0;
      if (e.pack) max_n = std::max(max_n, c.Cin * c.KH * c.KW * c.cout_pad);
    }
    pack_all_kernel<<<dim3(std::min(32, cdiv(max_n, 256)), t.n), 256, 0, st>>>(t, p->d_params, p->d_packed, p->d_ss_eval, p->d_ss_raw);
    AKE_LAUNCHED();
  }
  // The fp16 hi/lo operand images of the tensor-core path only serve eval-mode forwards: a training loop uploads new
  // parameters every step and never reads them, so after the first build they are rebuilt lazily (ensure_umma_images).
  if (p->umma) {
    if (!p->d_wimg_pc) pack_umma_images(p, st);
    else p->umma_dirty = true;
  }
  p->has_params = true;
}

static void pack_umma_images(ake_pcn* p, cudaStream_t st) {
  {
    if (!p->d_wimg) AKE_CUDA(cudaMalloc(&p->d_wimg, (size_t)kP2PWBytes * p->umma_convs.size()));
    for (size_t i = 0; i < p->umma_convs.size(); ++i) {
      const Conv& c = p->convs[p->umma_convs[i]];
      p2p_pack_weights_kernel<<<14, 256, 0, st>>>(p->d_params + c.w_off, c.Cout, c.Cin,
                                                   reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(p->d_wimg) + i * kP2PWBytes));
      AKE_LAUNCHED();
    }
    if (!p->umma_convs.empty()) {
      const Conv& c = p->convs[p->umma_convs[0]];
      if (c.Cin == 5 && c.Cout == 8) {
        if (!p->d_wimg_f1) AKE_CUDA(cudaMalloc(&p->d_wimg_f1, 4096 + kP2PWBytes));
        p2p1_pack_mel_kernel<<<2, 256, 0, st>>>(p->d_params + c.w_off, c.Cout, c.Cin, p->d_wimg_f1);
        AKE_LAUNCHED();
        p2p_pack_weights_sub_kernel<<<14, 256, 0, st>>>(p->d_params + c.w_off, c.Cout, c.Cin, 1, 4,
                                                       reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(p->d_wimg_f1) + 4096));
        AKE_LAUNCHED();
      }
    }
    {
      const Conv& cs = p->convs[p->layers[1].sem];
      if (cs.Cin == 8 && cs.Cout == 8) {
        if (!p->d_wimg_semi) AKE_CUDA(cudaMalloc(&p->d_wimg_semi, kSemiWBytes));
        semi_pack_weights_kernel<<<5, 128, 0, st>>>(p->d_params + cs.w_off, p->d_wimg_semi);
        AKE_LAUNCHED();
      }
    }
    {
      const std::vector<int>& l0 = p->layers[0].pc2pc;
      if (!p->d_wimg_l0) AKE_CUDA(cudaMalloc(&p->d_wimg_l0, (size_t)kPc8WBytes * l0.size()));
      for (size_t i = 0; i < l0.size(); ++i) {
        const Conv& c = p->convs[l0[i]];
        pc8_pack_weights_kernel<<<21, 256, 0, st>>>(p->d_params + c.w_off, c.Cout, c.Cin,
                                                     reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(p->d_wimg_l0) + i * kPc8WBytes));
        AKE_LAUNCHED();
      }
    }
    const std::vector<int>& pcs = p->layers[1].pc2pc;
    if (!p->d_wimg_pc) AKE_CUDA(cudaMalloc(&p->d_wimg_pc, (size_t)kPcWBytes * pcs.size()));
    for (size_t i = 0; i < pcs.size(); ++i) {
      const Conv& c = p->convs[pcs[i]];
      pc2pc_pack_weights_kernel<<<84, 256, 0, st>>>(p->d_params + c.w_off, c.Cout, c.Cin,
                                                     reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(p->d_wimg_pc) + i * kPcWBytes));
      AKE_LAUNCHED();
    }
    if (p->umma_heads) {
      const Conv& ct = p->convs[p->tonic_head[0]];
      const Conv& ck = p->convs[p->key_head[0]];
      if (!p->d_wimg_heads) {
        AKE_CUDA(cudaMalloc(&p->d_wimg_heads, (size_t)84 * 4096));
        AKE_CUDA(cudaMalloc(&p->d_ss_heads, sizeof(float) * 128));
      }
      equiv_pack_weights_kernel<<<168, 256, 0, st>>>(p->d_params + ct.w_off, p->d_params + ck.w_off, 32, 64, 16, 1, p->d_wimg_heads);
      AKE_LAUNCHED();
      if (!p->d_wimg_tail) {
        AKE_CUDA(cudaMalloc(&p->d_wimg_tail, (size_t)3 * 24 * 512));
        AKE_CUDA(cudaMalloc(&p->d_wimg_genre, (size_t)16384));
      }
      AKE_CUDA(cudaMemsetAsync(p->d_wimg_tail, 0, (size_t)3 * 24 * 512, st));
      {
        const Conv* tc[3] = {&p->convs[p->tonic_head[1]], &p->convs[p->key_head[1]], p->cfg.genre ? &p->convs[p->genre_head[1]] : nullptr};
        for (int h = 0; h < 3; ++h) {
          if (!tc[h]) continue;
          if (tc[h]->Cin != 32) fail(AKE_ERR_UNSUPPORTED, "internal: head tail expects 32 input channels");
          head_tail_pack_kernel<<<cdiv(tc[h]->KH * 256, 256), 256, 0, st>>>(
              p->d_params + tc[h]->w_off, tc[h]->KH, reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(p->d_wimg_tail) + (size_t)h * 24 * 512));
          AKE_LAUNCHED();
        }
        if (p->cfg.genre) {
          const Conv& cg = p->convs[p->genre_head[0]];
          equiv_pack_weights_kernel<<<16, 256, 0, st>>>(p->d_params + cg.w_off, p->d_params + cg.w_off, 32, 32, 16, 1, p->d_wimg_genre, 1);
          AKE_LAUNCHED();
        }
      }
      const Conv* hc[2] = {&ct, &ck};
      for (int h = 0; h < 2; ++h) {
        AKE_CUDA(cudaMemcpyAsync(p->d_ss_heads + 32 * h, p->d_ss_eval + hc[h]->ss_off, sizeof(float) * 32, cudaMemcpyDeviceToDevice, st));
        AKE_CUDA(cudaMemcpyAsync(p->d_ss_heads + 64 + 32 * h, p->d_ss_eval + p->n_ss + hc[h]->ss_off, sizeof(float) * 32,
                                 cudaMemcpyDeviceToDevice, st));
      }
    }
  }
  p->umma_dirty = false;
}

static inline void ensure_umma_images(ake_pcn* p, cudaStream_t st) {
  if (p->umma && p->umma_dirty) pack_umma_images(p, st);
}

}  // namespace ake

// =============================================================================== extern "C"
extern "C" {

int ake_abi_version(void) { return AKE_ABI_VERSION; }
const char* ake_last_error(void) { return g_last_error.c_str(); }
int64_t ake_launch_count(int reset) {
  int64_t v = g_launches;
  if (reset) g_launches = 0;
  return v;
}

int ake_profile_enable(int on) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_prof_on.store(on != 0);
  if (!on) {
    for (auto& e : g_prof_entries) {
      if (e.ev0) g_prof_pool.push_back(e.ev0);
      if (e.ev1) g_prof_pool.push_back(e.ev1);
    }
    g_prof_entries.clear();
  }
  return AKE_OK;
}

int ake_profile_collect(char* tags_out, int tag_stride, double* ms_out, int64_t* launches_out, int cap) {
  int n_tags = 0;
  int rc = guarded([&] {
    if (!tags_out || !ms_out || !launches_out || tag_stride < 8 || cap <= 0) fail(AKE_ERR_INVALID, "bad arguments");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    std::vector<std::string> tags;
    std::vector<double> ms;
    std::vector<int64_t> launches;
    for (auto& e : g_prof_entries) {
      float t = 0.f;
      if (e.ev0 && e.ev1) {
        AKE_CUDA(cudaEventSynchronize(e.ev1));
        AKE_CUDA(cudaEventElapsedTime(&t, e.ev0, e.ev1));
      }
      size_t k = 0;
      for (; k < tags.size(); ++k)
        if (tags[k] == e.tag) break;
      if (k == tags.size()) tags.push_back(e.tag), ms.push_back(0.0), launches.push_back(0);
      ms[k] += t, launches[k] += e.launches;
      if (e.ev0) g_prof_pool.push_back(e.ev0);
      if (e.ev1) g_prof_pool.push_back(e.ev1);
    }
    g_prof_entries.clear();
    if ((int)tags.size() > cap) fail(AKE_ERR_INVALID, "need room for %zu tags", tags.size());
    for (size_t k = 0; k < tags.size(); ++k) {
      snprintf(tags_out + k * tag_stride, tag_stride, "%s", tags[k].c_str());
      ms_out[k] = ms[k], launches_out[k] = launches[k];
    }
    n_tags = (int)tags.size();
  });
  return rc == AKE_OK ? n_tags : rc;
}

int ake_pcn_create(const ake_pcn_config* cfg, ake_pcn** out) {
  return guarded([&] {
    if (!cfg || !out) fail(AKE_ERR_INVALID, "null argument");
    ake_pcn* p = new ake_pcn();
    p->cfg = *cfg;
    try {
      build_plan(p);
    } catch (...) {
      delete p;
      throw;
    }
    *out = p;
  });
}

void ake_pcn_destroy(ake_pcn* p) {
  if (!p) return;
  free_device_state(p);
  for (auto& kv : p->tapes) delete kv.second;
  delete p;
}

int ake_pcn_num_tensors(const ake_pcn* p) { return p ? (int)p->tensors.size() : AKE_ERR_INVALID; }
const char* ake_pcn_tensor_name(const ake_pcn* p, int i) {
  return (p && i >= 0 && i < (int)p->tensors.size()) ? p->tensors[i].name.c_str() : nullptr;
}
int ake_pcn_tensor_shape(const ake_pcn* p, int i, int64_t shape4[4]) {
  if (!p || i < 0 || i >= (int)p->tensors.size()) return AKE_ERR_INVALID;
  for (int k = 0; k < 4; ++k) shape4[k] = p->tensors[i].shape[k];
  return p->tensors[i].ndim;
}
int64_t ake_pcn_param_floats(const ake_pcn* p) { return p ? p->n_params : AKE_ERR_INVALID; }
int ake_pcn_bn_channels(const ake_pcn* p) { return p ? p->n_bn_ch : AKE_ERR_INVALID; }
int ake_pcn_bn_counts(const ake_pcn* p, int64_t* counts_out, int cap) {
  if (!p || !counts_out || cap < (int)p->bns.size()) return AKE_ERR_INVALID;
  for (size_t i = 0; i < p->bns.size(); ++i) counts_out[i] = i < p->bn_count.size() ? p->bn_count[i] : 0;
  return (int)p->bns.size();
}
int ake_pcn_local_frames(const ake_pcn* p, int T, int* genre_frames_out) {
  if (!p || !p->cfg.local || T <= 0) return AKE_ERR_INVALID;
  const int k = p->cfg.kernel_size;
  const int Tf = T - p->cfg.head_layers * (k - 1);  // no time pooling with opt.local
  const int W = p->cfg.frames * p->cfg.loc_window_size - p->cfg.head_layers * (k - 1);
  if (genre_frames_out) *genre_frames_out = Tf;
  return (W >= 1 && Tf - W + 1 >= 1) ? Tf - W + 1 : AKE_ERR_INVALID;
}
int ake_pcn_get_config(const ake_pcn* p, ake_pcn_config* out) {
  if (!p || !out) return AKE_ERR_INVALID;
  *out = p->cfg;
  return AKE_OK;
}

int ake_pcn_set_params_f32(ake_pcn* p, const float* flat_dev, int64_t n, void* stream) {
  return guarded([&] {
    if (!p || !flat_dev) fail(AKE_ERR_INVALID, "null argument");
    upload_params(p, flat_dev, n, static_cast<cudaStream_t>(stream));
  });
}

size_t ake_pcn_workspace_bytes(const ake_pcn* p, int B, int T, int bn_mode) {
  if (!p || B <= 0 || T <= 0) return 0;
  try {
    Fwd f(const_cast<ake_pcn*>(p), B, T, bn_mode != 0, nullptr, 0, nullptr);
    f.seq_len = nullptr, f.bn_stats_out = nullptr;
    if (bn_mode == 2) {
      TrainTape tape;
      f.run_keep(nullptr, nullptr, nullptr, nullptr, tape);
      f.backward_keep(tape, nullptr, nullptr, nullptr, nullptr);
    } else {
      f.run(nullptr, nullptr, nullptr, nullptr);
    }
    return f.arena.off + 256;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return 0;
  }
}

int ake_pcn_forward_f32(ake_pcn* p, const float* mel_dev, int B, int T, const int32_t* seq_len_dev, int bn_mode,
                        float* key_out_dev, float* tonic_out_dev, float* genre_out_dev, float* bn_stats_out_dev,
                        void* ws_dev, size_t ws_bytes, void* stream) {
  return guarded([&] {
    if (!p || !mel_dev || !key_out_dev || !tonic_out_dev || !ws_dev) fail(AKE_ERR_INVALID, "null argument");
    if (p->cfg.genre && !genre_out_dev) fail(AKE_ERR_INVALID, "genre head enabled but genre_out_dev is NULL");
    if (B <= 0 || T <= 0) fail(AKE_ERR_INVALID, "B and T must be positive (got %d, %d)", B, T);
    if (!p->has_params) fail(AKE_ERR_INVALID, "ake_pcn_set_params_f32 has not been called");
    if (current_device() != p->device)
      fail(AKE_ERR_INVALID, "the plan's weights live on device %d but device %d is current: upload the parameters there first", p->device,
           current_device());
    if (bn_mode == 0) ensure_umma_images(p, static_cast<cudaStream_t>(stream));
    ProfScope prof("pcn.total", static_cast<cudaStream_t>(stream));
    Fwd f(p, B, T, bn_mode != 0, ws_dev, ws_bytes, static_cast<cudaStream_t>(stream));
    f.seq_len = seq_len_dev, f.bn_stats_out = bn_stats_out_dev;
    if (bn_mode == 2) {
      if (p->tapes.size() > 64) {  // forwards whose backward never came (dropped graphs): forget the oldest bookkeeping
        for (auto& kv : p->tapes) delete kv.second;
        p->tapes.clear();
      }
      TrainTape*& tape = p->tapes[ws_dev];
      if (!tape) tape = new TrainTape();
      tape->valid = false;
      f.run_keep(mel_dev, key_out_dev, tonic_out_dev, p->cfg.genre ? genre_out_dev : nullptr, *tape);
      tape->ws = ws_dev;
    } else {
      f.run(mel_dev, key_out_dev, tonic_out_dev, p->cfg.genre ? genre_out_dev : nullptr);
    }
  });
}

int ake_pcn_forward_rows_f32(ake_pcn* p, const float* mel_dev, int B, int T, const int32_t* seq_len_dev, float* rows_out_dev,
                             int32_t* ids_out_dev, void* ws_dev, size_t ws_bytes, void* stream) {
  return guarded([&] {
    if (!p || !mel_dev || !rows_out_dev || !ws_dev) fail(AKE_ERR_INVALID, "null argument");
    if (p->cfg.local) fail(AKE_ERR_UNSUPPORTED, "opt.local produces per-window outputs: use ake_pcn_forward_f32");
    if (B <= 0 || T <= 0) fail(AKE_ERR_INVALID, "B and T must be positive (got %d, %d)", B, T);
    if (!p->has_params) fail(AKE_ERR_INVALID, "ake_pcn_set_params_f32 has not been called");
    if (current_device() != p->device)
      fail(AKE_ERR_INVALID, "the plan's weights live on device %d but device %d is current: upload the parameters there first", p->device,
           current_device());
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ensure_umma_images(p, st);
    ProfScope prof("pcn.total", st);
    // without a genre head the genre columns stay zero
    if (!p->cfg.genre) AKE_CUDA(cudaMemsetAsync(rows_out_dev, 0, sizeof(float) * AKE_ROW_FLOATS * (size_t)B, st));
    Fwd f(p, B, T, false, ws_dev, ws_bytes, st);
    f.seq_len = seq_len_dev, f.bn_stats_out = nullptr;
    f.os_key = f.os_tonic = f.os_genre = AKE_ROW_FLOATS;
    f.run(mel_dev, rows_out_dev, rows_out_dev + 12, p->cfg.genre ? rows_out_dev + 24 : nullptr);
    if (ids_out_dev) {
      decode_kernel<<<cdiv(B, 4), 128, 0, st>>>(rows_out_dev, rows_out_dev + 12, p->cfg.genre ? rows_out_dev + 24 : nullptr, B, ids_out_dev,
                                                  ids_out_dev + B, ids_out_dev + 2 * B, AKE_ROW_FLOATS, AKE_ROW_FLOATS, AKE_ROW_FLOATS);
      AKE_LAUNCHED();
    }
  });
}

int ake_pcn_backward_f32(ake_pcn* p, const float* d_key_out_dev, const float* d_tonic_out_dev, const float* d_genre_out_dev,
                         float* grads_out_dev, int64_t n_floats, void* ws_dev, size_t ws_bytes, void* stream) {
  return guarded([&] {
    if (!p || !grads_out_dev || !ws_dev) fail(AKE_ERR_INVALID, "null argument");
    if (!d_key_out_dev && !d_tonic_out_dev && !d_genre_out_dev) fail(AKE_ERR_INVALID, "no output gradient given");
    if (n_floats != p->n_params) fail(AKE_ERR_INVALID, "grads_out_dev must hold %lld floats", (long long)p->n_params);
    auto it = p->tapes.find(ws_dev);
    if (it == p->tapes.end() || !it->second->valid)
      fail(AKE_ERR_INVALID, "no kept forward in this workspace: call ake_pcn_forward_f32 with bn_mode = 2 on it first (one backward per kept forward)");
    TrainTape* tape = it->second;
    ProfScope prof("pcn.backward", static_cast<cudaStream_t>(stream));
    Fwd f(p, tape->B, tape->T, true, ws_dev, ws_bytes, static_cast<cudaStream_t>(stream));
    f.arena.off = tape->ws_off;
    f.seq_len = tape->seq_len, f.bn_stats_out = nullptr;
    f.backward_keep(*tape, d_key_out_dev, d_tonic_out_dev, d_genre_out_dev, grads_out_dev);
    delete tape;  // the backward pass reuses nothing: one backward per kept forward
    p->tapes.erase(it);
  });
}

int ake_loss_f32(const float* key_out_dev, const float* tonic_out_dev, const float* genre_out_dev, const float* key_labels_dev,
                 const int32_t* tonic_idx_dev, const int32_t* genre_idx_dev, int B, float key_weight, float tonic_weight,
                 float genre_weight, float* loss_out_dev, float* d_key_out_dev, float* d_tonic_out_dev, float* d_genre_out_dev,
                 void* stream) {
  return guarded([&] {
    if (!key_out_dev || !tonic_out_dev || !key_labels_dev || !tonic_idx_dev || !loss_out_dev || !d_key_out_dev || !d_tonic_out_dev)
      fail(AKE_ERR_INVALID, "null argument");
    if (genre_out_dev && !d_genre_out_dev) fail(AKE_ERR_INVALID, "d_genre_out_dev is required with a genre head");
    if (B <= 0) fail(AKE_ERR_INVALID, "B must be positive");
    loss_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(key_out_dev, tonic_out_dev, genre_out_dev, key_labels_dev, tonic_idx_dev,
                                                                 genre_idx_dev, B, key_weight, tonic_weight, genre_weight, loss_out_dev,
                                                                 d_key_out_dev, d_tonic_out_dev, d_genre_out_dev);
    AKE_LAUNCHED();
  });
}

int64_t ake_pcn_get_tap(const ake_pcn* p, const char* name, float* out_dev, int64_t cap, void* stream) {
  int64_t n = AKE_ERR_INVALID;
  int rc = guarded([&] {
    if (!p || !name) fail(AKE_ERR_INVALID, "null argument");
    auto it = p->taps.find(name);
    if (it == p->taps.end()) fail(AKE_ERR_INVALID, "no tap named '%s' in the last forward", name);
    n = it->second.second;
    if (out_dev) {
      if (cap < n) fail(AKE_ERR_INVALID, "tap '%s' needs %lld floats", name, (long long)n);
      AKE_CUDA(cudaMemcpyAsync(out_dev, it->second.first, sizeof(float) * n, cudaMemcpyDeviceToDevice,
                               static_cast<cudaStream_t>(stream)));
    }
  });
  return rc == AKE_OK ? n : rc;
}

int ake_decode_f32(const float* key_out_dev, const float* tonic_out_dev, const float* genre_out_dev, int B,
                   int32_t* key_id_dev, int32_t* tonic_id_dev, int32_t* genre_id_dev, void* stream) {
  return guarded([&] {
    if (B <= 0) fail(AKE_ERR_INVALID, "B must be positive");
    if ((key_id_dev && !key_out_dev) || (tonic_id_dev && !tonic_out_dev)) fail(AKE_ERR_INVALID, "null argument");
    decode_kernel<<<cdiv(B, 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(key_out_dev, tonic_out_dev, genre_out_dev, B,
                                                                            key_id_dev, tonic_id_dev, genre_id_dev);
    AKE_LAUNCHED();
  });
}

int ake_mirex_f32(const float* key_out_dev, const float* tonic_out_dev, const float* key_labels_dev,
                  const float* tonic_labels_dev, const float* key_signature_id_dev, int sig_width, int B, uint64_t* counters_dev,
                  float* similarity_out_dev, int32_t* category_out_dev, void* stream) {
  return guarded([&] {
    if (B <= 0) fail(AKE_ERR_INVALID, "B must be positive");
    if (sig_width <= 0) fail(AKE_ERR_INVALID, "sig_width must be positive");
    if (!key_out_dev || !tonic_out_dev || !key_labels_dev || !tonic_labels_dev || !key_signature_id_dev || !counters_dev)
      fail(AKE_ERR_INVALID, "null argument");
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "counter width");
    mirex_kernel<<<cdiv(B, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
        key_out_dev, tonic_out_dev, key_labels_dev, tonic_labels_dev, key_signature_id_dev, sig_width, B,
        reinterpret_cast<unsigned long long*>(counters_dev), similarity_out_dev, category_out_dev);
    AKE_LAUNCHED();
  });
}

int ake_adam_step_f32(const float* flat_grads_dev, float* m_flat_dev, float* v_flat_dev, float* const* param_ptrs_dev,
                      const int64_t* offsets_dev, int n_tensors, int64_t total, float lr, float beta1, float beta2, float eps,
                      float weight_decay, float grad_scale, int step, void* stream) {
  return guarded([&] {
    if (!flat_grads_dev || !m_flat_dev || !v_flat_dev || !param_ptrs_dev || !offsets_dev) fail(AKE_ERR_INVALID, "null argument");
    if (n_tensors <= 0 || total <= 0 || step < 1) fail(AKE_ERR_INVALID, "n_tensors, total and step must be positive");
    static_assert(sizeof(long long) == sizeof(int64_t), "offset width");
    AdamArgs a{flat_grads_dev, m_flat_dev, v_flat_dev, param_ptrs_dev, reinterpret_cast<const long long*>(offsets_dev), n_tensors,
               (long long)total, lr, beta1, beta2, eps, weight_decay, grad_scale,
               (float)(1.0 - std::pow((double)beta1, step)), (float)std::sqrt(1.0 - std::pow((double)beta2, step))};
    adam_step_kernel<<<ew_blocks(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    AKE_LAUNCHED();
  });
}

}  // extern "C"
