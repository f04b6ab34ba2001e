// pipeline.cu -- host audio in, host predictions out (ake_estimate_host_f32).
// Chains the two hot-path stages exactly as the reference's eval loop does (eval.py:118-129 ->
// KeyDataset.get_all (KeyDataset.py:469-509) -> PitchClassNet.forward (models.py:747-817) ->
// argmax decode (models.py:1083-1085, 1096, 923)), with the H2D / D2H copies inside the call.
//
// The batch is cut into chunks of clips: chunk i+1 crosses PCIe on a private copy stream while chunk i
// runs CQT + forward on the caller's stream (clips are independent in eval mode, so chunking does not
// change any result).  PCIe is the bound of this path (5.76 MB per standard clip).
#include <algorithm>
#include <vector>

#include "common.cuh"

using namespace ake;

namespace {

struct Layout {
  float* audio;        // fp32 clips: the whole batch (fp32 host audio) or one chunk (16-bit PCM host audio)
  int16_t* audio16;    // 16-bit PCM staging of the whole batch (ake_estimate_host_i16), else NULL
  long long* d_len;
  float* mel;
  int* seq_len;
  float* key;
  float* tonic;
  float* genre;
  int* ids;
  void* cqt_ws;
  size_t cqt_ws_bytes;
  void* pcn_ws;
  size_t pcn_ws_bytes;
  long long stride;
  int T, chunk;
  size_t total;
};

int chunk_clips(int B) {
  // >= 8 chunks for large batches (deep overlap), chunks of at least 8 clips (enough CTAs per launch)
  int c = std::max(8, cdiv(B, 8));
  c = std::min(c, 64);
  return std::min(c, B);
}

Layout carve(const ake_cqt* cqt, const ake_pcn* pcn, int B, long long n_max, void* ws, size_t ws_bytes, int n_bins, bool pcm16) {
  Layout l{};
  Arena ar(ws, ws_bytes);
  l.stride = (long long)align_up((size_t)n_max, 8);
  l.T = ake_cqt_frames(cqt, n_max);
  if (l.T <= 0) fail(AKE_ERR_INVALID, "clip too short");
  l.chunk = chunk_clips(B);
  if (pcm16) {
    l.audio16 = ar.take<int16_t>((size_t)B * l.stride);
    l.audio = ar.take<float>((size_t)l.chunk * l.stride);
  } else {
    l.audio = ar.take<float>((size_t)B * l.stride);
  }
  l.d_len = ar.take<long long>(B);
  l.mel = ar.take<float>((size_t)l.chunk * n_bins * l.T);
  l.seq_len = ar.take<int>(B);
  l.key = ar.take<float>((size_t)B * 12);
  l.tonic = ar.take<float>((size_t)B * 12);
  l.genre = ar.take<float>((size_t)B * 11);
  l.ids = ar.take<int>((size_t)B * 3);
  l.cqt_ws_bytes = ake_cqt_workspace_bytes(cqt, l.chunk, n_max);
  l.cqt_ws = ar.take<char>(l.cqt_ws_bytes);
  l.pcn_ws_bytes = ake_pcn_workspace_bytes(pcn, l.chunk, l.T, 0);
  if (l.pcn_ws_bytes == 0) fail(AKE_ERR_INVALID, "%s", ake_last_error());
  l.pcn_ws = ar.take<char>(l.pcn_ws_bytes);
  l.total = ar.off;
  return l;
}

// 16-bit PCM -> fp32 exactly as torchaudio.load normalises it (KeyDataset.py:478-481: int16 / 32768, exact in fp32).
// One thread per 8 samples (16 B in, 32 B out); rows are 16-byte aligned (stride % 8 == 0).
__global__ void pcm16_to_f32_kernel(const int16_t* __restrict__ src, float* __restrict__ dst, long long stride, long long n8, int nb) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8 * nb) return;
  const long long b = i / n8, q = i - b * n8;
  const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + b * stride) + q);
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
  float o[8];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    o[2 * e] = (float)(int16_t)(w[e] & 0xFFFFu) * (1.f / 32768.f);
    o[2 * e + 1] = (float)(int16_t)(w[e] >> 16) * (1.f / 32768.f);
  }
  float4* d = reinterpret_cast<float4*>(dst + b * stride) + 2 * q;
  d[0] = make_float4(o[0], o[1], o[2], o[3]);
  d[1] = make_float4(o[4], o[5], o[6], o[7]);
}

// One private copy stream + event ring per host thread and device.
struct CopyLane {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t entry = nullptr;
  std::vector<cudaEvent_t> landed;
  long long* h_len = nullptr;  // page-locked staging of the per-clip lengths (free again when the call returns: it ends with a sync)
  size_t h_len_cap = 0;
  long long* lengths(size_t n) {
    if (h_len_cap < n) {
      if (h_len) cudaFreeHost(h_len);
      h_len = nullptr, h_len_cap = 0;
      AKE_CUDA(cudaMallocHost(&h_len, sizeof(long long) * n));
      h_len_cap = n;
    }
    return h_len;
  }
  void ensure(int dev, size_t n_events) {
    if (device != dev) {
      // (streams/events of a previous device are leaked deliberately: a host thread switching devices is rare)
      device = dev, stream = nullptr, entry = nullptr, landed.clear();
    }
    if (!stream) {
      AKE_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
      AKE_CUDA(cudaEventCreateWithFlags(&entry, cudaEventDisableTiming));
    }
    while (landed.size() < n_events) {
      cudaEvent_t e;
      AKE_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      landed.push_back(e);
    }
  }
};
thread_local CopyLane g_lane;

}  // namespace

namespace {

template <class Sample>
void estimate_host(ake_cqt* cqt, ake_pcn* pcn, const Sample* audio_host, int64_t stride, const int64_t* lengths_host, int B,
                   int64_t n_max, float* key_out_host, float* tonic_out_host, float* genre_out_host, int32_t* ids_host, void* ws_dev,
                   size_t ws_bytes, void* stream) {
  constexpr bool kPcm16 = sizeof(Sample) == 2;
  if (!cqt || !pcn || !audio_host || !ws_dev) fail(AKE_ERR_INVALID, "null argument");
  if (B <= 0 || n_max <= 0 || stride < n_max) fail(AKE_ERR_INVALID, "bad sizes");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n_bins = ake_cqt_n_bins(cqt);
  ake_pcn_config cfg;
  ake_pcn_get_config(pcn, &cfg);
  const bool genre = cfg.genre != 0;
  if (!genre && genre_out_host) fail(AKE_ERR_INVALID, "genre_out_host given but the plan has no genre head");
  if (cfg.pitches != n_bins) fail(AKE_ERR_INVALID, "CQT plan has %d bins, network expects %d", n_bins, cfg.pitches);
  Layout l = carve(cqt, pcn, B, n_max, ws_dev, ws_bytes, n_bins, kPcm16);
  // every clip is padded to the batch's longest frame count, as KeyDataset.py:242-254 pads a batch
  int T_batch = l.T;
  if (lengths_host) {
    int64_t longest = 0;
    for (int b = 0; b < B; ++b) {
      if (lengths_host[b] < 0 || lengths_host[b] > n_max) fail(AKE_ERR_INVALID, "lengths_host[%d] outside [0, n_max]", b);
      longest = std::max<int64_t>(longest, lengths_host[b]);
    }
    T_batch = ake_cqt_frames(cqt, longest);
    if (T_batch <= 0) fail(AKE_ERR_INVALID, "clips too short");
  }
  const int dev = current_device();
  // chunk boundaries: full chunks, then the last one cut into a half and two quarters (>= 8 clips each) -- the only compute the
  // PCIe transfer cannot hide is that of the final chunk, so it is kept short
  std::vector<int> cuts = {0};
  while (cuts.back() < B) {
    const int left = B - cuts.back();
    if (left > l.chunk) {
      cuts.push_back(cuts.back() + l.chunk);
    } else if (left >= 32 && cuts.size() > 1) {
      const int q = left / 4;
      cuts.push_back(cuts.back() + left - 2 * q), cuts.push_back(cuts.back() + q), cuts.push_back(B);
    } else {
      cuts.push_back(B);
    }
  }
  const int n_chunks = (int)cuts.size() - 1;
  g_lane.ensure(dev, (size_t)n_chunks);
  // the copy stream must not overwrite the audio buffer while earlier work on `st` still reads it
  AKE_CUDA(cudaEventRecord(g_lane.entry, st));
  AKE_CUDA(cudaStreamWaitEvent(g_lane.stream, g_lane.entry, 0));
  if (lengths_host) {
    // every clip's length crosses once, from page-locked memory, ahead of the first chunk (a pageable source would make each
    // chunk's copy wait for the previous chunk's kernels)
    long long* h = g_lane.lengths((size_t)B);
    for (int b = 0; b < B; ++b) h[b] = lengths_host[b];
    AKE_CUDA(cudaMemcpyAsync(l.d_len, h, sizeof(long long) * B, cudaMemcpyHostToDevice, g_lane.stream));
  }
  Sample* stage = kPcm16 ? reinterpret_cast<Sample*>(l.audio16) : reinterpret_cast<Sample*>(l.audio);
  for (int c = 0; c < n_chunks; ++c) {
    const int b0 = cuts[c], nb = cuts[c + 1] - b0;
    AKE_CUDA(cudaMemcpy2DAsync(stage + (size_t)b0 * l.stride, sizeof(Sample) * l.stride, audio_host + (size_t)b0 * stride,
                               sizeof(Sample) * stride, sizeof(Sample) * n_max, nb, cudaMemcpyHostToDevice, g_lane.stream));
    AKE_CUDA(cudaEventRecord(g_lane.landed[c], g_lane.stream));
  }
  for (int c = 0; c < n_chunks; ++c) {
    const int b0 = cuts[c], nb = cuts[c + 1] - b0;
    AKE_CUDA(cudaStreamWaitEvent(st, g_lane.landed[c], 0));
    const float* clips = l.audio + (size_t)b0 * l.stride;
    if (kPcm16) {
      const long long n8 = (n_max + 7) / 8;
      pcm16_to_f32_kernel<<<(unsigned)cdiv64(n8 * nb, 256), 256, 0, st>>>(l.audio16 + (size_t)b0 * l.stride, l.audio, l.stride, n8, nb);
      AKE_LAUNCHED();
      clips = l.audio;
    }
    {
      ProfScope prof("cqt.total", st);
      run_cqt(cqt, clips, l.stride, nullptr, lengths_host ? l.d_len + b0 : nullptr, nb, n_max, AKE_CQT_LOGMAG, l.mel, T_batch,
              l.seq_len + b0, l.cqt_ws, l.cqt_ws_bytes, st);
    }
    int rc = ake_pcn_forward_f32(pcn, l.mel, nb, T_batch, l.seq_len + b0, 0, l.key + (size_t)b0 * 12, l.tonic + (size_t)b0 * 12,
                                 l.genre + (size_t)b0 * 11, nullptr, l.pcn_ws, l.pcn_ws_bytes, st);
    if (rc != AKE_OK) fail(rc, "%s", ake_last_error());
  }
  if (ids_host) {
    int rc = ake_decode_f32(l.key, l.tonic, genre ? l.genre : nullptr, B, l.ids, l.ids + B, l.ids + 2 * B, st);
    if (rc != AKE_OK) fail(rc, "%s", ake_last_error());
    AKE_CUDA(cudaMemcpyAsync(ids_host, l.ids, sizeof(int) * 3 * B, cudaMemcpyDeviceToHost, st));
  }
  if (key_out_host) AKE_CUDA(cudaMemcpyAsync(key_out_host, l.key, sizeof(float) * 12 * B, cudaMemcpyDeviceToHost, st));
  if (tonic_out_host) AKE_CUDA(cudaMemcpyAsync(tonic_out_host, l.tonic, sizeof(float) * 12 * B, cudaMemcpyDeviceToHost, st));
  if (genre_out_host) AKE_CUDA(cudaMemcpyAsync(genre_out_host, l.genre, sizeof(float) * 11 * B, cudaMemcpyDeviceToHost, st));
  AKE_CUDA(cudaStreamSynchronize(st));
}

size_t workspace_bytes(const ake_cqt* cqt, const ake_pcn* pcn, int B, int64_t n_max, bool pcm16) {
  if (!cqt || !pcn || B <= 0 || n_max <= 0) return 0;
  try {
    return carve(cqt, pcn, B, n_max, nullptr, 0, ake_cqt_n_bins(cqt), pcm16).total + 256;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return 0;
  }
}

}  // namespace

extern "C" {

size_t ake_estimate_workspace_bytes(const ake_cqt* cqt, const ake_pcn* pcn, int B, int64_t n_max) {
  return workspace_bytes(cqt, pcn, B, n_max, false);
}
size_t ake_estimate_workspace_bytes_i16(const ake_cqt* cqt, const ake_pcn* pcn, int B, int64_t n_max) {
  return workspace_bytes(cqt, pcn, B, n_max, true);
}

int ake_estimate_host_f32(ake_cqt* cqt, ake_pcn* pcn, const float* audio_host, int64_t stride, const int64_t* lengths_host,
                          int B, int64_t n_max, float* key_out_host, float* tonic_out_host, float* genre_out_host,
                          int32_t* ids_host, void* ws_dev, size_t ws_bytes, void* stream) {
  return guarded([&] {
    estimate_host<float>(cqt, pcn, audio_host, stride, lengths_host, B, n_max, key_out_host, tonic_out_host, genre_out_host, ids_host,
                         ws_dev, ws_bytes, stream);
  });
}

int ake_estimate_host_i16(ake_cqt* cqt, ake_pcn* pcn, const int16_t* pcm_host, int64_t stride, const int64_t* lengths_host,
                          int B, int64_t n_max, float* key_out_host, float* tonic_out_host, float* genre_out_host,
                          int32_t* ids_host, void* ws_dev, size_t ws_bytes, void* stream) {
  return guarded([&] {
    estimate_host<int16_t>(cqt, pcn, pcm_host, stride, lengths_host, B, n_max, key_out_host, tonic_out_host, genre_out_host, ids_host,
                           ws_dev, ws_bytes, stream);
  });
}

}  // extern "C"
