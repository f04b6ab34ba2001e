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
  float* audio;
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

Layout carve(const ake_cqt* cqt, const ake_pcn* pcn, int B, long long n_max, void* ws, size_t ws_bytes, int n_bins) {
  Layout l{};
  Arena ar(ws, ws_bytes);
  l.stride = (long long)align_up((size_t)n_max, 4);
  l.T = ake_cqt_frames(cqt, n_max);
  if (l.T <= 0) fail(AKE_ERR_INVALID, "clip too short");
  l.chunk = chunk_clips(B);
  l.audio = ar.take<float>((size_t)B * l.stride);
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

// One private copy stream + event ring per host thread and device.
struct CopyLane {
  int device = -1;
  cudaStream_t stream = nullptr;
  cudaEvent_t entry = nullptr;
  std::vector<cudaEvent_t> landed;
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

extern "C" {

size_t ake_estimate_workspace_bytes(const ake_cqt* cqt, const ake_pcn* pcn, int B, int64_t n_max) {
  if (!cqt || !pcn || B <= 0 || n_max <= 0) return 0;
  try {
    return carve(cqt, pcn, B, n_max, nullptr, 0, ake_cqt_n_bins(cqt)).total + 256;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return 0;
  }
}

int ake_estimate_host_f32(ake_cqt* cqt, ake_pcn* pcn, const float* audio_host, int64_t stride, const int64_t* lengths_host,
                          int B, int64_t n_max, float* key_out_host, float* tonic_out_host, float* genre_out_host,
                          int32_t* ids_host, void* ws_dev, size_t ws_bytes, void* stream) {
  return guarded([&] {
    if (!cqt || !pcn || !audio_host || !ws_dev) fail(AKE_ERR_INVALID, "null argument");
    if (B <= 0 || n_max <= 0 || stride < n_max) fail(AKE_ERR_INVALID, "bad sizes");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int n_bins = ake_cqt_n_bins(cqt);
    ake_pcn_config cfg;
    ake_pcn_get_config(pcn, &cfg);
    const bool genre = cfg.genre != 0;
    if (!genre && genre_out_host) fail(AKE_ERR_INVALID, "genre_out_host given but the plan has no genre head");
    if (cfg.pitches != n_bins) fail(AKE_ERR_INVALID, "CQT plan has %d bins, network expects %d", n_bins, cfg.pitches);
    Layout l = carve(cqt, pcn, B, n_max, ws_dev, ws_bytes, n_bins);
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
    int dev = 0;
    AKE_CUDA(cudaGetDevice(&dev));
    const int n_chunks = cdiv(B, l.chunk);
    g_lane.ensure(dev, (size_t)n_chunks);
    // the copy stream must not overwrite the audio buffer while earlier work on `st` still reads it
    AKE_CUDA(cudaEventRecord(g_lane.entry, st));
    AKE_CUDA(cudaStreamWaitEvent(g_lane.stream, g_lane.entry, 0));
    for (int c = 0; c < n_chunks; ++c) {
      const int b0 = c * l.chunk, nb = std::min(l.chunk, B - b0);
      AKE_CUDA(cudaMemcpy2DAsync(l.audio + (size_t)b0 * l.stride, sizeof(float) * l.stride, audio_host + (size_t)b0 * stride,
                                 sizeof(float) * stride, sizeof(float) * n_max, nb, cudaMemcpyHostToDevice, g_lane.stream));
      AKE_CUDA(cudaEventRecord(g_lane.landed[c], g_lane.stream));
    }
    for (int c = 0; c < n_chunks; ++c) {
      const int b0 = c * l.chunk, nb = std::min(l.chunk, B - b0);
      AKE_CUDA(cudaStreamWaitEvent(st, g_lane.landed[c], 0));
      int rc = ake_cqt_run_f32(cqt, l.audio + (size_t)b0 * l.stride, l.stride, lengths_host ? lengths_host + b0 : nullptr, nb,
                               n_max, AKE_CQT_LOGMAG, l.mel, T_batch, l.seq_len + b0, l.cqt_ws, l.cqt_ws_bytes, st);
      if (rc != AKE_OK) fail(rc, "%s", ake_last_error());
      rc = ake_pcn_forward_f32(pcn, l.mel, nb, T_batch, l.seq_len + b0, 0, l.key + (size_t)b0 * 12, l.tonic + (size_t)b0 * 12,
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
  });
}

}  // extern "C"
