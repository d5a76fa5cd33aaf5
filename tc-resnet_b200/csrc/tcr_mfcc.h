// tcr_mfcc.h — argument block and launcher of the fused front-end kernel (tcr_mfcc.cu).
#pragma once
#include "tcr_device.cuh"

namespace tcr {

struct MfccArgs {
  const void* wav;          // [N, clip] fp32 samples, or int16 PCM when pcm16 != 0
  int pcm16;
  float* feat;              // [N, frames, features]
  int clip, window, stride, frames, features, mel_bins;
  int fpb;                  // frames per CTA (== warps per CTA)
  int magnitude;            // 0: power spectrogram (MFCC path), 1: magnitude (log-mel path)
  int use_dct;              // 1: MFCC, 0: log-mel output
  const float* window_tab;  // [window] periodic Hann
  const float2* tw;         // [fft/2]   exp(-2 pi i n / (fft/2))
  const float2* tw2;        // [fft/2+1] exp(-2 pi i k / fft)
  const int* mel_start;     // [mel_bins] first FFT bin with non-zero weight
  const int* mel_len;       // [mel_bins]
  const int* mel_off;       // [mel_bins] offset into mel_w
  const float* mel_w;       // packed non-zero weights
  const float* dct;         // [mel_bins, features]
};

size_t mfcc_smem_bytes(const MfccArgs& a, int nf2, int warps);
int mfcc_launch(const MfccArgs& a, int n, int fft_length, cudaStream_t stream);

}  // namespace tcr
