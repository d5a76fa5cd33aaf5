// tcr_augment.cu — the per-clip input stage in front of the MFCC kernel, on the device.
//
// Replaces, for one batch, what the reference's tf.data map does per file on host threads
// (datasets/augmentation_factory.py; called from datasets/audio_data_wrapper.py:37-58 / data_wrapper_base.py:59-89):
//   decode_wav: int16 / 32768 -> f32, crop or zero-pad to desired_samples              (augmentation_factory.py:146-158)
//   silent clips ("" filename): zeros                                                   (:172-178, :193-199)
//   _shift_audio: shift by s in [-L/10, L/10) with zero fill                            (:104-143)
//   _mix_background: clip(background[o : o+L] * volume + foreground, -1, 1)             (:30-101)
// The random draws (shift, background choice / offset / volume) are made by the host and passed per clip, so the result is
// bit-identical to the host restatement for the same draws (TF's own RNG streams are not reproducible).
// One elementwise pass: 2 B (pcm) + 4 B (background) read, 4 B written per sample -> HBM-bound.
#include "tcr_device.cuh"
#include "tcr_net.h"

namespace tcr {

struct AugArgs {
  const int16_t* pcm; int64_t pcm_stride;      // [n][pcm_stride] int16 samples (only `length` of a row are valid)
  const tcr_augment_clip* clips;               // [n]
  const float* background;                     // concatenated background recordings (may be null: no mixing)
  int64_t background_samples;                  // floats in `background` (0: unknown, offsets are trusted)
  float* out; int clip;                        // [n][clip]
};

__global__ void __launch_bounds__(256) augment_kernel(AugArgs a) {
  pdl_wait();
  const int n = blockIdx.y;
  const tcr_augment_clip c = a.clips[n];
  // decode_wav crops / zero-pads to the clip length; a row never holds more than pcm_stride samples
  const int valid = c.silent ? 0 : max(0, min(min(c.length, a.clip), (int)(a.pcm_stride < 0x7fffffff ? a.pcm_stride : 0x7fffffff)));
  const int16_t* src = a.pcm + (size_t)n * a.pcm_stride;
  // a crop that would run past the end of the bank is not mixed in (the host pads short recordings, device_input_stage.py)
  const bool mix = a.background != nullptr && c.bg_offset >= 0 &&
                   (a.background_samples <= 0 || c.bg_offset + (int64_t)a.clip <= a.background_samples);
  const float* bg = mix ? a.background + c.bg_offset : nullptr;
  float* dst = a.out + (size_t)n * a.clip;
  for (int i4 = blockIdx.x * blockDim.x + threadIdx.x; 4 * i4 < a.clip; i4 += gridDim.x * blockDim.x) {
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = 4 * i4 + k;
      const int j = i - c.shift;                                     // shift >= 0: zeros in front; shift < 0: zeros at the end
      float fg = 0.f;
      if (i < a.clip && j >= 0 && j < valid) fg = (float)src[j] * (1.0f / 32768.0f);
      float s = fg;
      if (mix && i < a.clip) s = __fadd_rn(__fmul_rn(bg[i], c.bg_volume), fg);   // multiply, then add: two roundings like tf.multiply + tf.add
      v[k] = fminf(fmaxf(s, -1.0f), 1.0f);
    }
    if (4 * i4 + 3 < a.clip) {
      st4(dst + 4 * i4, make_float4(v[0], v[1], v[2], v[3]));
    } else {
      for (int k = 0; k < 4 && 4 * i4 + k < a.clip; ++k) dst[4 * i4 + k] = v[k];
    }
  }
}

int augment_launch(const int16_t* pcm, int64_t pcm_stride, const tcr_augment_clip* clips, const float* background, int64_t background_samples,
                   float* out, int clip, int n, cudaStream_t s) {
  AugArgs a{pcm, pcm_stride, clips, background, background_samples, out, clip};
  const int bx = std::max(1, std::min(8, (clip / 4 + 255) / 256));
  TCR_LAUNCH("augment", augment_kernel, dim3(bx, n), dim3(256), 0, s, a);
  return 0;
}

}  // namespace tcr
