// Classic dual-band mode (pystripe `filter_streaks`).  Not part of the reference snapshot —
// its README explains why the authors left it (artifacts around bright cells) — but named by the north star:
//
//   T   = threshold (given, or skimage threshold_otsu of the plane)
//   bg  = min(img, T),  fg = max(img, T)
//   bgf = subband(bg, sigma_bg),  fgf = subband(fg, sigma_fg)
//         subband(x, s): log(1 + x) -> wavedec2(db3) -> cH_l <- irfft(rfft(cH_l) g_l) -> waverec2 -> exp(y) - 1
//         (pure notch: no Otsu mask, no median in-painting; g on the packed FFTPACK index like filtering.py:206-215)
//   f   = sigmoid((img - T) / crossover)                       (filtering.py:13-51)
//   out = clip((fgf f + bgf (1 - f) - dark) / flat, 0, 65535) -> uint16 (truncation)
//
// The two sub-band passes run through the ordinary pipeline (DSTR_FLAG_NOTCH_ONLY | DSTR_FLAG_EXPM1); this file
// holds the three small kernels around them: the clamp that forms bg / fg, the blend, and the per-plane histogram of
// a uint16 plane from which the host derives the Otsu threshold.
#pragma once
#include <cstdint>

namespace dstr {

// out[z][i] = upper ? max(in[z][i], T[z]) : min(in[z][i], T[z])   (float32 out: T need not be an integer)
template <typename IN_T>
__global__ void __launch_bounds__(256)
clamp_band_kernel(const IN_T* __restrict__ in, float* __restrict__ out, size_t plane_px, const float* __restrict__ thr, int upper) {
    const int z = blockIdx.y;
    const float T = thr[z];
    const IN_T* src = in + (size_t)z * plane_px;
    float* dst = out + (size_t)z * plane_px;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane_px; i += (size_t)gridDim.x * blockDim.x) {
        const float x = (float)src[i];
        dst[i] = upper ? fmaxf(x, T) : fminf(x, T);
    }
}

struct BlendArgs {
    const float* bgf;      // background band, filtered (nullptr: the image itself — sigma_bg == 0)
    const float* fgf;      // foreground band, filtered (nullptr: the image itself — sigma_fg == 0)
    const float* thr;      // [Z]
    const float* inv_flat; // nullable, [H*W], 1 / flat
    float crossover;
    float dark;
    int single;            // 1: out = fgf (sigma_fg == sigma_bg), no blend
};

template <typename IN_T>
__global__ void __launch_bounds__(256)
dual_band_blend_kernel(const IN_T* __restrict__ img, uint16_t* __restrict__ out, size_t plane_px, BlendArgs a) {
    const int z = blockIdx.y;
    const float T = a.thr[z];
    const size_t base = (size_t)z * plane_px;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < plane_px; i += (size_t)gridDim.x * blockDim.x) {
        const float x = (float)img[base + i];
        float v;
        if (a.single) {
            v = a.fgf[base + i];
        } else {
            const float f = 1.0f / (1.0f + expf(-(x - T) / a.crossover));
            const float fg = a.fgf ? a.fgf[base + i] : x;
            const float bg = a.bgf ? a.bgf[base + i] : x;
            v = fg * f + bg * (1.0f - f);
        }
        if (a.dark > 0.f) v -= a.dark;
        if (a.inv_flat) v *= a.inv_flat[i];
        v = fminf(fmaxf(v, 0.0f), 65535.0f);  // np.clip, then astype(uint16) truncates
        out[base + i] = (uint16_t)v;
    }
}

// Exact histogram of a uint16 plane, hist[z][65536].  A block counts at most 65535 pixels into 16-bit counters packed two
// per shared-memory word (128 KB), so no counter can overflow, then adds its non-zero bins to the global histogram.
constexpr int HU_PX_PER_BLOCK = 65535;
constexpr int HU_THREADS = 1024;
__global__ void __launch_bounds__(HU_THREADS)
hist_u16_kernel(const uint16_t* __restrict__ in, size_t plane_px, unsigned* __restrict__ hist) {
    extern __shared__ unsigned s_cnt[];  // 32768 words
    const int z = blockIdx.y;
    for (int i = threadIdx.x; i < 32768; i += HU_THREADS) s_cnt[i] = 0u;
    __syncthreads();
    const size_t p0 = (size_t)blockIdx.x * HU_PX_PER_BLOCK;
    const size_t p1 = p0 + HU_PX_PER_BLOCK < plane_px ? p0 + HU_PX_PER_BLOCK : plane_px;
    const uint16_t* src = in + (size_t)z * plane_px;
    for (size_t i = p0 + threadIdx.x; i < p1; i += HU_THREADS) {
        const unsigned v = src[i];
        atomicAdd(&s_cnt[v >> 1], (v & 1u) ? 0x10000u : 1u);
    }
    __syncthreads();
    unsigned* h = hist + (size_t)z * 65536;
    for (int i = threadIdx.x; i < 32768; i += HU_THREADS) {
        const unsigned w = s_cnt[i];
        if (w & 0xffffu) atomicAdd(&h[2 * i], w & 0xffffu);
        if (w >> 16) atomicAdd(&h[2 * i + 1], w >> 16);
    }
}

}  // namespace dstr
