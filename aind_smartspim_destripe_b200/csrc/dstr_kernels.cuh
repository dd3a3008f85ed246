// Device kernels of the B200-native destripe engine (sm_100a).
//
// Pipeline per chunk of Z planes (all planes batched in every launch), see DESIGN.md:
//   analysis_kernel   u16/f32 -> log(1+x) -> one db3 analysis level, keeps only cA_l and cH_l
//                     (cV, cD are never needed: the synthesis side works on *deltas*)
//   hist_kernel       256-bin np.histogram-compatible histogram of cH_l^2 per plane
//   otsu_kernel       skimage threshold_otsu arithmetic (float32, sequential) + threshold cap
//   filter_rows_kernel per row: mask, exact median, in-paint, x - irfft(rfft(x) g) as an exact
//                     time-domain operator, writes dH_l = cH'_l - cH_l in place
//   synth_kernel      dA_{l-1} = idwt2(dA_l, dH_l, 0, 0); final level fused with
//                     (1+x) * exp(delta) + 1 -> dark/flat -> clip -> truncate -> u16
//
// Reference semantics restated: /root/reference/code/aind_smartspim_destripe/filtering.py:139-224
// (log_space_fft_filtering), :54-88 (fg/bg means), :338-414 (flatfield_correction).
#pragma once
#include <cuda/ptx>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace dstr {

// ---- db3 filter bank (PyWavelets float32 path uses float32 copies of these taps) ------------
#define DSTR_LO0 0.035226291882100656f
#define DSTR_LO1 -0.08544127388224149f
#define DSTR_LO2 -0.13501102001039084f
#define DSTR_LO3 0.4598775021193313f
#define DSTR_LO4 0.8068915093133388f
#define DSTR_LO5 0.3326705529509569f

__device__ __forceinline__ float dec_lo(int j) {
    switch (j) {
        case 0: return DSTR_LO0;
        case 1: return DSTR_LO1;
        case 2: return DSTR_LO2;
        case 3: return DSTR_LO3;
        case 4: return DSTR_LO4;
        default: return DSTR_LO5;
    }
}
// dec_hi[k] = (-1)^(k+1) dec_lo[5-k]
__device__ __forceinline__ float dec_hi(int j) {
    switch (j) {
        case 0: return -DSTR_LO5;
        case 1: return DSTR_LO4;
        case 2: return -DSTR_LO3;
        case 3: return DSTR_LO2;
        case 4: return -DSTR_LO1;
        default: return DSTR_LO0;
    }
}
// rec_lo = dec_lo reversed, rec_hi = dec_hi reversed
__device__ __forceinline__ float rec_lo(int j) { return dec_lo(5 - j); }
__device__ __forceinline__ float rec_hi(int j) { return dec_hi(5 - j); }

struct LevelStat {
    unsigned qmin_inv;   // ~bits of min(cH^2)  (zero-initialised; atomicMax)
    unsigned qmax_bits;  // bits of max(cH^2)
    float otsu_raw;      // skimage threshold_otsu(cH^2)
    float thr;           // min(max_threshold, sqrt(otsu_raw))
    int otsu_bin;
    float thr_q;         // largest float q with sqrt_rn(q) <= thr:  sqrt(c*c) > thr  <=>  c*c > thr_q
    int pad[2];
    unsigned hist[256];
};

struct PlaneStat {
    double fg_sum;
    double bg_sum;
    unsigned long long fg_cnt;
    unsigned long long bg_cnt;
};

struct DispatchParams {
    float max_thr_cells;
    float max_thr_nocells;
    float high_int;
    int mode;  // 0: always no_cells; 1: per-plane dispatch; 2: always cells (second pass of the dual-band mode)
    int notch_only;  // 1: no Otsu mask / median in-painting, every coefficient goes through the notch (dual-band mode)
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // half-sample symmetric extension (x[-1-k] = x[k], x[n+k] = x[n-1-k]), repeated for short
    // signals; division-free: one reflection suffices unless the signal is shorter than the halo
    while (i < 0 || i >= n) i = (i < 0) ? (-1 - i) : (2 * n - 1 - i);
    return i;
}

__device__ __forceinline__ int plane_uses_cells(const PlaneStat& ps, const DispatchParams& dp) {
    // filtering.py:462  fore_mean > back_mean and fore_mean > microscope_high_int
    if (dp.mode == 0) return 0;
    if (dp.mode == 2) return 1;
    const double fg = ps.fg_cnt ? ps.fg_sum / (double)ps.fg_cnt : 0.0;
    const double bg = ps.bg_cnt ? ps.bg_sum / (double)ps.bg_cnt : 0.0;
    return (fg > bg && fg > (double)dp.high_int) ? 1 : 0;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- conversions that stay off the quarter-rate XU pipe (no I2F / F2I) --------------------------
// u16 -> f32: (0x4B000000 | u) is the float 2^23 + u; subtracting 2^23 is exact.
// f32 in [0, 65535] -> u16 with truncation: (r + 2^23) rounded toward zero has floor(r) in its
// low mantissa bits.
__device__ __forceinline__ float u16_to_f32(unsigned u) { return __uint_as_float(0x4B000000u | u) - 8388608.0f; }
__device__ __forceinline__ unsigned f32_to_u16_trunc(float r) {
    return __float_as_uint(__fadd_rz(r, 8388608.0f)) & 0xffffu;
}
__device__ __forceinline__ float to_f32(unsigned short v) { return u16_to_f32(v); }
__device__ __forceinline__ float to_f32(float v) { return v; }

// aligned two-element loads / stores (caller guarantees alignment)
__device__ __forceinline__ void load_pair(const unsigned short* p, float& a, float& b) {
    const unsigned u = *reinterpret_cast<const unsigned*>(p);
    a = u16_to_f32(u & 0xffffu);
    b = u16_to_f32(u >> 16);
}
__device__ __forceinline__ void load_pair(const float* p, float& a, float& b) {
    const float2 f = *reinterpret_cast<const float2*>(p);
    a = f.x;
    b = f.y;
}
// values already clipped to [0, 65535]
__device__ __forceinline__ void store_pair(unsigned short* p, float a, float b) {
    __stcs(reinterpret_cast<unsigned*>(p), f32_to_u16_trunc(a) | (f32_to_u16_trunc(b) << 16));  // final output: streaming
}
__device__ __forceinline__ void store_pair(float* p, float a, float b) {
    *reinterpret_cast<float2*>(p) = make_float2(a, b);
}
__device__ __forceinline__ void store_one(unsigned short* p, float a) { *p = (unsigned short)f32_to_u16_trunc(a); }
__device__ __forceinline__ void store_one(float* p, float a) { *p = a; }

template <typename T>
struct PairOf;
template <>
struct PairOf<unsigned short> {
    typedef unsigned type;
    __device__ static __forceinline__ unsigned short lo(unsigned w) { return (unsigned short)(w & 0xffffu); }
    __device__ static __forceinline__ unsigned short hi(unsigned w) { return (unsigned short)(w >> 16); }
};
template <>
struct PairOf<float> {
    typedef float2 type;
    __device__ static __forceinline__ float lo(float2 w) { return w.x; }
    __device__ static __forceinline__ float hi(float2 w) { return w.y; }
};

// log / exp: the MUFU-based intrinsics are as accurate as the library versions on this path's
// domain (log argument >= 1: absolute error ~1e-7 + final rounding; |delta| small for exp);
// -DDSTR_ACCURATE_MATH switches to logf / expf.
__device__ __forceinline__ float fast_log(float x) {  // x >= 1 on this path: no denormal scaling
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y * 0.6931471805599453f;
}
__device__ __forceinline__ float fast_exp(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x * 1.4426950408889634f));
    return y;
}
#ifdef DSTR_ACCURATE_MATH
#define DSTR_LOGF(x) logf(x)
#define DSTR_EXPF(x) expf(x)
#else
#define DSTR_LOGF(x) fast_log(x)
#define DSTR_EXPF(x) fast_exp(x)
#endif

// =============================================================================================
// analysis: one 2-D db3 level, symmetric mode, axis -2 first then axis -1 (pywt.dwt2), keeping
// cA ('aa') and cH ('da': high-pass along Y, low-pass along X).
//
// Register-tiled, no shared-memory staging: one warp owns a strip of 30 output columns x
// AN_TOY output rows.  Lane l holds input columns (2p, 2p+1), p = ox0 - 2 + l, and marches down
// the rows with a 6-row sliding window (axis -2 pass: 24 FMA per output row); the axis -1 pass
// takes the two neighbouring column pairs from lanes l-1 and l-2 by shuffle (8 SHFL + 12 FMA),
// so lanes 2..31 emit outputs ox0 .. ox0+29.  Halo cost: 4 extra input rows per 2*AN_TOY and
// 2 of 32 lanes.  The row loop is unrolled by 3 (the period of the 6-row window).
// =============================================================================================
#ifndef DSTR_AN_TOY
#define DSTR_AN_TOY 24
#endif
constexpr int AN_TOY = DSTR_AN_TOY;  // multiple of 3
constexpr int AN_OXW = 30;  // output columns per warp
constexpr int AN_WX = 2;    // warps of a block along x ...
constexpr int AN_WY = 4;    // ... and along y
constexpr int AN_WARPS = AN_WX * AN_WY;
constexpr int AN_THREADS = 32 * AN_WARPS;

// single reflection + clamp: exact for every index a valid output needs when n >= 8 (rows or
// columns past that are only touched by outputs that are discarded)
__device__ __forceinline__ int reflect_fast(int i, int n) {
    i = (i < 0) ? (-1 - i) : i;
    i = (i >= n) ? (2 * n - 1 - i) : i;
    return max(0, min(i, n - 1));
}

// raw (unconverted) two-column loads: the conversion is deferred to the point of use so that the
// in-order warp does not stall on the load right after issuing it
template <typename IN_T, bool VEC>
struct RawPair;
template <typename IN_T>
struct RawPair<IN_T, true> {
    typename PairOf<IN_T>::type w;
    __device__ __forceinline__ void load(const IN_T* row, int c0, int /*c1*/) {
        w = *reinterpret_cast<const typename PairOf<IN_T>::type*>(row + c0);
    }
    __device__ __forceinline__ void get(bool swp, float& v0, float& v1) const {
        const float a = to_f32(PairOf<IN_T>::lo(w)), b = to_f32(PairOf<IN_T>::hi(w));
        v0 = swp ? b : a;
        v1 = swp ? a : b;
    }
};
template <typename IN_T>
struct RawPair<IN_T, false> {
    IN_T a, b;
    __device__ __forceinline__ void load(const IN_T* row, int c0, int c1) {
        a = row[c0];
        b = row[c1];
    }
    __device__ __forceinline__ void get(bool, float& v0, float& v1) const {
        v0 = to_f32(a);
        v1 = to_f32(b);
    }
};

// VEC: Hs, Ws >= 8, Ws even, even pitch / plane stride: every (reflected) column pair is an aligned
// two-element word, possibly swapped, and one reflection per index suffices; otherwise two scalar
// loads per row with the general (repeated) reflection.
#ifndef DSTR_AN_MINB
#define DSTR_AN_MINB 1
#endif
#ifndef DSTR_SY_MINB
#define DSTR_SY_MINB 1
#endif
template <typename IN_T, bool FIRST, bool STATS, bool VEC>
__global__ void __launch_bounds__(AN_THREADS, DSTR_AN_MINB)
analysis_kernel(const IN_T* __restrict__ in, int Hs, int Ws, int in_pitch, size_t in_pstride,
                float* __restrict__ cA, float* __restrict__ cH, int Ho, int Wo, int out_pitch,
                size_t out_pstride, LevelStat* __restrict__ lstat, int stat_stride,
                PlaneStat* __restrict__ pstat, float fg_thr32) {
    __shared__ float s_red[2][AN_WARPS];
    __shared__ double s_dred[2][AN_WARPS];
    __shared__ unsigned s_cred[2][AN_WARPS];

    const int lane = threadIdx.x & 31;
    const int wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction: the tile test below is no divergence
    const int z = blockIdx.z;
    const int ox0 = (blockIdx.x * AN_WX + (wid % AN_WX)) * AN_OXW;
    const int oy0 = (blockIdx.y * AN_WY + (wid / AN_WX)) * AN_TOY;
    const IN_T* src = in + (size_t)z * in_pstride;

    const int p = ox0 - 2 + lane;  // column pair (2p, 2p+1)
    const int gx0 = 2 * p, gx1 = 2 * p + 1;
    int c0, c1;
    bool swp = false;
    if (VEC) {
        // reflected pair: (2p, 2p+1) -> (2q+1, 2q) with q = -1-p (left) or Ws-1-p (right, Ws even)
        int q = p;
        if (p < 0) {
            q = -1 - p;
            swp = true;
        } else if (gx1 >= Ws) {
            q = Ws - 1 - p;
            swp = true;
        }
        q = max(0, min(q, Ws / 2 - 1));  // lanes of strips past the right edge: any valid address
        c0 = 2 * q;
        c1 = c0 + 1;
    } else {
        c0 = reflect_idx(gx0, Ws);
        c1 = reflect_idx(gx1, Ws);
    }
    const bool own0 = (lane >= 2) && (gx0 < Ws);  // pixel ownership for the plane statistics
    const bool own1 = (lane >= 2) && (gx1 < Ws);

    float fg_s = 0.f, all_s = 0.f;  // per-thread, per-column partial sums (<= 48 pixels each: exact for integers)
    float fg_s1 = 0.f, all_s1 = 0.f;
    unsigned fg_c = 0, fg_c1 = 0, all_c = 0;  // all_c counts rows: both columns share it
    float qmin = __int_as_float(0x7f800000), qmax = 0.f;

    if (ox0 < Wo && oy0 < Ho) {  // warp-uniform
        typedef RawPair<IN_T, VEC> Raw;
        float w0[6], w1[6];
        const IN_T* lane_base = src + (VEC ? c0 : 0);
        // fetch: issue the global loads of input row r (raw); finish: convert, statistics, log.
        // The loads of output row oy+1 are issued before the arithmetic of output row oy so that
        // their latency is covered by the in-order instruction stream of the same warp.
        auto fetch = [&](int r, Raw& e) {
            const int gy = VEC ? reflect_fast(2 * oy0 - 4 + r, Hs) : reflect_idx(2 * oy0 - 4 + r, Hs);
            e.load(lane_base + (unsigned)(gy * in_pitch), VEC ? 0 : c0, c1);
        };
        auto finish = [&](int r, const Raw& e, float& v0, float& v1) {
            e.get(swp, v0, v1);
            if (FIRST) {
                if (STATS) {
                    const int gy0 = 2 * oy0 - 4 + r;
                    if (r >= 4 && gy0 < Hs) {  // rows owned by this tile: [2 oy0, 2 oy0 + 2 AN_TOY)
                        // per-column accumulators, no ownership selects in the loop: lanes / columns that
                        // belong to a neighbour are masked once at the end
                        all_s += v0;
                        all_s1 += v1;
                        ++all_c;
                        if (v0 >= fg_thr32) {
                            fg_s += v0;
                            ++fg_c;
                        }
                        if (v1 >= fg_thr32) {
                            fg_s1 += v1;
                            ++fg_c1;
                        }
                    }
                }
                v0 = DSTR_LOGF(__fadd_rn(1.0f, v0));  // np.log(1.0 + x) in float32
                v1 = DSTR_LOGF(__fadd_rn(1.0f, v1));
            }
        };
        Raw e[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) fetch(r, e[r]);
#pragma unroll
        for (int r = 0; r < 4; ++r) finish(r, e[r], w0[r], w1[r]);
        Raw n4 = e[4], n5 = e[5];  // rows 4, 5 in flight

        const int gox = ox0 + lane - 2;
        const bool col_ok = (lane >= 2) && (gox < Wo);
        float* oA = cA + (size_t)z * out_pstride + (size_t)oy0 * out_pitch + gox;
        float* oH = cH + (size_t)z * out_pstride + (size_t)oy0 * out_pitch + gox;
        for (int oy3 = 0; oy3 < AN_TOY; oy3 += 3) {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const int oy = oy3 + k;  // window slot of input row r is r % 6 == (2k + ..) % 6
                const Raw c4 = n4, c5 = n5;
                // prefetch the two rows of the next output row (row indices are reflected / clamped,
                // so the loads past the last output row are harmless and stay unconditional)
                fetch(2 * oy + 6, n4);
                fetch(2 * oy + 7, n5);
                finish(2 * oy + 4, c4, w0[(2 * k + 4) % 6], w1[(2 * k + 4) % 6]);
                finish(2 * oy + 5, c5, w0[(2 * k + 5) % 6], w1[(2 * k + 5) % 6]);
                // axis -2: tap j multiplies input row 2oy + 5 - j
                float a0 = 0.f, a1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int j = 0; j < 6; ++j) {
                    const float v0 = w0[(2 * k + 5 - j + 6) % 6], v1 = w1[(2 * k + 5 - j + 6) % 6];
                    a0 = fmaf(dec_lo(j), v0, a0);
                    a1 = fmaf(dec_lo(j), v1, a1);
                    d0 = fmaf(dec_hi(j), v0, d0);
                    d1 = fmaf(dec_hi(j), v1, d1);
                }
                // axis -1 (low-pass): taps 0,1 on this pair, 2,3 on lane-1, 4,5 on lane-2
                const float a0m1 = __shfl_up_sync(0xffffffffu, a0, 1), a1m1 = __shfl_up_sync(0xffffffffu, a1, 1);
                const float a0m2 = __shfl_up_sync(0xffffffffu, a0, 2), a1m2 = __shfl_up_sync(0xffffffffu, a1, 2);
                const float d0m1 = __shfl_up_sync(0xffffffffu, d0, 1), d1m1 = __shfl_up_sync(0xffffffffu, d1, 1);
                const float d0m2 = __shfl_up_sync(0xffffffffu, d0, 2), d1m2 = __shfl_up_sync(0xffffffffu, d1, 2);
                float ca = dec_lo(0) * a1, ch = dec_lo(0) * d1;
                ca = fmaf(dec_lo(1), a0, ca);
                ca = fmaf(dec_lo(2), a1m1, ca);
                ca = fmaf(dec_lo(3), a0m1, ca);
                ca = fmaf(dec_lo(4), a1m2, ca);
                ca = fmaf(dec_lo(5), a0m2, ca);
                ch = fmaf(dec_lo(1), d0, ch);
                ch = fmaf(dec_lo(2), d1m1, ch);
                ch = fmaf(dec_lo(3), d0m1, ch);
                ch = fmaf(dec_lo(4), d1m2, ch);
                ch = fmaf(dec_lo(5), d0m2, ch);
                const bool ok = col_ok && oy0 + oy < Ho;
                if (ok) {
                    *oA = ca;
                    *oH = ch;
                }
                const float q = __fmul_rn(ch, ch);
                qmin = ok ? fminf(qmin, q) : qmin;
                qmax = ok ? fmaxf(qmax, q) : qmax;
                oA += out_pitch;
                oH += out_pitch;
            }
        }
    }

    // block reductions -> one set of atomics per block
    qmin = warp_min(qmin);
    qmax = warp_max(qmax);
    if (lane == 0) {
        s_red[0][wid] = qmin;
        s_red[1][wid] = qmax;
    }
    if (FIRST && STATS) {
        // ownership mask applied once per lane and column
        const double fs = warp_sum((own0 ? (double)fg_s : 0.0) + (own1 ? (double)fg_s1 : 0.0));
        const double as = warp_sum((own0 ? (double)all_s : 0.0) + (own1 ? (double)all_s1 : 0.0));
        fg_c = __reduce_add_sync(0xffffffffu, (own0 ? fg_c : 0u) + (own1 ? fg_c1 : 0u));
        all_c = __reduce_add_sync(0xffffffffu, all_c * ((own0 ? 1u : 0u) + (own1 ? 1u : 0u)));
        if (lane == 0) {
            s_dred[0][wid] = fs;
            s_dred[1][wid] = as;
            s_cred[0][wid] = fg_c;
            s_cred[1][wid] = all_c;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float mn = s_red[0][0], mx = s_red[1][0];
        for (int w = 1; w < AN_WARPS; ++w) {
            mn = fminf(mn, s_red[0][w]);
            mx = fmaxf(mx, s_red[1][w]);
        }
        LevelStat* st = lstat + (size_t)z * stat_stride;
        if (mn <= mx) {  // at least one valid output in this block
            atomicMax(&st->qmin_inv, ~__float_as_uint(mn));
            atomicMax(&st->qmax_bits, __float_as_uint(mx));
        }
        if (FIRST && STATS) {
            double fs = 0.0, as = 0.0;
            unsigned long long fc = 0, ac = 0;
            for (int w = 0; w < AN_WARPS; ++w) {
                fs += s_dred[0][w];
                as += s_dred[1][w];
                fc += s_cred[0][w];
                ac += s_cred[1][w];
            }
            PlaneStat* ps = pstat + z;
            if (fc) {
                atomicAdd(&ps->fg_sum, fs);
                atomicAdd(&ps->fg_cnt, fc);
            }
            if (ac - fc) {
                atomicAdd(&ps->bg_sum, as - fs);
                atomicAdd(&ps->bg_cnt, ac - fc);
            }
        }
    }
}

// =============================================================================================
// Level-1 analysis with TMA staging.  Same arithmetic and lane layout as analysis_kernel, but the
// input rows are brought into a shared-memory ring by 1-D bulk copies (cp.async.bulk, completion
// on an mbarrier) issued by one thread, AT_STAGES-1 stages ahead of the consumers, so every CTA
// keeps ~15 KB of reads in flight without holding them in registers.  One CTA = 4 warps side by
// side = 120 output columns; it walks `rows_per_cta` output rows, 3 per stage (6 input rows, the
// period of the 6-row window).  Each staged row is an aligned window of AT_WIN input columns that
// contains every (reflected) column the CTA needs; rows are reflected when the copy is issued.
// Requirements (checked on the host): Ws >= AT_WIN, Ws * sizeof(IN_T) % 16 == 0.
// =============================================================================================
constexpr int AT_WARPS = 4;
constexpr int AT_THREADS = 32 * AT_WARPS;
constexpr int AT_OXB = AN_OXW * AT_WARPS;  // 120 output columns per CTA
constexpr int AT_WIN = 256;                // staged input columns per row
constexpr int AT_RS = 6;                   // input rows per stage
constexpr int AT_STAGES = 6;

template <typename IN_T, bool STATS>
__global__ void __launch_bounds__(AT_THREADS)
analysis_tma_kernel(const IN_T* __restrict__ in, int Hs, int Ws, size_t in_pstride, float* __restrict__ cA,
                    float* __restrict__ cH, int Ho, int Wo, int out_pitch, size_t out_pstride,
                    LevelStat* __restrict__ lstat, int stat_stride, PlaneStat* __restrict__ pstat,
                    float fg_thr32, int rows_per_cta) {
    __shared__ __align__(128) IN_T s_ring[AT_STAGES][AT_RS][AT_WIN];
    __shared__ __align__(8) uint64_t s_full[AT_STAGES];
    __shared__ float s_red[2][AT_WARPS];
    __shared__ double s_dred[2][AT_WARPS];
    __shared__ unsigned s_cred[2][AT_WARPS];

    namespace ptx = cuda::ptx;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int z = blockIdx.z;
    const int ox0b = blockIdx.x * AT_OXB;
    const int oy0 = blockIdx.y * rows_per_cta;
    const int R = min(rows_per_cta, Ho - oy0);       // output rows of this CTA (>= 1)
    const int n_in = 2 * R + 4;                      // input rows it consumes
    const int n_stages = (n_in + AT_RS - 1) / AT_RS;
    const IN_T* src = in + (size_t)z * in_pstride;

    // aligned window of staged columns: contains the reflected image of [2 ox0b - 4, 2 ox0b + 243]
    constexpr int ALIGN_EL = 16 / (int)sizeof(IN_T);
    int ws = 2 * ox0b - 8;
    ws = max(0, min(ws, Ws - AT_WIN));
    ws &= ~(ALIGN_EL - 1);

    const int ox0 = ox0b + wid * AN_OXW;
    const int p = ox0 - 2 + lane;  // column pair (2p, 2p+1)
    const int gx0 = 2 * p, gx1 = 2 * p + 1;
    // staged-row offsets of the two (reflected, clamped) columns of this lane
    const int i0 = max(0, min(reflect_fast(gx0, Ws) - ws, AT_WIN - 1));
    const int i1 = max(0, min(reflect_fast(gx1, Ws) - ws, AT_WIN - 1));
    const bool own0 = (lane >= 2) && (gx0 < Ws) && (ox0 < Wo);
    const bool own1 = (lane >= 2) && (gx1 < Ws) && (ox0 < Wo);

    auto issue_stage = [&](int s) {  // one thread: 6 bulk row copies completing on s_full[slot]
        const int slot = s % AT_STAGES;
        constexpr unsigned ROW_BYTES = AT_WIN * sizeof(IN_T);
        ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, &s_full[slot],
                                       AT_RS * ROW_BYTES);
#pragma unroll
        for (int rr = 0; rr < AT_RS; ++rr) {
            const int gy = reflect_fast(2 * oy0 - 4 + s * AT_RS + rr, Hs);
            ptx::cp_async_bulk(ptx::space_shared, ptx::space_global, &s_ring[slot][rr][0],
                               src + (size_t)gy * Ws + ws, ROW_BYTES, &s_full[slot]);
        }
    };

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < AT_STAGES; ++i) ptx::mbarrier_init(&s_full[i], 1);
        ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
    }
    __syncthreads();
    if (tid == 0) {
        for (int s = 0; s < AT_STAGES - 1 && s < n_stages; ++s) issue_stage(s);
    }

    float fg_s = 0.f, all_s = 0.f, fg_s1 = 0.f, all_s1 = 0.f;  // per column of the lane's pixel pair
    unsigned fg_c = 0, fg_c1 = 0, all_c = 0;                  // all_c counts rows: both columns share it
    float qmin = __int_as_float(0x7f800000), qmax = 0.f;
    float w0[6], w1[6];
    const int gox = ox0 + lane - 2;
    const bool col_ok = (lane >= 2) && (gox < Wo) && (ox0 < Wo);
    float* oA = cA + (size_t)z * out_pstride + (size_t)oy0 * out_pitch + gox;
    float* oH = cH + (size_t)z * out_pstride + (size_t)oy0 * out_pitch + gox;

    ptrdiff_t row_off = -2 * (ptrdiff_t)out_pitch;  // output row 3 s + k - 2 of stage s, k = 0
    for (int s = 0; s < n_stages; ++s) {
        __syncthreads();  // every warp is done with stage s-1: its slot may be refilled
        if (tid == 0 && s + AT_STAGES - 1 < n_stages) issue_stage(s + AT_STAGES - 1);
        const int slot = s % AT_STAGES;
        const unsigned parity = (unsigned)(s / AT_STAGES) & 1u;
        while (!ptx::mbarrier_try_wait_parity(&s_full[slot], parity)) {
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            // local input rows r = 6s + 2k, 6s + 2k + 1 enter the window (slot r % 6 == 2k, 2k + 1)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int r = s * AT_RS + 2 * k + h;
                float v0 = to_f32(s_ring[slot][2 * k + h][i0]);
                float v1 = to_f32(s_ring[slot][2 * k + h][i1]);
                if (STATS) {
                    const int gy0 = 2 * oy0 - 4 + r;
                    if (r >= 4 && r < 4 + 2 * R && gy0 < Hs) {  // rows owned by this CTA
                        // per-column accumulators, no ownership selects in the loop: lanes / columns that
                        // belong to a neighbour are masked once at the end
                        all_s += v0;
                        all_s1 += v1;
                        ++all_c;
                        if (v0 >= fg_thr32) {
                            fg_s += v0;
                            ++fg_c;
                        }
                        if (v1 >= fg_thr32) {
                            fg_s1 += v1;
                            ++fg_c1;
                        }
                    }
                }
                w0[2 * k + h] = DSTR_LOGF(__fadd_rn(1.0f, v0));  // np.log(1.0 + x) in float32
                w1[2 * k + h] = DSTR_LOGF(__fadd_rn(1.0f, v1));
            }
            // output row oy = 3s + k - 2 uses local input rows 2oy .. 2oy+5 = the window after this pair
            const int oy = 3 * s + k - 2;
            float a0 = 0.f, a1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
            for (int j = 0; j < 6; ++j) {
                // tap j multiplies local row 2oy + 5 - j = 6s + 2k + 1 - j  -> slot (2k + 1 - j) mod 6
                const float v0 = w0[(2 * k + 1 - j + 12) % 6], v1 = w1[(2 * k + 1 - j + 12) % 6];
                a0 = fmaf(dec_lo(j), v0, a0);
                a1 = fmaf(dec_lo(j), v1, a1);
                d0 = fmaf(dec_hi(j), v0, d0);
                d1 = fmaf(dec_hi(j), v1, d1);
            }
            const float a0m1 = __shfl_up_sync(0xffffffffu, a0, 1), a1m1 = __shfl_up_sync(0xffffffffu, a1, 1);
            const float a0m2 = __shfl_up_sync(0xffffffffu, a0, 2), a1m2 = __shfl_up_sync(0xffffffffu, a1, 2);
            const float d0m1 = __shfl_up_sync(0xffffffffu, d0, 1), d1m1 = __shfl_up_sync(0xffffffffu, d1, 1);
            const float d0m2 = __shfl_up_sync(0xffffffffu, d0, 2), d1m2 = __shfl_up_sync(0xffffffffu, d1, 2);
            float ca = dec_lo(0) * a1, ch = dec_lo(0) * d1;
            ca = fmaf(dec_lo(1), a0, ca);
            ca = fmaf(dec_lo(2), a1m1, ca);
            ca = fmaf(dec_lo(3), a0m1, ca);
            ca = fmaf(dec_lo(4), a1m2, ca);
            ca = fmaf(dec_lo(5), a0m2, ca);
            ch = fmaf(dec_lo(1), d0, ch);
            ch = fmaf(dec_lo(2), d1m1, ch);
            ch = fmaf(dec_lo(3), d0m1, ch);
            ch = fmaf(dec_lo(4), d1m2, ch);
            ch = fmaf(dec_lo(5), d0m2, ch);
            // predicated stores through a running row offset (no branch, no 64-bit multiply per row)
            const bool ok = col_ok && (unsigned)oy < (unsigned)R;
            const ptrdiff_t ro = row_off + k * (ptrdiff_t)out_pitch;
            if (ok) {
                oA[ro] = ca;
                oH[ro] = ch;
            }
            const float q = __fmul_rn(ch, ch);
            qmin = ok ? fminf(qmin, q) : qmin;
            qmax = ok ? fmaxf(qmax, q) : qmax;
        }
        row_off += 3 * (ptrdiff_t)out_pitch;
    }

    // block reductions -> one set of atomics per block
    qmin = warp_min(qmin);
    qmax = warp_max(qmax);
    if (lane == 0) {
        s_red[0][wid] = qmin;
        s_red[1][wid] = qmax;
    }
    if (STATS) {
        // ownership mask applied once per lane and column
        const double fs = warp_sum((own0 ? (double)fg_s : 0.0) + (own1 ? (double)fg_s1 : 0.0));
        const double as = warp_sum((own0 ? (double)all_s : 0.0) + (own1 ? (double)all_s1 : 0.0));
        fg_c = __reduce_add_sync(0xffffffffu, (own0 ? fg_c : 0u) + (own1 ? fg_c1 : 0u));
        all_c = __reduce_add_sync(0xffffffffu, all_c * ((own0 ? 1u : 0u) + (own1 ? 1u : 0u)));
        if (lane == 0) {
            s_dred[0][wid] = fs;
            s_dred[1][wid] = as;
            s_cred[0][wid] = fg_c;
            s_cred[1][wid] = all_c;
        }
    }
    __syncthreads();
    if (tid == 0) {
        float mn = s_red[0][0], mx = s_red[1][0];
        for (int w = 1; w < AT_WARPS; ++w) {
            mn = fminf(mn, s_red[0][w]);
            mx = fmaxf(mx, s_red[1][w]);
        }
        LevelStat* st = lstat + (size_t)z * stat_stride;
        if (mn <= mx) {
            atomicMax(&st->qmin_inv, ~__float_as_uint(mn));
            atomicMax(&st->qmax_bits, __float_as_uint(mx));
        }
        if (STATS) {
            double fs = 0.0, as = 0.0;
            unsigned long long fc = 0, ac = 0;
            for (int w = 0; w < AT_WARPS; ++w) {
                fs += s_dred[0][w];
                as += s_dred[1][w];
                fc += s_cred[0][w];
                ac += s_cred[1][w];
            }
            PlaneStat* ps = pstat + z;
            if (fc) {
                atomicAdd(&ps->fg_sum, fs);
                atomicAdd(&ps->fg_cnt, fc);
            }
            if (ac - fc) {
                atomicAdd(&ps->bg_sum, as - fs);
                atomicAdd(&ps->bg_cnt, ac - fc);
            }
        }
    }
}

// plane statistics only (get_foreground_background_mean, filtering.py:54-88)
template <typename IN_T>
__global__ void __launch_bounds__(256)
plane_stats_kernel(const IN_T* __restrict__ in, int H, int W, size_t pstride,
                   PlaneStat* __restrict__ pstat, float fg_thr32) {
    __shared__ double s_dred[2][8];
    __shared__ unsigned s_cred[2][8];
    const int z = blockIdx.y;
    const IN_T* src = in + (size_t)z * pstride;
    const size_t n = (size_t)H * W;
    double fg_s = 0.0, bg_s = 0.0;
    unsigned fg_c = 0, bg_c = 0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float v = to_f32(src[i]);
        if (v >= fg_thr32) {
            fg_s += (double)v;
            fg_c++;
        } else {
            bg_s += (double)v;
            bg_c++;
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    fg_s = warp_sum(fg_s);
    bg_s = warp_sum(bg_s);
    fg_c = __reduce_add_sync(0xffffffffu, fg_c);
    bg_c = __reduce_add_sync(0xffffffffu, bg_c);
    if (lane == 0) {
        s_dred[0][wid] = fg_s;
        s_dred[1][wid] = bg_s;
        s_cred[0][wid] = fg_c;
        s_cred[1][wid] = bg_c;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double fs = 0.0, bs = 0.0;
        unsigned long long fc = 0, bc = 0;
        for (int w = 0; w < 8; ++w) {
            fs += s_dred[0][w];
            bs += s_dred[1][w];
            fc += s_cred[0][w];
            bc += s_cred[1][w];
        }
        PlaneStat* ps = pstat + z;
        if (fc) {
            atomicAdd(&ps->fg_sum, fs);
            atomicAdd(&ps->fg_cnt, fc);
        }
        if (bc) {
            atomicAdd(&ps->bg_sum, bs);
            atomicAdd(&ps->bg_cnt, bc);
        }
    }
}

// =============================================================================================
// histogram of q = cH^2 with np.histogram(bins=256, range=(min,max)) semantics (numpy 1.26.4):
// float32 edges = float32( float64(i) * float64(step32) + float64(first) ), bin fixed by
// comparisons against the float32 edges, last bin right-closed.
// =============================================================================================
__device__ __forceinline__ float hist_edge(int i, float first, float last, float delta,
                                           float step) {
    if (i >= 256) return last;
    double y = (double)i;
    if (step == 0.f) {
        y = __ddiv_rn(y, 256.0);
        y = __dmul_rn(y, (double)delta);
    } else {
        y = __dmul_rn(y, (double)step);
    }
    y = __dadd_rn(y, (double)first);
    return __double2float_rn(y);
}

// One bin index with np.histogram semantics: arithmetic guess, then the +-1 fix-ups against the
// float32 edges decide (the guess only has to land within one bin of the answer, so a reciprocal
// multiply replaces numpy's division without changing the result).
__device__ __forceinline__ int hist_bin(float v, float first, float inv_width, const float* s_edges) {
    const float q = __fmul_rn(v, v);
    int idx = (int)(__fsub_rn(q, first) * inv_width);
    idx = max(0, min(idx, 255));
    if (q < s_edges[idx]) {
        idx = max(idx - 1, 0);
    } else if (idx != 255 && q >= s_edges[idx + 1]) {
        idx++;
    }
    return idx;
}

__global__ void __launch_bounds__(256)
hist_kernel(const float* __restrict__ cH, int Hl, int Wl, int pitch, size_t pstride,
            LevelStat* __restrict__ lstat, int stat_stride) {
    __shared__ float s_edges[257];
    __shared__ unsigned s_hist[256];
    const int tid = threadIdx.x;
    const int z = blockIdx.y;
    LevelStat* st = lstat + (size_t)z * stat_stride;
    const float first = __uint_as_float(~st->qmin_inv);
    const float last = __uint_as_float(st->qmax_bits);
    if (first == last) return;  // constant band: threshold_otsu returns the value itself
    const float delta = __fsub_rn(last, first);
    const float step = __fdiv_rn(delta, 256.0f);
    for (int i = tid; i < 257; i += 256) s_edges[i] = hist_edge(i, first, last, delta, step);
    s_hist[tid] = 0;
    __syncthreads();

    const float inv_width = 256.0f / delta;
    const float* src = cH + (size_t)z * pstride;
    const int lane = tid & 31;
    // cH^2 is extremely skewed: most samples fall into bin 0, which is counted in a register; the others take one
    // shared-memory atomic each.  (A divergent "bin 0 or general search" form executed the search for nearly every
    // warp anyway: 47 instructions per coefficient against 16 for the straight-line form below.)
    unsigned cnt0 = 0;
    // the band is walked as float4 quads over the padded rows (pitch % 4 == 0, 16-byte aligned)
    const int qpr = pitch >> 2;  // quads per row
    const int nquads = Hl * qpr;
    // (row, quad) advance by a fixed (dr, dq) per step: one division per thread instead of one per quad
    const int stride = gridDim.x * 256;
    const int dr = stride / qpr, dq = stride - dr * qpr;
    int r = (blockIdx.x * 256 + tid) / qpr;
    int cq = (blockIdx.x * 256 + tid) - r * qpr;
    for (int qi = blockIdx.x * 256 + tid; qi < nquads; qi += stride, r += dr, cq += dq) {
        if (cq >= qpr) {
            cq -= qpr;
            ++r;
        }
        const int c = cq << 2;
        const float4 v = *reinterpret_cast<const float4*>(src + (size_t)r * pitch + c);
        const float q[4] = {__fmul_rn(v.x, v.x), __fmul_rn(v.y, v.y), __fmul_rn(v.z, v.z), __fmul_rn(v.w, v.w)};
        // every sample through the bin search, no divergent slow path: guess, one fix-up against the two float32
        // edges of the guessed bin (identical to hist_bin), bin 0 counted in a register, the rest one predicated atomic
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const bool valid = c + k < Wl;
            int idx = (int)(__fsub_rn(q[k], first) * inv_width);
            idx = max(0, min(idx, 255));
            const float elo = s_edges[idx], ehi = s_edges[idx + 1];
            const bool dec = q[k] < elo;
            const bool inc = !dec && idx != 255 && q[k] >= ehi;
            idx = max(idx - (dec ? 1 : 0), 0) + (inc ? 1 : 0);
            cnt0 += (valid && idx == 0) ? 1u : 0u;
            if (valid && idx != 0) atomicAdd(&s_hist[idx], 1u);
        }
    }
    cnt0 = __reduce_add_sync(0xffffffffu, cnt0);
    if (lane == 0 && cnt0) atomicAdd(&s_hist[0], cnt0);
    __syncthreads();
    const unsigned h = s_hist[tid];
    if (h) atomicAdd(&st->hist[tid], h);
}

// =============================================================================================
// Otsu: skimage.filters.threshold_otsu arithmetic in float32 with sequential cumulative sums
// (np.cumsum order), first arg-max, bin centre; then min(max_threshold, sqrt(.)).
// One warp per (plane, level); lane 0 runs the two sequential sweeps.
// =============================================================================================
__global__ void __launch_bounds__(32)
otsu_kernel(LevelStat* __restrict__ lstat_base, size_t level_stride, int stat_stride,
            const PlaneStat* __restrict__ pstat, DispatchParams dp) {
    __shared__ float s_cnt[256], s_ctr[256], s_w1[256], s_m1[256];
    const int z = blockIdx.x;
    const int l = blockIdx.y;
    const int lane = threadIdx.x;
    LevelStat* st = lstat_base + (size_t)l * level_stride + (size_t)z * stat_stride;
    const float first = __uint_as_float(~st->qmin_inv);
    const float last = __uint_as_float(st->qmax_bits);
    const float max_thr = plane_uses_cells(pstat[z], dp) ? dp.max_thr_cells : dp.max_thr_nocells;
    if (dp.notch_only) {
        // nothing is masked: thr only scales the row-filter operands (max |cH|), the mask test never fires
        if (lane == 0) {
            st->otsu_raw = last;
            st->otsu_bin = -1;
            st->thr = __fsqrt_rn(last);
            st->thr_q = __int_as_float(0x7f800000);
        }
        return;
    }

    float otsu;
    int best_i = -1;
    if (first == last) {
        otsu = first;
    } else {
        const float delta = __fsub_rn(last, first);
        const float step = __fdiv_rn(delta, 256.0f);
        for (int i = lane; i < 256; i += 32) {
            const float e0 = hist_edge(i, first, last, delta, step);
            const float e1 = hist_edge(i + 1, first, last, delta, step);
            s_ctr[i] = __fdiv_rn(__fadd_rn(e0, e1), 2.0f);
            s_cnt[i] = (float)st->hist[i];
        }
        __syncwarp();
        if (lane == 0) {
            float w = 0.f, cs = 0.f;
            for (int i = 0; i < 256; ++i) {
                w = __fadd_rn(w, s_cnt[i]);
                cs = __fadd_rn(cs, __fmul_rn(s_cnt[i], s_ctr[i]));
                s_w1[i] = w;
                s_m1[i] = __fdiv_rn(cs, w);
            }
            float w2 = 0.f, cs2 = 0.f, best = -1.f;
            best_i = 0;
            for (int i = 255; i >= 1; --i) {
                w2 = __fadd_rn(w2, s_cnt[i]);
                cs2 = __fadd_rn(cs2, __fmul_rn(s_cnt[i], s_ctr[i]));
                const float m2 = __fdiv_rn(cs2, w2);
                const float d = __fsub_rn(s_m1[i - 1], m2);
                const float v = __fmul_rn(__fmul_rn(s_w1[i - 1], w2), __fmul_rn(d, d));
                if (v >= best) {  // descending scan + '>=' keeps the FIRST maximum
                    best = v;
                    best_i = i - 1;
                }
            }
        }
        best_i = __shfl_sync(0xffffffffu, best_i, 0);
        otsu = s_ctr[best_i];
    }
    if (lane == 0) {
        const float sq = __fsqrt_rn(otsu);
        st->otsu_raw = otsu;
        st->otsu_bin = best_i;
        const float thr = (sq < max_thr) ? sq : max_thr;  // python min(max_threshold, sqrt)
        st->thr = thr;
        // the mask test sqrt_rn(c*c) > thr (filtering.py:187-195) as a test on q = c*c: sqrt_rn is
        // monotone, so it equals q > T with T the largest float whose rounded root is <= thr
        float T = __fmul_rn(thr, thr);
        if (thr >= 0.f && thr < __int_as_float(0x7f800000)) {
            for (int it = 0; it < 8 && __fsqrt_rn(T) > thr; ++it) T = __uint_as_float(__float_as_uint(T) - 1u);
            for (int it = 0; it < 8; ++it) {
                const float Tn = __uint_as_float(__float_as_uint(T) + 1u);
                if (!(__fsqrt_rn(Tn) <= thr)) break;
                T = Tn;
            }
        }
        st->thr_q = T;
    }
}

// =============================================================================================
// Row filter.  For every row of cH_l (filtering.py:195-217):
//   m = sqrt(c*c) > thr;  bg = m ? 0 : c;  med = median(bg);  x = m ? med : c
//   bgf = irfft(rfft(x) * g) = x - B x      (g on the PACKED rfft index, scipy.fftpack layout)
//   cH' = m ? c : bgf   =>   dH = cH' - c = m ? 0 : -(B x)[t]
//
// B in the time domain.  With x_e / x_o the circularly even / odd parts of x about index 0,
//   B x = A x_e + Bo x_o,  A: cosine multipliers a_j = w(2j-1) (a_0 = 1),  Bo: sine multipliers
//   b_j = w(2j),  w(k) = exp(-k^2 / 2 s^2).
// Bo is a plain periodic Gaussian: a compact FIR `to`.  a_j has a kink at j = 0 (|2j| - 1), so
// its kernel has 1/u^2 tails; it is split on the host (double precision) into a smooth part G
// (compact FIR `te`) plus a remainder supported on the J lowest cosine modes, applied as a
// rank-J correction:  c_j = sum_v T1[v][j] x_e[v],  y_e[t] += sum_j c_j T2[j][t].
// Small bands use the dense kernels (J = 0, te / to of full circular length).  Both forms are
// evaluated on the half range t = 0..n/2 and mirrored:  y[t] = y_e + y_o,  y[n-t] = y_e - y_o.
//
// One warp per row for selection / in-painting / the c_j; the whole block for the register-tiled
// FIR.  FIR operands are stored with one pad word per 8 (phys = a + (a >> 3)) so that lanes whose
// 8-output windows are 8 apart hit distinct banks.
// =============================================================================================
constexpr int FR_ROWS = 4;
constexpr int FR_THREADS = 32 * FR_ROWS;

__device__ __forceinline__ unsigned f2key(float f) {
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    const unsigned b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(b);
}

#ifndef DSTR_FIR_UNROLL
#define DSTR_FIR_UNROLL 1
#endif
#define DSTR_PRAGMA(x) _Pragma(#x)
#define DSTR_UNROLL(n) DSTR_PRAGMA(unroll n)
// acc[i] += sum_k taps[k] * Xlog[8*m0 + i - k],  k = 0..ntap-1 (ntap % 8 == 0).
// Logical element a = 8 b + i lives at Xphys[BS * b + ES * i]  (padded rows: ES 1, BS 9).
template <int ES, int BS>
__device__ __forceinline__ void fir8(float (&acc)[8], const float* __restrict__ Xphys,
                                     const float* __restrict__ taps, int ntap, int m0) {
    const float* Xp = Xphys + BS * m0;
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = Xp[ES * i];
    const float4* t4 = reinterpret_cast<const float4*>(taps);
    DSTR_UNROLL(DSTR_FIR_UNROLL)
    for (int kk = 0; kk < ntap / 8; ++kk) {
        const float4 ta = t4[2 * kk], tb = t4[2 * kk + 1];
        const float tk[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
        Xp -= BS;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(tk[k], w[(i - k) & 7], acc[i]);
            w[7 - k] = Xp[ES * (7 - k)];
        }
    }
}

// ---- rank-J projection on the tensor path ----------------------------------------------------
// c[j][r] = sum_v T1[v][j] x_e[r][v] is a (Jpad x nv) x (nv x rows) product: legacy
// mma.sync m16n8k8 (HMMA.1688.F32.TF32; 278 TFLOP/s measured on this part, tools/probes) with the
// 3xTF32 split a = a_hi + a_lo (a_hi = the 19 leading bits, a_lo = a - a_hi exactly):
// a b ~ a_hi b_hi + a_lo b_hi + a_hi b_lo, i.e. float32-class products (the dropped a_lo b_lo term
// is 2^-22 relative) accumulated in float32.
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct NotchTables {
    const float* te;  // even-part FIR taps [ntap_e]; tap k <-> circular offset u = ue_lo + k
    const float* to;  // odd-part FIR taps  [ntap_o]; tap k <-> u = uo_lo + k
    const float* T1;  // omega_v cos(2 pi j v / n) in mma A-fragment order: [E block][Jpad / 16][lane][4]
    const float* T2;  // [J][nhp64]    rho_j   cos(2 pi j t / n)   (j-major; inside a row the two float4
                      //               halves of 8 consecutive 8-output segments are grouped: see t2_offset)
    int ntap_e, ue_lo, ntap_o, uo_lo, J, Jpad;
};

struct FilterLevelArgs {
    float* cH;
    int Hl, Wl, pitch;
    size_t pstride;
    const LevelStat* lstat;
    int stat_stride;
    NotchTables nt[2];  // [0] no_cells, [1] cells
    int nh;             // n / 2: the half range is t = 0..nh
    int nhp8;           // nh + 1 rounded up to a multiple of 8
    int n_pad8;
    int nhp64;          // T2 row stride: nhp8 rounded up to a multiple of 64
    int xlen_e_phys, xlen_o_phys;  // per-row physical floats (max over the two configs), = 8 mod 16
    int ntap_e_max, ntap_o_max, Jpad_max;
    int vec_ok;           // rows are 16-byte aligned: float4 stores allowed
    int prefetch_blocks;  // rows of the block this many blocks ahead are prefetched into L2 (0 = off)
    int ablate;  // DSTR_ABLATION builds only (timing experiments): 1 median, 2 projection, 4 expansion, 8 even FIR, 16 odd FIR
};

template <int EPL>
// 8 blocks per SM (64 registers) for rows up to 33 x 32 elements; wider rows keep more keys in
// registers and run 4 blocks per SM
#ifndef DSTR_FR_MINB
#define DSTR_FR_MINB 8
#endif
#ifndef DSTR_BUILD_UNROLL
#define DSTR_BUILD_UNROLL 1
#endif
#ifndef DSTR_LR1_UNROLL
#define DSTR_LR1_UNROLL 2
#endif
#ifndef DSTR_LR2_UNROLL
#define DSTR_LR2_UNROLL 2
#endif
__global__ void __launch_bounds__(FR_THREADS, (EPL <= 33 ? DSTR_FR_MINB : 4))
filter_rows_kernel(FilterLevelArgs a, const PlaneStat* __restrict__ pstat, DispatchParams dp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.Wl;
    float* s_te = reinterpret_cast<float*>(smem_raw);   // [ntap_e_max]
    float* s_to = s_te + a.ntap_e_max;                  // [ntap_o_max]
    float* s_E = s_to + a.ntap_o_max;                   // [FR_ROWS][xlen_e_phys]
    float* s_O = s_E + FR_ROWS * a.xlen_e_phys;         // [FR_ROWS][xlen_o_phys]
    // rank-J coefficients: accumulated as 64-bit fixed point (deterministic, order-free atomics),
    // then converted to float
    unsigned long long* s_c64 = reinterpret_cast<unsigned long long*>(s_O + FR_ROWS * a.xlen_o_phys);  // [FR_ROWS][Jpad_max]
    float* s_c = reinterpret_cast<float*>(s_c64 + max(FR_ROWS * a.Jpad_max, 128));                      // [FR_ROWS][Jpad_max + 8]
    const int cstride = a.Jpad_max + 8;  // rows 8 banks apart: the 4 rows' c_j are read in one wavefront
    unsigned* s_mask = reinterpret_cast<unsigned*>(s_c + FR_ROWS * cstride);                            // [FR_ROWS][EPL] mask bits

    const int tid = threadIdx.x, lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction
#ifdef DSTR_ABLATION
    const int abl = a.ablate;
#else
    constexpr int abl = 0;
#endif
    const int z = blockIdx.y;
    const int row0 = blockIdx.x * FR_ROWS;
    const int nrows = min(FR_ROWS, a.Hl - row0);
    // The blocks resident on the GPU hold ~19 MB of rows; asking L2 for the rows of the block that
    // will run in this slot one wave later turns the DRAM round trips of the load phase into L2 hits.
    if (a.prefetch_blocks > 0 && lane == 0) {
        const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + a.prefetch_blocks;
        const int pz = (int)(lin / gridDim.x);
        const int prow = (int)(lin - (long long)pz * gridDim.x) * FR_ROWS + wid;
        if (pz < (int)gridDim.y && prow < a.Hl) {
            const float* pp = a.cH + (size_t)pz * a.pstride + (size_t)prow * a.pitch;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pp), "r"(a.pitch * 4) : "memory");
        }
    }
    const int cfg = plane_uses_cells(pstat[z], dp);
    const NotchTables nt = cfg ? a.nt[1] : a.nt[0];
    // mask rule sqrt(c*c) > thr evaluated as c*c > thr_q (bit-identical, see otsu_kernel)
    const float thr_q = a.lstat[(size_t)z * a.stat_stride].thr_q;
    const int nh = a.nh;

    for (int i = tid; i < nt.ntap_e; i += FR_THREADS) s_te[i] = nt.te[i];
    for (int i = tid; i < nt.ntap_o; i += FR_THREADS) s_to[i] = nt.to[i];
    for (int i = tid; i < FR_ROWS * a.Jpad_max; i += FR_THREADS) s_c64[i] = 0ull;

    float* E = s_E + wid * a.xlen_e_phys;
    float* O = s_O + wid * a.xlen_o_phys;
    if (wid < nrows) {
        const float* grow = a.cH + (size_t)z * a.pstride + (size_t)(row0 + wid) * a.pitch;
        // ---- load, mask, keys ------------------------------------------------------------
        // every load of the row is issued before the first use (one memory round trip per row instead
        // of one per few elements): unconditional loads at immediate offsets from one base; the lanes
        // past the end of the row read into the next row / the slack behind the buffer and are discarded
        unsigned key[EPL];
        {
            const float* gl = grow + lane;
#pragma unroll
            for (int i = 0; i < EPL; ++i) key[i] = __float_as_uint(gl[32 * i]);
        }
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const int e = lane + 32 * i;
            const float c = __uint_as_float(key[i]);
            key[i] = 0xffffffffu;
            bool m = false;
            if (e < n) {
                m = __fmul_rn(c, c) > thr_q;
                key[i] = f2key(m ? 0.0f : (c + 0.0f));  // zero-filled background, canonical +0
            }
            // mask bits of elements 32i .. 32i+31, kept for the output stage
            const unsigned mbits = __ballot_sync(0xffffffffu, m);
            if (lane == 0) s_mask[wid * EPL + i] = mbits;
        }
        // ---- exact median of the zero-filled background (np.median, filtering.py:201) -------
        // Order statistics k1 = (n-1)/2 and k2 = n/2 of the keys.  The masked entries are exact
        // zeros sitting in the middle of a roughly symmetric distribution, so the median is very
        // often 0: test that first.  Otherwise bisect the key bits on the side that holds k1 and
        // stop as soon as the bracket isolates a single key.
        const unsigned KZ = 0x80000000u;  // f2key(+0.0f)
        const int k1 = (n - 1) >> 1, k2 = n >> 1;
        int cneg = 0, cle0 = 0;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            cneg += (key[i] < KZ) ? 1 : 0;
            cle0 += (key[i] <= KZ) ? 1 : 0;
        }
        cneg = __reduce_add_sync(0xffffffffu, cneg);
        cle0 = __reduce_add_sync(0xffffffffu, cle0);
        float med;
        if ((cneg <= k1 && k2 < cle0) || (abl & 1)) {
            med = 0.f;
        } else {
            unsigned res;
            int lo_cnt, hi_cnt;
            if (k1 < cneg) {  // negative side: keys in [0, 2^31)
                res = 0u;
                lo_cnt = 0;
                hi_cnt = cneg;
            } else {  // zero / positive side: keys in [2^31, 2^32)
                res = KZ;
                lo_cnt = cneg;
                hi_cnt = n;
            }
            bool unique = (hi_cnt - lo_cnt) == 1;
            for (int b = 30; b >= 0 && !unique; --b) {
                const unsigned trial = res | (1u << b);
                int cnt = 0;
#pragma unroll
                for (int i = 0; i < EPL; ++i) cnt += (key[i] < trial) ? 1 : 0;
                cnt = __reduce_add_sync(0xffffffffu, cnt);
                if (cnt <= k1) {
                    res = trial;
                    lo_cnt = cnt;
                } else {
                    hi_cnt = cnt;
                }
                unique = (hi_cnt - lo_cnt) == 1;
            }
            // res is now either the key itself (all bits decided) or a lower bound of the single
            // key left in the bracket: the smallest key >= res is the k1-th order statistic
            unsigned kk1 = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < EPL; ++i)
                if (key[i] >= res) kk1 = min(kk1, key[i]);
            kk1 = __reduce_min_sync(0xffffffffu, kk1);
            med = key2f(kk1);
            if (k2 != k1) {
                int cle = 0;
                unsigned nxt = 0xffffffffu;
#pragma unroll
                for (int i = 0; i < EPL; ++i) {
                    cle += (key[i] <= kk1) ? 1 : 0;
                    if (key[i] > kk1) nxt = min(nxt, key[i]);
                }
                cle = __reduce_add_sync(0xffffffffu, cle);
                nxt = __reduce_min_sync(0xffffffffu, nxt);
                const unsigned kk2 = (cle >= k1 + 2) ? kk1 : nxt;
                med = (key2f(kk1) + key2f(kk2)) * 0.5f;
            }
        }
        // ---- in-painted row x[t] = m ? med : c, split into circular even / odd parts and stored
        //      circularly extended for the FIRs:  E[tau + OFFe] = x_e[tau mod n], same for O.
        //      The row is re-read from global memory (L1-resident) instead of being staged.
        {
            const int OFFe = nt.ue_lo + nt.ntap_e, OFFo = nt.uo_lo + nt.ntap_o;
            const int len_e = a.nhp8 + nt.ntap_e, len_o = a.nhp8 + nt.ntap_o;
            const int tau_lo = -max(OFFe, OFFo);
            const int tau_hi = max(len_e - OFFe, len_o - OFFo);
            int t = (tau_lo + lane) % n;
            if (t < 0) t += n;
            const int step = 32 % n;
            // the row pointer is materialised once (one wide multiply-add per load instead of a 64-bit
            // index chain); stores are predicated on clamped indices, no branches in the loop
            const float* gp = grow;
            asm volatile("" : "+l"(gp));
            int ae = tau_lo + lane + OFFe, ao = tau_lo + lane + OFFo;
            DSTR_UNROLL(DSTR_BUILD_UNROLL)
            for (int tau = tau_lo + lane; tau < tau_hi && !(abl & 64); tau += 32, ae += 32, ao += 32) {
                const unsigned tr = (t == 0) ? 0u : (unsigned)(n - t);
                const float c1 = gp[(unsigned)t], c2 = gp[tr];
                const float x1 = (__fmul_rn(c1, c1) > thr_q) ? med : c1;
                const float x2 = (__fmul_rn(c2, c2) > thr_q) ? med : c2;
                const bool pe = (unsigned)ae < (unsigned)len_e, po = (unsigned)ao < (unsigned)len_o;
                const int ie = pe ? ae + (ae >> 3) : 0, io = po ? ao + (ao >> 3) : 0;
                const float ve = 0.5f * (x1 + x2), vo = 0.5f * (x1 - x2);
                if (pe) E[ie] = ve;
                if (po) O[io] = vo;
                t += step;
                if (t >= n) t -= n;
            }
        }
    }
    __syncthreads();  // every row's E / O is complete

    // ---- rank-J correction coefficients  c_j = sum_v T1[v][j] x_e[v], all rows at once ----------
    // x_e[v] sits at logical index v + OFFe of the padded E rows.  The index range is walked in the
    // 8-element blocks of the padded layout = the k-steps of the mma; warp w takes the w-th quarter
    // of the blocks.  A = T1 (16 modes x 8 elements per tile, fragment-ordered on the host, zero
    // outside 0 <= v <= nh and j < J so the partial first / last blocks need no predicates),
    // B = x_e (8 elements x 8 columns; columns 0..3 = the rows of the block, 4..7 zero).
    if (nt.J > 0 && !(abl & 2)) {
        const int OFFe = nt.ue_lo + nt.ntap_e;
        const int Jpad = nt.Jpad;
        const int nv = nh + 1;
        const int blk_lo = OFFe >> 3, blk_hi = (OFFe + nv + 7) >> 3;
        const int bw = (blk_hi - blk_lo + FR_ROWS - 1) / FR_ROWS;
        const int b_begin = blk_lo + wid * bw, b_end = min(blk_hi, b_begin + bw);
        const int g = lane >> 2, tig = lane & 3;
        const int mtiles = Jpad >> 4;
        for (int j0 = 0; j0 < Jpad; j0 += 32) {  // two 16-mode tiles per pass
            float acc[2][4];
#pragma unroll
            for (int m = 0; m < 2; ++m) acc[m][0] = acc[m][1] = acc[m][2] = acc[m][3] = 0.f;
            const float4* tf = reinterpret_cast<const float4*>(nt.T1) +
                               ((size_t)(b_begin - blk_lo) * mtiles + (j0 >> 4)) * 32 + lane;
            const float* eb = s_E + (g & 3) * a.xlen_e_phys + 9 * b_begin + tig;  // rows >= nrows: unused garbage
            DSTR_UNROLL(DSTR_LR1_UNROLL)
            for (int blk = b_begin; blk < b_end; ++blk) {
                const float x0 = (g < FR_ROWS) ? eb[0] : 0.f;
                const float x1 = (g < FR_ROWS) ? eb[4] : 0.f;
                unsigned bh0, bl0, bh1, bl1;
                split_tf32(x0, bh0, bl0);
                split_tf32(x1, bh1, bl1);
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const float4 t = __ldg(tf + m * 32);
                    unsigned ah[4], al[4];
                    split_tf32(t.x, ah[0], al[0]);
                    split_tf32(t.y, ah[1], al[1]);
                    split_tf32(t.z, ah[2], al[2]);
                    split_tf32(t.w, ah[3], al[3]);
                    mma_tf32(acc[m], al, bh0, bh1);
                    mma_tf32(acc[m], ah, bl0, bl1);
                    mma_tf32(acc[m], ah, bh0, bh1);
                }
                tf += mtiles * 32;
                eb += 9;
            }
            // each warp holds the partial sums of its range of v: combine in shared memory as
            // 2^-32 fixed point (|c_j| < 2^30 by far; integer adds commute, so the result does
            // not depend on the order in which the warps arrive)
            // accumulator fragment: acc[m][0..1] = mode 16 m + g, columns 2 tig, 2 tig + 1; acc[m][2..3] = mode + 8
            if (tig < FR_ROWS / 2) {
#pragma unroll
                for (int m = 0; m < 2; ++m)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int r = 2 * tig + (q & 1);
                        const int j = j0 + 16 * m + g + ((q & 2) ? 8 : 0);
                        atomicAdd(s_c64 + r * a.Jpad_max + j, (unsigned long long)__float2ll_rn(acc[m][q] * 4294967296.0f));
                    }
            }
        }
        __syncthreads();
        for (int i = tid; i < FR_ROWS * a.Jpad_max; i += FR_THREADS) {
            const int r = i / a.Jpad_max;
            s_c[r * cstride + (i - r * a.Jpad_max)] = __ll2float_rn((long long)s_c64[i]) * (1.0f / 4294967296.0f);
        }
        __syncthreads();
    }

    // ---- register-tiled FIRs + rank-J correction over (8-output segment, row) pairs -----------
    // The row index runs fastest: a warp covers 8 segments x FR_ROWS rows, so its T2 loads touch
    // 8 x 16 contiguous bytes (broadcast over the rows) and the rows' E / O windows sit in
    // different banks (row strides = 8 mod 16 words, segment stride 9 words).
    const int nseg = a.nhp8 >> 3;
    const int total = FR_ROWS * nseg;
    float* stage = reinterpret_cast<float*>(s_c64) + wid * 64;  // free after the c_j conversion; >= 1 KB
    for (int w0 = wid * 32; w0 < total; w0 += FR_THREADS) {  // warp-uniform: the stores are warp-collective
        const int w = w0 + lane;
        const int seg = w / FR_ROWS;
        const int r = w - seg * FR_ROWS;
        const bool active = (w < total) && (r < nrows);
        float ye[8], yo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ye[i] = yo[i] = 0.f;
        if (active) {
            if (!(abl & 8)) fir8<1, 9>(ye, s_E + r * a.xlen_e_phys, s_te, nt.ntap_e, seg + (nt.ntap_e >> 3));
            if (!(abl & 16)) fir8<1, 9>(yo, s_O + r * a.xlen_o_phys, s_to, nt.ntap_o, seg + (nt.ntap_o >> 3));
            if (nt.J > 0 && !(abl & 4)) {
                const float* cp = s_c + r * cstride;
                // outputs 8 seg .. 8 seg + 3 at float4 index 16 (seg / 8) + seg % 8, the next four 8 further
                const float4* t2 = reinterpret_cast<const float4*>(nt.T2) + 16 * (seg >> 3) + (seg & 7);
                const int stride4 = a.nhp64 >> 2;
                DSTR_UNROLL(DSTR_LR2_UNROLL)
                for (int j = 0; j < nt.J; ++j) {
                    const float c = cp[j];
                    const float4 u0 = __ldg(t2);
                    const float4 u1 = __ldg(t2 + 8);
                    t2 += stride4;
                    ye[0] = fmaf(c, u0.x, ye[0]);
                    ye[1] = fmaf(c, u0.y, ye[1]);
                    ye[2] = fmaf(c, u0.z, ye[2]);
                    ye[3] = fmaf(c, u0.w, ye[3]);
                    ye[4] = fmaf(c, u1.x, ye[4]);
                    ye[5] = fmaf(c, u1.y, ye[5]);
                    ye[6] = fmaf(c, u1.z, ye[6]);
                    ye[7] = fmaf(c, u1.w, ye[7]);
                }
            }
        }
        if (abl & 32) continue;
        // ---- dH[t] = masked ? 0 : -(B x)[t], t = 8 seg + i (direct half) and n - t (mirrored half);
        //      mask bits come from the selection phase -------------------------------------------
        const int t0 = 8 * seg;
        float vd[8], vm[8];
        if (active) {
            const unsigned* mrow = s_mask + r * EPL;
            const unsigned dbits = mrow[t0 >> 5] >> (t0 & 31);  // bit i <-> t0 + i (t0 % 8 == 0: one word)
            // bits of n - t0 - 7 .. n - t0 (bit 7 - i <-> n - t0 - i); entries below 0 are never stored
            const int lo = n - t0 - 7, loc = max(lo, 0);
            const int wi = loc >> 5;
            const unsigned mbits = __funnelshift_r(mrow[wi], mrow[min(wi + 1, EPL - 1)], loc & 31) << (loc - lo);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                vd[i] = ((dbits >> i) & 1u) ? 0.f : -(ye[i] + yo[i]);
                vm[i] = ((mbits >> (7 - i)) & 1u) ? 0.f : -(ye[i] - yo[i]);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) vd[i] = vm[i] = 0.f;
        }
        float* orow = a.cH + (size_t)z * a.pstride + (size_t)(row0 + r) * a.pitch;
        // direct half: lanes l and l ^ 4 hold neighbouring segments of one row; they swap half a segment
        // so that each 16-byte store of the pair completes a 32-byte sector
        {
            const bool fast = active && a.vec_ok && (t0 + 7 <= nh);
            const int partner_fast = __shfl_xor_sync(0xffffffffu, (int)fast, 4);
            const bool pair = fast && partner_fast;
            const bool odd = (lane >> 2) & 1;
            float rcv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) rcv[k] = __shfl_xor_sync(0xffffffffu, odd ? vd[k] : vd[4 + k], 4);
            if (pair) {
                float4 q1, q2;
                if (!odd) {
                    q1 = make_float4(vd[0], vd[1], vd[2], vd[3]);
                    q2 = make_float4(rcv[0], rcv[1], rcv[2], rcv[3]);
                } else {
                    q1 = make_float4(rcv[0], rcv[1], rcv[2], rcv[3]);
                    q2 = make_float4(vd[4], vd[5], vd[6], vd[7]);
                }
                float* p1 = orow + (odd ? t0 - 4 : t0);      // even: own 0..3          odd: partner's 4..7
                float* p2 = orow + (odd ? t0 + 4 : t0 + 8);  // even: partner's 0..3    odd: own 4..7
                *reinterpret_cast<float4*>(p1) = q1;
                *reinterpret_cast<float4*>(p2) = q2;
            } else if (active) {
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    if (t0 + i <= nh) orow[t0 + i] = vd[i];
            }
        }
        // mirrored half: row by row through a 64-float staging line, written out as 2 x 128 contiguous bytes
        {
            const int s = lane >> 2;
            const int tm_lo = n - (w0 / FR_ROWS) * 8 - 63;  // lowest address of the warp's 8 segments
#pragma unroll
            for (int rr = 0; rr < FR_ROWS; ++rr) {
                if (active && r == rr) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) stage[8 * (7 - s) + (7 - i)] = vm[i];
                }
                __syncwarp();
                if (rr < nrows) {
                    float* orr = a.cH + (size_t)z * a.pstride + (size_t)(row0 + rr) * a.pitch;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int p = lane + 32 * h;
                        const int tm = tm_lo + p;
                        // n - tm = t in 1..nh with tm != t, inside the segments that exist
                        if (tm > nh && tm < n) orr[tm] = stage[p];
                    }
                }
                __syncwarp();
            }
        }
    }
}

// =============================================================================================
// synthesis of the deltas: out = idwt2(dA, (dH, 0, 0)) (axis -1 first with rec_lo for both
// bands, then axis -2 with rec_lo on the dA branch and rec_hi on the dH branch), trimmed to the
// parent level's shape (pywt.waverec2).  FINAL fuses the inverse log and the epilogue.
//
// Register-tiled: one warp owns 64 output columns x SY_TY output rows.  Lane l owns output
// columns (2m, 2m+1), m = x0/2 + l, which need coefficient columns m, m+1, m+2 (three coalesced,
// overlapping loads per band and coefficient row) and marches down with a 3-row window (row
// loop unrolled by 3).
// =============================================================================================
constexpr int SY_TX = 64;
#ifndef DSTR_SY_TY
#define DSTR_SY_TY 48
#endif
constexpr int SY_TY = DSTR_SY_TY;  // multiple of 6
#ifndef DSTR_SY_WARPS
#define DSTR_SY_WARPS 4
#endif
constexpr int SY_WARPS = DSTR_SY_WARPS;
constexpr int SY_THREADS = 32 * SY_WARPS;

struct EpilogueArgs {
    const float* inv_flat;  // nullable; 1 / flatfield
    const float* dark;      // nullable
    int shadow;             // apply dark/flat
    int expm1;              // exp(y) - 1 instead of exp(y) + 1
};

// VEC (FINAL only): Wo and the plane stride are even, so the two pixels of a lane are one aligned
// word of the image, of the output and of the dark / flat fields.
// One warp tile of the synthesis.  INTERIOR: every coefficient / pixel the tile touches lies inside the
// band and the image, so all clamps and boundary predicates fold away (95 % of the tiles of a 2048^2 plane).
template <bool FINAL, typename IN_T, typename OUT_T, bool VEC, bool INTERIOR>
__device__ __forceinline__ void synth_tile(const float* __restrict__ dA, const float* __restrict__ dH, int Hl, int Wl, int pitch_l,
             size_t pstride_l, float* __restrict__ outA, int Ho, int Wo, int pitch_o,
             size_t pstride_o, const IN_T* __restrict__ img, OUT_T* __restrict__ out,
             size_t img_pstride, EpilogueArgs ep, int x0, int y0) {
    const int lane = threadIdx.x & 31;
    const int z = blockIdx.z;
    const int m = (x0 >> 1) + lane;
    const int cy0 = y0 >> 1;
    // clamped coefficient columns: out-of-range columns only feed discarded outputs or are
    // multiplied by the zero that replaces them below
    const bool c0 = INTERIOR || m < Wl, c1 = INTERIOR || m + 1 < Wl, c2 = INTERIOR || m + 2 < Wl;
    const int mc = INTERIOR ? m : min(m, Wl - 1);
    const float* pA = dA ? dA + (size_t)z * pstride_l + mc : nullptr;
    const float* pH = dH ? dH + (size_t)z * pstride_l + mc : nullptr;
    const int o1 = c1 ? 1 : 0, o2 = c2 ? 2 : 0;

    // fetch: issue the six coefficient loads of coefficient row r (row r+1 is fetched before row r
    // is consumed: software pipelining of the in-order warp); xpass: axis -1 synthesis of that row
    // -> (L0, L1) from dA, (G0, G1) from dH for output columns 2m, 2m+1.  All loads are
    // unconditional on clamped addresses (the level buffers carry a few floats of slack), values
    // outside the band are replaced by zero afterwards, so the compiler can hoist every load.
    auto fetch = [&](int r, float (&c)[6]) {
        const int gy = INTERIOR ? cy0 + r : min(cy0 + r, Hl - 1);
        const unsigned o = (unsigned)(gy * pitch_l);
        if (pA) {
            const float* q = pA + o;
            c[0] = q[0];
            c[1] = q[1];
            c[2] = q[2];
        } else {
            c[0] = c[1] = c[2] = 0.f;
        }
        if (pH) {
            const float* q = pH + o;
            c[3] = q[0];
            c[4] = q[1];
            c[5] = q[2];
        } else {
            c[3] = c[4] = c[5] = 0.f;
        }
    };
    auto xpass = [&](int r, const float (&c)[6], float& L0, float& L1, float& G0, float& G1) {
        const bool rok = INTERIOR || (cy0 + r) < Hl;
        const float a0 = (rok && c0) ? c[0] : 0.f, a1 = (rok && c1) ? c[1] : 0.f, a2 = (rok && c2) ? c[2] : 0.f;
        const float h0 = (rok && c0) ? c[3] : 0.f, h1 = (rok && c1) ? c[4] : 0.f, h2 = (rok && c2) ? c[5] : 0.f;
        // x = 2m + px: sum_j rec_lo[2j + px] * c[m + 2 - j]
        L0 = fmaf(rec_lo(0), a2, fmaf(rec_lo(2), a1, rec_lo(4) * a0));
        L1 = fmaf(rec_lo(1), a2, fmaf(rec_lo(3), a1, rec_lo(5) * a0));
        G0 = fmaf(rec_lo(0), h2, fmaf(rec_lo(2), h1, rec_lo(4) * h0));
        G1 = fmaf(rec_lo(1), h2, fmaf(rec_lo(3), h1, rec_lo(5) * h0));
    };

    float L0[3], L1[3], G0[3], G1[3];
    float cn[6], cn2[6];  // coefficient rows in flight (two steps ahead: more bytes in flight per warp)
    {
        float ca[6], cb[6];
        fetch(0, ca);
        fetch(1, cb);
        fetch(2, cn);
        fetch(3, cn2);
        xpass(0, ca, L0[0], L1[0], G0[0], G1[0]);
        xpass(1, cb, L0[1], L1[1], G0[1], G1[1]);
    }
    const int gx = 2 * m;
    const bool v0ok = INTERIOR || gx < Wo, v1ok = INTERIOR || gx + 1 < Wo;
    const float one = ep.expm1 ? -1.0f : 1.0f;
    typedef RawPair<IN_T, VEC> Raw;
    const int gxc = INTERIOR ? gx : min(gx, Wo - (VEC ? 2 : 1));  // clamped columns (loads only)
    const int gx1c = INTERIOR ? gx + 1 : min(gx + 1, Wo - 1);
    const IN_T* img_z = FINAL ? img + (size_t)z * img_pstride : nullptr;
    OUT_T* out_z = FINAL ? out + (size_t)z * img_pstride : nullptr;
    float* outA_z = FINAL ? nullptr : outA + (size_t)z * pstride_o + gx;
    // raw image pixels one step ahead of their use (FINAL): the synthesis arithmetic of a step is too short to cover
    // the DRAM latency of loads issued inside it (ncu: 59 % of the stall samples sat on their first use)
    Raw px_next[2];
    auto fetch_px = [&](int my_) {
#pragma unroll
        for (int py = 0; py < 2; ++py) {
            const int gyc = min(y0 + 2 * my_ + py, Ho - 1);
            const IN_T* rowp = img_z + (unsigned)(gyc * Wo);
            px_next[py].load(rowp + (VEC ? gxc : 0), VEC ? 0 : gxc, gx1c);
        }
    };
    if (FINAL) fetch_px(0);
    for (int my3 = 0; my3 < SY_TY / 2; my3 += 3) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int my = my3 + k;
            float cc[6];
#pragma unroll
            for (int i = 0; i < 6; ++i) {
                cc[i] = cn[i];
                cn[i] = cn2[i];
            }
            fetch(my + 4, cn2);
            // raw image pixels and dark / flat of the two output rows of this step (FINAL)
            Raw px[2];
            float dk[2][2] = {{0.f, 0.f}, {0.f, 0.f}}, ifl[2][2] = {{1.f, 1.f}, {1.f, 1.f}};
            unsigned pixc[2];
            if (FINAL) {
#pragma unroll
                for (int py = 0; py < 2; ++py) {
                    const int gyc = INTERIOR ? y0 + 2 * my + py : min(y0 + 2 * my + py, Ho - 1);
                    pixc[py] = (unsigned)(gyc * Wo);
                    px[py] = px_next[py];
                }
                fetch_px(my + 1);
                if (ep.shadow) {
#pragma unroll
                    for (int py = 0; py < 2; ++py) {
                        if (VEC) {
                            load_pair(ep.dark + pixc[py] + gxc, dk[py][0], dk[py][1]);
                            load_pair(ep.inv_flat + pixc[py] + gxc, ifl[py][0], ifl[py][1]);
                        } else {
                            dk[py][0] = ep.dark[pixc[py] + gxc];
                            dk[py][1] = ep.dark[pixc[py] + gx1c];
                            ifl[py][0] = ep.inv_flat[pixc[py] + gxc];
                            ifl[py][1] = ep.inv_flat[pixc[py] + gx1c];
                        }
                    }
                }
            }
            // coefficient row my+2 enters the window; row r lives in slot r % 3
            xpass(my + 2, cc, L0[(k + 2) % 3], L1[(k + 2) % 3], G0[(k + 2) % 3], G1[(k + 2) % 3]);
#pragma unroll
            for (int py = 0; py < 2; ++py) {
                const int gy = y0 + 2 * my + py;
                float v0 = 0.f, v1 = 0.f;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const float fl = py ? rec_lo(2 * j + 1) : rec_lo(2 * j);
                    const float fh = py ? rec_hi(2 * j + 1) : rec_hi(2 * j);
                    const int sl = (k + 2 - j + 3) % 3;
                    v0 = fmaf(fl, L0[sl], v0);
                    v0 = fmaf(fh, G0[sl], v0);
                    v1 = fmaf(fl, L1[sl], v1);
                    v1 = fmaf(fh, G1[sl], v1);
                }
                const bool row_ok = INTERIOR || ((gy < Ho) && v0ok);
                if (!FINAL) {
                    // pitch_o is a multiple of 4 and gx is even: the pair store stays inside the row
                    if (row_ok) store_pair(outA_z + (unsigned)(gy * pitch_o), v0, v1);
                } else {
                    float xin0, xin1;
                    px[py].get(false, xin0, xin1);
                    // exp(log(1+x) + delta) + 1 == (1+x) * exp(delta) + 1   (filtering.py:175,222)
                    float r0 = fmaf(__fadd_rn(1.0f, xin0), DSTR_EXPF(v0), one);
                    float r1 = fmaf(__fadd_rn(1.0f, xin1), DSTR_EXPF(v1), one);
                    if (ep.shadow) {  // flatfield_correction, filtering.py:399-412
                        // where(r <= dark, 0, r - dark) == max(r - dark, 0) for finite values
                        r0 = fmaxf(r0 - dk[py][0], 0.f) * ifl[py][0];
                        r1 = fmaxf(r1 - dk[py][1], 0.f) * ifl[py][1];
                    }
                    if (sizeof(OUT_T) == 2 || ep.shadow) {
                        r0 = fminf(fmaxf(r0, 0.f), 65535.f);  // np.clip; the u16 conversion truncates
                        r1 = fminf(fmaxf(r1, 0.f), 65535.f);
                        if (sizeof(OUT_T) != 2) {
                            r0 = truncf(r0);
                            r1 = truncf(r1);
                        }
                    }
                    if (row_ok) {
                        OUT_T* o = out_z + pixc[py] + gx;  // gy < Ho: pixc is the unclamped row
                        if (VEC) {
                            store_pair(o, r0, r1);
                        } else {
                            store_one(o, r0);
                            if (v1ok) store_one(o + 1, r1);
                        }
                    }
                }
            }
        }
    }
}

template <bool FINAL, typename IN_T, typename OUT_T, bool VEC>
__global__ void __launch_bounds__(SY_THREADS, DSTR_SY_MINB)
synth_kernel(const float* __restrict__ dA, const float* __restrict__ dH, int Hl, int Wl, int pitch_l,
             size_t pstride_l, float* __restrict__ outA, int Ho, int Wo, int pitch_o,
             size_t pstride_o, const IN_T* __restrict__ img, OUT_T* __restrict__ out,
             size_t img_pstride, EpilogueArgs ep) {
    const int wid = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);  // warp-uniform by construction
    const int x0 = (blockIdx.x * SY_WARPS + wid) * SY_TX;
    const int y0 = blockIdx.y * SY_TY;
    if (x0 >= Wo) return;  // no block-level synchronisation below
    // the big final kernel gets a predicate-free path for tiles away from the right / bottom borders
    // (no measurable gain for the small deep-level launches):
    // coefficient columns x0/2 .. x0/2 + 33, coefficient rows y0/2 .. y0/2 + SY_TY/2 + 3, SY_TX x SY_TY pixels
    const bool interior = FINAL && (x0 + SY_TX <= Wo) && ((x0 >> 1) + 34 <= Wl) && (y0 + SY_TY <= Ho) &&
                          ((y0 >> 1) + SY_TY / 2 + 4 <= Hl);
    if (interior)
        synth_tile<FINAL, IN_T, OUT_T, VEC, true>(dA, dH, Hl, Wl, pitch_l, pstride_l, outA, Ho, Wo, pitch_o, pstride_o,
                                                  img, out, img_pstride, ep, x0, y0);
    else
        synth_tile<FINAL, IN_T, OUT_T, VEC, false>(dA, dH, Hl, Wl, pitch_l, pstride_l, outA, Ho, Wo, pitch_o, pstride_o,
                                                   img, out, img_pstride, ep, x0, y0);
}

// standalone flatfield_correction (filtering.py:338-414): elementwise over n_outer x n_inner
__global__ void __launch_bounds__(256)
flatfield_kernel(const float* __restrict__ img, const float* __restrict__ flat,
                 const float* __restrict__ dark, const float* __restrict__ baseline,
                 unsigned short* __restrict__ out, size_t n_outer, size_t n_inner) {
    const size_t n = n_outer * n_inner;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        float v = img[i];
        const float dk = dark[i];
        v = (v <= dk) ? 0.f : (v - dk);
        v = v / flat[i];
        if (baseline) v -= baseline[i / n_inner];
        v = fminf(fmaxf(v, 0.f), 65535.f);
        out[i] = (unsigned short)v;
    }
}

// 2x2x2 windowed mean of uint16 volumes with truncation (xarray_multiscale.windowed_mean +
// preserve_dtype, zarr_destriper.py:399-405): out = floor(sum of 8 / 8); odd trailing planes, rows
// and columns are cropped.  One thread per output voxel pair along x when Wo is even.
__global__ void __launch_bounds__(256)
downscale2x_kernel(const unsigned short* __restrict__ in, int Z, int H, int W,
                   unsigned short* __restrict__ out, int Zo, int Ho, int Wo) {
    const size_t n = (size_t)Zo * Ho * Wo;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const int x = (int)(i % Wo);
        const size_t t = i / Wo;
        const int y = (int)(t % Ho), z = (int)(t / Ho);
        unsigned sum = 0;
#pragma unroll
        for (int dz = 0; dz < 2; ++dz)
#pragma unroll
            for (int dy = 0; dy < 2; ++dy) {
                const unsigned short* p = in + ((size_t)(2 * z + dz) * H + (2 * y + dy)) * W + 2 * x;
                sum += (unsigned)p[0] + (unsigned)p[1];
            }
        out[i] = (unsigned short)(sum >> 3);
    }
}

}  // namespace dstr
