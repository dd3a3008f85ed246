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
    int pad[3];
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
    int mode;  // 0: always no_cells; 1: per-plane dispatch
};

__device__ __forceinline__ int reflect_idx(int i, int n) {
    // half-sample symmetric extension, repeated for short signals
    if (i < 0 || i >= n) {
        const int p = 2 * n;
        i %= p;
        if (i < 0) i += p;
        if (i >= n) i = p - 1 - i;
    }
    return i;
}

__device__ __forceinline__ int plane_uses_cells(const PlaneStat& ps, const DispatchParams& dp) {
    // filtering.py:462  fore_mean > back_mean and fore_mean > microscope_high_int
    if (dp.mode == 0) return 0;
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

// =============================================================================================
// analysis: one 2-D db3 level, symmetric mode, axis -2 first then axis -1 (pywt.dwt2), keeping
// cA ('aa') and cH ('da': high-pass along Y, low-pass along X).
// =============================================================================================
constexpr int AN_TOX = 64;
constexpr int AN_TOY = 16;
constexpr int AN_INW = 2 * AN_TOX + 4;  // 132
constexpr int AN_INH = 2 * AN_TOY + 4;  // 36
constexpr int AN_THREADS = 256;

template <typename IN_T, bool FIRST>
__global__ void __launch_bounds__(AN_THREADS)
analysis_kernel(const IN_T* __restrict__ in, int Hs, int Ws, int in_pitch, size_t in_pstride,
                float* __restrict__ cA, float* __restrict__ cH, int Ho, int Wo, int out_pitch,
                size_t out_pstride, LevelStat* __restrict__ lstat, int stat_stride,
                PlaneStat* __restrict__ pstat, float fg_half_thr) {
    __shared__ float s_in[AN_INH][AN_INW + 1];
    __shared__ float s_a[AN_TOY][AN_INW + 1];
    __shared__ float s_d[AN_TOY][AN_INW + 1];
    __shared__ float s_red[2][AN_THREADS / 32];
    __shared__ double s_dred[2][AN_THREADS / 32];
    __shared__ unsigned s_cred[2][AN_THREADS / 32];

    const int tid = threadIdx.x;
    const int z = blockIdx.z;
    const int ox0 = blockIdx.x * AN_TOX;
    const int oy0 = blockIdx.y * AN_TOY;
    const IN_T* src = in + (size_t)z * in_pstride;

    double fg_s = 0.0, bg_s = 0.0;
    unsigned fg_c = 0, bg_c = 0;

    for (int idx = tid; idx < AN_INH * AN_INW; idx += AN_THREADS) {
        const int r = idx / AN_INW;
        const int c = idx - r * AN_INW;
        const int gy0 = 2 * oy0 - 4 + r;
        const int gx0 = 2 * ox0 - 4 + c;
        const int gy = reflect_idx(gy0, Hs);
        const int gx = reflect_idx(gx0, Ws);
        float v = (float)src[(size_t)gy * in_pitch + gx];
        if (FIRST) {
            // plane statistics: every pixel exactly once (tile interior, un-reflected)
            if (r >= 4 && r < 4 + 2 * AN_TOY && c >= 4 && c < 4 + 2 * AN_TOX && gy0 < Hs &&
                gx0 < Ws) {
                const float hv = __half2float(__float2half_rn(v));
                if (hv >= fg_half_thr) {
                    fg_s += (double)v;
                    fg_c++;
                } else {
                    bg_s += (double)v;
                    bg_c++;
                }
            }
            v = logf(__fadd_rn(1.0f, v));  // np.log(1.0 + x) in float32
        }
        s_in[r][c] = v;
    }
    __syncthreads();

    // axis -2 (Y)
    for (int idx = tid; idx < AN_TOY * AN_INW; idx += AN_THREADS) {
        const int oy = idx / AN_INW;
        const int c = idx - oy * AN_INW;
        float a = 0.f, d = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const float v = s_in[2 * oy + 5 - j][c];
            a = fmaf(dec_lo(j), v, a);
            d = fmaf(dec_hi(j), v, d);
        }
        s_a[oy][c] = a;
        s_d[oy][c] = d;
    }
    __syncthreads();

    // axis -1 (X), low-pass only
    float qmin = __int_as_float(0x7f800000), qmax = 0.f;
    float* dA = cA + (size_t)z * out_pstride;
    float* dH = cH + (size_t)z * out_pstride;
    for (int idx = tid; idx < AN_TOY * AN_TOX; idx += AN_THREADS) {
        const int oy = idx / AN_TOX;
        const int ox = idx - oy * AN_TOX;
        float a = 0.f, h = 0.f;
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            a = fmaf(dec_lo(j), s_a[oy][2 * ox + 5 - j], a);
            h = fmaf(dec_lo(j), s_d[oy][2 * ox + 5 - j], h);
        }
        const int goy = oy0 + oy, gox = ox0 + ox;
        if (goy < Ho && gox < Wo) {
            dA[(size_t)goy * out_pitch + gox] = a;
            dH[(size_t)goy * out_pitch + gox] = h;
            const float q = __fmul_rn(h, h);
            qmin = fminf(qmin, q);
            qmax = fmaxf(qmax, q);
        }
    }

    // block reductions -> one atomic per block
    const int lane = tid & 31, wid = tid >> 5;
    qmin = warp_min(qmin);
    qmax = warp_max(qmax);
    if (lane == 0) {
        s_red[0][wid] = qmin;
        s_red[1][wid] = qmax;
    }
    if (FIRST) {
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
    }
    __syncthreads();
    if (tid == 0) {
        float mn = s_red[0][0], mx = s_red[1][0];
        for (int w = 1; w < AN_THREADS / 32; ++w) {
            mn = fminf(mn, s_red[0][w]);
            mx = fmaxf(mx, s_red[1][w]);
        }
        LevelStat* st = lstat + (size_t)z * stat_stride;
        if (mn <= mx) {  // at least one valid output in this block
            atomicMax(&st->qmin_inv, ~__float_as_uint(mn));
            atomicMax(&st->qmax_bits, __float_as_uint(mx));
        }
        if (FIRST) {
            double fs = 0.0, bs = 0.0;
            unsigned long long fc = 0, bc = 0;
            for (int w = 0; w < AN_THREADS / 32; ++w) {
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
}

// plane statistics only (get_foreground_background_mean, filtering.py:54-88)
template <typename IN_T>
__global__ void __launch_bounds__(256)
plane_stats_kernel(const IN_T* __restrict__ in, int H, int W, size_t pstride,
                   PlaneStat* __restrict__ pstat, float fg_half_thr) {
    __shared__ double s_dred[2][8];
    __shared__ unsigned s_cred[2][8];
    const int z = blockIdx.y;
    const IN_T* src = in + (size_t)z * pstride;
    const size_t n = (size_t)H * W;
    double fg_s = 0.0, bg_s = 0.0;
    unsigned fg_c = 0, bg_c = 0;
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
        const float v = (float)src[i];
        const float hv = __half2float(__float2half_rn(v));
        if (hv >= fg_half_thr) {
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

    const float* src = cH + (size_t)z * pstride;
    const int lane = tid & 31;
    const int wl_up = (Wl + 255) & ~255;
    for (int r = blockIdx.x; r < Hl; r += gridDim.x) {
        const float* row = src + (size_t)r * pitch;
        for (int c = tid; c < wl_up; c += 256) {
            int idx = -1;
            if (c < Wl) {
                const float v = row[c];
                const float q = __fmul_rn(v, v);
                const float f = __fmul_rn(__fdiv_rn(__fsub_rn(q, first), delta), 256.0f);
                idx = (int)f;
                idx = max(0, min(idx, 255));
                if (q < s_edges[idx]) {
                    idx = max(idx - 1, 0);
                } else if (idx != 255 && q >= s_edges[idx + 1]) {
                    idx++;
                }
            }
            const unsigned peers = __match_any_sync(0xffffffffu, idx);
            if (idx >= 0 && lane == (__ffs(peers) - 1)) atomicAdd(&s_hist[idx], __popc(peers));
        }
    }
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
        st->thr = (sq < max_thr) ? sq : max_thr;  // python min(max_threshold, sqrt)
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

// acc[i] += sum_k taps[k] * Xlog[8*m0 + i - k],  k = 0..ntap-1 (ntap % 8 == 0)
__device__ __forceinline__ void fir8(float (&acc)[8], const float* __restrict__ Xphys,
                                     const float* __restrict__ taps, int ntap, int m0) {
    const float* Xp = Xphys + 9 * m0;
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) w[i] = Xp[i];
    const float4* t4 = reinterpret_cast<const float4*>(taps);
    for (int kk = 0; kk < ntap / 8; ++kk) {
        const float4 ta = t4[2 * kk], tb = t4[2 * kk + 1];
        const float tk[8] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
        Xp -= 9;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(tk[k], w[(i - k) & 7], acc[i]);
            w[7 - k] = Xp[7 - k];
        }
    }
}

struct NotchTables {
    const float* te;  // even-part FIR taps [ntap_e]; tap k <-> circular offset u = ue_lo + k
    const float* to;  // odd-part FIR taps  [ntap_o]; tap k <-> u = uo_lo + k
    const float* T1;  // [nhp4][Jpad]  omega_v cos(2 pi j v / n)   (v-major)
    const float* T2;  // [J][nhp8]     rho_j   cos(2 pi j t / n)   (j-major)
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
    int xlen_e_phys, xlen_o_phys;  // per-row physical floats (max over the two configs)
    int ntap_e_max, ntap_o_max, Jpad_max;
};

template <int EPL>
__global__ void __launch_bounds__(FR_THREADS)
filter_rows_kernel(FilterLevelArgs a, const PlaneStat* __restrict__ pstat, DispatchParams dp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.Wl;
    float* s_te = reinterpret_cast<float*>(smem_raw);                 // [ntap_e_max]
    float* s_to = s_te + a.ntap_e_max;                                // [ntap_o_max]
    float* s_E = s_to + a.ntap_o_max;                                 // [FR_ROWS][xlen_e_phys]
    float* s_O = s_E + FR_ROWS * a.xlen_e_phys;                       // [FR_ROWS][xlen_o_phys]
    float* s_x = s_O + FR_ROWS * a.xlen_o_phys;                       // [FR_ROWS][n_pad8]
    float* s_c = s_x + FR_ROWS * a.n_pad8;                            // [FR_ROWS][Jpad_max]
    unsigned char* s_m = reinterpret_cast<unsigned char*>(s_c + FR_ROWS * a.Jpad_max);  // [FR_ROWS][n_pad8]

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int z = blockIdx.y;
    const int row0 = blockIdx.x * FR_ROWS;
    const int nrows = min(FR_ROWS, a.Hl - row0);
    const int cfg = plane_uses_cells(pstat[z], dp);
    const NotchTables nt = cfg ? a.nt[1] : a.nt[0];
    const float thr = a.lstat[(size_t)z * a.stat_stride].thr;
    const int nh = a.nh;

    for (int i = tid; i < nt.ntap_e; i += FR_THREADS) s_te[i] = nt.te[i];
    for (int i = tid; i < nt.ntap_o; i += FR_THREADS) s_to[i] = nt.to[i];

    if (wid < nrows) {
        float* grow = a.cH + (size_t)z * a.pstride + (size_t)(row0 + wid) * a.pitch;
        // ---- load, mask, keys ------------------------------------------------------------
        unsigned key[EPL];
        float* xs = s_x + wid * a.n_pad8;
        unsigned char* ms = s_m + wid * a.n_pad8;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const int e = lane + 32 * i;
            key[i] = 0xffffffffu;
            if (e < n) {
                const float c = grow[e];
                const float p = __fsqrt_rn(__fmul_rn(c, c));
                const bool m = p > thr;
                const float bg = m ? 0.0f : (c + 0.0f);  // canonical +0
                key[i] = f2key(bg);
                xs[e] = c;
                ms[e] = m ? 1 : 0;
            }
        }
        // ---- exact median of the zero-filled background (np.median, filtering.py:201) -------
        const int k1 = (n - 1) >> 1;
        unsigned res = 0;
        for (int b = 31; b >= 0; --b) {
            const unsigned trial = res | (1u << b);
            int cnt = 0;
#pragma unroll
            for (int i = 0; i < EPL; ++i) cnt += (key[i] < trial) ? 1 : 0;
            cnt = __reduce_add_sync(0xffffffffu, cnt);
            if (cnt <= k1) res = trial;
        }
        float med = key2f(res);
        if ((n & 1) == 0) {
            int cle = 0;
            unsigned nxt = 0xffffffffu;
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                cle += (key[i] <= res) ? 1 : 0;
                if (key[i] > res) nxt = min(nxt, key[i]);
            }
            cle = __reduce_add_sync(0xffffffffu, cle);
            nxt = __reduce_min_sync(0xffffffffu, nxt);
            const unsigned k2 = (cle >= k1 + 2) ? res : nxt;
            med = (key2f(res) + key2f(k2)) * 0.5f;
        }
        __syncwarp();
        // ---- in-paint ---------------------------------------------------------------------
        for (int e = lane; e < n; e += 32)
            if (ms[e]) xs[e] = med;
        __syncwarp();
        // ---- even / odd parts, circularly extended:  E[a] = x_e[(a - OFFe) mod n] ----------
        float* E = s_E + wid * a.xlen_e_phys;
        float* O = s_O + wid * a.xlen_o_phys;
        {
            const int OFF = nt.ue_lo + nt.ntap_e;
            const int xlen_log = a.nhp8 + nt.ntap_e;
            int t = (lane - OFF) % n;
            if (t < 0) t += n;
            const int step = 32 % n;
            for (int al = lane; al < xlen_log; al += 32) {
                const int tr = (t == 0) ? 0 : n - t;
                E[al + (al >> 3)] = 0.5f * (xs[t] + xs[tr]);
                t += step;
                if (t >= n) t -= n;
            }
        }
        {
            const int OFF = nt.uo_lo + nt.ntap_o;
            const int xlen_log = a.nhp8 + nt.ntap_o;
            int t = (lane - OFF) % n;
            if (t < 0) t += n;
            const int step = 32 % n;
            for (int al = lane; al < xlen_log; al += 32) {
                const int tr = (t == 0) ? 0 : n - t;
                O[al + (al >> 3)] = 0.5f * (xs[t] - xs[tr]);
                t += step;
                if (t >= n) t -= n;
            }
        }
        __syncwarp();
        // ---- rank-J correction coefficients  c_j = sum_v T1[v][j] x_e[v] ---------------------
        if (nt.J > 0) {
            const int OFF = nt.ue_lo + nt.ntap_e;
            const int nhp4 = (nh + 4) & ~3;
            for (int v = lane; v < nhp4; v += 32) {
                const int al = v + OFF;
                xs[v] = (v <= nh) ? E[al + (al >> 3)] : 0.f;
            }
            __syncwarp();
            const float4* x4 = reinterpret_cast<const float4*>(xs);
            const int Jpad = nt.Jpad;
            float* cp = s_c + wid * a.Jpad_max;
            for (int j0 = 0; j0 < Jpad; j0 += 64) {
                const bool two = (j0 + 32) < Jpad;
                const float* t1 = nt.T1 + j0 + lane;
                float acc0 = 0.f, acc1 = 0.f;
                for (int v4 = 0; v4 < (nhp4 >> 2); ++v4) {
                    const float4 xv = x4[v4];
                    const float* tt = t1 + (size_t)(4 * v4) * Jpad;
                    acc0 = fmaf(xv.x, __ldg(tt), acc0);
                    acc0 = fmaf(xv.y, __ldg(tt + Jpad), acc0);
                    acc0 = fmaf(xv.z, __ldg(tt + 2 * Jpad), acc0);
                    acc0 = fmaf(xv.w, __ldg(tt + 3 * Jpad), acc0);
                    if (two) {
                        acc1 = fmaf(xv.x, __ldg(tt + 32), acc1);
                        acc1 = fmaf(xv.y, __ldg(tt + Jpad + 32), acc1);
                        acc1 = fmaf(xv.z, __ldg(tt + 2 * Jpad + 32), acc1);
                        acc1 = fmaf(xv.w, __ldg(tt + 3 * Jpad + 32), acc1);
                    }
                }
                cp[j0 + lane] = acc0;
                if (two) cp[j0 + 32 + lane] = acc1;
            }
        }
    }
    __syncthreads();

    // ---- register-tiled FIRs + rank-J correction over (row, 8-output segment) pairs -----------
    const int nseg = a.nhp8 >> 3;
    for (int w = tid; w < nrows * nseg; w += FR_THREADS) {
        const int r = w / nseg;
        const int seg = w - r * nseg;
        float ye[8], yo[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) ye[i] = yo[i] = 0.f;
        fir8(ye, s_E + r * a.xlen_e_phys, s_te, nt.ntap_e, seg + (nt.ntap_e >> 3));
        fir8(yo, s_O + r * a.xlen_o_phys, s_to, nt.ntap_o, seg + (nt.ntap_o >> 3));
        if (nt.J > 0) {
            const float* cp = s_c + r * a.Jpad_max;
            const float4* t2 = reinterpret_cast<const float4*>(nt.T2 + 8 * seg);
            const int stride4 = a.nhp8 >> 2;
            for (int j = 0; j < nt.J; ++j) {
                const float c = cp[j];
                const float4 u0 = __ldg(t2 + (size_t)j * stride4);
                const float4 u1 = __ldg(t2 + (size_t)j * stride4 + 1);
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
        float* orow = a.cH + (size_t)z * a.pstride + (size_t)(row0 + r) * a.pitch;
        const unsigned char* ms = s_m + r * a.n_pad8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int t = 8 * seg + i;
            if (t <= nh) {
                orow[t] = ms[t] ? 0.f : -(ye[i] + yo[i]);
                const int tm = n - t;
                if (t != 0 && tm != t) orow[tm] = ms[tm] ? 0.f : -(ye[i] - yo[i]);
            }
        }
    }
}

// =============================================================================================
// synthesis of the deltas: out = idwt2(dA, (dH, 0, 0)) (axis -1 first with rec_lo for both
// bands, then axis -2 with rec_lo on the dA branch and rec_hi on the dH branch), trimmed to the
// parent level's shape (pywt.waverec2).  FINAL fuses the inverse log and the epilogue.
// =============================================================================================
constexpr int SY_TX = 64;
constexpr int SY_TY = 32;
constexpr int SY_CW = SY_TX / 2 + 2;  // 34
constexpr int SY_CH = SY_TY / 2 + 2;  // 18
constexpr int SY_THREADS = 256;

struct EpilogueArgs {
    const float* flat;  // nullable
    const float* dark;  // nullable
    int shadow;         // apply dark/flat
    int expm1;          // exp(y) - 1 instead of exp(y) + 1
};

template <bool FINAL, typename IN_T, typename OUT_T>
__global__ void __launch_bounds__(SY_THREADS)
synth_kernel(const float* __restrict__ dA, const float* __restrict__ dH, int Hl, int Wl, int pitch_l,
             size_t pstride_l, float* __restrict__ outA, int Ho, int Wo, int pitch_o,
             size_t pstride_o, const IN_T* __restrict__ img, OUT_T* __restrict__ out,
             size_t img_pstride, EpilogueArgs ep) {
    __shared__ float s_A[SY_CH][SY_CW + 1];
    __shared__ float s_H[SY_CH][SY_CW + 1];
    __shared__ float s_L[SY_CH][SY_TX + 1];
    __shared__ float s_G[SY_CH][SY_TX + 1];
    const int tid = threadIdx.x;
    const int z = blockIdx.z;
    const int x0 = blockIdx.x * SY_TX, y0 = blockIdx.y * SY_TY;
    const int cx0 = x0 >> 1, cy0 = y0 >> 1;
    const float* pA = dA ? dA + (size_t)z * pstride_l : nullptr;
    const float* pH = dH ? dH + (size_t)z * pstride_l : nullptr;

    for (int idx = tid; idx < SY_CH * SY_CW; idx += SY_THREADS) {
        const int r = idx / SY_CW, c = idx - r * SY_CW;
        const int gy = cy0 + r, gx = cx0 + c;
        const bool ok = (gy < Hl) && (gx < Wl);
        s_A[r][c] = (ok && pA) ? pA[(size_t)gy * pitch_l + gx] : 0.f;
        s_H[r][c] = (ok && pH) ? pH[(size_t)gy * pitch_l + gx] : 0.f;
    }
    __syncthreads();
    // axis -1
    for (int idx = tid; idx < SY_CH * SY_TX; idx += SY_THREADS) {
        const int r = idx / SY_TX, x = idx - r * SY_TX;
        const int mx = x >> 1, px = x & 1;
        float l = 0.f, g = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float f = px ? rec_lo(2 * j + 1) : rec_lo(2 * j);
            l = fmaf(f, s_A[r][mx + 2 - j], l);
            g = fmaf(f, s_H[r][mx + 2 - j], g);
        }
        s_L[r][x] = l;
        s_G[r][x] = g;
    }
    __syncthreads();
    // axis -2 (+ epilogue)
    for (int idx = tid; idx < SY_TY * SY_TX; idx += SY_THREADS) {
        const int y = idx / SY_TX, x = idx - y * SY_TX;
        const int gy = y0 + y, gx = x0 + x;
        if (gy >= Ho || gx >= Wo) continue;
        const int my = y >> 1, py = y & 1;
        float v = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float fl = py ? rec_lo(2 * j + 1) : rec_lo(2 * j);
            const float fh = py ? rec_hi(2 * j + 1) : rec_hi(2 * j);
            v = fmaf(fl, s_L[my + 2 - j][x], v);
            v = fmaf(fh, s_G[my + 2 - j][x], v);
        }
        if (!FINAL) {
            outA[(size_t)z * pstride_o + (size_t)gy * pitch_o + gx] = v;
        } else {
            // exp(log(1+x) + delta) + 1 == (1+x) * exp(delta) + 1   (filtering.py:175,222)
            const size_t pix = (size_t)gy * Wo + gx;
            const float xin = (float)img[(size_t)z * img_pstride + pix];
            float r = __fmul_rn(__fadd_rn(1.0f, xin), expf(v));
            r = ep.expm1 ? (r - 1.0f) : (r + 1.0f);
            if (ep.shadow) {  // flatfield_correction, filtering.py:399-412
                const float dk = ep.dark[pix];
                r = (r <= dk) ? 0.f : (r - dk);
                r = r / ep.flat[pix];
            }
            if (sizeof(OUT_T) == 2) {
                r = fminf(fmaxf(r, 0.f), 65535.f);  // np.clip
                out[(size_t)z * img_pstride + pix] = (OUT_T)(unsigned short)r;  // truncation
            } else {
                if (ep.shadow) r = truncf(fminf(fmaxf(r, 0.f), 65535.f));
                out[(size_t)z * img_pstride + pix] = (OUT_T)r;
            }
        }
    }
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

}  // namespace dstr
