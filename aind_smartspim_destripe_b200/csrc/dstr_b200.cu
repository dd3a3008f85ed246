// libdstr_b200.so - host side of the B200-native destripe engine: C-ABI (include/dstr_b200.h),
// workspace management, kernel orchestration and the H2D / compute / D2H stream pipeline that
// replaces the reference's multiprocessing producer/consumer
// (/root/reference/code/aind_smartspim_destripe/zarr_destriper.py:797-906).
#include "dstr_kernels.cuh"
#include "dstr_notch_umma.cuh"
#include "dstr_rows_mma.cuh"
#include "dstr_dual_band.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/dstr_b200.h"

using namespace dstr;

namespace {

constexpr int kMaxLevels = 16;
constexpr int kFilterTaps = 6;  // db3
constexpr int kSideStreams = 3;

std::mutex g_err_mutex;
std::string g_last_error;

void set_global_error(const std::string& s) {
    std::lock_guard<std::mutex> lk(g_err_mutex);
    g_last_error = s;
}

int max_level_1d(int n) {
    // pywt.dwt_max_level: floor(log2(n / (F - 1)))
    if (n < kFilterTaps - 1) return 0;
    int lvl = 0;
    while ((long long)(kFilterTaps - 1) << (lvl + 1) <= (long long)n) ++lvl;
    return lvl;
}

struct LevelGeom {
    int H, W, pitch;
    size_t pstride;  // floats per plane
};

// Device tables of one (level, config): FIR taps and the rank-J cosine correction.
struct NotchDevice {
    float sigma = -1.f;  // sigma the tables were built for
    double eps = -1.0;
    float* d_buf = nullptr;  // one allocation: te | to | T1 | T2
    NotchTables nt = {};
    // tcgen05 path (dstr_notch_umma.cuh): fp16 hi / lo Hankel tables [TBh | TBl | TRh | TRl]
    uint4* d_umma = nullptr;
    UmmaCfg um = {};  // table geometry of this config (tables = d_umma)
    // warp-level tensor path (dstr_rows_mma.cuh): one allocation [tr | T1f | T2f]
    unsigned char* d_mma = nullptr;
    MmaCfg mm = {};
};

// Geometry of the tcgen05 row filter for a band of width n (depends on n only).
struct UmmaGeom {
    int nh = 0, nout = 0, P = 0, Nt = 0, NC = 0, Kpad = 0, tr_bytes = 0, mw = 0;
    size_t smem_fixed = 0;  // ring + mask bits (the tables come on top, per config)
    bool ok = false;
};

struct TapTable {
    NotchDevice cfg[2];  // [0] no_cells, [1] cells
};

struct TimerSpan {
    int id;
    cudaEvent_t a, b;
};

}  // namespace

struct dstr_ctx {
    int device = 0;
    int sm_count = 148;
    int zcap = 0;
    int H = 0, W = 0;
    int Lmax = 0;    // pywt.dwtn_max_level of the plane shape (what level=None means)
    int Lalloc = 0;  // levels with workspace: deeper explicit levels are allowed like in pywt (it only warns)
    LevelGeom geom[kMaxLevels + 1];
    float* d_A[kMaxLevels + 1] = {};
    float* d_H[kMaxLevels + 1] = {};
    LevelStat* d_lstat = nullptr;  // [Lmax][zcap]
    PlaneStat* d_pstat = nullptr;  // [zcap]
    TapTable taps[kMaxLevels + 1];
    float* d_flat = nullptr;
    float* d_dark = nullptr;
    bool have_flat_dark = false;
    // staging for host buffers (double buffered)
    void* d_in[2] = {nullptr, nullptr};
    void* d_out[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    // optional fused pyramid outputs (levels 1 and 2) of the chunk being destriped
    void* pyr_out[2] = {nullptr, nullptr};
    uint16_t* d_pyr[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [level][buffer]
    size_t pyr_bytes = 0;
    int subchunk = 0;
    cudaStream_t s_h2d = nullptr, s_comp = nullptr, s_d2h = nullptr;
    cudaStream_t s_side[4] = {};                       // per-level hist/otsu/filter branches
    cudaEvent_t ev_an[kMaxLevels + 1] = {}, ev_flt[kMaxLevels + 1] = {};
    bool overlap = true;
    bool use_tma = true;  // level-1 analysis through the TMA-staged kernel when the plane shape allows
    // row filter on tcgen05 where the band geometry allows (dstr_notch_umma.cuh).  Off by default: measured on
    // B200 it is 6-10 % slower than the CUDA-core kernel; DSTR_UMMA=1 or dstr_set_umma turn it on
    bool use_umma = false;
    int row_filter = 1;  // 0: FMA kernel (filter_rows_kernel), 1: mma.sync kernel (filter_rows_mma_kernel)
    // per-CTA operand scratch of the tcgen05 row filter, one per level (the levels' filters run
    // concurrently on the side streams)
    uint8_t* d_um_scratch[kMaxLevels + 1] = {};
    // dual-band mode (dstr_dual_band_chunk): [0] background band filtered, [1] foreground band filtered, [2] clamped
    // band (float32, zcap planes each), per-plane thresholds, 1 / flat, staging for host chunks
    float* d_db[3] = {nullptr, nullptr, nullptr};
    float* d_db_thr = nullptr;
    float* d_db_invflat = nullptr;
    void* d_db_in = nullptr;
    uint16_t* d_db_out = nullptr;
    size_t um_scratch_bytes[kMaxLevels + 1] = {};
    cudaEvent_t ev_h2d[2] = {}, ev_comp[2] = {}, ev_d2h[2] = {};
    unsigned long long host_it = 0;  // sub-chunks streamed so far (staging buffer = host_it & 1), across calls
    bool async_pending = false;      // a DSTR_FLAG_NO_SYNC call with host buffers is still in flight
    float fg_half_thr = 384.f;
    float fg_thr32 = 384.f;  // the same rule on the float32 value (fg_threshold_f32)
    double notch_eps = 1e-6;  // truncation tolerance of the hybrid notch operator (0 = dense)
    // instrumentation
    bool profiling = false;
    int debug_stop = DSTR_STAGE_NONE;
    int last_levels = 0;
    int last_z = 0;
    DispatchParams last_dp = {};  // dispatch rule of the last pass (dstr_debug_fetch reports use_cells)
    double timers[DSTR_NUM_TIMERS] = {};
    uint64_t launches = 0;
    std::vector<TimerSpan> spans;
    std::vector<cudaEvent_t> ev_pool;
    std::string err;
};

namespace {

int fail(dstr_ctx* ctx, int code, const std::string& msg) {
    if (ctx) ctx->err = msg;
    set_global_error(msg);
    return code;
}

#define CK(ctx, call)                                                                         \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess) {                                                              \
            return fail(ctx, (int)e_,                                                         \
                        std::string(#call) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + \
                            ":" + std::to_string(__LINE__) + ")");                            \
        }                                                                                     \
    } while (0)

// ---- float16 foreground rule (filtering.py:78-81) -----------------------------------------------
// mask = sigmoid((float16(v) - 400) / 20) > 0.3 evaluated in float16.  The rule is monotone in
// float16(v); the smallest float16 that satisfies it is found by emulating the float16
// arithmetic with correctly rounded conversions (exp evaluated in float32 then rounded, as
// numpy does for float16 ufuncs).
float half_round(float x) { return __half2float(__float2half_rn(x)); }

float half_from_bits(unsigned bits) {
    __half_raw hr;
    hr.x = (unsigned short)bits;
    return __half2float(__half(hr));
}

bool fg_rule(float h, float thr) {
    float zf = half_round(h - 400.0f);
    zf = half_round(zf / 20.0f);
    const float e = half_round(expf(-zf));
    const float den = half_round(1.0f + e);
    const float f = half_round(1.0f / den);
    return f > thr;
}

// smallest float16 value (ascending over all finite float16) for which the rule holds;
// -inf if it holds everywhere, +inf if nowhere
float find_fg_half_threshold(float threshold_mask) {
    const float thr = half_round(threshold_mask);
    // negatives, ascending: bit patterns 0xfbff (-65504) down to 0x8000 (-0)
    for (unsigned bits = 0xfbffu; bits >= 0x8000u; --bits) {
        if (fg_rule(half_from_bits(bits), thr)) return bits == 0xfbffu ? -INFINITY : half_from_bits(bits);
    }
    for (unsigned bits = 0; bits <= 0x7bffu; ++bits) {
        if (fg_rule(half_from_bits(bits), thr)) return half_from_bits(bits);
    }
    return INFINITY;
}

// The device applies the rule to the float32 pixel value directly: float16 rounding is monotone,
// so "float16(v) >= h" is "v >= t" for the smallest float32 t that rounds to a float16 >= h
// (found by bisection over the ordered float32 bit patterns).  NaN encodes "never".
float fg_threshold_f32(float half_thr) {
    if (std::isinf(half_thr)) return half_thr < 0.f ? -INFINITY : NAN;
    auto key = [](float f) {
        uint32_t b;
        std::memcpy(&b, &f, 4);
        return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    };
    auto unkey = [](uint32_t k) {
        const uint32_t b = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
        float f;
        std::memcpy(&f, &b, 4);
        return f;
    };
    uint32_t hi = key(half_thr);  // satisfies the rule
    uint32_t lo = key(-INFINITY); // float16(-inf) = -inf < half_thr
    while (hi - lo > 1) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (half_round(unkey(mid)) >= half_thr)
            hi = mid;
        else
            lo = mid;
    }
    return unkey(hi);
}

// ---- time-domain form of the packed-rfft notch (filtering.py:206-215, scipy.fftpack layout) --
// Y_pk[k] = w[k] X_pk[k], w[k] = exp(-k^2 / (2 s^2)) on the packed index k:
//   Re X_j gets a_j = w[2j-1], Im X_j gets b_j = w[2j], DC gets w[0], Nyquist (n even) w[n-1].
//   p_j = (a_j + b_j)/2 multiplies X_j, q_j = (a_j - b_j)/2 multiplies conj(X_j).
void notch_kernels_host(int n, double s, std::vector<double>& hp, std::vector<double>& hq) {
    hp.assign(n, 0.0);
    hq.assign(n, 0.0);
    auto w = [&](int k) { return std::exp(-((double)k * (double)k) / (2.0 * s * s)); };
    const int nh = n / 2;
    const int J = (n % 2 == 0) ? nh + 1 : (n + 1) / 2;
    std::vector<double> p(J), q(J), wt(J);
    p[0] = w(0);
    q[0] = 0.0;
    wt[0] = 1.0;
    for (int j = 1; j < J; ++j) {
        if (n % 2 == 0 && j == nh) {
            p[j] = w(n - 1);
            q[j] = 0.0;
            wt[j] = 1.0;
        } else {
            const double a = w(2 * j - 1), b = w(2 * j);
            p[j] = 0.5 * (a + b);
            q[j] = 0.5 * (a - b);
            wt[j] = 2.0;
        }
    }
    // cos(2 pi j u / n) via a table indexed by (j*u) mod n
    std::vector<double> ct(n);
    for (int i = 0; i < n; ++i) ct[i] = std::cos(2.0 * M_PI * (double)i / (double)n);
    for (int u = 0; u < n; ++u) {
        double sp = 0.0, sq = 0.0;
        long long idx = 0;
        for (int j = 0; j < J; ++j) {
            const double c = ct[idx];
            sp += wt[j] * p[j] * c;
            sq += wt[j] * q[j] * c;
            idx += u;
            if (idx >= n) idx -= n;
        }
        hp[u] = sp / n;
        hq[u] = sq / n;
    }
}

// ---- hybrid design of the even-part operator -------------------------------------------------
// a_j = w(2|j| - 1) (a_0 = 1) = G(j) + r(j):  G(j) = w(2 phi(j) - 1) with the smooth absolute
// value phi(j) = j erf(j / beta) + beta / sqrt(pi) exp(-j^2 / beta^2); r is supported on the
// lowest ~4.5 beta modes.  The kernel of G is compact (radius R_G), r is applied as a rank-J
// cosine correction.  beta is chosen per band to minimise the work; tolerances are relative
// to the L1 norm of the kernel (i.e. to the operator's gain on a bounded row).
struct NotchHost {
    std::vector<float> te, to, T1, T2;
    std::vector<float> T1f;  // T1 in mma A-fragment order for the device: [E block][Jpad / 16][lane][4]
    int ntap_e = 0, ue_lo = 0, ntap_o = 0, uo_lo = 0, J = 0, Jpad = 0;
};

void idft_even(int n, const std::vector<double>& coef, const std::vector<double>& ct,
               std::vector<double>& h) {
    // h[u] = (1/n) [c_0 + 2 sum_{1<=j<n/2} c_j cos(2 pi j u / n) + (n even) c_{n/2} cos(pi u)]
    const int J = (int)coef.size();
    h.assign(n, 0.0);
    for (int u = 0; u < n; ++u) {
        double acc = 0.0;
        long long idx = 0;
        for (int j = 0; j < J; ++j) {
            const double wt = (j == 0 || (n % 2 == 0 && j == n / 2)) ? 1.0 : 2.0;
            acc += wt * coef[j] * ct[idx];
            idx += u;
            if (idx >= n) idx -= n;
        }
        h[u] = acc / n;
    }
}

int tail_radius(int n, const std::vector<double>& h, double eps) {
    // smallest R such that sum_{circular distance > R} |h| < eps * sum |h|
    const int nh = n / 2;
    double tot = 0.0;
    for (double v : h) tot += std::fabs(v);
    std::vector<double> ring(nh + 1, 0.0);
    for (int u = 0; u < n; ++u) ring[std::min(u, n - u)] += std::fabs(h[u]);
    double tail = 0.0;
    int R = nh;
    for (int d = nh; d >= 1; --d) {
        tail += ring[d];
        if (tail >= eps * tot) break;
        R = d - 1;
    }
    return R;
}

void pack_taps(int n, const std::vector<double>& h, int R, bool dense, std::vector<float>& taps,
               int& u_lo, int& ntap_pad) {
    int ntap;
    if (dense) {
        u_lo = -(n / 2);
        ntap = n;
    } else {
        u_lo = -R;
        ntap = 2 * R + 1;
    }
    ntap_pad = (ntap + 7) & ~7;
    taps.assign(ntap_pad, 0.f);
    for (int k = 0; k < ntap; ++k) {
        const int u = u_lo + k;
        taps[k] = (float)h[((u % n) + n) % n];
    }
}

// cost-model weights (overridable for experiments: DSTR_NOTCH_COST_J / DSTR_NOTCH_COST_JPAD)
double env_or(const char* name, double dflt) {
    const char* v = std::getenv(name);
    return (v && *v) ? std::atof(v) : dflt;
}
double notch_cost_j() {
    static const double v = env_or("DSTR_NOTCH_COST_J", 3.0);
    return v;
}
double notch_cost_jpad() {
    static const double v = env_or("DSTR_NOTCH_COST_JPAD", 0.5);
    return v;
}

void design_notch(int n, double s, double eps, NotchHost& out) {
    const int nh = n / 2;
    const int Jn = (n % 2 == 0) ? nh + 1 : (n + 1) / 2;  // cosine modes j = 0..Jn-1
    auto w = [&](double k) { return std::exp(-(k * k) / (2.0 * s * s)); };
    std::vector<double> ct(n);
    for (int i = 0; i < n; ++i) ct[i] = std::cos(2.0 * M_PI * (double)i / (double)n);
    std::vector<double> a(Jn), b(Jn);
    for (int j = 0; j < Jn; ++j) {
        a[j] = (j == 0) ? w(0) : w(2.0 * j - 1.0);
        b[j] = w(2.0 * j);
    }
    if (n % 2 == 0) a[nh] = w(n - 1.0);
    std::vector<double> hb;
    idft_even(n, b, ct, hb);
    const int nhp8 = (nh + 1 + 7) & ~7;
    const int nhp64 = (nhp8 + 63) & ~63;
    const int nhp4 = (nh + 4) & ~3;

    // dense fallback: exact kernels of full circular length
    std::vector<double> ha;
    idft_even(n, a, ct, ha);
    double best_cost = 2.0 * ((n + 7) & ~7);
    bool hybrid = false;
    double best_beta = 0.0;
    int best_RG = 0, best_J = 0;
    int Rb = nh;
    std::vector<double> best_G, best_hG;
    if (eps > 0.0 && n >= 48) {
        Rb = tail_radius(n, hb, eps);
        const double betas[] = {6, 8, 10, 12, 14, 16, 20, 24, 32};
        for (double beta : betas) {
            std::vector<double> G(Jn), hG;
            for (int j = 0; j < Jn; ++j) {
                const double x = (double)j;
                const double phi = x * std::erf(x / beta) + beta / std::sqrt(M_PI) * std::exp(-(x / beta) * (x / beta));
                G[j] = w(2.0 * phi - 1.0);
            }
            idft_even(n, G, ct, hG);
            const int RG = tail_radius(n, hG, eps);
            // number of low modes to keep: sum of dropped 2|r_j| below eps
            double tail = 0.0;
            int J = Jn;
            for (int j = Jn - 1; j >= 0; --j) {
                tail += 2.0 * std::fabs(a[j] - G[j]);
                if (tail >= eps) break;
                J = j;
            }
            if (2 * RG + 1 >= n || 2 * Rb + 1 >= n || J >= nh) continue;
            // work in units of one FIR tap: the rank-J stages issue more loads per FMA than the
            // register-tiled FIR (cost per mode ~3), and the projection runs on 32-lane groups of modes
            const double cost = (double)(((2 * RG + 1 + 7) & ~7) + ((2 * Rb + 1 + 7) & ~7)) +
                                notch_cost_j() * J + notch_cost_jpad() * ((J + 31) & ~31);
            if (cost < best_cost) {
                best_cost = cost;
                hybrid = true;
                best_beta = beta;
                best_RG = RG;
                best_J = J;
                best_G = G;
                best_hG = hG;
            }
        }
    }
    (void)best_beta;
    if (!hybrid) {
        pack_taps(n, ha, 0, true, out.te, out.ue_lo, out.ntap_e);
        pack_taps(n, hb, 0, true, out.to, out.uo_lo, out.ntap_o);
        out.J = 0;
        out.Jpad = 0;
        out.T1.clear();
        out.T1f.clear();
        out.T2.clear();
        return;
    }
    pack_taps(n, best_hG, best_RG, false, out.te, out.ue_lo, out.ntap_e);
    pack_taps(n, hb, Rb, false, out.to, out.uo_lo, out.ntap_o);
    const int J = best_J;
    out.J = J;
    out.Jpad = (J + 31) & ~31;
    out.T1.assign((size_t)(nhp4 + 16) * out.Jpad, 0.f);  // 8 zero rows of margin before v = 0 and after v = nh
    out.T2.assign((size_t)std::max(J, 1) * nhp64, 0.f);
    for (int j = 0; j < J; ++j) {
        const double r = a[j] - best_G[j];
        const double rho = ((j == 0) ? 1.0 : 2.0) * r / n;  // J < n/2: no Nyquist mode here
        for (int v = 0; v <= nh; ++v) {
            const double omega = (v == 0 || (n % 2 == 0 && v == nh)) ? 1.0 : 2.0;
            const double c = ct[(int)(((long long)j * v) % n)];
            out.T1[(size_t)(v + 8) * out.Jpad + j] = (float)(omega * c);
            // row layout: groups of 8 segments (64 outputs); first the float4 of outputs 0..3 of each of
            // the 8 segments, then the float4 of outputs 4..7 (a warp's load is 128 contiguous bytes)
            const int seg = v >> 3, k = v & 7;
            const size_t off = (size_t)(seg >> 3) * 64 + (size_t)(k >> 2) * 32 + (size_t)(seg & 7) * 4 + (k & 3);
            out.T2[(size_t)j * nhp64 + off] = (float)(rho * c);
        }
    }
    // Device copy of T1: the projection runs as m16n8k8 mma tiles over the 8-element blocks of the
    // padded E row (logical index a = OFFe + v, block a >> 3).  Tile (block tb, modes 16 mt .. 16 mt + 15):
    // lane (g = lane / 4, tig = lane % 4) holds a0 = (mode g, element tig), a1 = (mode g + 8, element
    // tig), a2 = (mode g, element tig + 4), a3 = (mode g + 8, element tig + 4).
    {
        const int OFFe = out.ue_lo + out.ntap_e;
        const int blk_lo = OFFe >> 3, blk_hi = (OFFe + nh + 1 + 7) >> 3;
        const int nblk = blk_hi - blk_lo, mtiles = out.Jpad / 16;
        out.T1f.assign((size_t)nblk * mtiles * 32 * 4, 0.f);
        auto t1 = [&](int v, int j) -> float {
            if (v < 0 || v > nh || j >= J) return 0.f;
            return out.T1[(size_t)(v + 8) * out.Jpad + j];
        };
        for (int tb = 0; tb < nblk; ++tb)
            for (int mt = 0; mt < mtiles; ++mt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, tig = lane & 3;
                    const int v0 = 8 * (blk_lo + tb) - OFFe;
                    float* dst = &out.T1f[(((size_t)tb * mtiles + mt) * 32 + lane) * 4];
                    dst[0] = t1(v0 + tig, 16 * mt + g);
                    dst[1] = t1(v0 + tig, 16 * mt + g + 8);
                    dst[2] = t1(v0 + tig + 4, 16 * mt + g);
                    dst[3] = t1(v0 + tig + 4, 16 * mt + g + 8);
                }
    }
}

// ---- tcgen05 row filter: geometry and tables (dstr_notch_umma.cuh) ---------------------------
UmmaGeom umma_geom(int n) {
    UmmaGeom g;
    g.nh = n / 2;
    g.nout = g.nh + 1;
    g.P = (g.nout + UM_NT_MAX - 1) / UM_NT_MAX;
    g.Nt = (((g.nout + g.P - 1) / g.P) + 15) & ~15;
    g.NC = (n + UM_KC - 1) / UM_KC;
    g.Kpad = UM_KC * g.NC;
    const int L = g.P * g.Nt + g.Kpad + 8;  // largest u + v (+ the 7 shifted copies)
    g.tr_bytes = ((L + 7) / 8) * 128;
    g.mw = g.NC;
    g.smem_fixed = (size_t)UM_STAGES * 4 * UM_CHUNK_BYTES + (size_t)2 * UM_ROWS * g.mw * 4;
    // (2n - Rb window of um_band: 1.5 n + 48 < 2 n - n / 2 always holds for n >= 96, so it never occurs)
    g.ok = n >= 96 && n <= 32 * 33 && g.NC <= 64 && g.Nt <= UM_NT_MAX;
    return g;
}

constexpr size_t kUmmaSmemMax = 226 * 1024;  // 227 KB minus the kernel's static shared memory

uint16_t half_bits(float v) {
    const __half h = __float2half_rn(v);
    return *reinterpret_cast<const uint16_t*>(&h);
}
float half_value(uint16_t b) {
    __half_raw hr;
    hr.x = b;
    return __half2float(__half(hr));
}

// Tables of one (band width n, notch width s): core-matrix blocks T[blk][r][e] = f(8 blk + r + e) of the
// periodic sequences f = 256 * 1/2 hb restricted to circular distance <= Rb ("TB") and
// f = 256 * 1/2 (ha - that) ("TR"; the full ha when the remainder needs the three-product split anyway),
// as fp16 hi and lo.  Layout [TRh | TRl (r3 only) | TBh | TBl]; TB is stored as the two windows of k the
// band chunks touch when that is smaller (UmmaCfg, um_tb_off).
struct UmmaHost {
    std::vector<uint16_t> tabs;
    UmmaCfg cfg = {};
    size_t bytes = 0;
};

void build_umma_host(int n, double s, const UmmaGeom& g, UmmaHost& out) {
    const int nh = n / 2;
    const int Jn = (n % 2 == 0) ? nh + 1 : (n + 1) / 2;
    auto w = [&](double k) { return std::exp(-(k * k) / (2.0 * s * s)); };
    std::vector<double> ct(n);
    for (int i = 0; i < n; ++i) ct[i] = std::cos(2.0 * M_PI * (double)i / (double)n);
    std::vector<double> a(Jn), b(Jn), ha, hb;
    for (int j = 0; j < Jn; ++j) {
        a[j] = (j == 0) ? w(0) : w(2.0 * j - 1.0);
        b[j] = w(2.0 * j);
    }
    if (n % 2 == 0) a[nh] = w(n - 1.0);
    idft_even(n, a, ct, ha);
    idft_even(n, b, ct, hb);
    // band radius: the odd part is hb alone, so its truncated tail (relative L1 mass 2e-7) is the
    // truncation error of the operator; for the even part the cut goes into the remainder exactly
    int Rb = tail_radius(n, hb, 2e-7);
    if (2 * Rb + 1 >= n) Rb = nh;
    std::vector<double> fb(n), fr(n);
    double r2 = 0.0;
    for (int k = 0; k < n; ++k) {
        const int d = std::min(k, n - k);
        fb[k] = (d <= Rb) ? hb[k] : 0.0;
        fr[k] = ha[k] - fb[k];
        r2 += fr[k] * fr[k];
    }
    UmmaCfg& c = out.cfg;
    c = UmmaCfg{};
    c.Rb = Rb;
    c.r3 = std::sqrt(r2) > 0.004 ? 1 : 0;  // single-product error ~ |r|_2 2^-12 |x|
    if (c.r3)
        for (int k = 0; k < n; ++k) fr[k] = ha[k];  // E runs entirely on the split remainder table
    c.tr_bytes = g.tr_bytes;
    // TB windows: a band chunk starting at k = lo reads k in [lo, lo + 16 + Nt + 22]
    const int K0 = (Rb + g.Nt + 48 + 7) & ~7;
    const int k1 = std::max(0, (n - Rb - g.Nt - 32)) & ~7;
    const int len1 = ((n + Rb + g.Nt + 40 - k1) + 7) & ~7;
    c.tb_compact = (K0 + len1) * 16 < g.tr_bytes ? 1 : 0;
    c.tb_k1 = k1;
    c.tb_w1_off = (K0 / 8) * 128;
    c.tb_bytes = c.tb_compact ? ((K0 + len1) / 8) * 128 : g.tr_bytes;
    const size_t total_bytes = (size_t)c.tr_bytes * (c.r3 ? 2 : 1) + (size_t)2 * c.tb_bytes;
    out.bytes = (total_bytes + 15) & ~(size_t)15;
    out.tabs.assign(out.bytes / 2, 0);
    auto fill = [&](size_t off_hi_bytes, size_t off_lo_bytes, bool want_lo, int nblk, int kbase, const std::vector<double>& f) {
        for (int blk = 0; blk < nblk; ++blk)
            for (int r = 0; r < 8; ++r)
                for (int e = 0; e < 8; ++e) {
                    const int k = (kbase + 8 * blk + r + e) % n;
                    const size_t o = (size_t)blk * 64 + r * 8 + e;
                    const float v = (float)(0.5 * f[k] * (double)UM_TABLE_SCALE);
                    const uint16_t h = half_bits(v);
                    out.tabs[off_hi_bytes / 2 + o] = h;
                    if (want_lo) out.tabs[off_lo_bytes / 2 + o] = half_bits(v - half_value(h));
                }
    };
    const size_t oTRh = 0, oTRl = c.tr_bytes, oTBh = (size_t)c.tr_bytes * (c.r3 ? 2 : 1), oTBl = oTBh + c.tb_bytes;
    fill(oTRh, oTRl, c.r3 != 0, c.tr_bytes / 128, 0, fr);
    if (c.tb_compact) {
        fill(oTBh, oTBl, true, K0 / 8, 0, fb);
        fill(oTBh + c.tb_w1_off, oTBl + c.tb_w1_off, true, len1 / 8, k1, fb);
    } else {
        fill(oTBh, oTBl, true, c.tb_bytes / 128, 0, fb);
    }
    c.need_band = 0;
    for (int p = 0; p < g.P; ++p)
        for (int cc = 0; cc < g.NC; ++cc)
            if (um_band(p * g.Nt, g.Nt, cc, n, Rb)) c.need_band |= 1ull << cc;
}

// ---- tables of the mma.sync row filter (dstr_rows_mma.cuh) from the hybrid design -------------
struct MmaHost {
    std::vector<uint16_t> tr;    // [E: hi0 | hi1 | lo0 | lo1][trlen_e], [O: ...][trlen_o]
    std::vector<uint32_t> fq;    // tap fragments as register quads: [E: S_e][hi, lo][14][4], [O: S_o][hi, lo][14][4]
    std::vector<uint16_t> T1f;   // [nblk][Jpad / 16][32][16 halfs: 8 hi, 8 lo]
    std::vector<uint16_t> T2f;   // [nseg16][Jpad / 16][32][16 halfs]
    MmaCfg cfg = {};
};

void build_mma_host(int n, const NotchHost& h, MmaHost& out) {
    const int nh = n / 2;
    MmaCfg& c = out.cfg;
    c = MmaCfg{};
    auto pad16 = [](int v) { return (v + 15) & ~15; };
    c.ntap_e = pad16(h.ntap_e);
    c.ntap_o = pad16(h.ntap_o);
    c.ue_lo = h.ue_lo;
    c.uo_lo = h.uo_lo;
    c.S_e = c.ntap_e / 16 + 1;
    c.S_o = c.ntap_o / 16 + 1;
    c.trlen_e = 16 * c.S_e + 32;
    c.trlen_o = 16 * c.S_o + 32;
    out.tr.assign((size_t)4 * (c.trlen_e + c.trlen_o), 0);
    auto fill_tr = [&](size_t base, int trlen, int ntapP, const std::vector<float>& taps, int ntap) {
        // tr[q] = taps[ntapP + 15 - q] (zero outside the real taps), scaled; copy 1 is shifted by one entry
        std::vector<float> tr(trlen + 1, 0.f);
        for (int q = 0; q <= trlen; ++q) {
            const int kk = ntapP + 15 - q;
            if (kk >= 0 && kk < ntap && kk < (int)taps.size()) tr[q] = taps[kk] * RM_TAP_SCALE;
        }
        for (int q = 0; q < trlen; ++q)
            for (int sh = 0; sh < 2; ++sh) {
                const float v = tr[q + sh];
                const uint16_t hi = half_bits(v);
                out.tr[base + (size_t)sh * trlen + q] = hi;
                out.tr[base + (size_t)(2 + sh) * trlen + q] = half_bits(v - half_value(hi));
            }
    };
    fill_tr(0, c.trlen_e, c.ntap_e, h.te, h.ntap_e);
    fill_tr((size_t)4 * c.trlen_e, c.trlen_o, c.ntap_o, h.to, h.ntap_o);
    // the same taps as the register quads the kernel loads: lane (g, tig) reads q = 2 tig - g + 15 (8 .. 21)
    out.fq.assign((size_t)(c.S_e + c.S_o) * 28 * 4, 0u);
    auto fill_fq = [&](size_t qbase, size_t trbase, int trlen, int S) {
        for (int s = 0; s < S; ++s)
            for (int part = 0; part < 2; ++part) {
                const uint16_t* t = out.tr.data() + trbase + (size_t)(part ? 2 : 0) * trlen;  // unshifted hi / lo copy
                auto w = [&](int i) -> uint32_t { return (uint32_t)t[i] | ((uint32_t)t[i + 1] << 16); };
                for (int q = 8; q <= 21; ++q) {
                    uint32_t* dst = out.fq.data() + (qbase + ((size_t)s * 2 + part) * 14 + (q - 8)) * 4;
                    dst[0] = w(16 * s + q);
                    dst[1] = w(16 * s + q - 8);
                    dst[2] = w(16 * s + q + 8);
                    dst[3] = dst[0];
                }
            }
    };
    fill_fq(0, 0, c.trlen_e, c.S_e);
    fill_fq((size_t)c.S_e * 28, (size_t)4 * c.trlen_e, c.trlen_o, c.S_o);
    c.J = h.J;
    c.Jpad = pad16(h.J);
    c.cs = 1.0f;
    c.inv_x = 1.0f;
    const int nseg16 = (nh + 1 + 15) / 16;
    if (h.J > 0) {
        const int OFFe = c.ue_lo + c.ntap_e;
        c.blk_lo = OFFe >> 4;
        const int blk_hi = (OFFe + nh + 1 + 15) >> 4;
        c.nblk = blk_hi - c.blk_lo;
        const int mtiles = c.Jpad / 16;
        auto t1 = [&](int v, int j) -> float {  // omega_v cos(2 pi j v / n)
            if (v < 0 || v > nh || j >= h.J) return 0.f;
            return h.T1[(size_t)(v + 8) * h.Jpad + j];
        };
        const int nhp64 = ((((nh + 1 + 7) & ~7) + 63) & ~63);
        auto t2 = [&](int j, int t) -> float {  // rho_j cos(2 pi j t / n), device layout of the CUDA-core kernel
            if (t > nh || j >= h.J) return 0.f;
            const int seg = t >> 3, k = t & 7;
            const size_t off = (size_t)(seg >> 3) * 64 + (size_t)(k >> 2) * 32 + (size_t)(seg & 7) * 4 + (k & 3);
            return h.T2[(size_t)j * nhp64 + off];
        };
        auto put = [](std::vector<uint16_t>& dst, size_t at, float v0, float v1) {  // one register: (hi pair) and, 8 halfs on, (lo pair)
            const uint16_t h0 = half_bits(v0), h1 = half_bits(v1);
            dst[at] = h0;
            dst[at + 1] = h1;
            dst[at + 8] = half_bits(v0 - half_value(h0));
            dst[at + 9] = half_bits(v1 - half_value(h1));
        };
        out.T1f.assign((size_t)c.nblk * mtiles * 32 * 16, 0);
        for (int blk = 0; blk < c.nblk; ++blk)
            for (int mt = 0; mt < mtiles; ++mt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, tig = lane & 3;
                    const int v0 = 16 * (c.blk_lo + blk) - OFFe;  // element of x_e at k = 0 of this block
                    const size_t at = (((size_t)blk * mtiles + mt) * 32 + lane) * 16;
                    // a0: (mode g, k 2tig..+1)  a1: (mode g + 8, k 2tig..+1)  a2: (mode g, k 2tig + 8..)  a3: (mode g + 8, k 2tig + 8..)
                    const int m0 = 16 * mt + g, m1 = m0 + 8, k0 = v0 + 2 * tig;
                    put(out.T1f, at + 0, t1(k0, m0), t1(k0 + 1, m0));
                    put(out.T1f, at + 2, t1(k0, m1), t1(k0 + 1, m1));
                    put(out.T1f, at + 4, t1(k0 + 8, m0), t1(k0 + 9, m0));
                    put(out.T1f, at + 6, t1(k0 + 8, m1), t1(k0 + 9, m1));
                }
        // power-of-two scales: |c_j| <= 2 (nh + 1) 2^14 before scaling; T2 entries are tiny
        int ce = 0;
        std::frexp(2.0f * (float)(nh + 1), &ce);
        c.cs = std::ldexp(1.0f, -ce);
        float t2max = 0.f;
        for (int j = 0; j < h.J; ++j)
            for (int t = 0; t <= nh; ++t) t2max = std::max(t2max, std::fabs(t2(j, t)));
        int te = 0;
        std::frexp(std::max(t2max, 1e-30f), &te);
        const float ts = std::ldexp(1.0f, 12 - te);  // largest entry in [2^11, 2^12)
        c.inv_x = 1.0f / (c.cs * ts);
        out.T2f.assign((size_t)nseg16 * mtiles * 32 * 16, 0);
        for (int seg = 0; seg < nseg16; ++seg)
            for (int kt = 0; kt < mtiles; ++kt)
                for (int lane = 0; lane < 32; ++lane) {
                    const int g = lane >> 2, tig = lane & 3;
                    const size_t at = (((size_t)seg * mtiles + kt) * 32 + lane) * 16;
                    const int tA = 16 * seg + g, tB = tA + 8, j0 = 16 * kt + 2 * tig;
                    // A[m = output][k = mode]
                    put(out.T2f, at + 0, t2(j0, tA) * ts, t2(j0 + 1, tA) * ts);
                    put(out.T2f, at + 2, t2(j0, tB) * ts, t2(j0 + 1, tB) * ts);
                    put(out.T2f, at + 4, t2(j0 + 8, tA) * ts, t2(j0 + 9, tA) * ts);
                    put(out.T2f, at + 6, t2(j0 + 8, tB) * ts, t2(j0 + 9, tB) * ts);
                }
    }
}

int build_taps_cfg(dstr_ctx* ctx, int level, int cfg, float sigma) {
    NotchDevice& D = ctx->taps[level].cfg[cfg];
    if (D.d_buf && D.sigma == sigma && D.eps == ctx->notch_eps) return 0;
    const int n = ctx->geom[level].W;
    const int Hl = ctx->geom[level].H;
    // s = rows of this band * sigma / min(H, W)   (filtering.py:180,208-213)
    const double s = (double)Hl * ((double)sigma / (double)std::min(ctx->H, ctx->W));
    NotchHost hst;
    design_notch(n, s, ctx->notch_eps, hst);
    auto pad32 = [](size_t v) { return (v + 31) & ~(size_t)31; };  // 128-byte aligned sub-tables
    const size_t o_te = 0, o_to = o_te + pad32(hst.te.size()), o_T1 = o_to + pad32(hst.to.size()),
                 o_T2 = o_T1 + pad32(hst.T1f.size()), total = o_T2 + pad32(hst.T2.size());
    std::vector<float> host(total, 0.f);
    std::copy(hst.te.begin(), hst.te.end(), host.begin() + o_te);
    std::copy(hst.to.begin(), hst.to.end(), host.begin() + o_to);
    std::copy(hst.T1f.begin(), hst.T1f.end(), host.begin() + o_T1);
    std::copy(hst.T2.begin(), hst.T2.end(), host.begin() + o_T2);
    if (D.d_buf) {
        CK(ctx, cudaStreamSynchronize(ctx->s_comp));
        cudaFree(D.d_buf);
        D.d_buf = nullptr;
    }
    CK(ctx, cudaMalloc(&D.d_buf, total * sizeof(float)));
    CK(ctx, cudaMemcpyAsync(D.d_buf, host.data(), total * sizeof(float), cudaMemcpyHostToDevice, ctx->s_comp));
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    D.nt.te = D.d_buf + o_te;
    D.nt.to = D.d_buf + o_to;
    D.nt.T1 = D.d_buf + o_T1;
    D.nt.T2 = D.d_buf + o_T2;
    D.nt.ntap_e = hst.ntap_e;
    D.nt.ue_lo = hst.ue_lo;
    D.nt.ntap_o = hst.ntap_o;
    D.nt.uo_lo = hst.uo_lo;
    D.nt.J = hst.J;
    D.nt.Jpad = hst.Jpad;
    D.sigma = sigma;
    D.eps = ctx->notch_eps;
    {
        MmaHost mh;
        build_mma_host(n, hst, mh);
        if (D.d_mma) {
            cudaFree(D.d_mma);
            D.d_mma = nullptr;
        }
        auto pad256 = [](size_t v) { return (v + 255) & ~(size_t)255; };
        const size_t b_tr = pad256(mh.tr.size() * 2), b_t1 = pad256(mh.T1f.size() * 2), b_t2 = pad256(mh.T2f.size() * 2);
        const size_t b_fq = pad256(mh.fq.size() * 4);
        CK(ctx, cudaMalloc(&D.d_mma, b_tr + b_t1 + b_t2 + b_fq + 256));
        CK(ctx, cudaMemcpyAsync(D.d_mma + b_tr + b_t1 + b_t2, mh.fq.data(), mh.fq.size() * 4, cudaMemcpyHostToDevice, ctx->s_comp));
        CK(ctx, cudaMemcpyAsync(D.d_mma, mh.tr.data(), mh.tr.size() * 2, cudaMemcpyHostToDevice, ctx->s_comp));
        if (!mh.T1f.empty())
            CK(ctx, cudaMemcpyAsync(D.d_mma + b_tr, mh.T1f.data(), mh.T1f.size() * 2, cudaMemcpyHostToDevice, ctx->s_comp));
        if (!mh.T2f.empty())
            CK(ctx, cudaMemcpyAsync(D.d_mma + b_tr + b_t1, mh.T2f.data(), mh.T2f.size() * 2, cudaMemcpyHostToDevice, ctx->s_comp));
        CK(ctx, cudaStreamSynchronize(ctx->s_comp));
        D.mm = mh.cfg;
        D.mm.tr = reinterpret_cast<const __half*>(D.d_mma);
        D.mm.T1f = reinterpret_cast<const uint4*>(D.d_mma + b_tr);
        D.mm.T2f = reinterpret_cast<const uint4*>(D.d_mma + b_tr + b_t1);
        D.mm.fq = reinterpret_cast<const uint4*>(D.d_mma + b_tr + b_t1 + b_t2);
    }
    {
        const UmmaGeom g = umma_geom(n);
        if (D.d_umma) {
            cudaFree(D.d_umma);
            D.d_umma = nullptr;
        }
        if (g.ok) {
            UmmaHost uh;
            build_umma_host(n, s, g, uh);
            if (g.smem_fixed + uh.bytes <= kUmmaSmemMax) {
                CK(ctx, cudaMalloc(&D.d_umma, uh.bytes));
                CK(ctx, cudaMemcpyAsync(D.d_umma, uh.tabs.data(), uh.bytes, cudaMemcpyHostToDevice, ctx->s_comp));
                CK(ctx, cudaStreamSynchronize(ctx->s_comp));
                D.um = uh.cfg;
                D.um.tables = D.d_umma;
            }
        }
    }
    return 0;
}

int build_taps(dstr_ctx* ctx, int level, float sigma_cells, float sigma_nocells) {
    int rc = build_taps_cfg(ctx, level, 0, sigma_nocells);
    if (rc) return rc;
    return build_taps_cfg(ctx, level, 1, sigma_cells);
}

cudaEvent_t get_event(dstr_ctx* ctx) {
    if (!ctx->ev_pool.empty()) {
        cudaEvent_t e = ctx->ev_pool.back();
        ctx->ev_pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
}

struct ScopedTimer {
    dstr_ctx* ctx;
    int id;
    cudaEvent_t a = nullptr;
    ScopedTimer(dstr_ctx* c, int i) : ctx(c), id(i) {
        if (ctx->profiling) {
            a = get_event(ctx);
            cudaEventRecord(a, ctx->s_comp);
        }
    }
    ~ScopedTimer() {
        if (ctx->profiling) {
            cudaEvent_t b = get_event(ctx);
            cudaEventRecord(b, ctx->s_comp);
            ctx->spans.push_back({id, a, b});
        }
    }
};

void resolve_timers(dstr_ctx* ctx) {
    for (auto& sp : ctx->spans) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) ctx->timers[sp.id] += ms;
        ctx->ev_pool.push_back(sp.a);
        ctx->ev_pool.push_back(sp.b);
    }
    ctx->spans.clear();
}

template <int EPL>
int launch_filter(dstr_ctx* ctx, const FilterLevelArgs& fa, int Z, size_t smem, const DispatchParams& dp,
                  cudaStream_t st) {
    smem += sizeof(unsigned) * FR_ROWS * EPL;  // mask bit words
    // opt in to the full dynamic shared memory once per instantiation and device (never lowered, so
    // concurrent contexts in other host threads cannot invalidate each other's launches)
    static std::mutex mtx;
    static bool done[64] = {};
    {
        std::lock_guard<std::mutex> lk(mtx);
        const int dev = ctx->device & 63;
        if (!done[dev]) {
            CK(ctx, cudaFuncSetAttribute(filter_rows_kernel<EPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         227 * 1024));
            done[dev] = true;
        }
    }
    dim3 grid((fa.Hl + FR_ROWS - 1) / FR_ROWS, Z);
    filter_rows_kernel<EPL><<<grid, FR_THREADS, smem, st>>>(fa, ctx->d_pstat, dp);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return 0;
}

// Everything one chunk pass needs to launch its kernels.
// The row-filter kernels load a row as EPL x 32 lanes without a bounds test (EPL = template width >= ceil(W_l / 32)): the
// over-read of the last row of the last plane must stay inside the allocation.  Largest template: 65 (tcgen05: 33).
constexpr int kHSlack = 32 * 65 + 32;
static_assert(kHSlack >= 32 * 65, "row-filter over-read slack");

constexpr int kModeForceCells = 2;  // internal: every plane uses the `cells` tables (second band of the dual-band mode)

struct Pass {
    dstr_ctx* ctx;
    const void* d_in;
    void* d_out;
    int in_dtype, out_dtype, z, L, flags, stat_stride;
    size_t level_stride;
    DispatchParams dp;
};

int launch_analysis(const Pass& P, int l, cudaStream_t st) {
    dstr_ctx* ctx = P.ctx;
    const int H = ctx->H, W = ctx->W, z = P.z;
    const LevelGeom& gs = ctx->geom[l - 1];
    const LevelGeom& go = ctx->geom[l];
    const int cols_per_block = AN_OXW * AN_WX, rows_per_block = AN_TOY * AN_WY;
    dim3 grid((go.W + cols_per_block - 1) / cols_per_block, (go.H + rows_per_block - 1) / rows_per_block, z);
    LevelStat* ls = ctx->d_lstat + (size_t)(l - 1) * P.level_stride;
    const bool stats = P.dp.mode == 1;
#define LAUNCH_AN1(IN_T, ST, VEC)                                                                       \
    analysis_kernel<IN_T, true, ST, VEC><<<grid, AN_THREADS, 0, st>>>(                                   \
        (const IN_T*)P.d_in, H, W, W, (size_t)H * W, ctx->d_A[1], ctx->d_H[1], go.H, go.W, go.pitch,     \
        go.pstride, ls, P.stat_stride, ctx->d_pstat, ctx->fg_thr32)
#define LAUNCH_AN1V(IN_T, ST)                \
    do {                                     \
        if (vec) LAUNCH_AN1(IN_T, ST, true); \
        else LAUNCH_AN1(IN_T, ST, false);    \
    } while (0)
    // aligned two-element loads need an even width, pitch and plane stride
    const bool vec = (gs.W % 2 == 0) && (gs.W >= 8) && (gs.H >= 8) && (gs.pitch % 2 == 0) && (gs.pstride % 2 == 0);
    const size_t esz = P.in_dtype == DSTR_U16 ? 2 : 4;
    if (l == 1 && ctx->use_tma && W >= AT_WIN && (W * esz) % 16 == 0 && ((size_t)H * W * esz) % 16 == 0 && H >= 8 &&
        ((uintptr_t)P.d_in % 16) == 0) {
        const int rows_per_cta = 126;  // multiple of 3; 2.4 % halo rows
        dim3 gt((go.W + AT_OXB - 1) / AT_OXB, (go.H + rows_per_cta - 1) / rows_per_cta, z);
#define LAUNCH_TMA(IN_T, ST)                                                                              \
    analysis_tma_kernel<IN_T, ST><<<gt, AT_THREADS, 0, st>>>((const IN_T*)P.d_in, H, W, (size_t)H * W, ctx->d_A[1], \
                                                              ctx->d_H[1], go.H, go.W, go.pitch, go.pstride, ls,  \
                                                              P.stat_stride, ctx->d_pstat, ctx->fg_thr32, rows_per_cta)
        if (P.in_dtype == DSTR_U16) {
            if (stats) LAUNCH_TMA(uint16_t, true);
            else LAUNCH_TMA(uint16_t, false);
        } else {
            if (stats) LAUNCH_TMA(float, true);
            else LAUNCH_TMA(float, false);
        }
#undef LAUNCH_TMA
    } else if (l == 1) {
        if (P.in_dtype == DSTR_U16) {
            if (stats) LAUNCH_AN1V(uint16_t, true);
            else LAUNCH_AN1V(uint16_t, false);
        } else {
            if (stats) LAUNCH_AN1V(float, true);
            else LAUNCH_AN1V(float, false);
        }
    } else if (vec) {
        analysis_kernel<float, false, false, true><<<grid, AN_THREADS, 0, st>>>(
            ctx->d_A[l - 1], gs.H, gs.W, gs.pitch, gs.pstride, ctx->d_A[l], ctx->d_H[l], go.H, go.W, go.pitch,
            go.pstride, ls, P.stat_stride, ctx->d_pstat, ctx->fg_thr32);
    } else {
        analysis_kernel<float, false, false, false><<<grid, AN_THREADS, 0, st>>>(
            ctx->d_A[l - 1], gs.H, gs.W, gs.pitch, gs.pstride, ctx->d_A[l], ctx->d_H[l], go.H, go.W, go.pitch,
            go.pstride, ls, P.stat_stride, ctx->d_pstat, ctx->fg_thr32);
    }
#undef LAUNCH_AN1V
#undef LAUNCH_AN1
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return 0;
}

int launch_hist(const Pass& P, int l, cudaStream_t st) {
    dstr_ctx* ctx = P.ctx;
    const LevelGeom& g = ctx->geom[l];
    const int nblk = std::max(1, (g.H * (g.pitch / 4) + 2047) / 2048);  // ~8 quads per thread
    dim3 grid(nblk, P.z);
    hist_kernel<<<grid, 256, 0, st>>>(ctx->d_H[l], g.H, g.W, g.pitch, g.pstride,
                                      ctx->d_lstat + (size_t)(l - 1) * P.level_stride, P.stat_stride);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return 0;
}

// Otsu thresholds of levels l0 .. l0+nl-1
int launch_otsu(const Pass& P, int l0, int nl, cudaStream_t st) {
    dstr_ctx* ctx = P.ctx;
    const bool stack = (P.flags & DSTR_FLAG_STACK_OTSU) != 0;
    dim3 grid(stack ? 1 : P.z, nl);
    otsu_kernel<<<grid, 32, 0, st>>>(ctx->d_lstat + (size_t)(l0 - 1) * P.level_stride, P.level_stride,
                                     P.stat_stride, ctx->d_pstat, P.dp);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return 0;
}

template <int EPL>
int launch_umma(dstr_ctx* ctx, const UmmaLevelArgs& ua, int grid, size_t smem, const DispatchParams& dp, cudaStream_t st) {
    static std::mutex mtx;
    static bool done[64] = {};
    {
        std::lock_guard<std::mutex> lk(mtx);
        const int dev = ctx->device & 63;
        if (!done[dev]) {
            CK(ctx, cudaFuncSetAttribute(notch_umma_kernel<EPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUmmaSmemMax));
            done[dev] = true;
        }
    }
    notch_umma_kernel<EPL><<<grid, UM_THREADS, smem, st>>>(ua, ctx->d_pstat, dp);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return 0;
}

// Row filter of level l on the tensor cores; returns -1000 when the geometry does not qualify.
int launch_filter_umma(const Pass& P, int l, cudaStream_t st) {
    dstr_ctx* ctx = P.ctx;
    const LevelGeom& g = ctx->geom[l];
    const TapTable& T = ctx->taps[l];
    const UmmaGeom ug = umma_geom(g.W);
    if (!ug.ok || !T.cfg[0].d_umma || !T.cfg[1].d_umma) return -1000;
    if ((g.pitch % 4) != 0 || (g.pstride % 4) != 0) return -1000;
    UmmaLevelArgs ua;
    ua.cH = ctx->d_H[l];
    ua.Hl = g.H;
    ua.n = g.W;
    ua.pitch = g.pitch;
    ua.pstride = g.pstride;
    ua.lstat = ctx->d_lstat + (size_t)(l - 1) * P.level_stride;
    ua.stat_stride = P.stat_stride;
    ua.nh = ug.nh;
    ua.nout = ug.nout;
    ua.P = ug.P;
    ua.Nt = ug.Nt;
    ua.NC = ug.NC;
    ua.Kpad = ug.Kpad;
    ua.items_per_plane = (g.H + UM_ROWS - 1) / UM_ROWS;
    ua.rows_per_item = std::min(UM_ROWS, (((g.H + ua.items_per_plane - 1) / ua.items_per_plane) + 7) & ~7);
    ua.items_per_plane = (g.H + ua.rows_per_item - 1) / ua.rows_per_item;
    ua.n_items = ua.items_per_plane * P.z;
    ua.mw = ug.mw;
    ua.vec_ok = (reinterpret_cast<uintptr_t>(ua.cH) % 16 == 0) ? 1 : 0;
    size_t tab_max = 0;
    for (int c = 0; c < 2; ++c) {
        ua.cfg[c] = T.cfg[c].um;
        const size_t tb = (size_t)ua.cfg[c].tr_bytes * (ua.cfg[c].r3 ? 2 : 1) + (size_t)2 * ua.cfg[c].tb_bytes;
        tab_max = std::max(tab_max, (tb + 1023) & ~(size_t)1023);
    }
    ua.tab_max = (int)tab_max;
    const size_t smem = tab_max + ug.smem_fixed;
    if (smem > kUmmaSmemMax) return -1000;
    const int grid = std::min(ua.n_items, ctx->sm_count);
    ua.scratch_stride = (size_t)4 * ug.NC * UM_CHUNK_BYTES;
    const size_t need = 2 * ua.scratch_stride * (size_t)ctx->sm_count;  // two item buffers per CTA
    if (ctx->um_scratch_bytes[l] < need) {
        // (re)allocation only happens on the first pass over a new geometry: drain everything first
        CK(ctx, cudaDeviceSynchronize());
        if (ctx->d_um_scratch[l]) cudaFree(ctx->d_um_scratch[l]);
        ctx->d_um_scratch[l] = nullptr;
        ctx->um_scratch_bytes[l] = 0;
        CK(ctx, cudaMalloc(&ctx->d_um_scratch[l], need));
        // rows / chunks never written must stay finite.  Ordered on the launching stream: the engine's streams are
        // non-blocking, so a memset on the legacy default stream could still be running when the kernel starts
        CK(ctx, cudaMemsetAsync(ctx->d_um_scratch[l], 0, need, st));
        ctx->um_scratch_bytes[l] = need;
    }
    ua.scratch = ctx->d_um_scratch[l];
    ua.prof = nullptr;
    static const bool um_prof = env_or("DSTR_UMMA_PROF", 0.0) != 0.0;
    long long* d_prof = nullptr;
    if (um_prof && l == 1) {
        CK(ctx, cudaMalloc(&d_prof, sizeof(long long) * 256 * grid));
        CK(ctx, cudaMemsetAsync(d_prof, 0, sizeof(long long) * 256 * grid, st));
        ua.prof = d_prof;
    }
    const int epl = (g.W + 31) / 32;
    int rc;
    if (epl <= 5) rc = launch_umma<5>(ctx, ua, grid, smem, P.dp, st);
    else if (epl <= 9) rc = launch_umma<9>(ctx, ua, grid, smem, P.dp, st);
    else if (epl <= 17) rc = launch_umma<17>(ctx, ua, grid, smem, P.dp, st);
    else rc = launch_umma<33>(ctx, ua, grid, smem, P.dp, st);
    if (d_prof) {
        // diagnostic build path (DSTR_UMMA_PROF=1): per-role wait / total cycles of the level-1 launch, CTA 0 and the mean
        std::vector<long long> h((size_t)256 * grid);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), d_prof, sizeof(long long) * h.size(), cudaMemcpyDeviceToHost);
        cudaFree(d_prof);
        for (int w = 0; w < UM_THREADS / 32; ++w) {
            const char* role = w == 0 ? "tma" : w == 1 ? "mma" : (w >= 4 && w < 8) ? "epi" : "prep";
            if (w > 8 && w < UM_THREADS / 32 - 1) continue;  // the prep warps behave alike
            double v[6] = {0, 0, 0, 0, 0, 0};
            for (int b = 0; b < grid; ++b)
                for (int k = 0; k < 6; ++k) v[k] += (double)h[((size_t)b * 32 + w) * 8 + k] / grid;
            fprintf(stderr, "[umma prof L1] warp %2d %-4s wait %9.0f total %9.0f | pass1 %9.0f median %9.0f pass2 %9.0f | items %.1f\n",
                    w, role, v[0], v[3], v[1], v[2], v[4], v[5]);
        }
    }
    return rc;
}

template <int EPL, int NT>
int launch_rows_mma(dstr_ctx* ctx, const RowsMmaArgs& ra, int Z, size_t smem, const DispatchParams& dp, cudaStream_t st) {
    static std::mutex mtx;
    static bool done[64] = {};
    {
        std::lock_guard<std::mutex> lk(mtx);
        const int dev = ctx->device & 63;
        if (!done[dev]) {
            CK(ctx, cudaFuncSetAttribute(filter_rows_mma_kernel<EPL, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
            done[dev] = true;
        }
    }
    constexpr int rows = 4 * NT;
    dim3 grid((ra.Hl + rows - 1) / rows, Z);
    filter_rows_mma_kernel<EPL, NT><<<grid, 32 * rows, smem, st>>>(ra, ctx->d_pstat, dp);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return 0;
}

// Row filter of level l with the contractions on mma.sync; returns -1000 when the band does not qualify.
int launch_filter_rows_mma(const Pass& P, int l, cudaStream_t st) {
    dstr_ctx* ctx = P.ctx;
    const LevelGeom& g = ctx->geom[l];
    const TapTable& T = ctx->taps[l];
    if (!T.cfg[0].d_mma || !T.cfg[1].d_mma || g.W < 16) return -1000;
    RowsMmaArgs ra;
    ra.cH = ctx->d_H[l];
    ra.Hl = g.H;
    ra.Wl = g.W;
    ra.pitch = g.pitch;
    ra.pstride = g.pstride;
    ra.lstat = ctx->d_lstat + (size_t)(l - 1) * P.level_stride;
    ra.stat_stride = P.stat_stride;
    ra.cfg[0] = T.cfg[0].mm;
    ra.cfg[1] = T.cfg[1].mm;
    ra.nh = g.W / 2;
    ra.nseg16 = (ra.nh + 1 + 15) / 16;
    auto len_of = [&](int S) {  // entries the MMAs read, padded so that (len / 2) = 4 (mod 8) words: the 8 operand columns of a
        int len = 16 * ra.nseg16 + 16 * S;  // warp start 4 * odd banks apart, i.e. on 8 different 4-bank groups (conflict-free)
        len += (8 - len % 16 + 16) % 16;
        return len;
    };
    ra.len_e = len_of(std::max(ra.cfg[0].S_e, ra.cfg[1].S_e));
    ra.len_o = len_of(std::max(ra.cfg[0].S_o, ra.cfg[1].S_o));
    ra.trlen_e_max = std::max(ra.cfg[0].trlen_e, ra.cfg[1].trlen_e);
    ra.trlen_o_max = std::max(ra.cfg[0].trlen_o, ra.cfg[1].trlen_o);
    ra.Jpad_max = std::max(ra.cfg[0].Jpad, ra.cfg[1].Jpad);
    ra.S_e_max = std::max(ra.cfg[0].S_e, ra.cfg[1].S_e);
    ra.S_o_max = std::max(ra.cfg[0].S_o, ra.cfg[1].S_o);
    const int epl = (g.W + 31) / 32;
    const int tmpl = epl <= 2 ? 2 : epl <= 5 ? 5 : epl <= 9 ? 9 : epl <= 17 ? 17 : epl <= 33 ? 33 : 65;
    if (epl > 65) return -1000;
    if (32 * tmpl > g.W + kHSlack) return fail(ctx, DSTR_E_STATE, "row-filter over-read exceeds the band slack");
    auto smem_of = [&](int rows) {
        return (size_t)16 * 28 * (ra.S_e_max + ra.S_o_max) + 2 * ((size_t)rows * 2 * (ra.len_e + ra.len_o)) +
               8 * (size_t)std::max(rows * ra.Jpad_max, 4) + 2 * (size_t)rows * 2 * (ra.Jpad_max + 8) + 4 * (size_t)rows * tmpl;
    };
    static const size_t extra_smem = (size_t)env_or("DSTR_RM_EXTRA_SMEM", 0.0);  // occupancy experiments
    // 8 rows per block (two n8 tiles per A fragment) while at least two blocks fit an SM; DSTR_RM_ROWS=4 forces the 4-row form
    static const int force_rows = (int)env_or("DSTR_RM_ROWS", 0.0);
    const int nt = (force_rows == 4 || ctx->row_filter == 2 || smem_of(8) + extra_smem > 110 * 1024 || g.H < 8) ? 1 : 2;
    const size_t smem = smem_of(4 * nt) + extra_smem;
    if (smem > 227 * 1024) return -1000;
    {
        // L2 prefetch one wave of resident blocks ahead; DSTR_FILTER_PREFETCH overrides (0 = off)
        static const int pf = (int)env_or("DSTR_FILTER_PREFETCH", -1.0);
        ra.prefetch_blocks = pf >= 0 ? pf : (8 / nt) * ctx->sm_count;
        if ((g.pitch * 4) % 16 != 0) ra.prefetch_blocks = 0;
    }
#define RM_DISPATCH(E)                                                        \
    do {                                                                      \
        if (nt == 2) return launch_rows_mma<E, 2>(ctx, ra, P.z, smem, P.dp, st); \
        return launch_rows_mma<E, 1>(ctx, ra, P.z, smem, P.dp, st);           \
    } while (0)
    if (tmpl == 2) RM_DISPATCH(2);
    if (tmpl == 5) RM_DISPATCH(5);
    if (tmpl == 9) RM_DISPATCH(9);
    if (tmpl == 17) RM_DISPATCH(17);
    if (tmpl == 33) RM_DISPATCH(33);
    RM_DISPATCH(65);
#undef RM_DISPATCH
}

int launch_filter_level(const Pass& P, int l, cudaStream_t st) {
    dstr_ctx* ctx = P.ctx;
    if (ctx->use_umma) {
        const int rcu = launch_filter_umma(P, l, st);
        if (rcu != -1000) return rcu;
    }
    if (ctx->row_filter >= 1) {
        const int rcm = launch_filter_rows_mma(P, l, st);
        if (rcm != -1000) return rcm;
    }
    const LevelGeom& g = ctx->geom[l];
    const TapTable& T = ctx->taps[l];
    FilterLevelArgs fa;
    fa.cH = ctx->d_H[l];
    fa.Hl = g.H;
    fa.Wl = g.W;
    fa.pitch = g.pitch;
    fa.pstride = g.pstride;
    fa.lstat = ctx->d_lstat + (size_t)(l - 1) * P.level_stride;
    fa.stat_stride = P.stat_stride;
    fa.nt[0] = T.cfg[0].nt;
    fa.nt[1] = T.cfg[1].nt;
    fa.nh = g.W / 2;
    fa.nhp8 = (fa.nh + 1 + 7) & ~7;
    fa.n_pad8 = (g.W + 7) & ~7;
    fa.ntap_e_max = std::max(fa.nt[0].ntap_e, fa.nt[1].ntap_e);
    fa.ntap_o_max = std::max(fa.nt[0].ntap_o, fa.nt[1].ntap_o);
    fa.Jpad_max = std::max(fa.nt[0].Jpad, fa.nt[1].Jpad);
    fa.nhp64 = (fa.nhp8 + 63) & ~63;
    {
        // one wave of resident blocks ahead (8 blocks per SM); DSTR_FILTER_PREFETCH overrides (0 = off)
        static const int pf = (int)env_or("DSTR_FILTER_PREFETCH", -1.0);
        fa.prefetch_blocks = pf >= 0 ? pf : 8 * ctx->sm_count;
        if ((g.pitch * 4) % 16 != 0) fa.prefetch_blocks = 0;
    }
    fa.vec_ok = (g.pitch % 4 == 0) && (g.pstride % 4 == 0) && (reinterpret_cast<uintptr_t>(fa.cH) % 16 == 0);
    fa.ablate = 0;
#ifdef DSTR_ABLATION
    fa.ablate = (int)env_or("DSTR_ABLATE", 0.0);
#endif
    // row strides = 8 (mod 16) words: with the 9-word segment stride the 8 segments x 4 rows of a warp
    // fall into 32 distinct banks
    auto bank_pad = [](int v) { return v + ((8 - v) & 15); };
    fa.xlen_e_phys = bank_pad((fa.nhp8 + fa.ntap_e_max) / 8 * 9);
    fa.xlen_o_phys = bank_pad((fa.nhp8 + fa.ntap_o_max) / 8 * 9);
    const size_t smem = sizeof(float) * ((size_t)fa.ntap_e_max + fa.ntap_o_max +
                                         (size_t)FR_ROWS * (fa.xlen_e_phys + fa.xlen_o_phys) +
                                         (size_t)2 * std::max(FR_ROWS * fa.Jpad_max, 128) +  // 64-bit accumulators (>= 1 KB: reused as store staging)
                                         (size_t)FR_ROWS * fa.Jpad_max +                     // float copy
                                         (size_t)8 * FR_ROWS);                // row padding of the float copy
    if (smem > 227 * 1024) return fail(ctx, DSTR_E_SHAPE, "row too long for filter kernel");
    const int epl = (g.W + 31) / 32;
    if (epl <= 2) return launch_filter<2>(ctx, fa, P.z, smem, P.dp, st);
    if (epl <= 5) return launch_filter<5>(ctx, fa, P.z, smem, P.dp, st);
    if (epl <= 9) return launch_filter<9>(ctx, fa, P.z, smem, P.dp, st);
    if (epl <= 17) return launch_filter<17>(ctx, fa, P.z, smem, P.dp, st);
    if (epl <= 33) return launch_filter<33>(ctx, fa, P.z, smem, P.dp, st);
    if (epl <= 65) return launch_filter<65>(ctx, fa, P.z, smem, P.dp, st);
    return fail(ctx, DSTR_E_SHAPE, "row too long for filter kernel");
}

// dA_{l-1} = idwt2(dA_l, dH_l)
int launch_synth(const Pass& P, int l, cudaStream_t st) {
    dstr_ctx* ctx = P.ctx;
    const LevelGeom& g = ctx->geom[l];
    const LevelGeom& go = ctx->geom[l - 1];
    EpilogueArgs ep0 = {};
    dim3 grid((go.W + SY_TX * SY_WARPS - 1) / (SY_TX * SY_WARPS), (go.H + SY_TY - 1) / SY_TY, P.z);
    const float* dA = (l == P.L) ? nullptr : ctx->d_A[l];
    synth_kernel<false, float, float, false><<<grid, SY_THREADS, 0, st>>>(
        dA, ctx->d_H[l], g.H, g.W, g.pitch, g.pstride, ctx->d_A[l - 1], go.H, go.W, go.pitch, go.pstride, nullptr,
        nullptr, 0, ep0);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return 0;
}

int launch_final(const Pass& P, cudaStream_t st) {
    dstr_ctx* ctx = P.ctx;
    const int H = ctx->H, W = ctx->W, L = P.L;
    EpilogueArgs ep;
    ep.inv_flat = ctx->d_flat;  // stored as the correctly rounded reciprocal
    ep.dark = ctx->d_dark;
    ep.shadow = (P.flags & DSTR_FLAG_SHADOW) ? 1 : 0;
    ep.expm1 = (P.flags & DSTR_FLAG_EXPM1) ? 1 : 0;
    const LevelGeom& g = ctx->geom[L > 0 ? 1 : 0];
    const float* dA = (L >= 2) ? ctx->d_A[1] : nullptr;
    const float* dH = (L >= 1) ? ctx->d_H[1] : nullptr;
    dim3 grid((W + SY_TX * SY_WARPS - 1) / (SY_TX * SY_WARPS), (H + SY_TY - 1) / SY_TY, P.z);
    const size_t ps = (size_t)H * W;
#define LAUNCH_FINAL_V(IN_T, OUT_T, VEC)                                                                   \
    synth_kernel<true, IN_T, OUT_T, VEC><<<grid, SY_THREADS, 0, st>>>(                                     \
        dA, dH, g.H, g.W, g.pitch, g.pstride, nullptr, H, W, W, ps, (const IN_T*)P.d_in, (OUT_T*)P.d_out, \
        ps, ep)
#define LAUNCH_FINAL(IN_T, OUT_T)                    \
    do {                                             \
        if (vecf) LAUNCH_FINAL_V(IN_T, OUT_T, true); \
        else LAUNCH_FINAL_V(IN_T, OUT_T, false);     \
    } while (0)
    const bool vecf = (W % 2 == 0) && (W >= 2) && (ps % 2 == 0);
    if (P.in_dtype == DSTR_U16 && P.out_dtype == DSTR_U16) LAUNCH_FINAL(uint16_t, uint16_t);
    else if (P.in_dtype == DSTR_U16 && P.out_dtype == DSTR_F32) LAUNCH_FINAL(uint16_t, float);
    else if (P.in_dtype == DSTR_F32 && P.out_dtype == DSTR_U16) LAUNCH_FINAL(float, uint16_t);
    else LAUNCH_FINAL(float, float);
#undef LAUNCH_FINAL_V
#undef LAUNCH_FINAL
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return 0;
}

#define RC(call)             \
    do {                     \
        int rc_ = (call);    \
        if (rc_) return rc_; \
    } while (0)

// Process z planes resident on the device (d_in -> d_out).
//
// Dependency graph of one pass:  A_1 -> A_2 -> ... -> A_L (analysis chain);  per level
// A_l -> hist_l -> otsu_l -> filter_l;  synthesis chain S_L -> ... -> S_2 -> final, where S_l needs
// filter_l.  The per-level branches are independent of the analysis chain below them, so in the
// normal mode they run on side streams (the long level-1 row filter overlaps the whole deep
// pyramid); the compute stream carries the two chains and joins the branches by events.  With
// profiling or a debug stop the pass is issued stage by stage on the compute stream alone.
int process_device(dstr_ctx* ctx, const void* d_in, int in_dtype, void* d_out, int out_dtype, int z,
                   const dstr_params& cells, const dstr_params& no_cells, float high_int, int mode,
                   int flags, int L) {
    cudaStream_t st = ctx->s_comp;
    Pass P;
    P.ctx = ctx;
    P.d_in = d_in;
    P.d_out = d_out;
    P.in_dtype = in_dtype;
    P.out_dtype = out_dtype;
    P.z = z;
    P.L = L;
    P.flags = flags;
    P.stat_stride = (flags & DSTR_FLAG_STACK_OTSU) ? 0 : 1;
    P.level_stride = (size_t)ctx->zcap;
    P.dp.max_thr_cells = cells.max_threshold;
    P.dp.max_thr_nocells = no_cells.max_threshold;
    P.dp.high_int = high_int;
    P.dp.mode = (mode == DSTR_MODE_DISPATCH) ? 1 : (mode == kModeForceCells ? 2 : 0);
    P.dp.notch_only = (flags & DSTR_FLAG_NOTCH_ONLY) ? 1 : 0;
    ctx->last_levels = L;
    ctx->last_z = z;
    ctx->last_dp = P.dp;

    ScopedTimer t_all(ctx, 7);
    if (L > 0) CK(ctx, cudaMemsetAsync(ctx->d_lstat, 0, sizeof(LevelStat) * P.level_stride * L, st));
    CK(ctx, cudaMemsetAsync(ctx->d_pstat, 0, sizeof(PlaneStat) * ctx->zcap, st));

    const bool staged = ctx->profiling || ctx->debug_stop != DSTR_STAGE_NONE || !ctx->overlap;
    if (staged) {
        for (int l = 1; l <= L; ++l) {
            ScopedTimer t(ctx, l == 1 ? 0 : 1);
            RC(launch_analysis(P, l, st));
        }
        if (ctx->debug_stop == DSTR_STAGE_ANALYSIS) return 0;
        if (L > 0) {
            {
                ScopedTimer t(ctx, 2);
                if (!P.dp.notch_only)
                    for (int l = 1; l <= L; ++l) RC(launch_hist(P, l, st));
            }
            {
                ScopedTimer t(ctx, 3);
                RC(launch_otsu(P, 1, L, st));
            }
            if (ctx->debug_stop == DSTR_STAGE_OTSU) return 0;
            {
                ScopedTimer t(ctx, 4);
                {
                    ScopedTimer t1(ctx, 8);
                    RC(launch_filter_level(P, 1, st));
                }
                for (int l = 2; l <= L; ++l) RC(launch_filter_level(P, l, st));
            }
            if (ctx->debug_stop == DSTR_STAGE_FILTER) return 0;
            {
                ScopedTimer t(ctx, 5);
                for (int l = L; l >= 2; --l) RC(launch_synth(P, l, st));
            }
            if (ctx->debug_stop == DSTR_STAGE_SYNTH) return 0;
        }
        ScopedTimer t(ctx, 6);
        RC(launch_final(P, st));
        return 0;
    }

    // ---- overlapped issue ---------------------------------------------------------------------------
    for (int l = 1; l <= L; ++l) {
        RC(launch_analysis(P, l, st));
        CK(ctx, cudaEventRecord(ctx->ev_an[l], st));
        cudaStream_t side = ctx->s_side[(l - 1) % kSideStreams];
        CK(ctx, cudaStreamWaitEvent(side, ctx->ev_an[l], 0));
        if (!P.dp.notch_only) RC(launch_hist(P, l, side));
        RC(launch_otsu(P, l, 1, side));
        RC(launch_filter_level(P, l, side));
        CK(ctx, cudaEventRecord(ctx->ev_flt[l], side));
    }
    for (int l = L; l >= 2; --l) {
        CK(ctx, cudaStreamWaitEvent(st, ctx->ev_flt[l], 0));
        RC(launch_synth(P, l, st));
    }
    if (L >= 1) CK(ctx, cudaStreamWaitEvent(st, ctx->ev_flt[1], 0));
    RC(launch_final(P, st));
    return 0;
}
#undef RC

int launch_downscale(dstr_ctx* ctx, const uint16_t* in, int Z, int H, int W, uint16_t* out, cudaStream_t st) {
    const int Zo = Z / 2, Ho = H / 2, Wo = W / 2;
    const size_t n = (size_t)Zo * Ho * Wo;
    if (n == 0) return 0;
    const int blocks = (int)std::min<size_t>((n + 255) / 256, (size_t)148 * 32);
    downscale2x_kernel<<<blocks, 256, 0, st>>>(in, Z, H, W, out, Zo, Ho, Wo);
    ctx->launches++;
    CK(ctx, cudaGetLastError());
    return 0;
}

size_t dtype_size(int dt) { return dt == DSTR_U16 ? 2 : 4; }

bool is_device_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

int ensure_stage(dstr_ctx* ctx, int planes) {
    const size_t need = (size_t)planes * ctx->H * ctx->W * 4;
    if (ctx->stage_bytes >= need) return 0;
    for (int i = 0; i < 2; ++i) {
        if (ctx->d_in[i]) cudaFree(ctx->d_in[i]);
        if (ctx->d_out[i]) cudaFree(ctx->d_out[i]);
        ctx->d_in[i] = ctx->d_out[i] = nullptr;
    }
    ctx->stage_bytes = 0;
    for (int i = 0; i < 2; ++i) {
        CK(ctx, cudaMalloc(&ctx->d_in[i], need));
        CK(ctx, cudaMalloc(&ctx->d_out[i], need));
    }
    ctx->stage_bytes = need;
    return 0;
}

int ensure_pyramid_stage(dstr_ctx* ctx, int planes) {
    const size_t need = (size_t)(planes / 2 + 1) * (ctx->H / 2 + 1) * (ctx->W / 2 + 1) * sizeof(uint16_t);
    if (ctx->pyr_bytes >= need) return 0;
    for (int l = 0; l < 2; ++l)
        for (int b = 0; b < 2; ++b) {
            if (ctx->d_pyr[l][b]) cudaFree(ctx->d_pyr[l][b]);
            ctx->d_pyr[l][b] = nullptr;
        }
    ctx->pyr_bytes = 0;
    for (int l = 0; l < 2; ++l)
        for (int b = 0; b < 2; ++b) CK(ctx, cudaMalloc(&ctx->d_pyr[l][b], need));
    ctx->pyr_bytes = need;
    return 0;
}

// Fused pyramid levels of a destriped batch that is still resident (next-row f3): level 1 =
// 2x2x2 windowed mean of the uint16 output, level 2 = the same of level 1.  z0 is a multiple of 4.
// Returns through *h1 / *h2 the device staging pointers that still have to be copied to host
// targets (nullptr when the kernel wrote the target directly).
int emit_pyramid(dstr_ctx* ctx, const uint16_t* d_out, int z0, int zn, int b, cudaStream_t st,
                 uint16_t** h1, uint16_t** h2) {
    *h1 = *h2 = nullptr;
    const int H1 = ctx->H / 2, W1 = ctx->W / 2, H2 = H1 / 2, W2 = W1 / 2;
    if (!ctx->pyr_out[0]) return 0;
    uint16_t* t1 = (uint16_t*)ctx->pyr_out[0] + (size_t)(z0 / 2) * H1 * W1;
    uint16_t* w1 = t1;
    if (!is_device_ptr(ctx->pyr_out[0])) {
        w1 = ctx->d_pyr[0][b];
        *h1 = w1;
    }
    int rc = launch_downscale(ctx, d_out, zn, ctx->H, ctx->W, w1, st);
    if (rc) return rc;
    if (!ctx->pyr_out[1]) return 0;
    uint16_t* t2 = (uint16_t*)ctx->pyr_out[1] + (size_t)(z0 / 4) * H2 * W2;
    uint16_t* w2 = t2;
    if (!is_device_ptr(ctx->pyr_out[1])) {
        w2 = ctx->d_pyr[1][b];
        *h2 = w2;
    }
    return launch_downscale(ctx, w1, zn / 2, H1, W1, w2, st);
}

}  // namespace

// =================================================================================================
extern "C" {

const char* dstr_last_error(const dstr_ctx* ctx) {
    if (ctx) return ctx->err.c_str();
    std::lock_guard<std::mutex> lk(g_err_mutex);
    static thread_local std::string copy;
    copy = g_last_error;
    return copy.c_str();
}

int dstr_max_level(int H, int W) {
    if (H <= 0 || W <= 0) return DSTR_E_ARG;
    return std::min(max_level_1d(H), max_level_1d(W));
}

int dstr_level_shape(int H, int W, int level, int* H_l, int* W_l) {
    if (H <= 0 || W <= 0 || level < 0 || !H_l || !W_l) return DSTR_E_ARG;
    int h = H, w = W;
    for (int l = 0; l < level; ++l) {
        h = (h + kFilterTaps - 1) / 2;
        w = (w + kFilterTaps - 1) / 2;
    }
    *H_l = h;
    *W_l = w;
    return 0;
}

float dstr_foreground_threshold(float threshold_mask) { return find_fg_half_threshold(threshold_mask); }
float dstr_foreground_threshold_f32(float threshold_mask) {
    return fg_threshold_f32(find_fg_half_threshold(threshold_mask));
}

int dstr_notch_kernels(int n, double s, double* hp, double* hq) {
    if (n <= 0 || !(s > 0.0) || !hp || !hq) return DSTR_E_ARG;
    std::vector<double> a, b;
    notch_kernels_host(n, s, a, b);
    std::memcpy(hp, a.data(), sizeof(double) * n);
    std::memcpy(hq, b.data(), sizeof(double) * n);
    return 0;
}

int dstr_create(int device, int max_planes, int H, int W, dstr_ctx** out) {
    if (!out || max_planes <= 0 || H <= 0 || W <= 0) return fail(nullptr, DSTR_E_ARG, "dstr_create: bad argument");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0)
        return fail(nullptr, e != cudaSuccess ? (int)e : (int)cudaErrorNoDevice,
                    std::string("dstr_create: no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, DSTR_E_ARG, "dstr_create: bad device index");
    dstr_ctx* ctx = new dstr_ctx();
    ctx->device = device;
    ctx->zcap = max_planes;
    ctx->H = H;
    ctx->W = W;
    ctx->Lmax = std::min(std::min(max_level_1d(H), max_level_1d(W)), kMaxLevels);
    ctx->Lalloc = std::min(kMaxLevels, ctx->Lmax + 4);
    *out = nullptr;
#define CKC(call)                                                                         \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            int rc_ = fail(nullptr, (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_)); \
            dstr_destroy(ctx);                                                            \
            return rc_;                                                                   \
        }                                                                                 \
    } while (0)
    CKC(cudaSetDevice(device));
    CKC(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    ctx->geom[0] = {H, W, W, (size_t)H * W};
    for (int l = 1; l <= ctx->Lalloc; ++l) {
        LevelGeom g;
        g.H = (ctx->geom[l - 1].H + kFilterTaps - 1) / 2;
        g.W = (ctx->geom[l - 1].W + kFilterTaps - 1) / 2;
        g.pitch = (g.W + 3) & ~3;
        g.pstride = (size_t)g.H * g.pitch;
        ctx->geom[l] = g;
        // slack: the synthesis kernel reads columns m+1, m+2 unconditionally (8 floats), the row filter
        // loads whole 32-lane groups of a row for a templated element count (up to 32 x 65 floats past the last row start)
        CKC(cudaMalloc(&ctx->d_A[l], sizeof(float) * (g.pstride * max_planes + 8)));
        CKC(cudaMalloc(&ctx->d_H[l], sizeof(float) * (g.pstride * max_planes + kHSlack)));
    }
    if (ctx->Lalloc > 0) CKC(cudaMalloc(&ctx->d_lstat, sizeof(LevelStat) * (size_t)ctx->Lalloc * max_planes));
    CKC(cudaMalloc(&ctx->d_pstat, sizeof(PlaneStat) * (size_t)max_planes));
    CKC(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    {
        // the compute stream carries the dependency chains: give it the highest priority so that
        // its short kernels are not queued behind the long row filters of the side streams
        int prio_lo = 0, prio_hi = 0;
        CKC(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        CKC(cudaStreamCreateWithPriority(&ctx->s_comp, cudaStreamNonBlocking, prio_hi));
        for (int i = 0; i < kSideStreams; ++i)
            CKC(cudaStreamCreateWithPriority(&ctx->s_side[i], cudaStreamNonBlocking, prio_lo));
        for (int l = 0; l <= kMaxLevels; ++l) {
            CKC(cudaEventCreateWithFlags(&ctx->ev_an[l], cudaEventDisableTiming));
            CKC(cudaEventCreateWithFlags(&ctx->ev_flt[l], cudaEventDisableTiming));
        }
    }
    CKC(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CKC(cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
        CKC(cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming));
    }
#undef CKC
    ctx->use_umma = env_or("DSTR_UMMA", 0.0) != 0.0;
    ctx->row_filter = (int)env_or("DSTR_ROW_FILTER", 1.0);
    ctx->fg_half_thr = find_fg_half_threshold(0.3f);
    ctx->fg_thr32 = fg_threshold_f32(ctx->fg_half_thr);
    ctx->subchunk = 0;
    *out = ctx;
    return 0;
}

int dstr_destroy(dstr_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    if (ctx->s_comp) cudaStreamSynchronize(ctx->s_comp);
    if (ctx->s_h2d) cudaStreamSynchronize(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamSynchronize(ctx->s_d2h);
    for (int l = 0; l <= kMaxLevels; ++l) {
        if (ctx->d_A[l]) cudaFree(ctx->d_A[l]);
        if (ctx->d_H[l]) cudaFree(ctx->d_H[l]);
        for (int c = 0; c < 2; ++c)
        {
                if (ctx->taps[l].cfg[c].d_buf) cudaFree(ctx->taps[l].cfg[c].d_buf);
                if (ctx->taps[l].cfg[c].d_umma) cudaFree(ctx->taps[l].cfg[c].d_umma);
                if (ctx->taps[l].cfg[c].d_mma) cudaFree(ctx->taps[l].cfg[c].d_mma);
            }
    }
    for (int l = 0; l < 2; ++l)
        for (int b2 = 0; b2 < 2; ++b2)
            if (ctx->d_pyr[l][b2]) cudaFree(ctx->d_pyr[l][b2]);
    for (int l = 0; l <= kMaxLevels; ++l)
        if (ctx->d_um_scratch[l]) cudaFree(ctx->d_um_scratch[l]);
    for (int i = 0; i < 3; ++i)
        if (ctx->d_db[i]) cudaFree(ctx->d_db[i]);
    if (ctx->d_db_thr) cudaFree(ctx->d_db_thr);
    if (ctx->d_db_invflat) cudaFree(ctx->d_db_invflat);
    if (ctx->d_db_in) cudaFree(ctx->d_db_in);
    if (ctx->d_db_out) cudaFree(ctx->d_db_out);
    if (ctx->d_lstat) cudaFree(ctx->d_lstat);
    if (ctx->d_pstat) cudaFree(ctx->d_pstat);
    if (ctx->d_flat) cudaFree(ctx->d_flat);  // d_dark lives in the same allocation
    for (int i = 0; i < 2; ++i) {
        if (ctx->d_in[i]) cudaFree(ctx->d_in[i]);
        if (ctx->d_out[i]) cudaFree(ctx->d_out[i]);
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_comp[i]) cudaEventDestroy(ctx->ev_comp[i]);
        if (ctx->ev_d2h[i]) cudaEventDestroy(ctx->ev_d2h[i]);
    }
    for (auto& sp : ctx->spans) {
        cudaEventDestroy(sp.a);
        cudaEventDestroy(sp.b);
    }
    for (auto e : ctx->ev_pool) cudaEventDestroy(e);
    for (int i = 0; i < kSideStreams; ++i)
        if (ctx->s_side[i]) {
            cudaStreamSynchronize(ctx->s_side[i]);
            cudaStreamDestroy(ctx->s_side[i]);
        }
    for (int l = 0; l <= kMaxLevels; ++l) {
        if (ctx->ev_an[l]) cudaEventDestroy(ctx->ev_an[l]);
        if (ctx->ev_flt[l]) cudaEventDestroy(ctx->ev_flt[l]);
    }
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_comp) cudaStreamDestroy(ctx->s_comp);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    delete ctx;
    return 0;
}

// Work enqueued by an asynchronous host-buffer call must finish before anything else touches the
// staging buffers, the workspace or the tables.
static int drain_async(dstr_ctx* ctx) {
    if (!ctx->async_pending) return 0;
    CK(ctx, cudaStreamSynchronize(ctx->s_h2d));
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    CK(ctx, cudaStreamSynchronize(ctx->s_d2h));
    ctx->async_pending = false;
    return 0;
}

int dstr_set_flat_dark(dstr_ctx* ctx, const float* flat, const float* dark) {
    if (!ctx) return DSTR_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    {
        int rcd = drain_async(ctx);
        if (rcd) return rcd;
    }
    if (!flat || !dark) {
        ctx->have_flat_dark = false;
        return 0;
    }
    const size_t bytes = sizeof(float) * (size_t)ctx->H * ctx->W;
    if (!ctx->d_flat) {
        // one allocation [1/flat | dark]: a single L2 access-policy window can cover both fields
        CK(ctx, cudaMalloc(&ctx->d_flat, 2 * bytes));
        ctx->d_dark = ctx->d_flat + (size_t)ctx->H * ctx->W;
    }
    // the epilogue multiplies by 1/flat (division stays off the device's XU pipe); the
    // reciprocal is rounded once from double, so x * (1/flat) is within 1 ulp of x / flat
    std::vector<float> inv((size_t)ctx->H * ctx->W);
    for (size_t i = 0; i < inv.size(); ++i) inv[i] = (float)(1.0 / (double)flat[i]);
    CK(ctx, cudaMemcpyAsync(ctx->d_flat, inv.data(), bytes, cudaMemcpyHostToDevice, ctx->s_comp));
    CK(ctx, cudaMemcpyAsync(ctx->d_dark, dark, bytes, cudaMemcpyHostToDevice, ctx->s_comp));
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    {
        // The two fields are re-read for every plane while GBs of plane data stream through L2:
        // keep them resident with a persisting access-policy window on the compute stream (the
        // final synthesis kernel runs there).  Best effort: failures only cost performance.
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess && prop.persistingL2CacheMaxSize > 0 &&
            prop.accessPolicyMaxWindowSize > 0) {
            const size_t win = std::min<size_t>(2 * bytes, (size_t)prop.accessPolicyMaxWindowSize);
            const size_t carve = std::min<size_t>(win, (size_t)prop.persistingL2CacheMaxSize);
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve) == cudaSuccess) {
                cudaStreamAttrValue attr = {};
                attr.accessPolicyWindow.base_ptr = ctx->d_flat;
                attr.accessPolicyWindow.num_bytes = win;
                attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)carve / (double)win);
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                cudaStreamSetAttribute(ctx->s_comp, cudaStreamAttributeAccessPolicyWindow, &attr);
            }
            cudaGetLastError();
        }
    }
    ctx->have_flat_dark = true;
    return 0;
}

int dstr_filter_chunk(dstr_ctx* ctx, const void* in, int in_dtype, void* out, int out_dtype, int Z,
                      const dstr_params* cells, const dstr_params* no_cells, float high_int,
                      int mode, int flags) {
    if (!ctx) return DSTR_E_ARG;
    if (!in || !out || Z <= 0 || !no_cells) return fail(ctx, DSTR_E_ARG, "dstr_filter_chunk: bad argument");
    if ((in_dtype != DSTR_U16 && in_dtype != DSTR_F32) || (out_dtype != DSTR_U16 && out_dtype != DSTR_F32))
        return fail(ctx, DSTR_E_ARG, "dstr_filter_chunk: bad dtype");
    if (mode != DSTR_MODE_LOGSPACE && mode != DSTR_MODE_DISPATCH)
        return fail(ctx, DSTR_E_ARG, "dstr_filter_chunk: bad mode");
    if (mode == DSTR_MODE_DISPATCH && !cells)
        return fail(ctx, DSTR_E_ARG, "dstr_filter_chunk: dispatch mode needs both parameter sets");
    if ((flags & DSTR_FLAG_SHADOW) && !ctx->have_flat_dark)
        return fail(ctx, DSTR_E_STATE, "dstr_filter_chunk: shadow correction requested without flat/dark");
    if ((flags & DSTR_FLAG_STACK_OTSU) && mode == DSTR_MODE_DISPATCH)
        return fail(ctx, DSTR_E_UNSUPPORTED, "stack-wide Otsu is only defined for log-space mode");
    dstr_params pc = cells ? *cells : *no_cells;
    dstr_params pn = *no_cells;
    if (mode == DSTR_MODE_LOGSPACE) pc = pn;
    if (!(pn.sigma > 0.f) || !(pc.sigma > 0.f))
        return fail(ctx, DSTR_E_ARG, "sigma must be positive");  // notch(): filtering.py:111-112
    int L = pn.level < 0 ? ctx->Lmax : pn.level;
    const int Lc = pc.level < 0 ? ctx->Lmax : pc.level;
    if (Lc != L)
        return fail(ctx, DSTR_E_UNSUPPORTED,
                    "cells/no_cells use different decomposition levels: split the chunk by config");
    if (L > ctx->Lalloc)
        return fail(ctx, DSTR_E_UNSUPPORTED, "level exceeds pywt.dwtn_max_level for this plane shape by more than 4");
    CK(ctx, cudaSetDevice(ctx->device));

    for (int l = 1; l <= L; ++l) {
        int rc = build_taps(ctx, l, pc.sigma, pn.sigma);
        if (rc) return rc;
    }

    const bool in_dev = is_device_ptr(in);
    const bool out_dev = is_device_ptr(out);
    const size_t plane_px = (size_t)ctx->H * ctx->W;
    const size_t in_pb = plane_px * dtype_size(in_dtype);
    const size_t out_pb = plane_px * dtype_size(out_dtype);
    const bool stack = (flags & DSTR_FLAG_STACK_OTSU) != 0;
    if (stack && Z > ctx->zcap)
        return fail(ctx, DSTR_E_SHAPE, "stack-wide Otsu needs the whole chunk within max_planes");

    const bool pyramid = ctx->pyr_out[0] != nullptr;
    // DSTR_FLAG_NO_SYNC with host buffers: the copies and kernels are enqueued and the call returns; the
    // staging buffers are handed from call to call through the same events, so the H2D of the next chunk
    // overlaps the D2H of this one.  The caller owns both host buffers until dstr_synchronize.
    const bool async = (flags & DSTR_FLAG_NO_SYNC) && !ctx->profiling && !pyramid && !stack && !in_dev && !out_dev;
    if (!async) {
        int rcd = drain_async(ctx);
        if (rcd) return rcd;
    }
    if (pyramid) {
        if (out_dtype != DSTR_U16) return fail(ctx, DSTR_E_ARG, "pyramid outputs need a uint16 chunk output");
        if (ctx->zcap < 4) return fail(ctx, DSTR_E_STATE, "pyramid outputs need max_planes >= 4");
        int rcp = ensure_pyramid_stage(ctx, ctx->zcap);
        if (rcp) return rcp;
    }
    int rc = 0;
    if (in_dev && out_dev) {
        const int zstep = pyramid ? (ctx->zcap & ~3) : ctx->zcap;
        for (int z0 = 0; z0 < Z; z0 += zstep) {
            const int zn = std::min(zstep, Z - z0);
            rc = process_device(ctx, (const char*)in + (size_t)z0 * in_pb, in_dtype,
                                (char*)out + (size_t)z0 * out_pb, out_dtype, zn, pc, pn, high_int, mode,
                                flags, L);
            if (rc) return rc;
            if (pyramid) {
                uint16_t *h1, *h2;
                rc = emit_pyramid(ctx, (const uint16_t*)((char*)out + (size_t)z0 * out_pb), z0, zn, 0, ctx->s_comp, &h1, &h2);
                if (rc) return rc;
                const int H1 = ctx->H / 2, W1 = ctx->W / 2;
                if (h1) CK(ctx, cudaMemcpyAsync((uint16_t*)ctx->pyr_out[0] + (size_t)(z0 / 2) * H1 * W1, h1,
                                                sizeof(uint16_t) * (size_t)(zn / 2) * H1 * W1, cudaMemcpyDeviceToHost, ctx->s_comp));
                if (h2) CK(ctx, cudaMemcpyAsync((uint16_t*)ctx->pyr_out[1] + (size_t)(z0 / 4) * (H1 / 2) * (W1 / 2), h2,
                                                sizeof(uint16_t) * (size_t)(zn / 4) * (H1 / 2) * (W1 / 2), cudaMemcpyDeviceToHost, ctx->s_comp));
                if (h1 || h2) CK(ctx, cudaStreamSynchronize(ctx->s_comp));  // staging buffer 0 is reused by the next batch
            }
        }
        if (!(flags & DSTR_FLAG_NO_SYNC) || ctx->profiling) {
            CK(ctx, cudaStreamSynchronize(ctx->s_comp));
            resolve_timers(ctx);
        }
        return 0;
    }

    // ---- host buffers: three-stream pipeline over sub-chunks -----------------------------------------
    int sub = ctx->subchunk > 0 ? ctx->subchunk : std::min(ctx->zcap, 4);
    sub = std::min(sub, ctx->zcap);
    if (pyramid) sub = std::max(4, sub & ~3);
    if (stack) sub = Z;
    rc = ensure_stage(ctx, sub);
    if (rc) return rc;
    for (int z0 = 0; z0 < Z; z0 += sub, ++ctx->host_it) {
        const int zn = std::min(sub, Z - z0);
        const unsigned long long it = ctx->host_it;
        const int b = (int)(it & 1);
        const void* src = (const char*)in + (size_t)z0 * in_pb;
        void* dst = (char*)out + (size_t)z0 * out_pb;
        const void* dsrc = src;
        void* ddst = dst;
        if (!in_dev) {
            // the previous use of staging buffer b must have been consumed by compute
            if (it >= 2) CK(ctx, cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_comp[b], 0));
            CK(ctx, cudaMemcpyAsync(ctx->d_in[b], src, in_pb * zn, cudaMemcpyHostToDevice, ctx->s_h2d));
            CK(ctx, cudaEventRecord(ctx->ev_h2d[b], ctx->s_h2d));
            CK(ctx, cudaStreamWaitEvent(ctx->s_comp, ctx->ev_h2d[b], 0));
            dsrc = ctx->d_in[b];
        }
        if (!out_dev || pyramid) {
            if (it >= 2) CK(ctx, cudaStreamWaitEvent(ctx->s_comp, ctx->ev_d2h[b], 0));
        }
        if (!out_dev) ddst = ctx->d_out[b];
        rc = process_device(ctx, dsrc, in_dtype, ddst, out_dtype, zn, pc, pn, high_int, mode, flags, L);
        if (rc) return rc;
        uint16_t *h1 = nullptr, *h2 = nullptr;
        if (pyramid) {
            rc = emit_pyramid(ctx, (const uint16_t*)ddst, z0, zn, b, ctx->s_comp, &h1, &h2);
            if (rc) return rc;
        }
        CK(ctx, cudaEventRecord(ctx->ev_comp[b], ctx->s_comp));
        if (!out_dev || h1 || h2) {
            CK(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_comp[b], 0));
            if (!out_dev)
                CK(ctx, cudaMemcpyAsync(dst, ctx->d_out[b], out_pb * zn, cudaMemcpyDeviceToHost, ctx->s_d2h));
            const int H1 = ctx->H / 2, W1 = ctx->W / 2;
            if (h1) CK(ctx, cudaMemcpyAsync((uint16_t*)ctx->pyr_out[0] + (size_t)(z0 / 2) * H1 * W1, h1,
                                            sizeof(uint16_t) * (size_t)(zn / 2) * H1 * W1, cudaMemcpyDeviceToHost, ctx->s_d2h));
            if (h2) CK(ctx, cudaMemcpyAsync((uint16_t*)ctx->pyr_out[1] + (size_t)(z0 / 4) * (H1 / 2) * (W1 / 2), h2,
                                            sizeof(uint16_t) * (size_t)(zn / 4) * (H1 / 2) * (W1 / 2), cudaMemcpyDeviceToHost, ctx->s_d2h));
            CK(ctx, cudaEventRecord(ctx->ev_d2h[b], ctx->s_d2h));
        }
    }
    if (async) {
        ctx->async_pending = true;
        return 0;
    }
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    CK(ctx, cudaStreamSynchronize(ctx->s_d2h));
    CK(ctx, cudaStreamSynchronize(ctx->s_h2d));
    ctx->async_pending = false;
    resolve_timers(ctx);
    return 0;
}

int dstr_plane_stats(dstr_ctx* ctx, const void* in, int in_dtype, int Z, double* fg_mean,
                     double* bg_mean, int* use_cells, float high_int, float threshold_mask) {
    if (!ctx) return DSTR_E_ARG;
    if (!in || Z <= 0 || !fg_mean || !bg_mean) return fail(ctx, DSTR_E_ARG, "dstr_plane_stats: bad argument");
    if (in_dtype != DSTR_U16 && in_dtype != DSTR_F32) return fail(ctx, DSTR_E_ARG, "bad dtype");
    CK(ctx, cudaSetDevice(ctx->device));
    {
        int rcd = drain_async(ctx);
        if (rcd) return rcd;
    }
    const bool in_dev = is_device_ptr(in);
    const size_t plane_px = (size_t)ctx->H * ctx->W;
    const size_t in_pb = plane_px * dtype_size(in_dtype);
    int sub = std::min(ctx->zcap, Z);
    if (!in_dev) {
        int rc = ensure_stage(ctx, sub);
        if (rc) return rc;
    }
    std::vector<PlaneStat> hs(sub);
    const float fg_thr = (threshold_mask == 0.3f) ? ctx->fg_thr32 : fg_threshold_f32(find_fg_half_threshold(threshold_mask));
    for (int z0 = 0; z0 < Z; z0 += sub) {
        const int zn = std::min(sub, Z - z0);
        const void* src = (const char*)in + (size_t)z0 * in_pb;
        if (!in_dev) {
            CK(ctx, cudaMemcpyAsync(ctx->d_in[0], src, in_pb * zn, cudaMemcpyHostToDevice, ctx->s_comp));
            src = ctx->d_in[0];
        }
        CK(ctx, cudaMemsetAsync(ctx->d_pstat, 0, sizeof(PlaneStat) * zn, ctx->s_comp));
        dim3 grid(std::max(1, (int)std::min<size_t>(plane_px / 4096 + 1, 592)), zn);
        if (in_dtype == DSTR_U16)
            plane_stats_kernel<uint16_t><<<grid, 256, 0, ctx->s_comp>>>((const uint16_t*)src, ctx->H, ctx->W,
                                                                         plane_px, ctx->d_pstat, fg_thr);
        else
            plane_stats_kernel<float><<<grid, 256, 0, ctx->s_comp>>>((const float*)src, ctx->H, ctx->W, plane_px,
                                                                      ctx->d_pstat, fg_thr);
        ctx->launches++;
        CK(ctx, cudaGetLastError());
        CK(ctx, cudaMemcpyAsync(hs.data(), ctx->d_pstat, sizeof(PlaneStat) * zn, cudaMemcpyDeviceToHost, ctx->s_comp));
        CK(ctx, cudaStreamSynchronize(ctx->s_comp));
        for (int i = 0; i < zn; ++i) {
            const double fg = hs[i].fg_cnt ? hs[i].fg_sum / (double)hs[i].fg_cnt : 0.0;
            const double bg = hs[i].bg_cnt ? hs[i].bg_sum / (double)hs[i].bg_cnt : 0.0;
            fg_mean[z0 + i] = fg;
            bg_mean[z0 + i] = bg;
            if (use_cells) use_cells[z0 + i] = (fg > bg && fg > (double)high_int) ? 1 : 0;
        }
    }
    return 0;
}

int dstr_flatfield_correction(int device, const float* img, const float* flat, const float* dark,
                              const float* baseline, uint16_t* out, int n_outer, int n_inner_h,
                              int n_inner_w) {
    if (!img || !flat || !dark || !out || n_outer <= 0 || n_inner_h <= 0 || n_inner_w <= 0)
        return fail(nullptr, DSTR_E_ARG, "dstr_flatfield_correction: bad argument");
    dstr_ctx* ctx = nullptr;
    CK(ctx, cudaSetDevice(device));
    const size_t inner = (size_t)n_inner_h * n_inner_w, n = inner * n_outer;
    float *d_img = nullptr, *d_flat = nullptr, *d_dark = nullptr, *d_base = nullptr;
    unsigned short* d_out = nullptr;
    int rc = 0;
    auto cleanup = [&]() {
        cudaFree(d_img);
        cudaFree(d_flat);
        cudaFree(d_dark);
        cudaFree(d_base);
        cudaFree(d_out);
    };
#define CKF(call)                                                                          \
    do {                                                                                   \
        cudaError_t e_ = (call);                                                           \
        if (e_ != cudaSuccess) {                                                           \
            rc = fail(nullptr, (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_)); \
            cleanup();                                                                     \
            return rc;                                                                     \
        }                                                                                  \
    } while (0)
    CKF(cudaMalloc(&d_img, n * 4));
    CKF(cudaMalloc(&d_flat, n * 4));
    CKF(cudaMalloc(&d_dark, n * 4));
    CKF(cudaMalloc(&d_out, n * 2));
    CKF(cudaMemcpy(d_img, img, n * 4, cudaMemcpyHostToDevice));
    CKF(cudaMemcpy(d_flat, flat, n * 4, cudaMemcpyHostToDevice));
    CKF(cudaMemcpy(d_dark, dark, n * 4, cudaMemcpyHostToDevice));
    if (baseline) {
        CKF(cudaMalloc(&d_base, (size_t)n_outer * 4));
        CKF(cudaMemcpy(d_base, baseline, (size_t)n_outer * 4, cudaMemcpyHostToDevice));
    }
    const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
    flatfield_kernel<<<blocks, 256>>>(d_img, d_flat, d_dark, d_base, d_out, (size_t)n_outer, inner);
    CKF(cudaGetLastError());
    CKF(cudaMemcpy(out, d_out, n * 2, cudaMemcpyDeviceToHost));
#undef CKF
    cleanup();
    return 0;
}

int dstr_host_alloc(void** ptr, uint64_t bytes) {
    if (!ptr || bytes == 0) return DSTR_E_ARG;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess) return fail(nullptr, (int)e, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
    return 0;
}
int dstr_host_free(void* ptr) {
    if (!ptr) return 0;
    cudaError_t e = cudaFreeHost(ptr);
    return e == cudaSuccess ? 0 : fail(nullptr, (int)e, std::string("cudaFreeHost: ") + cudaGetErrorString(e));
}
int dstr_host_register(void* ptr, uint64_t bytes) {
    if (!ptr || bytes == 0) return DSTR_E_ARG;
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
    return e == cudaSuccess ? 0 : fail(nullptr, (int)e, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
}
int dstr_host_unregister(void* ptr) {
    if (!ptr) return 0;
    cudaError_t e = cudaHostUnregister(ptr);
    return e == cudaSuccess ? 0 : fail(nullptr, (int)e, std::string("cudaHostUnregister: ") + cudaGetErrorString(e));
}

int dstr_device_alloc(dstr_ctx* ctx, void** ptr, uint64_t bytes) {
    if (!ctx || !ptr || bytes == 0) return DSTR_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMalloc(ptr, bytes));
    return 0;
}
int dstr_device_free(dstr_ctx* ctx, void* ptr) {
    if (!ctx) return DSTR_E_ARG;
    if (!ptr) return 0;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaFree(ptr));
    return 0;
}
int dstr_memcpy_h2d(dstr_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
    if (!ctx || !dst || !src) return DSTR_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, ctx->s_comp));
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    return 0;
}
int dstr_memcpy_d2h(dstr_ctx* ctx, void* dst, const void* src, uint64_t bytes) {
    if (!ctx || !dst || !src) return DSTR_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->s_comp));
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    return 0;
}
int dstr_synchronize(dstr_ctx* ctx) {
    if (!ctx) return DSTR_E_ARG;
    CK(ctx, cudaSetDevice(ctx->device));
    CK(ctx, cudaStreamSynchronize(ctx->s_h2d));
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    CK(ctx, cudaStreamSynchronize(ctx->s_d2h));
    ctx->async_pending = false;
    resolve_timers(ctx);
    return 0;
}
void* dstr_compute_stream(dstr_ctx* ctx) { return ctx ? (void*)ctx->s_comp : nullptr; }

int dstr_set_profiling(dstr_ctx* ctx, int enabled) {
    if (!ctx) return DSTR_E_ARG;
    ctx->profiling = enabled != 0;
    return 0;
}
int dstr_get_timers(dstr_ctx* ctx, double* ms_out, uint64_t* launches_out) {
    if (!ctx) return DSTR_E_ARG;
    if (ms_out) std::memcpy(ms_out, ctx->timers, sizeof(ctx->timers));
    if (launches_out) *launches_out = ctx->launches;
    return 0;
}
int dstr_reset_timers(dstr_ctx* ctx) {
    if (!ctx) return DSTR_E_ARG;
    std::memset(ctx->timers, 0, sizeof(ctx->timers));
    ctx->launches = 0;
    return 0;
}
int dstr_set_debug_stop(dstr_ctx* ctx, int stage) {
    if (!ctx || stage < DSTR_STAGE_NONE || stage > DSTR_STAGE_SYNTH) return DSTR_E_ARG;
    ctx->debug_stop = stage;
    return 0;
}
int dstr_set_notch_tolerance(dstr_ctx* ctx, double eps) {
    if (!ctx || !(eps >= 0.0) || eps > 1e-2) return DSTR_E_ARG;
    ctx->notch_eps = eps;
    return 0;
}

int dstr_notch_design(int n, double s, double eps, int* info /*[6]*/) {
    if (n <= 0 || !(s > 0.0) || !info) return DSTR_E_ARG;
    NotchHost h;
    design_notch(n, s, eps, h);
    info[0] = h.ntap_e;
    info[1] = h.ue_lo;
    info[2] = h.ntap_o;
    info[3] = h.uo_lo;
    info[4] = h.J;
    info[5] = h.Jpad;
    return 0;
}

int dstr_notch_apply_host(int n, double s, double eps, const double* x, double* y) {
    // y = B x evaluated on the host with exactly the tables the device uses (float32 taps,
    // double accumulation): lets CPU tests bound the truncation error of the hybrid design.
    if (n <= 0 || !(s > 0.0) || !x || !y) return DSTR_E_ARG;
    NotchHost h;
    design_notch(n, s, eps, h);
    const int nh = n / 2;
    const int nhp64 = (((nh + 1 + 7) & ~7) + 63) & ~63;
    auto t2_off = [](int t) {  // the device layout of a T2 row (design_notch)
        const int seg = t >> 3, k = t & 7;
        return (size_t)(seg >> 3) * 64 + (size_t)(k >> 2) * 32 + (size_t)(seg & 7) * 4 + (k & 3);
    };
    std::vector<double> xe(n), xo(n);
    for (int t = 0; t < n; ++t) {
        const int tr = (t == 0) ? 0 : n - t;
        xe[t] = 0.5 * (x[t] + x[tr]);
        xo[t] = 0.5 * (x[t] - x[tr]);
    }
    std::vector<double> c(std::max(h.J, 1), 0.0);
    for (int j = 0; j < h.J; ++j)
        for (int v = 0; v <= nh; ++v) c[j] += (double)h.T1[(size_t)(v + 8) * h.Jpad + j] * xe[v];
    for (int t = 0; t <= nh; ++t) {
        double ye = 0.0, yo = 0.0;
        for (int k = 0; k < h.ntap_e; ++k) ye += (double)h.te[k] * xe[(((t - (h.ue_lo + k)) % n) + n) % n];
        for (int k = 0; k < h.ntap_o; ++k) yo += (double)h.to[k] * xo[(((t - (h.uo_lo + k)) % n) + n) % n];
        for (int j = 0; j < h.J; ++j) ye += c[j] * (double)h.T2[(size_t)j * nhp64 + t2_off(t)];
        y[t] = ye + yo;
        const int tm = n - t;
        if (t != 0 && tm != t) y[tm] = ye - yo;
    }
    return 0;
}

int dstr_set_row_filter(dstr_ctx* ctx, int kind) {
    if (!ctx || kind < 0 || kind > 2) return DSTR_E_ARG;
    ctx->row_filter = kind;
    return 0;
}

int dstr_set_umma(dstr_ctx* ctx, int enabled) {
    if (!ctx) return DSTR_E_ARG;
    ctx->use_umma = enabled != 0;
    return 0;
}

int dstr_notch_umma_info(int n, double s, int* info /*[8]*/) {
    if (n <= 0 || !info) return DSTR_E_ARG;
    const UmmaGeom g = umma_geom(n);
    size_t tab = 0;
    if (g.ok && s > 0.0) {
        UmmaHost uh;
        build_umma_host(n, s, g, uh);
        tab = (uh.bytes + 1023) & ~(size_t)1023;
    }
    info[0] = (g.ok && g.smem_fixed + tab <= kUmmaSmemMax) ? 1 : 0;
    info[1] = g.P;
    info[2] = g.Nt;
    info[3] = g.NC;
    info[4] = (int)tab;
    info[5] = (int)(g.smem_fixed + tab);
    info[6] = g.nout;
    info[7] = g.Kpad;
    return 0;
}

int dstr_notch_umma_apply_host(int n, double s, double thr, const double* x, double* y, int* info /*[4]*/) {
    // y = B x evaluated on the host through the very data path of notch_umma_kernel: the fp16 hi / lo
    // Hankel tables addressed like the UMMA descriptors do (block of the k step + u_local / 8 + k / 8,
    // shifted copy u_local % 8, compact band windows), pre-scaled fp16 hi / lo operands, the band /
    // remainder product lists; accumulation in double.  Lets CPU tests check table layout, banding,
    // scaling and the split.
    if (n <= 0 || !(s > 0.0) || !x || !y) return DSTR_E_ARG;
    const UmmaGeom g = umma_geom(n);
    if (!g.ok) return DSTR_E_UNSUPPORTED;
    UmmaHost uh;
    build_umma_host(n, s, g, uh);
    const UmmaCfg& uc = uh.cfg;
    const size_t oTRh = 0, oTRl = uc.tr_bytes, oTBh = (size_t)uc.tr_bytes * (uc.r3 ? 2 : 1), oTBl = oTBh + uc.tb_bytes;
    float scale = 1.0f;
    if (thr > 0.0 && thr < 1e30) {
        int e;
        std::frexp((float)thr, &e);
        scale = std::ldexp(1.0f, std::max(-24, std::min(14 - e, 40)));
    }
    std::vector<float> EH(g.Kpad, 0.f), EL(g.Kpad, 0.f), OH(g.Kpad, 0.f), OL(g.Kpad, 0.f);
    for (int v = 0; v < n; ++v) {
        const float xv = (float)x[v] * scale, xp = (float)x[(n - v) % n] * scale;
        const float e = xv + xp, o = xv - xp;
        EH[v] = half_value(half_bits(e));
        EL[v] = half_value(half_bits(e - EH[v]));
        OH[v] = half_value(half_bits(o));
        OL[v] = half_value(half_bits(o - OH[v]));
    }
    auto tab = [&](size_t base_bytes, uint32_t blk_off_bytes, int ul, int k) {
        // element (u_local, k) of the B operand whose descriptor starts at base + blk_off: SBO = LBO = 128 bytes
        const size_t byte = base_bytes + blk_off_bytes + (size_t)((ul >> 3) + (k >> 3)) * 128 + (size_t)(ul & 7) * 16 + (size_t)(k & 7) * 2;
        return (double)half_value(uh.tabs[byte / 2]);
    };
    const double inv = 1.0 / ((double)scale * (double)UM_TABLE_SCALE);
    long long mmas = 0;
    for (int p = 0; p < g.P; ++p) {
        const int u0 = p * g.Nt;
        for (int ul = 0; ul < g.Nt; ++ul) {
            const int u = u0 + ul;
            if (u >= g.nout) break;
            double dE = 0.0, dO = 0.0;
            for (int c = 0; c < g.NC; ++c) {
                const bool band = um_band(u0, g.Nt, c, n, uc.Rb);
                const int lo = u0 + UM_KC * c;
                for (int j = 0; j < 2; ++j) {
                    if (ul == 0) mmas += (uc.r3 ? 3 : 1) + (band ? (uc.r3 ? 3 : 6) : 0);
                    const int k0 = lo + 16 * j;
                    const uint32_t roff = (uint32_t)(k0 >> 3) * 128u;
                    const uint32_t boff = band ? um_tb_off(uc, lo, k0) : 0u;
                    for (int k = 0; k < 16; ++k) {
                        const int v = UM_KC * c + 16 * j + k;
                        const double trh = tab(oTRh, roff, ul, k);
                        dE += (double)EH[v] * trh;
                        if (uc.r3) dE += (double)EL[v] * trh + (double)EH[v] * tab(oTRl, roff, ul, k);
                        if (band) {
                            const double tbh = tab(oTBh, boff, ul, k), tbl = tab(oTBl, boff, ul, k);
                            if (!uc.r3) dE += (double)EH[v] * tbh + (double)EL[v] * tbh + (double)EH[v] * tbl;
                            dO += (double)OH[v] * tbh + (double)OL[v] * tbh + (double)OH[v] * tbl;
                        }
                    }
                }
            }
            const double E = dE * inv, O = dO * inv;
            y[u] = E - O;
            if (u >= 1 && 2 * u != n) y[n - u] = E + O;
        }
    }
    if (info) {
        info[0] = uc.Rb;
        info[1] = uc.r3;
        info[2] = (int)mmas;
        info[3] = uc.tb_compact;
    }
    return 0;
}

int dstr_set_pyramid_outputs(dstr_ctx* ctx, void* level1, void* level2) {
    if (!ctx) return DSTR_E_ARG;
    if (!level1 && level2) return fail(ctx, DSTR_E_ARG, "level 2 needs level 1");
    ctx->pyr_out[0] = level1;
    ctx->pyr_out[1] = level2;
    return 0;
}

int dstr_downscale2x(dstr_ctx* ctx, const uint16_t* in, int Z, int H, int W, uint16_t* out) {
    if (!ctx || !in || !out || Z < 2 || H < 2 || W < 2) return fail(ctx, DSTR_E_ARG, "dstr_downscale2x: bad argument");
    CK(ctx, cudaSetDevice(ctx->device));
    {
        int rcd = drain_async(ctx);
        if (rcd) return rcd;
    }
    const size_t n_in = (size_t)Z * H * W, n_out = (size_t)(Z / 2) * (H / 2) * (W / 2);
    const bool in_dev = is_device_ptr(in), out_dev = is_device_ptr(out);
    uint16_t *d_in = (uint16_t*)in, *d_out = out;
    int rc = 0;
    if (!in_dev) {
        CK(ctx, cudaMalloc(&d_in, n_in * 2));
        cudaError_t e = cudaMemcpyAsync(d_in, in, n_in * 2, cudaMemcpyHostToDevice, ctx->s_comp);
        if (e != cudaSuccess) rc = fail(ctx, (int)e, cudaGetErrorString(e));
    }
    if (!rc && !out_dev) {
        cudaError_t e = cudaMalloc(&d_out, n_out * 2);
        if (e != cudaSuccess) rc = fail(ctx, (int)e, cudaGetErrorString(e));
    }
    if (!rc) rc = launch_downscale(ctx, d_in, Z, H, W, d_out, ctx->s_comp);
    if (!rc && !out_dev) {
        cudaError_t e = cudaMemcpyAsync(out, d_out, n_out * 2, cudaMemcpyDeviceToHost, ctx->s_comp);
        if (e != cudaSuccess) rc = fail(ctx, (int)e, cudaGetErrorString(e));
    }
    cudaError_t es = cudaStreamSynchronize(ctx->s_comp);
    if (!rc && es != cudaSuccess) rc = fail(ctx, (int)es, cudaGetErrorString(es));
    if (!in_dev && d_in) cudaFree(d_in);
    if (!out_dev && d_out) cudaFree(d_out);
    return rc;
}

// Classic dual-band mode (dstr_dual_band.cuh).  in / out: host or device pointers, thresholds and flat: host pointers.
int dstr_dual_band_chunk(dstr_ctx* ctx, const void* in, int in_dtype, uint16_t* out, int Z, float sigma_fg,
                         float sigma_bg, int level, const float* thresholds, float crossover, float dark,
                         const float* flat) {
    if (!ctx) return DSTR_E_ARG;
    if (!in || !out || Z <= 0 || !thresholds) return fail(ctx, DSTR_E_ARG, "dstr_dual_band_chunk: bad argument");
    if (in_dtype != DSTR_U16 && in_dtype != DSTR_F32) return fail(ctx, DSTR_E_ARG, "dstr_dual_band_chunk: bad dtype");
    if (sigma_fg < 0.f || sigma_bg < 0.f || !(crossover > 0.f)) return fail(ctx, DSTR_E_ARG, "dstr_dual_band_chunk: bad sigma / crossover");
    const int L = level < 0 ? ctx->Lmax : level;
    if (L > ctx->Lalloc) return fail(ctx, DSTR_E_UNSUPPORTED, "level exceeds pywt.dwtn_max_level for this plane shape by more than 4");
    CK(ctx, cudaSetDevice(ctx->device));
    {
        int rcd = drain_async(ctx);
        if (rcd) return rcd;
    }
    cudaStream_t st = ctx->s_comp;
    const size_t plane_px = (size_t)ctx->H * ctx->W;
    const int zcap = ctx->zcap;
    const bool in_dev = is_device_ptr(in), out_dev = is_device_ptr(out);
    const size_t in_pb = plane_px * dtype_size(in_dtype);
    const bool any = sigma_fg > 0.f || sigma_bg > 0.f;
    const bool single = sigma_fg > 0.f && sigma_fg == sigma_bg;
    for (int i = 0; i < 3; ++i)
        if (!ctx->d_db[i]) CK(ctx, cudaMalloc(&ctx->d_db[i], sizeof(float) * plane_px * zcap));
    if (!ctx->d_db_thr) CK(ctx, cudaMalloc(&ctx->d_db_thr, sizeof(float) * zcap));
    if (!in_dev && !ctx->d_db_in) CK(ctx, cudaMalloc(&ctx->d_db_in, 4 * plane_px * zcap));
    if (!out_dev && !ctx->d_db_out) CK(ctx, cudaMalloc(&ctx->d_db_out, 2 * plane_px * zcap));
    const float* d_invflat = nullptr;
    if (flat) {
        if (!ctx->d_db_invflat) CK(ctx, cudaMalloc(&ctx->d_db_invflat, sizeof(float) * plane_px));
        std::vector<float> inv(plane_px);
        for (size_t i = 0; i < plane_px; ++i) inv[i] = 1.0f / flat[i];
        CK(ctx, cudaMemcpyAsync(ctx->d_db_invflat, inv.data(), sizeof(float) * plane_px, cudaMemcpyHostToDevice, st));
        CK(ctx, cudaStreamSynchronize(st));
        d_invflat = ctx->d_db_invflat;
    }
    // pystripe: s_l = H_l sigma / H (rows of the image); the tables here use min(H, W) (filtering.py:180)
    const float fix = (float)((double)std::min(ctx->H, ctx->W) / (double)ctx->H);
    dstr_params pb = {(sigma_bg > 0.f ? sigma_bg : sigma_fg) * fix, 0.f, L};
    dstr_params pf = {(sigma_fg > 0.f ? sigma_fg : sigma_bg) * fix, 0.f, L};
    if (any)
        for (int l = 1; l <= L; ++l) {
            int rc = build_taps(ctx, l, pf.sigma, pb.sigma);  // cells tables = foreground band, no_cells = background
            if (rc) return rc;
        }
    const int blocks = (int)std::min<size_t>((plane_px + 255) / 256, (size_t)ctx->sm_count * 8);
    const int band_flags = DSTR_FLAG_NOTCH_ONLY | DSTR_FLAG_EXPM1 | DSTR_FLAG_NO_SYNC;
    for (int z0 = 0; z0 < Z; z0 += zcap) {
        const int zn = std::min(zcap, Z - z0);
        const void* src = (const char*)in + (size_t)z0 * in_pb;
        if (!in_dev) {
            CK(ctx, cudaMemcpyAsync(ctx->d_db_in, src, in_pb * zn, cudaMemcpyHostToDevice, st));
            src = ctx->d_db_in;
        }
        CK(ctx, cudaMemcpyAsync(ctx->d_db_thr, thresholds + z0, sizeof(float) * zn, cudaMemcpyHostToDevice, st));
        dim3 grid(blocks, zn);
        BlendArgs ba;
        ba.bgf = nullptr;
        ba.fgf = nullptr;
        ba.thr = ctx->d_db_thr;
        ba.inv_flat = d_invflat;
        ba.crossover = crossover;
        ba.dark = dark;
        ba.single = single ? 1 : 0;
        int rc = 0;
        if (single) {
            rc = process_device(ctx, src, in_dtype, ctx->d_db[1], DSTR_F32, zn, pf, pb, 0.f, kModeForceCells, band_flags, L);
            if (rc) return rc;
            ba.fgf = ctx->d_db[1];
        } else {
            if (sigma_bg > 0.f) {
                if (in_dtype == DSTR_U16) clamp_band_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)src, ctx->d_db[2], plane_px, ctx->d_db_thr, 0);
                else clamp_band_kernel<float><<<grid, 256, 0, st>>>((const float*)src, ctx->d_db[2], plane_px, ctx->d_db_thr, 0);
                ctx->launches++;
                rc = process_device(ctx, ctx->d_db[2], DSTR_F32, ctx->d_db[0], DSTR_F32, zn, pf, pb, 0.f, DSTR_MODE_LOGSPACE, band_flags, L);
                if (rc) return rc;
                ba.bgf = ctx->d_db[0];
            }
            if (sigma_fg > 0.f) {
                if (in_dtype == DSTR_U16) clamp_band_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)src, ctx->d_db[2], plane_px, ctx->d_db_thr, 1);
                else clamp_band_kernel<float><<<grid, 256, 0, st>>>((const float*)src, ctx->d_db[2], plane_px, ctx->d_db_thr, 1);
                ctx->launches++;
                rc = process_device(ctx, ctx->d_db[2], DSTR_F32, ctx->d_db[1], DSTR_F32, zn, pf, pb, 0.f, kModeForceCells, band_flags, L);
                if (rc) return rc;
                ba.fgf = ctx->d_db[1];
            }
        }
        uint16_t* dst = out_dev ? out + (size_t)z0 * plane_px : ctx->d_db_out;
        if (in_dtype == DSTR_U16) dual_band_blend_kernel<uint16_t><<<grid, 256, 0, st>>>((const uint16_t*)src, dst, plane_px, ba);
        else dual_band_blend_kernel<float><<<grid, 256, 0, st>>>((const float*)src, dst, plane_px, ba);
        ctx->launches++;
        CK(ctx, cudaGetLastError());
        if (!out_dev)
            CK(ctx, cudaMemcpyAsync(out + (size_t)z0 * plane_px, dst, 2 * plane_px * zn, cudaMemcpyDeviceToHost, st));
        // the next batch reuses the staging buffers and thresholds
        CK(ctx, cudaStreamSynchronize(st));
    }
    return 0;
}

int dstr_histogram_u16(dstr_ctx* ctx, const uint16_t* in, int Z, uint32_t* hist) {
    if (!ctx) return DSTR_E_ARG;
    if (!in || !hist || Z <= 0) return fail(ctx, DSTR_E_ARG, "dstr_histogram_u16: bad argument");
    CK(ctx, cudaSetDevice(ctx->device));
    {
        int rcd = drain_async(ctx);
        if (rcd) return rcd;
    }
    static std::mutex mtx;
    static bool done[64] = {};
    {
        std::lock_guard<std::mutex> lk(mtx);
        if (!done[ctx->device & 63]) {
            CK(ctx, cudaFuncSetAttribute(hist_u16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
            done[ctx->device & 63] = true;
        }
    }
    cudaStream_t st = ctx->s_comp;
    const size_t plane_px = (size_t)ctx->H * ctx->W;
    const bool in_dev = is_device_ptr(in);
    uint16_t* d_in = nullptr;
    unsigned* d_hist = nullptr;
    int rc = 0;
    auto ck = [&](cudaError_t e) {
        if (e != cudaSuccess && !rc) rc = fail(ctx, (int)e, cudaGetErrorString(e));
    };
    ck(cudaMalloc(&d_hist, sizeof(unsigned) * 65536 * (size_t)Z));
    if (!rc && !in_dev) {
        ck(cudaMalloc(&d_in, 2 * plane_px * Z));
        if (!rc) ck(cudaMemcpyAsync(d_in, in, 2 * plane_px * Z, cudaMemcpyHostToDevice, st));
    }
    if (!rc) {
        ck(cudaMemsetAsync(d_hist, 0, sizeof(unsigned) * 65536 * (size_t)Z, st));
        dim3 grid((unsigned)((plane_px + HU_PX_PER_BLOCK - 1) / HU_PX_PER_BLOCK), Z);
        hist_u16_kernel<<<grid, HU_THREADS, 131072, st>>>(in_dev ? in : d_in, plane_px, d_hist);
        ctx->launches++;
        ck(cudaGetLastError());
        ck(cudaMemcpyAsync(hist, d_hist, sizeof(unsigned) * 65536 * (size_t)Z, cudaMemcpyDeviceToHost, st));
    }
    ck(cudaStreamSynchronize(st));
    if (d_in) cudaFree(d_in);
    if (d_hist) cudaFree(d_hist);
    return rc;
}

int dstr_set_tma(dstr_ctx* ctx, int enabled) {
    if (!ctx) return DSTR_E_ARG;
    ctx->use_tma = enabled != 0;
    return 0;
}

int dstr_set_overlap(dstr_ctx* ctx, int enabled) {
    if (!ctx) return DSTR_E_ARG;
    ctx->overlap = enabled != 0;
    return 0;
}

int dstr_set_subchunk(dstr_ctx* ctx, int planes) {
    if (!ctx || planes < 0) return DSTR_E_ARG;
    ctx->subchunk = planes;
    return 0;
}

int dstr_debug_fetch(dstr_ctx* ctx, int what, int level, void* host_buf, uint64_t host_bytes) {
    if (!ctx || !host_buf) return DSTR_E_ARG;
    if (level < 1 || level > ctx->Lalloc) return fail(ctx, DSTR_E_ARG, "dstr_debug_fetch: bad level");
    CK(ctx, cudaSetDevice(ctx->device));
    {
        int rcd = drain_async(ctx);
        if (rcd) return rcd;
    }
    CK(ctx, cudaStreamSynchronize(ctx->s_comp));
    const int Z = ctx->last_z;
    const LevelGeom& g = ctx->geom[level];
    if (what == DSTR_FETCH_CA || what == DSTR_FETCH_CH) {
        // the workspace holds the LAST sub-chunk only (last_z planes): a buffer of any other size means the caller
        // expects planes that are not there
        const size_t need = sizeof(float) * (size_t)Z * g.H * g.W;
        if (host_bytes != need)
            return fail(ctx, DSTR_E_ARG, "dstr_debug_fetch: buffer does not match the planes of the last sub-chunk "
                                         "(dstr_set_subchunk >= Z keeps a whole chunk resident)");
        const float* src = (what == DSTR_FETCH_CA) ? ctx->d_A[level] : ctx->d_H[level];
        for (int z = 0; z < Z; ++z) {
            CK(ctx, cudaMemcpy2D((float*)host_buf + (size_t)z * g.H * g.W, sizeof(float) * g.W,
                                 src + (size_t)z * g.pstride, sizeof(float) * g.pitch, sizeof(float) * g.W,
                                 g.H, cudaMemcpyDeviceToHost));
        }
        return 0;
    }
    std::vector<LevelStat> ls(Z);
    std::vector<PlaneStat> ps(Z);
    CK(ctx, cudaMemcpy(ls.data(), ctx->d_lstat + (size_t)(level - 1) * ctx->zcap, sizeof(LevelStat) * Z,
                       cudaMemcpyDeviceToHost));
    CK(ctx, cudaMemcpy(ps.data(), ctx->d_pstat, sizeof(PlaneStat) * Z, cudaMemcpyDeviceToHost));
    if (what == DSTR_FETCH_STATS) {
        if (host_bytes < sizeof(float) * 8 * (size_t)Z) return fail(ctx, DSTR_E_ARG, "buffer too small");
        float* o = (float*)host_buf;
        for (int z = 0; z < Z; ++z) {
            unsigned mn = ~ls[z].qmin_inv, mx = ls[z].qmax_bits;
            float fmn, fmx;
            std::memcpy(&fmn, &mn, 4);
            std::memcpy(&fmx, &mx, 4);
            const double fg = ps[z].fg_cnt ? ps[z].fg_sum / (double)ps[z].fg_cnt : 0.0;
            const double bg = ps[z].bg_cnt ? ps[z].bg_sum / (double)ps[z].bg_cnt : 0.0;
            o[8 * z + 0] = fmn;
            o[8 * z + 1] = fmx;
            o[8 * z + 2] = ls[z].otsu_raw;
            o[8 * z + 3] = ls[z].thr;
            // filtering.py:462 as the device evaluated it for this plane (plane_uses_cells)
            o[8 * z + 4] = (ctx->last_dp.mode != 0 && fg > bg && fg > (double)ctx->last_dp.high_int) ? 1.f : 0.f;
            o[8 * z + 5] = (float)fg;
            o[8 * z + 6] = (float)bg;
            o[8 * z + 7] = (float)ls[z].otsu_bin;
        }
        return 0;
    }
    if (what == DSTR_FETCH_HIST) {
        if (host_bytes < sizeof(unsigned) * 256 * (size_t)Z) return fail(ctx, DSTR_E_ARG, "buffer too small");
        for (int z = 0; z < Z; ++z) std::memcpy((unsigned*)host_buf + 256 * z, ls[z].hist, sizeof(unsigned) * 256);
        return 0;
    }
    return fail(ctx, DSTR_E_ARG, "dstr_debug_fetch: unknown item");
}

}  // extern "C"
