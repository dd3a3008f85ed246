// Row filter with the linear part on the warp-level tensor path (mma.sync m16n8k16, fp16 operand pairs).
//
// Same row arithmetic and the same operator design as filter_rows_kernel (dstr_kernels.cuh: mask, exact
// median, in-painting; B x = A x_e + Bo x_o with compact even / odd FIRs `te` / `to` plus the rank-J cosine
// correction), but every contraction runs as small matrix products instead of register-tiled FMA loops:
//
//   FIR       y[t0 + m] = sum_k A[m][k] * X[j0 + k],  A[m][k] = taps[m - k + const]  (a 16 x 16 Toeplitz tile,
//             fragments read from a reversed tap table in shared memory), 16 outputs x 16 taps per MMA
//   projection c[j] = sum_v T1[v][j] x_e[v]            (A = T1 tiles from global memory, fragment ordered)
//   expansion  y_e[t] += sum_j c[j] T2[j][t]           (A = T2 tiles, B = c)
//
// The N = 8 columns of every product are the block's 4 rows x {hi, lo}: operands are fp16 pairs
// x = hi + lo (power-of-two pre-scaling keeps them in range), so  A_hi * [X_hi | X_lo]  gives hi*hi and
// hi*lo in one MMA and  A_lo * [X_hi | X_lo]  adds lo*hi (lo*lo is 2^-22 and harmless); the two columns
// of a row are summed in the accumulator registers.  fp32 accumulation.  Thread (g, tig) of a warp ends
// up with the outputs t0 + g and t0 + g + 8 of row tig.
#pragma once
#include <type_traits>
#include "dstr_kernels.cuh"

namespace dstr {

struct MmaCfg {
    const __half* tr;   // reversed tap tables [E: hi0 | hi1 | lo0 | lo1] (trlen_e each) then [O: ...] (trlen_o each)
    const uint4* fq;    // tap A fragments as whole register quads: [E: S_e][hi, lo][14][4 words] then [O: S_o][hi, lo][14][4]
    const uint4* T1f;   // [nblk][Jpad / 16][32 lanes][hi, lo]  A fragments of T1 (modes x 16 elements)
    const uint4* T2f;   // [nseg16][Jpad / 16][32 lanes][hi, lo] A fragments of T2 (16 outputs x 16 modes)
    int ntap_e, ue_lo, ntap_o, uo_lo;  // tap counts padded to 16; tap k <-> circular offset u = u?_lo + k
    int S_e, S_o;                      // k steps of the FIRs (ntap / 16 + 1)
    int trlen_e, trlen_o;              // halfs per table copy
    int J, Jpad;                       // Jpad: multiple of 16
    int blk_lo, nblk;                  // 16-element blocks of the E index that hold x_e[0 .. nh]
    float cs;                          // power-of-two scale applied to the c_j before the fp16 split
    float inv_x;                       // 1 / (cs * T2 scale)
};

struct RowsMmaArgs {
    float* cH;
    int Hl, Wl, pitch;
    size_t pstride;
    const LevelStat* lstat;
    int stat_stride;
    MmaCfg cfg[2];  // [0] no_cells, [1] cells
    int nh;         // n / 2: outputs t = 0 .. nh
    int nseg16;     // 16-output segments
    int len_e, len_o;  // halfs per (row, part) operand array (max over the configs; = 8 mod 16)
    int trlen_e_max, trlen_o_max, Jpad_max;
    int S_e_max, S_o_max;
    int prefetch_blocks;
};

constexpr float RM_TAP_SCALE = 256.0f;
#ifndef DSTR_RM_SG
#define DSTR_RM_SG 4
#endif
constexpr int RM_SG = DSTR_RM_SG;  // 16-output segments a warp processes per tap-fragment load

__device__ __forceinline__ void mma_f16(float (&d)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

#ifndef DSTR_RM_MINB
#define DSTR_RM_MINB 8
#endif
// NT = n8 tiles of the operand matrices: a block filters ROWS = 4 NT rows (one warp each for the selection / operand
// phase) and every A fragment — tap quads from shared memory, rank-J table fragments from L2 — feeds NT MMAs.  NT = 2
// halves the table traffic per row (0.29 ms of the 1.87 ms row filter were those loads with NT = 1).
template <int EPL, int NT>
__global__ void __launch_bounds__(128 * NT, (EPL <= 33 ? DSTR_RM_MINB / NT : 4 / NT))
filter_rows_mma_kernel(RowsMmaArgs a, const PlaneStat* __restrict__ pstat, DispatchParams dp) {
    constexpr int ROWS = 4 * NT, THREADS = 32 * ROWS;
    constexpr int GF = RM_SG;  // segments per full group (NT GF accumulator tiles per thread; measured: 4 beats 2 and 3 for NT = 2)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.Wl;
    // tap fragments: entry (s, part, q - 8) is the register quad {t[16 s + q], t[16 s + q - 8], t[16 s + q + 8], t[16 s + q]}
    // (half pairs of the reversed, pre-scaled tap table) that a lane with q = 2 tig - g + 15 feeds to the MMA as
    // operand A of k step s — one 16-byte shared-memory load per fragment, nothing to assemble in registers
    uint4* s_fqe = reinterpret_cast<uint4*>(smem_raw);              // [S_e_max][2][14]
    uint4* s_fqo = s_fqe + a.S_e_max * 28;                          // [S_o_max][2][14]
    __half* s_E = reinterpret_cast<__half*>(s_fqo + a.S_o_max * 28);  // [ROWS * 2][len_e]
    __half* s_O = s_E + ROWS * 2 * a.len_e;                      // [ROWS * 2][len_o]
    unsigned long long* s_c64 = reinterpret_cast<unsigned long long*>(s_O + ROWS * 2 * a.len_o);  // [ROWS][Jpad_max]
    const int chs = a.Jpad_max + 8;                                 // c operand stride: = 8 (mod 16) halfs
    __half* s_ch = reinterpret_cast<__half*>(s_c64 + max(ROWS * a.Jpad_max, 4));  // [ROWS * 2][chs]
    unsigned* s_mask = reinterpret_cast<unsigned*>(s_ch + ROWS * 2 * chs);        // [ROWS][EPL]

    const int tid = threadIdx.x, lane = tid & 31;
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);  // same value, but known to be warp-uniform: no divergence handling
    const int z = blockIdx.y;
    const int row0 = blockIdx.x * ROWS;
    const int nrows = min(ROWS, a.Hl - row0);
    if (a.prefetch_blocks > 0 && lane == 0) {  // rows of the block that will run in this slot one wave later: DRAM -> L2
        const long long lin = (long long)blockIdx.y * gridDim.x + blockIdx.x + a.prefetch_blocks;
        const int pz = (int)(lin / gridDim.x);
        const int prow = (int)(lin - (long long)pz * gridDim.x) * ROWS + wid;
        if (pz < (int)gridDim.y && prow < a.Hl) {
            const float* pp = a.cH + (size_t)pz * a.pstride + (size_t)prow * a.pitch;
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pp), "r"(a.pitch * 4) : "memory");
        }
    }
    // the row of this warp, requested before anything else: the DRAM / L2 latency of these loads runs under the
    // table copies below (a warp past the end of the band reads the last row again and drops it)
    unsigned key[EPL];
    {
        const float* gl = a.cH + (size_t)z * a.pstride + (size_t)(row0 + min(wid, nrows - 1)) * a.pitch + lane;
#pragma unroll
        for (int i = 0; i < EPL; ++i) key[i] = __float_as_uint(gl[32 * i]);
    }
    const int cfgi = plane_uses_cells(pstat[z], dp);
    const MmaCfg mc = cfgi ? a.cfg[1] : a.cfg[0];
    const LevelStat* st = a.lstat + (size_t)z * a.stat_stride;
    const float thr_q = st->thr_q;
    // power-of-two pre-scale: |x| <= thr, so the scaled operands stay below 2^15 (fp16 range)
    float scale = 1.0f;
    bool coarse_ok = false;  // |x| scale < 2^15 for every unmasked x: the fp16 images are finite
    {
        const float thr = st->thr;
        if (thr > 0.f && thr < 1e30f) {
            int e;
            frexpf(thr, &e);
            scale = ldexpf(1.0f, max(-24, min(14 - e, 40)));
            coarse_ok = sqrtf(thr_q) * scale < 32000.0f;
        }
    }
    const int nh = a.nh;

    // tap fragments of this plane's config -> shared memory
    for (int i = tid; i < mc.S_e * 28; i += THREADS) s_fqe[i] = __ldg(mc.fq + i);
    for (int i = tid; i < mc.S_o * 28; i += THREADS) s_fqo[i] = __ldg(mc.fq + mc.S_e * 28 + i);
    for (int i = tid; i < ROWS * a.Jpad_max; i += THREADS) s_c64[i] = 0ull;

    const int OFFe = mc.ue_lo + mc.ntap_e, OFFo = mc.uo_lo + mc.ntap_o;
    const int use_e = 16 * a.nseg16 + 16 * mc.S_e, use_o = 16 * a.nseg16 + 16 * mc.S_o;  // operand entries the MMAs read
    if (wid < nrows) {
        const float* grow = a.cH + (size_t)z * a.pstride + (size_t)(row0 + wid) * a.pitch;
        // ---- load, mask, keys (see filter_rows_kernel) --------------------------------------------
        // fp16 images of the scaled zero-filled row, two per register (+inf past the end): the median search below
        // does its first, coarse bisection on these with packed compares (any rounding is monotone, so the order
        // statistics of the images bracket those of the floats)
        constexpr int HP = (EPL + 1) / 2;
        __half2 hx[HP];
        float hprev = 0.f;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const int e = lane + 32 * i;
            const float c = __uint_as_float(key[i]);
            key[i] = 0xffffffffu;
            bool m = false;
            float hv = __int_as_float(0x7f800000);
            if (e < n) {
                m = __fmul_rn(c, c) > thr_q;
                const float x = m ? 0.0f : (c + 0.0f);
                key[i] = f2key(x);  // zero-filled background, canonical +0
                hv = x * scale;
            }
            if (i & 1) hx[i >> 1] = __floats2half2_rn(hprev, hv);
            else if (i == EPL - 1) hx[i >> 1] = __floats2half2_rn(hv, __int_as_float(0x7f800000));
            hprev = hv;
            const unsigned mbits = __ballot_sync(0xffffffffu, m);
            if (lane == 0) s_mask[wid * EPL + i] = mbits;
        }
        // ---- exact median of the zero-filled background (np.median, filtering.py:201) -------------
        const unsigned KZ = 0x80000000u;
        const int k1 = (n - 1) >> 1, k2 = n >> 1;
        int cneg = 0, cle0 = 0;
        const bool nomask = thr_q == __int_as_float(0x7f800000);  // notch-only pass: the in-painting value is never used
        if (!nomask) {
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                cneg += (key[i] < KZ) ? 1 : 0;
                cle0 += (key[i] <= KZ) ? 1 : 0;
            }
            cneg = __reduce_add_sync(0xffffffffu, cneg);
            cle0 = __reduce_add_sync(0xffffffffu, cle0);
        }
        float med;
        if (nomask || (cneg <= k1 && k2 < cle0)) {
            med = 0.f;
        } else {
            unsigned kk1 = 0xffffffffu;  // key of the order statistic k1
            if (coarse_ok) {
                // stage 1: bisection over the 16-bit ordered keys of the fp16 images; invariant
                //   #{h < t(res16)} = lo_cnt <= k1 < hi_cnt = #{h < t(hi16)};  ends when one element is bracketed
                unsigned res16 = 0u, hi16 = 0xfc00u;  // t(0xfc00) = +inf
                int lo_cnt = 0, hi_cnt = n;
                for (int b = 15; b >= 0 && (hi_cnt - lo_cnt) != 1; --b) {
                    const unsigned trial = res16 | (1u << b);
                    const __half2 t2 = __half2half2(__ushort_as_half((unsigned short)((trial & 0x8000u) ? (trial & 0x7fffu) : (~trial & 0xffffu))));
                    __half2 a0 = __half2half2(__ushort_as_half((unsigned short)0)), a1 = a0;
#pragma unroll
                    for (int i = 0; i < HP; ++i) {
                        if (i & 1) a1 = __hadd2(a1, __hlt2(hx[i], t2));
                        else a0 = __hadd2(a0, __hlt2(hx[i], t2));
                    }
                    a0 = __hadd2(a0, a1);  // at most EPL per half: exact in fp16
                    int cnt = __float2int_rn(__low2float(a0) + __high2float(a0));
                    cnt = __reduce_add_sync(0xffffffffu, cnt);
                    if (cnt <= k1) {
                        res16 = trial;
                        lo_cnt = cnt;
                    } else {
                        hi16 = trial;
                        hi_cnt = cnt;
                    }
                }
                // stage 2: the bracketed elements are contiguous in float order; rank r among them
                const __half2 lo2 = __half2half2(__ushort_as_half((unsigned short)((res16 & 0x8000u) ? (res16 & 0x7fffu) : (~res16 & 0xffffu))));
                const __half2 hi2 = __half2half2(__ushort_as_half((unsigned short)((hi16 & 0x8000u) ? (hi16 & 0x7fffu) : (~hi16 & 0xffffu))));
                unsigned cmin = 0xffffffffu, cmax = 0u;
#pragma unroll
                for (int i = 0; i < HP; ++i) {
                    const unsigned mk = __hge2_mask(hx[i], lo2) & __hlt2_mask(hx[i], hi2);
                    if (mk & 0xffffu) {
                        cmin = min(cmin, key[2 * i]);
                        cmax = max(cmax, key[2 * i]);
                    }
                    if (2 * i + 1 < EPL && (mk >> 16)) {
                        cmin = min(cmin, key[2 * i + 1]);
                        cmax = max(cmax, key[2 * i + 1]);
                    }
                }
                cmin = __reduce_min_sync(0xffffffffu, cmin);
                cmax = __reduce_max_sync(0xffffffffu, cmax);
                const int mc_ = hi_cnt - lo_cnt, r = k1 - lo_cnt;
                if (r == 0) {
                    kk1 = cmin;
                } else if (r == mc_ - 1) {
                    kk1 = cmax;
                } else {  // several distinct floats inside one fp16 step: smallest key v with #{key <= v} > k1
                    unsigned lo = cmin, hi = cmax;
                    while (lo < hi) {
                        const unsigned mid = lo + ((hi - lo) >> 1);
                        int cnt = 0;
#pragma unroll
                        for (int i = 0; i < EPL; ++i) cnt += (key[i] <= mid) ? 1 : 0;
                        cnt = __reduce_add_sync(0xffffffffu, cnt);
                        if (cnt > k1) hi = mid;
                        else lo = mid + 1u;
                    }
                    kk1 = lo;
                }
            } else {
                unsigned res;
                int lo_cnt, hi_cnt;
                if (k1 < cneg) {
                    res = 0u;
                    lo_cnt = 0;
                    hi_cnt = cneg;
                } else {
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
#pragma unroll
                for (int i = 0; i < EPL; ++i)
                    if (key[i] >= res) kk1 = min(kk1, key[i]);
                kk1 = __reduce_min_sync(0xffffffffu, kk1);
            }
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
        // ---- in-painted row x = m ? med : c; circular even / odd parts, pre-scaled, as fp16 hi / lo:
        //      E[tau + OFFe] = x_e[tau mod n] for every entry the MMAs read, same for O
        {
            __half* Eh = s_E + (wid * 2) * a.len_e;
            __half* El = Eh + a.len_e;
            __half* Oh = s_O + (wid * 2) * a.len_o;
            __half* Ol = Oh + a.len_o;
            // two consecutive entries per lane and step: the E pair is one aligned half2 store per part (the first
            // entry of a pair is even); the O pair too when OFFo - OFFe is even, else four 16-bit stores
            int tau_lo = -max(OFFe, OFFo);
            tau_lo -= (tau_lo + OFFe) & 1;
            const int tau_hi = max(use_e - OFFe, use_o - OFFo);
            const bool o_aligned = ((OFFo - OFFe) & 1) == 0;
            int t0 = (tau_lo + 2 * lane) % n;
            if (t0 < 0) t0 += n;
            const int step = 64 % n;
            const float* gp = grow;
            asm volatile("" : "+l"(gp));
            const float hs = 0.5f * scale;
            int ae = tau_lo + 2 * lane + OFFe, ao = tau_lo + 2 * lane + OFFo;
            for (int tau = tau_lo + 2 * lane; tau < tau_hi; tau += 64, ae += 64, ao += 64) {
                const int t1 = (t0 + 1 == n) ? 0 : t0 + 1;
                const unsigned m0 = (t0 == 0) ? 0u : (unsigned)(n - t0), m1 = (t1 == 0) ? 0u : (unsigned)(n - t1);
                const float c00 = gp[(unsigned)t0], c01 = gp[(unsigned)t1], c10 = gp[m0], c11 = gp[m1];
                const float x00 = (__fmul_rn(c00, c00) > thr_q) ? med : c00;
                const float x01 = (__fmul_rn(c01, c01) > thr_q) ? med : c01;
                const float x10 = (__fmul_rn(c10, c10) > thr_q) ? med : c10;
                const float x11 = (__fmul_rn(c11, c11) > thr_q) ? med : c11;
                const float ve0 = hs * (x00 + x10), vo0 = hs * (x00 - x10);
                const float ve1 = hs * (x01 + x11), vo1 = hs * (x01 - x11);
                if ((unsigned)ae < (unsigned)use_e) {  // ae and use_e are even: ae + 1 is in range as well
                    const __half2 h = __floats2half2_rn(ve0, ve1);
                    const float2 back = __half22float2(h);
                    *reinterpret_cast<__half2*>(Eh + ae) = h;
                    *reinterpret_cast<__half2*>(El + ae) = __floats2half2_rn(ve0 - back.x, ve1 - back.y);
                }
                {
                    const __half2 h = __floats2half2_rn(vo0, vo1);
                    const float2 back = __half22float2(h);
                    const __half2 l = __floats2half2_rn(vo0 - back.x, vo1 - back.y);
                    if (o_aligned) {
                        if ((unsigned)ao < (unsigned)use_o) {
                            *reinterpret_cast<__half2*>(Oh + ao) = h;
                            *reinterpret_cast<__half2*>(Ol + ao) = l;
                        }
                    } else {
                        if ((unsigned)ao < (unsigned)use_o) {
                            Oh[ao] = __low2half(h);
                            Ol[ao] = __low2half(l);
                        }
                        if ((unsigned)(ao + 1) < (unsigned)use_o) {
                            Oh[ao + 1] = __high2half(h);
                            Ol[ao + 1] = __high2half(l);
                        }
                    }
                }
                t0 += step;
                if (t0 >= n) t0 -= n;
            }
        }
    }
    __syncthreads();  // every row's operands are complete

    const int g = lane >> 2, tig = lane & 3;
    // ---- rank-J coefficients  c_j = sum_v T1[v][j] x_e[v]  (scaled like the operands) ------------
    if (mc.J > 0) {
        const int mtiles = mc.Jpad >> 4;
        const int bw = (mc.nblk + ROWS - 1) / ROWS;
        const int b_begin = wid * bw, b_end = min(mc.nblk, b_begin + bw);
        const __half* xcol = s_E + g * a.len_e + 16 * mc.blk_lo + 2 * tig;  // column 8 t + g = (row 4 t + (g >> 1), part g & 1)
        for (int mt = 0; mt < mtiles; ++mt) {
            float acc[NT][4];
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
            const uint4* tf = mc.T1f + ((size_t)b_begin * mtiles + mt) * 64 + 2 * lane;
            for (int blk = b_begin; blk < b_end; ++blk) {
                const uint4 fh = __ldg(tf), fl = __ldg(tf + 1);
                const unsigned ah[4] = {fh.x, fh.y, fh.z, fh.w}, al[4] = {fl.x, fl.y, fl.z, fl.w};
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const unsigned b0 = *reinterpret_cast<const unsigned*>(xcol + 8 * t * a.len_e + 16 * blk);
                    const unsigned b1 = *reinterpret_cast<const unsigned*>(xcol + 8 * t * a.len_e + 16 * blk + 8);
                    mma_f16(acc[t], ah, b0, b1);
                    mma_f16(acc[t], al, b0, b1);
                }
                tf += (size_t)mtiles * 64;
            }
            // acc[0] + acc[1]: mode 16 mt + g of row 4 t + tig (hi and lo operand columns); acc[2] + acc[3]: mode + 8.
            // Partial sums of the warps are combined as 2^-24 fixed point (order-free integer adds: deterministic)
            if (b_begin < b_end) {
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    unsigned long long* dst = s_c64 + (4 * t + tig) * a.Jpad_max + 16 * mt + g;
                    atomicAdd(dst, (unsigned long long)__float2ll_rn((acc[t][0] + acc[t][1]) * 16777216.0f));
                    atomicAdd(dst + 8, (unsigned long long)__float2ll_rn((acc[t][2] + acc[t][3]) * 16777216.0f));
                }
            }
        }
        __syncthreads();
        for (int i = tid; i < ROWS * mc.Jpad; i += THREADS) {
            const int r = i / mc.Jpad, j = i - r * mc.Jpad;
            const float c = __ll2float_rn((long long)s_c64[r * a.Jpad_max + j]) * (1.0f / 16777216.0f) * mc.cs;
            const __half h = __float2half_rn(c);
            s_ch[(2 * r) * chs + j] = h;
            s_ch[(2 * r + 1) * chs + j] = __float2half_rn(c - __half2float(h));
        }
        __syncthreads();
    }

    // ---- FIRs + rank-J expansion + output ----------------------------------------------------------
    // A warp takes groups of RM_SG consecutive 16-output segments.  The tap fragments of a k step are loaded once
    // and used for every segment of the group (the shared-memory pipe, not the tensor pipe, is what limits this
    // phase: 2 + 6 / RM_SG loads per pair of MMAs instead of 8).  The expansion runs first into the accumulator
    // of the even part and is rescaled (a power of two) into the units of the FIR before the taps are added.
    const float inv_f = 1.0f / (scale * RM_TAP_SCALE);
    const float x_to_f = mc.inv_x * RM_TAP_SCALE;  // (1 / (cs ts)) / (1 / 256)
    // q = 2 tig - g + 15 selects the lane's fragment of a k step
    const uint4* fqe = s_fqe + (2 * tig - g + 15 - 8);
    const uint4* fqo = s_fqo + (2 * tig - g + 15 - 8);
    // the rows whose outputs this thread ends up with: 4 t + tig
    const __half* ecol = s_E + g * a.len_e + 2 * tig;  // operand column 8 t + g = (row 4 t + (g >> 1), part g & 1)
    const __half* ocol = s_O + g * a.len_o + 2 * tig;
    // every warp owns a contiguous range of segments (sizes differ by at most one)
    const int seg_lo = (a.nseg16 * wid) / ROWS, seg_hi = (a.nseg16 * (wid + 1)) / ROWS;
    // One group of G consecutive segments, G a compile-time constant: full groups of GF share the tap fragments of a
    // k step; the leftover segments of a warp run as groups of one.  (A run-time count would put every mma.sync
    // under a predicate, which costs a WARPSYNC + NOP per MMA.)
    auto do_group = [&](auto Gc, const int sg0) {
        constexpr int G = decltype(Gc)::value;
        float acc[G][NT][4];
        float ye[G][NT][2];
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
            for (int t = 0; t < NT; ++t) acc[i][t][0] = acc[i][t][1] = acc[i][t][2] = acc[i][t][3] = 0.f;
        if (mc.J > 0) {
            const int ktiles = mc.Jpad >> 4;
            const __half* ccol = s_ch + g * chs + 2 * tig;
            for (int kt = 0; kt < ktiles; ++kt) {
                unsigned b0[NT], b1[NT];
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    b0[t] = *reinterpret_cast<const unsigned*>(ccol + 8 * t * chs + 16 * kt);
                    b1[t] = *reinterpret_cast<const unsigned*>(ccol + 8 * t * chs + 16 * kt + 8);
                }
#pragma unroll
                for (int i = 0; i < G; ++i) {
                    const int seg = sg0 + i;
                    const uint4* tf = mc.T2f + ((size_t)seg * ktiles + kt) * 64 + 2 * lane;
                    const uint4 fh = __ldg(tf), fl = __ldg(tf + 1);
                    const unsigned ah[4] = {fh.x, fh.y, fh.z, fh.w}, al[4] = {fl.x, fl.y, fl.z, fl.w};
#pragma unroll
                    for (int t = 0; t < NT; ++t) {
                        mma_f16(acc[i][t], ah, b0[t], b1[t]);
                        mma_f16(acc[i][t], al, b0[t], b1[t]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < G; ++i)
#pragma unroll
                for (int t = 0; t < NT; ++t)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[i][t][k] *= x_to_f;
        }
        for (int s = 0; s < mc.S_e; ++s) {
            const uint4 fh = fqe[s * 28], fl = fqe[s * 28 + 14];
            const unsigned ah[4] = {fh.x, fh.y, fh.z, fh.w}, al[4] = {fl.x, fl.y, fl.z, fl.w};
#pragma unroll
            for (int i = 0; i < G; ++i) {
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const __half* xc = ecol + 8 * t * a.len_e + 16 * (sg0 + i + s);
                    const unsigned b0 = *reinterpret_cast<const unsigned*>(xc);
                    const unsigned b1 = *reinterpret_cast<const unsigned*>(xc + 8);
                    mma_f16(acc[i][t], ah, b0, b1);
                    mma_f16(acc[i][t], al, b0, b1);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
            for (int t = 0; t < NT; ++t) {
                ye[i][t][0] = acc[i][t][0] + acc[i][t][1];
                ye[i][t][1] = acc[i][t][2] + acc[i][t][3];
                acc[i][t][0] = acc[i][t][1] = acc[i][t][2] = acc[i][t][3] = 0.f;
            }
        for (int s = 0; s < mc.S_o; ++s) {
            const uint4 fh = fqo[s * 28], fl = fqo[s * 28 + 14];
            const unsigned ah[4] = {fh.x, fh.y, fh.z, fh.w}, al[4] = {fl.x, fl.y, fl.z, fl.w};
#pragma unroll
            for (int i = 0; i < G; ++i) {
#pragma unroll
                for (int t = 0; t < NT; ++t) {
                    const __half* xc = ocol + 8 * t * a.len_o + 16 * (sg0 + i + s);
                    const unsigned b0 = *reinterpret_cast<const unsigned*>(xc);
                    const unsigned b1 = *reinterpret_cast<const unsigned*>(xc + 8);
                    mma_f16(acc[i][t], ah, b0, b1);
                    mma_f16(acc[i][t], al, b0, b1);
                }
            }
        }
        // dH[t] = masked ? 0 : -(B x)[t];  (B x)[t] = y_e + y_o,  (B x)[n - t] = y_e - y_o
#pragma unroll
        for (int tl = 0; tl < NT; ++tl) {
            const int r = 4 * tl + tig;
            if (r >= nrows) continue;
            float* orow = a.cH + (size_t)z * a.pstride + (size_t)(row0 + r) * a.pitch;
            const unsigned* mrow = s_mask + r * EPL;
#pragma unroll
            for (int i = 0; i < G; ++i) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int t = 16 * (sg0 + i) + g + 8 * h;
                    if (t > nh) continue;
                    const float yev = ye[i][tl][h] * inv_f;
                    const float yo = (acc[i][tl][2 * h] + acc[i][tl][2 * h + 1]) * inv_f;
                    const bool md = (mrow[t >> 5] >> (t & 31)) & 1u;
                    orow[t] = md ? 0.f : -(yev + yo);
                    const int tm = n - t;
                    if (t >= 1 && tm != t) {
                        const bool mm = (mrow[tm >> 5] >> (tm & 31)) & 1u;
                        orow[tm] = mm ? 0.f : -(yev - yo);
                    }
                }
            }
        }
    };
    int sg0 = seg_lo;
    for (; sg0 + GF <= seg_hi; sg0 += GF) do_group(std::integral_constant<int, GF>{}, sg0);
    for (; sg0 < seg_hi; ++sg0) do_group(std::integral_constant<int, 1>{}, sg0);
}

}  // namespace dstr
