// Row filter on the 5th-generation tensor cores (tcgen05 / TMEM / TMA bulk copies), sm_100a.
//
// Same arithmetic as filter_rows_kernel (filtering.py:195-217):
//   m = sqrt(c*c) > thr;  x = m ? median(zero-filled background) : c;
//   dH = m ? 0 : -(B x),  B x = x - irfft(rfft(x) * g)  (g on the PACKED rfft index).
// With x_e[v] = x[v] + x[n-v], x_o[v] = x[v] - x[n-v] (circular, v = 0..n-1) and the even circular
// kernels ha (cosine multipliers a_j) and hb (sine multipliers b_j) of the operator,
//   (B x)[u]     = yE[u] + yO[u],   (B x)[n-u] = yE[u] - yO[u],   u = 0..n/2
//   yE[u] =  sum_v 1/2 ha(u+v) x_e[v]          (x_e even:  ha(u-v) -> ha(u+v))
//   yO[u] = -sum_v 1/2 hb(u+v) x_o[v]          (x_o odd)
// i.e. two products  D[row][u] = sum_v X[row][v] * T[u + v]  whose right operand is a HANKEL matrix
// of one periodic sequence.  The UMMA shared-memory descriptor (K-major, no swizzle) addresses
// 8 x 16-byte core matrices by  start + (u / 8) * SBO + (v / 8) * LBO;  a Hankel operand only
// depends on u / 8 + v / 8 (+ the in-core shift u % 8, materialised as 8 shifted copies), so with
// LBO = SBO = 128 bytes ONE table of 16 n bytes serves every (output tile, k step): the
// operator of a 1026-wide band lives in shared memory (4 tables, 104 KB) instead of being streamed.
//
// Precision: operands are fp16 pairs (hi + lo, power-of-two pre-scaling), products
// hi*hi + lo*hi + hi*lo accumulated in fp32 in TMEM (measured 1.5e-6 of max |B x| on random rows,
// tools/probes/umma_probe.cu).  ha = hb_band + r:  hb is a compact Gaussian (radius Rb: only the
// k chunks within Rb of u + v = 0 (mod n) are multiplied), the remainder r (1/u^2 tails of the
// kink of a_j at j = 0, |r|_2 ~ 3e-3) runs over the full circle, with a single fp16 product where
// that is below the tolerance (level 1) and with the three-product split otherwise.
//
// One persistent CTA per SM; an item = up to 128 consecutive rows of one plane (MMA M = 128):
//   A  (8 warps)  one warp per row: load, mask, exact median, in-paint, pre-scale -> staging rows
//                 (forward and index-reversed); then X_e / X_o as fp16 hi / lo in UMMA chunk order to
//                 a per-CTA scratch in global memory (stays in L2)
//   B  (warp 0)   TMA bulk copies scratch -> 3-stage shared-memory ring (mbarrier complete_tx)
//      (warp 1)   one thread issues tcgen05.mma kind::f16 (M 128, N = outputs of the pass, K 16),
//                 tcgen05.commit frees the ring slot / publishes the accumulators
//   C  (8 warps)  tcgen05.ld the E and O accumulators, combine, apply the mask, store dH in place
// B and C repeat per pass (a pass = up to 256 outputs: E and O accumulators share the 512 TMEM columns).
#pragma once
#include "dstr_kernels.cuh"

namespace dstr {

constexpr int UM_THREADS = 256;
constexpr int UM_WARPS = UM_THREADS / 32;
constexpr int UM_ROWS = 128;                          // rows of an item = MMA M
constexpr int UM_KC = 32;                             // k elements per chunk (two K = 16 MMAs)
constexpr int UM_CHUNK_BYTES = UM_ROWS * UM_KC * 2;   // 8 KB: [k / 8][row / 8][row % 8][8 halfs]
constexpr int UM_STAGES = 3;
constexpr int UM_BATCH = 8;                           // rows staged per phase-A batch (one per warp)
constexpr float UM_TABLE_SCALE = 256.0f;

struct UmmaCfg {
    const uint4* tables;            // [TBh | TBl | TRh | TRl], tab_bytes each (aliased Hankel layout)
    int Rb;                         // band radius of the compact kernel
    int r3;                         // remainder kernel with the three-product split
    unsigned long long need_band;   // chunks that are in band in at least one pass
};

struct UmmaLevelArgs {
    float* cH;
    int Hl, n, pitch;
    size_t pstride;
    const LevelStat* lstat;
    int stat_stride;
    int nh, nout;       // n / 2, outputs u = 0..nh
    int P, Nt;          // passes, outputs per pass (multiple of 16, <= 256)
    int NC, Kpad;       // k chunks, 32 NC >= n
    int rows_per_item, items_per_plane, n_items;
    int tab_bytes;
    int xs_stride;      // staging row stride in floats (= 4 mod 32)
    int mw;             // mask words per row
    int vec_ok;
    UmmaCfg cfg[2];
    uint8_t* scratch;
    size_t scratch_stride;  // bytes per CTA: 4 arrays x NC chunks
};

__host__ __device__ __forceinline__ bool um_band(int u0, int Nt, int c, int n, int Rb) {
    // does { (u + v) mod n : u in [u0, u0 + Nt), v in chunk c } come within Rb of 0 ?
    const int lo = u0 + UM_KC * c, hi = lo + Nt + UM_KC - 2;
    return (lo <= Rb) || (hi >= n - Rb && lo <= n + Rb) || (hi >= 2 * n - Rb && lo <= 2 * n + Rb);
}

__device__ __forceinline__ uint32_t um_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor: K-major, no swizzle, descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t um_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}

__device__ __forceinline__ void um_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void um_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(um_smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void um_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = um_smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}

__device__ __forceinline__ void um_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

// fp16 hi / lo of eight pre-scaled values, packed for one 16-byte core-matrix row
__device__ __forceinline__ void um_split8(const float (&v)[8], uint4& hi, uint4& lo) {
    unsigned h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 back = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * i] - back.x, v[2 * i + 1] - back.y);
        h[i] = *reinterpret_cast<const unsigned*>(&hh);
        l[i] = *reinterpret_cast<const unsigned*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int EPL>
__global__ void __launch_bounds__(UM_THREADS, 1)
notch_umma_kernel(UmmaLevelArgs a, const PlaneStat* __restrict__ pstat, DispatchParams dp) {
    extern __shared__ __align__(1024) uint8_t um_smem[];
    __shared__ __align__(8) uint64_t s_full[UM_STAGES], s_empty[UM_STAGES], s_acc;
    __shared__ uint32_t s_tmem;
    namespace ptx = cuda::ptx;

    uint8_t* s_tab = um_smem;                                                            // 4 tables
    uint8_t* s_ring = s_tab + 4 * (size_t)a.tab_bytes;                                   // [stage][EH|EL|OH|OL]
    unsigned* s_mask = reinterpret_cast<unsigned*>(s_ring + UM_STAGES * 4 * UM_CHUNK_BYTES);  // [128][mw]
    float* xs = reinterpret_cast<float*>(s_ring);       // phase A staging (the ring is idle then): forward rows
    float* xr = xs + UM_BATCH * a.xs_stride;            // ... and index-reversed rows  xr[v] = x[(n - v) mod n]

    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = a.n;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < UM_STAGES; ++i) {
            ptx::mbarrier_init(&s_full[i], 1);
            ptx::mbarrier_init(&s_empty[i], 1);
        }
        ptx::mbarrier_init(&s_acc, 1);
        ptx::fence_mbarrier_init(ptx::sem_release, ptx::scope_cluster);
    }
    if (wid == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(um_smem_u32(&s_tmem)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = s_tmem;
    const uint32_t tE = tmem_base, tO = tmem_base + 256;

    uint32_t it_p = 0, it_c = 0, acc_phase = 0;  // ring / accumulator barrier phases run across passes and items
    int cur_cfg = -1;
    uint8_t* scr = a.scratch + (size_t)blockIdx.x * a.scratch_stride;
    const size_t arr_stride = (size_t)a.NC * UM_CHUNK_BYTES;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(a.Nt >> 3) << 17) | ((uint32_t)(UM_ROWS >> 4) << 24);

    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const int z = item / a.items_per_plane;
        const int row0 = (item - z * a.items_per_plane) * a.rows_per_item;
        const int nrows = min(a.rows_per_item, a.Hl - row0);
        const int cfg = plane_uses_cells(pstat[z], dp);
        const UmmaCfg uc = cfg ? a.cfg[1] : a.cfg[0];
        if (cfg != cur_cfg) {  // CTA-uniform; every MMA of the previous item has completed (accumulator barrier)
            const int nvec = a.tab_bytes / 4;  // uint4 per ... 4 tables * tab_bytes / 16
            uint4* dst = reinterpret_cast<uint4*>(s_tab);
            for (int i = tid; i < nvec; i += UM_THREADS) dst[i] = __ldg(uc.tables + i);
            cur_cfg = cfg;
        }
        const LevelStat* st = a.lstat + (size_t)z * a.stat_stride;
        const float thr_q = st->thr_q;
        // power-of-two pre-scale: |x| <= thr, so |x_e|, |x_o| <= 2 thr < 2^15 after scaling (fp16 range)
        float scale = 1.0f;
        {
            const float thr = st->thr;
            if (thr > 0.f && thr < 1e30f) {
                int e;
                frexpf(thr, &e);  // thr < 2^e
                scale = ldexpf(1.0f, max(-24, min(14 - e, 40)));
            }
        }
        const float inv = 1.0f / (scale * UM_TABLE_SCALE);

        // ================= phase A: selection, in-painting, X_e / X_o operands =========================
        for (int b0 = 0; b0 < nrows; b0 += UM_BATCH) {
            const int rl = b0 + wid;
            float* xrow = xs + wid * a.xs_stride;
            float* rrow = xr + wid * a.xs_stride;
            if (rl < nrows) {
                const float* grow = a.cH + (size_t)z * a.pstride + (size_t)(row0 + rl) * a.pitch;
                unsigned key[EPL];
                {
                    const float* gl = grow + lane;  // lanes past the row end read slack and are discarded
#pragma unroll
                    for (int i = 0; i < EPL; ++i) key[i] = __float_as_uint(gl[32 * i]);
                }
                unsigned long long mm = 0ull;  // mask bits of this lane's elements
#pragma unroll
                for (int i = 0; i < EPL; ++i) {
                    const int e = lane + 32 * i;
                    const float c = __uint_as_float(key[i]);
                    key[i] = 0xffffffffu;
                    bool m = false;
                    if (e < n) {
                        m = __fmul_rn(c, c) > thr_q;  // sqrt(c*c) > thr, bit-identical (otsu_kernel)
                        key[i] = f2key(m ? 0.0f : (c + 0.0f));
                    }
                    mm |= (unsigned long long)(m ? 1u : 0u) << i;
                    const unsigned mbits = __ballot_sync(0xffffffffu, m);
                    if (lane == 0 && i < a.mw) s_mask[rl * a.mw + i] = mbits;
                }
                // exact median of the zero-filled background (np.median, filtering.py:201): see filter_rows_kernel
                const unsigned KZ = 0x80000000u;
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
                if (cneg <= k1 && k2 < cle0) {
                    med = 0.f;
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
                // in-painted, pre-scaled row: forward and index-reversed copies (both aligned for phase A2)
#pragma unroll
                for (int i = 0; i < EPL; ++i) {
                    const int e = lane + 32 * i;
                    if (e < n) {
                        const float x = (((mm >> i) & 1ull) ? med : key2f(key[i])) * scale;
                        xrow[e] = x;
                        rrow[e == 0 ? 0 : n - e] = x;
                    }
                }
                for (int e = n + lane; e < a.Kpad; e += 32) {
                    xrow[e] = 0.f;
                    rrow[e] = 0.f;
                }
            }
            __syncthreads();
            {
                const int brows = min(UM_BATCH, nrows - b0);
                const int r8 = lane & 7, oq = lane >> 3;
                for (int c = wid; c < a.NC; c += UM_WARPS) {
                    if (r8 < brows) {
                        const int off = r8 * a.xs_stride + UM_KC * c + 8 * oq;
                        const float4 f0 = *reinterpret_cast<const float4*>(xs + off);
                        const float4 f1 = *reinterpret_cast<const float4*>(xs + off + 4);
                        const float4 g0 = *reinterpret_cast<const float4*>(xr + off);
                        const float4 g1 = *reinterpret_cast<const float4*>(xr + off + 4);
                        const float fv[8] = {f0.x, f0.y, f0.z, f0.w, f1.x, f1.y, f1.z, f1.w};
                        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                        float ev[8], ov[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            ev[i] = fv[i] + gv[i];
                            ov[i] = fv[i] - gv[i];
                        }
                        const bool band = (uc.need_band >> c) & 1ull;
                        uint8_t* dstp = scr + (size_t)c * UM_CHUNK_BYTES + (size_t)oq * 2048 + (size_t)(b0 >> 3) * 128 + r8 * 16;
                        uint4 hi, lo;
                        um_split8(ev, hi, lo);
                        *reinterpret_cast<uint4*>(dstp) = hi;
                        if (band || uc.r3) *reinterpret_cast<uint4*>(dstp + arr_stride) = lo;
                        if (band) {
                            um_split8(ov, hi, lo);
                            *reinterpret_cast<uint4*>(dstp + 2 * arr_stride) = hi;
                            *reinterpret_cast<uint4*>(dstp + 3 * arr_stride) = lo;
                        }
                    }
                }
            }
            __syncthreads();  // the staging rows are rewritten by the next batch
        }
        // generic-proxy writes (tables in shared memory, operands in global memory, staging in the ring)
        // -> visible to / ordered before the async proxy (TMA bulk copies, tcgen05.mma operand reads)
        asm volatile("fence.proxy.async;" ::: "memory");
        __syncthreads();

        // ================= phases B / C per pass ========================================================
        for (int p = 0; p < a.P; ++p) {
            const int u0 = p * a.Nt;
            if (wid == 0) {
                // ---- TMA producer ----
                for (int c = 0; c < a.NC; ++c, ++it_p) {
                    const uint32_t s = it_p % UM_STAGES, ph = (it_p / UM_STAGES) & 1u;
                    um_wait(&s_empty[s], ph ^ 1u);
                    if (lane == 0) {
                        const bool band = um_band(u0, a.Nt, c, n, uc.Rb);
                        const bool lo = band || uc.r3;
                        const uint32_t bytes = UM_CHUNK_BYTES * (1u + (lo ? 1u : 0u) + (band ? 2u : 0u));
                        ptx::mbarrier_arrive_expect_tx(ptx::sem_release, ptx::scope_cta, ptx::space_shared, &s_full[s], bytes);
                        uint8_t* dst = s_ring + (size_t)s * 4 * UM_CHUNK_BYTES;
                        const uint8_t* src = scr + (size_t)c * UM_CHUNK_BYTES;
                        ptx::cp_async_bulk(ptx::space_shared, ptx::space_global, dst, src, UM_CHUNK_BYTES, &s_full[s]);
                        if (lo)
                            ptx::cp_async_bulk(ptx::space_shared, ptx::space_global, dst + UM_CHUNK_BYTES, src + arr_stride,
                                               UM_CHUNK_BYTES, &s_full[s]);
                        if (band) {
                            ptx::cp_async_bulk(ptx::space_shared, ptx::space_global, dst + 2 * UM_CHUNK_BYTES,
                                               src + 2 * arr_stride, UM_CHUNK_BYTES, &s_full[s]);
                            ptx::cp_async_bulk(ptx::space_shared, ptx::space_global, dst + 3 * UM_CHUNK_BYTES,
                                               src + 3 * arr_stride, UM_CHUNK_BYTES, &s_full[s]);
                        }
                    }
                    __syncwarp();
                }
            } else if (wid == 1) {
                // ---- MMA issuer ----
                uint32_t e_on = 0, o_on = 0;
                const uint32_t tab0 = um_smem_u32(s_tab);
                const uint32_t tb = (uint32_t)a.tab_bytes;
                for (int c = 0; c < a.NC; ++c, ++it_c) {
                    const uint32_t s = it_c % UM_STAGES, ph = (it_c / UM_STAGES) & 1u;
                    um_wait(&s_full[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    if (lane == 0) {
                        const bool band = um_band(u0, a.Nt, c, n, uc.Rb);
                        const uint32_t rs = um_smem_u32(s_ring) + s * 4 * UM_CHUNK_BYTES;
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            const uint32_t aoff = j * 4096;
                            const uint32_t boff = (uint32_t)((u0 + UM_KC * c + 16 * j) >> 3) * 128u;
                            const uint64_t dEH = um_desc(rs + aoff, 2048u, 128u);
                            const uint64_t dEL = um_desc(rs + UM_CHUNK_BYTES + aoff, 2048u, 128u);
                            const uint64_t dTRh = um_desc(tab0 + 2 * tb + boff, 128u, 128u);
                            um_mma(tE, dEH, dTRh, idesc, e_on);
                            e_on = 1;
                            if (uc.r3) {
                                const uint64_t dTRl = um_desc(tab0 + 3 * tb + boff, 128u, 128u);
                                um_mma(tE, dEL, dTRh, idesc, 1u);
                                um_mma(tE, dEH, dTRl, idesc, 1u);
                            }
                            if (band) {
                                const uint64_t dOH = um_desc(rs + 2 * UM_CHUNK_BYTES + aoff, 2048u, 128u);
                                const uint64_t dOL = um_desc(rs + 3 * UM_CHUNK_BYTES + aoff, 2048u, 128u);
                                const uint64_t dTBh = um_desc(tab0 + boff, 128u, 128u);
                                const uint64_t dTBl = um_desc(tab0 + tb + boff, 128u, 128u);
                                if (!uc.r3) {  // with r3 the TR table already holds the whole even kernel
                                    um_mma(tE, dEH, dTBh, idesc, 1u);
                                    um_mma(tE, dEL, dTBh, idesc, 1u);
                                    um_mma(tE, dEH, dTBl, idesc, 1u);
                                }
                                um_mma(tO, dOH, dTBh, idesc, o_on);
                                o_on = 1;
                                um_mma(tO, dOL, dTBh, idesc, 1u);
                                um_mma(tO, dOH, dTBl, idesc, 1u);
                            }
                        }
                        um_commit(&s_empty[s]);  // the slot is free once these MMAs have read it
                    }
                    __syncwarp();
                }
                if (lane == 0) um_commit(&s_acc);  // accumulators of the pass complete
                __syncwarp();
            }
            // ---- epilogue: every warp; warp w reads TMEM lanes 32 (w % 4) .. + 31 (= rows) ----
            um_wait(&s_acc, acc_phase);
            acc_phase ^= 1u;
            asm volatile("tcgen05.fence::after_thread_sync;");
            {
                bool o_valid = false;
                for (int c = 0; c < a.NC; ++c) o_valid = o_valid || um_band(u0, a.Nt, c, n, uc.Rb);
                const int q = wid & 3, hsel = wid >> 2;
                const int rloc = 32 * q + lane;
                const bool rvalid = rloc < nrows;
                float* orow = a.cH + (size_t)z * a.pstride + (size_t)(row0 + (rvalid ? rloc : 0)) * a.pitch;
                const unsigned* mrow = s_mask + (rvalid ? rloc : 0) * a.mw;
                const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
                for (int g16 = hsel; g16 < (a.Nt >> 4); g16 += 2) {
                    const int ua = u0 + 16 * g16;
                    if (ua >= a.nout) break;  // warp-uniform
                    uint32_t ve[16], vo[16];
                    um_tmem_ld16(lane_base + 16 * g16, ve);
                    if (o_valid) {
                        um_tmem_ld16(lane_base + 256 + 16 * g16, vo);
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) vo[j] = 0u;
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (rvalid) {
                        const unsigned dbits = mrow[ua >> 5] >> (ua & 31);  // bit j <-> t = ua + j
                        // bits of t = n - ua - 15 .. n - ua  (bit 15 - j <-> t = n - ua - j)
                        const int lo = n - ua - 15, loc = max(lo, 0);
                        const int wi = loc >> 5;
                        const unsigned mbits = __funnelshift_r(mrow[wi], mrow[min(wi + 1, a.mw - 1)], loc & 31) << (loc - lo);
                        float vd[16], vm[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float E = __uint_as_float(ve[j]) * inv, O = __uint_as_float(vo[j]) * inv;
                            // (B x)[u] = E - O,  (B x)[n - u] = E + O;  dH = masked ? 0 : -(B x)
                            vd[j] = ((dbits >> j) & 1u) ? 0.f : (O - E);
                            vm[j] = ((mbits >> (15 - j)) & 1u) ? 0.f : -(E + O);
                        }
                        if (a.vec_ok && ua + 15 < a.nout) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)
                                *reinterpret_cast<float4*>(orow + ua + 4 * k) = make_float4(vd[4 * k], vd[4 * k + 1], vd[4 * k + 2], vd[4 * k + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (ua + j < a.nout) orow[ua + j] = vd[j];
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int u = ua + j;
                            if (u >= 1 && u < a.nout && 2 * u != n) orow[n - u] = vm[j];
                        }
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            __syncthreads();  // accumulators drained: the next pass may overwrite them
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

}  // namespace dstr
