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
// LBO = SBO = 128 bytes ONE table of 16 n bytes serves every (output tile, k step): the operator
// of a 1026-wide band lives in shared memory (49 KB) instead of being streamed per tile.
//
// Precision: operands are fp16 pairs (hi + lo, power-of-two pre-scaling), products
// hi*hi + lo*hi + hi*lo accumulated in fp32 in TMEM (measured 1.5e-6 of max |B x| on random rows,
// tools/probes/umma_probe.cu).  ha = hb_band + r:  hb is a compact Gaussian (radius Rb: only the
// k chunks within Rb of u + v = 0 (mod n) are multiplied, and only those windows of its table are
// stored), the remainder r (1/u^2 tails of the kink of a_j at j = 0, |r|_2 ~ 3e-3) runs over the
// full circle, with a single fp16 product where that is below the tolerance (level 1) and with the
// three-product split otherwise.
//
// One persistent, warp-specialised CTA per SM (512 threads); an item = up to 128 consecutive rows
// of one plane (MMA M = 128); items are double-buffered so that the roles overlap:
//   prep (10 warps)  one warp per row: load, mask, exact median, in-paint, pre-scale (row rewritten in
//                    place), then X_e / X_o as fp16 hi / lo in UMMA chunk order -> per-CTA scratch in
//                    global memory (L2-resident), mask bits -> shared memory
//   TMA  (warp 0)    cp.async.bulk scratch -> 3-stage shared-memory ring (mbarrier complete_tx)
//   MMA  (warp 1)    one thread issues tcgen05.mma kind::f16 (M 128, N = outputs of the pass <= 128,
//                    K 16); tcgen05.commit frees ring slots / publishes the accumulators
//   epilogue (warps 4-7)  tcgen05.ld of the E and O accumulators, combine, mask, dH stored in place
// The E and O accumulators of a pass take 2 x 128 TMEM columns; two sets alternate, so the MMAs of
// pass p + 1 run while pass p is drained.
#pragma once
#include "dstr_kernels.cuh"

namespace dstr {

#ifndef DSTR_UM_THREADS
#define DSTR_UM_THREADS 640
#endif
constexpr int UM_THREADS = DSTR_UM_THREADS;
constexpr int UM_ROWS = 128;                          // rows of an item = MMA M
constexpr int UM_KC = 32;                             // k elements per chunk (two K = 16 MMAs)
constexpr int UM_CHUNK_BYTES = UM_ROWS * UM_KC * 2;   // 8 KB: [k / 8][row / 8][row % 8][8 halfs]
constexpr int UM_STAGES = 3;
constexpr int UM_NT_MAX = 128;                        // outputs per pass (two E + O accumulator sets in 512 columns)
constexpr int UM_NPREP = UM_THREADS / 32 - 6;        // prep warps: 2, 3 and 8 .. (all but TMA, MMA and the epilogue)
constexpr int UM_NEPI = 4;                            // epilogue warps: 4..7 (warp % 4 = TMEM lane quarter)
constexpr float UM_TABLE_SCALE = 256.0f;

// Tables of one config in device memory, copied verbatim to shared memory:
//   [TRh (tr_bytes) | TRl (tr_bytes, only if r3) | TBh (tb_bytes) | TBl (tb_bytes)]
// TR is addressed by k directly (block k / 8).  TB either the same (tb_compact = 0) or as two windows:
// k in [0, 8 tb_k0_blocks) at offset 0 and k in [tb_k1, ...) at byte offset tb_w1_off.
struct UmmaCfg {
    const uint4* tables;
    int Rb;                         // band radius of the compact kernel
    int r3;                         // remainder kernel with the three-product split (then TR holds all of ha)
    int tr_bytes, tb_bytes;
    int tb_compact, tb_k1, tb_w1_off;
    unsigned long long need_band;   // chunks that are in band in at least one pass
};

struct UmmaLevelArgs {
    float* cH;
    int Hl, n, pitch;
    size_t pstride;
    const LevelStat* lstat;
    int stat_stride;
    int nh, nout;       // n / 2, outputs u = 0..nh
    int P, Nt;          // passes, outputs per pass (multiple of 16, <= 128)
    int NC, Kpad;       // k chunks, 32 NC >= n
    int rows_per_item, items_per_plane, n_items;
    int tab_max;        // shared-memory bytes reserved for the tables (max over the configs)
    int mw;             // mask words per row
    int vec_ok;
    UmmaCfg cfg[2];
    uint8_t* scratch;
    size_t scratch_stride;  // bytes per (CTA, buffer): 4 arrays x NC chunks; a CTA owns two buffers
    long long* prof;        // optional [grid][32 warps][8] cycle counters (DSTR_UMMA_PROF=1)
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

// one elected lane of a converged warp
__device__ __forceinline__ bool um_elect() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void um_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(um_smem_u32(bar))
                 : "memory");
}

// Wait for the phase with the given parity.  A failed probe is followed by a short sleep: a warp that
// spins on try_wait competes for issue slots with the warps doing the work (40 % of the issued instructions
// of the first version of this kernel were such spins).
template <unsigned SLEEP_NS>
__device__ __forceinline__ void um_wait_ns(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = um_smem_u32(bar);
    uint32_t done;
    // watchdog: a launch of this kernel takes a few milliseconds; a wait that lasts seconds means a lost arrival
    // (a bug), and a trap that fails the launch is better than a kernel that never returns
    unsigned long long t0 = 0;
    for (unsigned spins = 0;; ++spins) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) break;
        if (SLEEP_NS) __nanosleep(SLEEP_NS);
        if ((spins & 0x3ffu) == 0x3ffu) {
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ull) __trap();  // 4 s
        }
    }
}
__device__ __forceinline__ void um_wait(uint64_t* bar, uint32_t parity) { um_wait_ns<1000>(bar, parity); }
__device__ __forceinline__ void um_wait_fast(uint64_t* bar, uint32_t parity) { um_wait_ns<64>(bar, parity); }

__device__ __forceinline__ void um_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}

__device__ __forceinline__ void um_tmem_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
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

__device__ __forceinline__ void um_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(um_smem_u32(bar)) : "memory");
}

// byte offset of the TB core-matrix block that holds k (k multiple of 8) for a band chunk starting at k = lo
__host__ __device__ __forceinline__ uint32_t um_tb_off(const UmmaCfg& uc, int lo, int k) {
    if (!uc.tb_compact || lo <= uc.Rb) return (uint32_t)(k >> 3) * 128u;
    return (uint32_t)uc.tb_w1_off + (uint32_t)((k - uc.tb_k1) >> 3) * 128u;
}

// chunks c of a pass whose k range comes within Rb of u + v = 0 (mod n): c <= ca or cb <= c <= cc
// (same set as um_band; the 2n window never occurs for eligible geometries)
struct UmBand {
    int ca, cb, cc;
    __host__ __device__ __forceinline__ bool in(int c) const { return c <= ca || (c >= cb && c <= cc); }
};
__host__ __device__ __forceinline__ UmBand um_band_range(int u0, int Nt, int n, int Rb) {
    UmBand b;
    b.ca = (u0 <= Rb) ? (Rb - u0) / UM_KC : -1;
    const int t = n - Rb - Nt - (UM_KC - 2) - u0;  // lo >= t
    b.cb = t <= 0 ? 0 : (t + UM_KC - 1) / UM_KC;
    b.cc = (n + Rb - u0) / UM_KC;
    return b;
}

struct UmItem {
    int z, row0, nrows;
    float thr_q, scale, inv;
};

__device__ __forceinline__ UmItem um_item(const UmmaLevelArgs& a, int item) {
    UmItem it;
    it.z = item / a.items_per_plane;
    it.row0 = (item - it.z * a.items_per_plane) * a.rows_per_item;
    it.nrows = min(a.rows_per_item, a.Hl - it.row0);
    const LevelStat* st = a.lstat + (size_t)it.z * a.stat_stride;
    it.thr_q = st->thr_q;
    // power-of-two pre-scale: |x| <= thr, so |x_e|, |x_o| <= 2 thr < 2^15 after scaling (fp16 range)
    float scale = 1.0f;
    const float thr = st->thr;
    if (thr > 0.f && thr < 1e30f) {
        int e;
        frexpf(thr, &e);  // thr < 2^e
        scale = ldexpf(1.0f, max(-24, min(14 - e, 40)));
    }
    it.scale = scale;
    it.inv = 1.0f / (scale * UM_TABLE_SCALE);
    return it;
}

template <int EPL>
__global__ void __launch_bounds__(UM_THREADS, 1)
notch_umma_kernel(UmmaLevelArgs a, const PlaneStat* __restrict__ pstat, DispatchParams dp) {
    extern __shared__ __align__(1024) uint8_t um_smem[];
    __shared__ __align__(8) uint64_t s_full[UM_STAGES], s_empty[UM_STAGES];
    __shared__ __align__(8) uint64_t s_acc_full[2], s_acc_empty[2];
    __shared__ __align__(8) uint64_t s_prep_done[2], s_scratch_free[2], s_mask_free[2];
    __shared__ uint32_t s_tmem;
    namespace ptx = cuda::ptx;

    uint8_t* s_tab = um_smem;                                                                 // tables of the current config
    uint8_t* s_ring = s_tab + a.tab_max;                                                      // [stage][EH|EL|OH|OL]
    unsigned* s_mask = reinterpret_cast<unsigned*>(s_ring + UM_STAGES * 4 * UM_CHUNK_BYTES);  // [2][128][mw]

    const int tid = threadIdx.x, lane = tid & 31;
    // the warp index through a shuffle: provably warp-uniform for the compiler, so the role branches are uniform
    // branches and the TMA / MMA issue loops run on the uniform datapath (descriptors in uniform registers)
    const int wid = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int n = a.n;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < UM_STAGES; ++i) {
            ptx::mbarrier_init(&s_full[i], 1);
            ptx::mbarrier_init(&s_empty[i], 1);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            ptx::mbarrier_init(&s_acc_full[i], 1);
            ptx::mbarrier_init(&s_acc_empty[i], UM_NEPI);
            ptx::mbarrier_init(&s_prep_done[i], UM_NPREP);
            ptx::mbarrier_init(&s_scratch_free[i], 1);
            ptx::mbarrier_init(&s_mask_free[i], UM_NEPI);
        }
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

    // sequence counters: every role walks the same items in the same order, so the barrier phases
    // follow from the counters alone (they keep running across the two config sweeps)
    long long t_wait = 0, t_busy = 0, t_items = 0, t_x = 0;  // role profile (only stored when a.prof is set)
    const long long t_begin = clock64();
#define UM_TIMED_WAIT_(fn, bar, par)             \
    do {                                         \
        if (a.prof) {                            \
            const long long t0_ = clock64();     \
            fn(bar, par);                        \
            t_wait += clock64() - t0_;           \
        } else {                                 \
            fn(bar, par);                        \
        }                                        \
    } while (0)
#define UM_TIMED_WAIT(bar, par) UM_TIMED_WAIT_(um_wait, bar, par)
#define UM_TIMED_WAIT_FAST(bar, par) UM_TIMED_WAIT_(um_wait_fast, bar, par)
    uint32_t seq = 0;    // items processed by this role
    uint32_t pseq = 0;   // passes (accumulator set = pseq & 1)
    uint32_t cseq = 0;   // ring chunks
    uint8_t* scr0 = a.scratch + (size_t)blockIdx.x * 2 * a.scratch_stride;
    const size_t arr_stride = (size_t)a.NC * UM_CHUNK_BYTES;
    const uint32_t idesc = (1u << 4) | ((uint32_t)(a.Nt >> 3) << 17) | ((uint32_t)(UM_ROWS >> 4) << 24);
    const int ncfg = dp.mode ? 2 : 1;

    for (int cfg = 0; cfg < ncfg; ++cfg) {
        const UmmaCfg uc = a.cfg[cfg];
        // does this CTA have an item of this config at all? (CTA-uniform)
        bool any = false;
        for (int item = blockIdx.x; item < a.n_items && !any; item += gridDim.x)
            any = plane_uses_cells(pstat[item / a.items_per_plane], dp) == cfg;
        if (!any) continue;
        {
            const int nvec = (uc.tr_bytes * (uc.r3 ? 2 : 1) + 2 * uc.tb_bytes) / 16;
            uint4* dst = reinterpret_cast<uint4*>(s_tab);
            for (int i = tid; i < nvec; i += UM_THREADS) dst[i] = __ldg(uc.tables + i);
            asm volatile("fence.proxy.async;" ::: "memory");  // table stores -> visible to the MMA operand reads
            __syncthreads();
        }
        const uint32_t tabTRh = um_smem_u32(s_tab);
        const uint32_t tabTRl = tabTRh + uc.tr_bytes;
        const uint32_t tabTBh = tabTRh + uc.tr_bytes * (uc.r3 ? 2 : 1);
        const uint32_t tabTBl = tabTBh + uc.tb_bytes;
        (void)tabTRl;
        (void)tabTBl;
        const uint64_t dA0 = um_desc(um_smem_u32(s_ring), 2048u, 128u);  // X chunk of ring slot 0 (EH), k step 0
        const uint64_t dTRh0 = um_desc(tabTRh, 128u, 128u);
        const uint64_t dTBh0 = um_desc(tabTBh, 128u, 128u);

        if (wid == 0) {
            // ======================= TMA producer =======================
            for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
                if (plane_uses_cells(pstat[item / a.items_per_plane], dp) != cfg) continue;
                const uint32_t buf = seq & 1u, use = seq >> 1;
                UM_TIMED_WAIT(&s_prep_done[buf], use & 1u);
                const uint8_t* scr = scr0 + (size_t)buf * a.scratch_stride;
                for (int p = 0; p < a.P; ++p) {
                    const int u0 = p * a.Nt;
                    const UmBand bd = um_band_range(u0, a.Nt, n, uc.Rb);
                    uint32_t s = cseq % UM_STAGES, ph = (cseq / UM_STAGES) & 1u;
                    for (int c = 0; c < a.NC; ++c, ++cseq) {
                        UM_TIMED_WAIT_FAST(&s_empty[s], ph ^ 1u);
                        if (um_elect()) {
                            const bool band = bd.in(c);
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
                        if (++s == UM_STAGES) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                }
                ++seq;
            }
        } else if (wid == 1) {
            // ======================= MMA issuer =======================
            for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
                if (plane_uses_cells(pstat[item / a.items_per_plane], dp) != cfg) continue;
                const uint32_t buf = seq & 1u;
                for (int p = 0; p < a.P; ++p, ++pseq) {
                    const int u0 = p * a.Nt;
                    const uint32_t set = pseq & 1u, use = pseq >> 1;
                    const uint32_t tE = tmem_base + set * 256u, tO = tE + 128u;
                    UM_TIMED_WAIT(&s_acc_empty[set], (use & 1u) ^ 1u);  // the epilogue has drained this set
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    uint32_t e_on = 0, o_on = 0;
                    const UmBand bd = um_band_range(u0, a.Nt, n, uc.Rb);
                    uint32_t s = cseq % UM_STAGES, ph = (cseq / UM_STAGES) & 1u;
                    uint32_t roff16 = (uint32_t)(u0 >> 3) * 8u;  // TR block offset of the k step, in 16-byte units
                    for (int c = 0; c < a.NC; ++c, ++cseq) {
                        UM_TIMED_WAIT_FAST(&s_full[s], ph);
                        asm volatile("tcgen05.fence::after_thread_sync;");
                        if (um_elect()) {
                            const bool band = bd.in(c);
                            const uint64_t aS = (uint64_t)(s * (4 * UM_CHUNK_BYTES / 16));  // ring slot, 16-byte units
                            const int lo = u0 + UM_KC * c;
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                // descriptors differ from the constant bases only in the 14-bit start-address field
                                const uint64_t dEH = dA0 + aS + (uint64_t)(j * 256);
                                const uint64_t dTRh = dTRh0 + roff16 + (uint64_t)(j * 16);
                                um_mma(tE, dEH, dTRh, idesc, (e_on | (uint32_t)j) ? 1u : 0u);
                                if (uc.r3) {
                                    um_mma(tE, dEH + (UM_CHUNK_BYTES / 16), dTRh, idesc, 1u);
                                    um_mma(tE, dEH, dTRh + (uint64_t)(uc.tr_bytes >> 4), idesc, 1u);
                                }
                                if (band) {
                                    const uint64_t boff16 = um_tb_off(uc, lo, lo + 16 * j) >> 4;
                                    const uint64_t dEL = dEH + (UM_CHUNK_BYTES / 16);
                                    const uint64_t dOH = dEH + 2 * (UM_CHUNK_BYTES / 16), dOL = dEH + 3 * (UM_CHUNK_BYTES / 16);
                                    const uint64_t dTBh = dTBh0 + boff16, dTBl = dTBh + (uint64_t)(uc.tb_bytes >> 4);
                                    if (!uc.r3) {  // with r3 the TR table already holds the whole even kernel
                                        um_mma(tE, dEH, dTBh, idesc, 1u);
                                        um_mma(tE, dEL, dTBh, idesc, 1u);
                                        um_mma(tE, dEH, dTBl, idesc, 1u);
                                    }
                                    um_mma(tO, dOH, dTBh, idesc, (o_on | (uint32_t)j) ? 1u : 0u);
                                    um_mma(tO, dOL, dTBh, idesc, 1u);
                                    um_mma(tO, dOH, dTBl, idesc, 1u);
                                }
                            }
                            um_commit(&s_empty[s]);  // the slot is free once these MMAs have read it
                            // every bulk copy of the item has landed: its scratch buffer may be rewritten
                            if (p == a.P - 1 && c == a.NC - 1) um_arrive(&s_scratch_free[buf]);
                        }
                        __syncwarp();
                        // accumulate flags are kept by every lane (whichever lane is elected next sees them)
                        e_on = 1;
                        if (bd.in(c)) o_on = 1;
                        roff16 += 32u;  // 32 k = 4 blocks of 128 bytes
                        if (++s == UM_STAGES) {
                            s = 0;
                            ph ^= 1u;
                        }
                    }
                    if (um_elect()) um_commit(&s_acc_full[set]);  // accumulators of the pass complete
                    __syncwarp();
                }
                ++seq;
            }
        } else if (wid >= 4 && wid < 8) {
            // ======================= epilogue =======================
            const int q = wid & 3;
            const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
            for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
                if (plane_uses_cells(pstat[item / a.items_per_plane], dp) != cfg) continue;
                const UmItem it = um_item(a, item);
                const uint32_t buf = seq & 1u;
                UM_TIMED_WAIT(&s_prep_done[buf], (seq >> 1) & 1u);  // mask bits of the item are in place
                const int rloc = 32 * q + lane;
                const bool rvalid = rloc < it.nrows;
                float* orow = a.cH + (size_t)it.z * a.pstride + (size_t)(it.row0 + (rvalid ? rloc : 0)) * a.pitch;
                const unsigned* mrow = s_mask + ((size_t)buf * UM_ROWS + (rvalid ? rloc : 0)) * a.mw;
                for (int p = 0; p < a.P; ++p, ++pseq) {
                    const int u0 = p * a.Nt;
                    const uint32_t set = pseq & 1u, use = pseq >> 1;
                    const UmBand bd = um_band_range(u0, a.Nt, n, uc.Rb);
                    const bool o_valid = bd.ca >= 0 || (bd.cb <= bd.cc && bd.cb < a.NC);
                    UM_TIMED_WAIT(&s_acc_full[set], use & 1u);
                    asm volatile("tcgen05.fence::after_thread_sync;");
                    for (int g8 = 0; g8 < (a.Nt >> 3); ++g8) {
                        const int ua = u0 + 8 * g8;
                        if (ua >= a.nout) break;  // warp-uniform
                        uint32_t ve[8], vo[8];
                        um_tmem_ld8(lane_base + set * 256u + 8 * g8, ve);
                        if (o_valid) {
                            um_tmem_ld8(lane_base + set * 256u + 128u + 8 * g8, vo);
                        } else {
#pragma unroll
                            for (int j = 0; j < 8; ++j) vo[j] = 0u;
                        }
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                        if (rvalid) {
                            const unsigned dbits = mrow[ua >> 5] >> (ua & 31);  // bit j <-> t = ua + j
                            // bits of t = n - ua - 7 .. n - ua  (bit 7 - j <-> t = n - ua - j)
                            const int lo = n - ua - 7, loc = max(lo, 0);
                            const int wi = loc >> 5;
                            const unsigned mbits = __funnelshift_r(mrow[wi], mrow[min(wi + 1, a.mw - 1)], loc & 31) << (loc - lo);
                            float vd[8], vm[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float E = __uint_as_float(ve[j]) * it.inv, O = __uint_as_float(vo[j]) * it.inv;
                                // (B x)[u] = E - O,  (B x)[n - u] = E + O;  dH = masked ? 0 : -(B x)
                                vd[j] = ((dbits >> j) & 1u) ? 0.f : (O - E);
                                vm[j] = ((mbits >> (7 - j)) & 1u) ? 0.f : -(E + O);
                            }
                            if (a.vec_ok && ua + 7 < a.nout) {
                                *reinterpret_cast<float4*>(orow + ua) = make_float4(vd[0], vd[1], vd[2], vd[3]);
                                *reinterpret_cast<float4*>(orow + ua + 4) = make_float4(vd[4], vd[5], vd[6], vd[7]);
                            } else {
#pragma unroll
                                for (int j = 0; j < 8; ++j)
                                    if (ua + j < a.nout) orow[ua + j] = vd[j];
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int u = ua + j;
                                if (u >= 1 && u < a.nout && 2 * u != n) orow[n - u] = vm[j];
                            }
                        }
                    }
                    asm volatile("tcgen05.fence::before_thread_sync;");
                    __syncwarp();
                    if (lane == 0) {
                        um_arrive(&s_acc_empty[set]);
                        if (p == a.P - 1) um_arrive(&s_mask_free[buf]);
                    }
                }
                ++seq;
            }
        } else {
            // ======================= prep: selection, in-painting, X_e / X_o operands =======================
            const int pw = (wid < 4) ? wid - 2 : wid - 6;  // 0 .. UM_NPREP - 1
            const int rho = (n - 7) & 3;  // alignment of the reversed octet x[n - v0 - 7 .. n - v0]
            for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
                if (plane_uses_cells(pstat[item / a.items_per_plane], dp) != cfg) continue;
                const UmItem it = um_item(a, item);
                const uint32_t buf = seq & 1u, use = seq >> 1;
                UM_TIMED_WAIT(&s_scratch_free[buf], (use & 1u) ^ 1u);
                UM_TIMED_WAIT(&s_mask_free[buf], (use & 1u) ^ 1u);
                uint8_t* scr = scr0 + (size_t)buf * a.scratch_stride;
                unsigned* mbase = s_mask + (size_t)buf * UM_ROWS * a.mw;
                for (int rl = pw; rl < it.nrows; rl += UM_NPREP) {
                    const long long tp0 = a.prof ? clock64() : 0;
                    const float* grow = a.cH + (size_t)it.z * a.pstride + (size_t)(it.row0 + rl) * a.pitch;
                    if (lane == 0 && rl + UM_NPREP < it.nrows)  // the warp's next row: DRAM -> L2 while this one is processed
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(grow + (size_t)UM_NPREP * a.pitch), "r"(a.pitch * 4)
                                     : "memory");
                    // ---- pass 1: mask bits and the sign counts that decide whether the median is zero ----
                    // b = m ? 0 : c is the zero-filled background (filtering.py:197); with k1 = (n-1)/2, k2 = n/2 its
                    // median is 0 iff  #{b < 0} <= k1  and  #{b <= 0} > k2
                    int cneg = 0, cle0 = 0;
                    {
                        const float* gl = grow + lane;  // lanes past the row end read slack and are discarded
                        constexpr int H0 = (EPL + 1) / 2;  // two batches of loads bound the live registers
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int i0 = h ? H0 : 0, i1 = h ? EPL : H0;
                            float cv[H0];
#pragma unroll
                            for (int i = i0; i < i1; ++i) cv[i - i0] = gl[32 * i];
#pragma unroll
                            for (int i = i0; i < i1; ++i) {
                                const bool valid = lane + 32 * i < n;
                                const float c = cv[i - i0];
                                const bool m = valid && (__fmul_rn(c, c) > it.thr_q);  // sqrt(c*c) > thr, bit-identical (otsu_kernel)
                                cneg += (valid && !m && c < 0.f) ? 1 : 0;
                                cle0 += (valid && (m || c <= 0.f)) ? 1 : 0;
                                const unsigned mbits = __ballot_sync(0xffffffffu, m);
                                if (lane == 0 && i < a.mw) mbase[rl * a.mw + i] = mbits;
                            }
                        }
                    }
                    cneg = __reduce_add_sync(0xffffffffu, cneg);
                    cle0 = __reduce_add_sync(0xffffffffu, cle0);
                    const long long tp1 = a.prof ? clock64() : 0;
                    const int k1 = (n - 1) >> 1, k2 = n >> 1;
                    float med = 0.f;
                    if (!(cneg <= k1 && k2 < cle0)) {
                        // exact median (np.median, filtering.py:201) by bisection on order-preserving keys: rare
                        // (the masked zeros sit in the middle of a roughly symmetric distribution)
                        unsigned key[EPL];
                        {
                            const float* gl = grow + lane;
#pragma unroll
                            for (int i = 0; i < EPL; ++i) {
                                const float c = gl[32 * i];
                                const bool m = __fmul_rn(c, c) > it.thr_q;
                                key[i] = (lane + 32 * i < n) ? f2key(m ? 0.0f : (c + 0.0f)) : 0xffffffffu;
                            }
                        }
                        const unsigned KZ = 0x80000000u;
                        unsigned kk1;
                        const bool negside = k1 < cneg;
                        if (!negside && k1 < cle0) {
                            kk1 = KZ;  // the k1-th order statistic is one of the zeros (k2 is the first positive entry)
                        } else {
                            // the k1-th order statistic is strictly negative / strictly positive.  The keys of that side
                            // share their leading bits (sign, most of the exponent): bisect below the common prefix of
                            // the smallest and the largest of them
                            unsigned res;
                            int lo_cnt = negside ? 0 : cle0, hi_cnt = negside ? cneg : n;
                            int b0;
                            {
                                unsigned kmn = 0xffffffffu, kmx = 0u;
#pragma unroll
                                for (int i = 0; i < EPL; ++i) {
                                    const bool on = negside ? (key[i] < KZ) : (key[i] > KZ && key[i] != 0xffffffffu);
                                    kmn = on ? min(kmn, key[i]) : kmn;
                                    kmx = on ? max(kmx, key[i]) : kmx;
                                }
                                kmn = __reduce_min_sync(0xffffffffu, kmn);
                                kmx = __reduce_max_sync(0xffffffffu, kmx);
                                const unsigned diff = kmn ^ kmx;
                                b0 = diff ? 31 - __clz((int)diff) : -1;  // highest differing bit (<= 30: same sign)
                                res = (b0 >= 0) ? (kmn & ~((2u << b0) - 1u)) : kmn;
                            }
                            bool unique = (hi_cnt - lo_cnt) == 1;
                            for (int b = b0; b >= 0 && !unique; --b) {
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
                            // res is the key itself or a lower bound of the single key left in the bracket; on the
                            // positive side the bracket starts above the zeros (lo_cnt = cle0), the prefix may not
                            const unsigned lower = negside ? res : max(res, KZ + 1u);
                            kk1 = 0xffffffffu;
#pragma unroll
                            for (int i = 0; i < EPL; ++i)
                                if (key[i] >= lower) kk1 = min(kk1, key[i]);
                            kk1 = __reduce_min_sync(0xffffffffu, kk1);
                        }
                        med = key2f(kk1);
                        if (k2 != k1) {
                            int cle = 0;
                            unsigned nx2 = 0xffffffffu;
#pragma unroll
                            for (int i = 0; i < EPL; ++i) {
                                cle += (key[i] <= kk1) ? 1 : 0;
                                if (key[i] > kk1) nx2 = min(nx2, key[i]);
                            }
                            cle = __reduce_add_sync(0xffffffffu, cle);
                            nx2 = __reduce_min_sync(0xffffffffu, nx2);
                            const unsigned kk2 = (cle >= k1 + 2) ? kk1 : nx2;
                            med = (key2f(kk1) + key2f(kk2)) * 0.5f;
                        }
                    }
                    const long long tp2 = a.prof ? clock64() : 0;
                    // ---- pass 2: X_e[v] = x[v] + x[(n - v) mod n], X_o[v] = x[v] - x[(n - v) mod n] with the in-painted
                    // x = m ? med : c (pre-scaled), v = 0 .. Kpad - 1 (zero past n).  The row is re-read (L1 / L2) forwards
                    // and index-reversed; one 8-element core-matrix row (16 bytes of fp16) per lane and step.
                    uint8_t* dst_row = scr + (size_t)(rl >> 3) * 128 + (size_t)(rl & 7) * 16;
                    const int NO = a.Kpad >> 3;
                    // interior octets (all 16 inputs inside the row, no wrap): two forward quads and the three aligned
                    // quads that cover x[n - v0 - 7 .. n - v0]; their loads are issued one step ahead
                    float4 nq[5];
                    bool nfast;
                    auto issue = [&](int o, float4(&q)[5]) -> bool {
                        const int v0 = 8 * o;
                        const bool f = (o >= 1) && (v0 + 8 <= n);
                        if (f) {
                            const int rb4 = (n - v0 - 7) & ~3;
                            q[0] = *reinterpret_cast<const float4*>(grow + v0);
                            q[1] = *reinterpret_cast<const float4*>(grow + v0 + 4);
                            q[2] = *reinterpret_cast<const float4*>(grow + rb4);
                            q[3] = *reinterpret_cast<const float4*>(grow + rb4 + 4);
                            q[4] = *reinterpret_cast<const float4*>(grow + rb4 + 8);
                        }
                        return f;
                    };
                    nfast = issue(lane, nq);
                    for (int o = lane; o < NO; o += 32) {
                        const int v0 = 8 * o;
                        float fv[8], gv[8];
                        const bool fast = nfast;
                        float4 q[5];
#pragma unroll
                        for (int k = 0; k < 5; ++k) q[k] = nq[k];
                        nfast = (o + 32 < NO) ? issue(o + 32, nq) : false;
                        if (fast) {
                            fv[0] = q[0].x, fv[1] = q[0].y, fv[2] = q[0].z, fv[3] = q[0].w;
                            fv[4] = q[1].x, fv[5] = q[1].y, fv[6] = q[1].z, fv[7] = q[1].w;
                            const float t[12] = {q[2].x, q[2].y, q[2].z, q[2].w, q[3].x, q[3].y, q[3].z, q[3].w, q[4].x, q[4].y, q[4].z, q[4].w};
                            // gv[i] = c[n - v0 - i] = t[rho + 7 - i]
                            switch (rho) {
                                case 0:
#pragma unroll
                                    for (int i = 0; i < 8; ++i) gv[i] = t[7 - i];
                                    break;
                                case 1:
#pragma unroll
                                    for (int i = 0; i < 8; ++i) gv[i] = t[8 - i];
                                    break;
                                case 2:
#pragma unroll
                                    for (int i = 0; i < 8; ++i) gv[i] = t[9 - i];
                                    break;
                                default:
#pragma unroll
                                    for (int i = 0; i < 8; ++i) gv[i] = t[10 - i];
                                    break;
                            }
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                fv[i] = (__fmul_rn(fv[i], fv[i]) > it.thr_q) ? med : fv[i];
                                gv[i] = (__fmul_rn(gv[i], gv[i]) > it.thr_q) ? med : gv[i];
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const int v = v0 + i;
                                const bool in = v < n;
                                const float cf = in ? grow[v] : 0.f;
                                const float cr = in ? grow[v == 0 ? 0 : n - v] : 0.f;
                                fv[i] = in ? ((__fmul_rn(cf, cf) > it.thr_q) ? med : cf) : 0.f;
                                gv[i] = in ? ((__fmul_rn(cr, cr) > it.thr_q) ? med : cr) : 0.f;
                            }
                        }
                        float ev[8], ov[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            ev[i] = (fv[i] + gv[i]) * it.scale;
                            ov[i] = (fv[i] - gv[i]) * it.scale;
                        }
                        const int c = o >> 2;
                        const bool band = (uc.need_band >> c) & 1ull;
                        uint8_t* dstp = dst_row + (size_t)c * UM_CHUNK_BYTES + (size_t)(o & 3) * 2048;
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
                    if (a.prof) {
                        const long long tp3 = clock64();
                        t_busy += tp1 - tp0;   // pass 1
                        t_items += tp2 - tp1;  // median
                        t_x += tp3 - tp2;      // pass 2
                    }
                }
                // operand stores (generic proxy, global) -> ordered before the TMA reads (async proxy)
                asm volatile("fence.proxy.async;" ::: "memory");
                __syncwarp();
                if (lane == 0) um_arrive(&s_prep_done[buf]);
                ++seq;
            }
        }
        // the roles of this sweep are done issuing; the next sweep reloads the tables, so every MMA that reads
        // them must have completed: the epilogue warps have seen every accumulator of the sweep
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
        // (every role advanced its own counters by the same amounts, so the phases stay aligned in the next sweep)
    }
    if (a.prof && lane == 0) {
        long long* pr = a.prof + ((size_t)blockIdx.x * 32 + wid) * 8;
        pr[0] = t_wait;
        pr[1] = t_busy;
        pr[2] = t_items;
        pr[3] = clock64() - t_begin;
        pr[4] = t_x;
        pr[5] = (long long)seq;
    }
#undef UM_TIMED_WAIT
#undef UM_TIMED_WAIT_FAST
#undef UM_TIMED_WAIT_
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (wid == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

}  // namespace dstr
