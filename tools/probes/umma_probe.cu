// tcgen05 building blocks of the notch operator, checked in isolation on a B200 (sm_100a):
//   D[128 rows x N] = sum_k X[r][k] * f[n + k + off]          (a Hankel-structured B operand)
// with kind::f16 MMAs (M = 128, K = 16), the A operand (rows of X, K-major, no swizzle) written by
// ordinary stores, the B operand ALIASED out of one table that stores 8 shifted copies of f
// (core matrix (n / 8, k / 8) depends on n / 8 + k / 8 only: LBO = SBO = 128 bytes), accumulators in
// TMEM read back with tcgen05.ld.  Also: the fp16 hi/lo three-product scheme against a double
// reference, and cycles per MMA for the shapes the row filter uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CHECK(x)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (x);                                                            \
        if (e_ != cudaSuccess) {                                                         \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            exit(2);                                                                     \
        }                                                                                \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, no swizzle (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;
}

__device__ __forceinline__ void mma_f16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}

struct Args {
    const float* X;   // [128][K]
    const float* f;   // [Lt] table values
    float* D;         // [naccum][128][N]
    long long* cycles;
    int K, N, off, Lt;
    int variant;      // 0: LBO = K-direction stride, SBO = 8-row-group stride (CUTLASS reading); 1: swapped
    int split;        // 0: single fp16 product; 1: hi/lo three-product scheme
    int naccum;       // accumulators (1 or 2: the E and O bands of the row filter)
    int repeat;       // repeat the MMA sequence (timing)
    float sx, st;     // power-of-two scales of X and f before the fp16 conversion
};

constexpr int KC = 32;  // k elements per A chunk

// A chunk layout: [kc = k / 8 (4)][row group (16)][row % 8][8 halfs]  -> LBO 2048, SBO 128
__global__ void __launch_bounds__(128, 1) umma_probe(Args a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nchunk = a.K / KC;
    const int chunk_bytes = 128 * KC * 2;  // 8 KB
    __half* A_hi = reinterpret_cast<__half*>(smem);
    __half* A_lo = reinterpret_cast<__half*>(smem + (size_t)nchunk * chunk_bytes);
    const int tab_bytes = ((a.Lt + 7) / 8) * 128;
    __half* T_hi = reinterpret_cast<__half*>(smem + (size_t)2 * nchunk * chunk_bytes);
    __half* T_lo = reinterpret_cast<__half*>(smem + (size_t)2 * nchunk * chunk_bytes + tab_bytes);

    // ---- operands by ordinary (generic proxy) stores ------------------------------------------
    for (int i = tid; i < 128 * a.K; i += 128) {
        const int r = i / a.K, k = i - r * a.K;
        const float x = a.X[i] * a.sx;
        const __half h = __float2half_rn(x);
        const __half l = __float2half_rn(x - __half2float(h));
        const int c = k / KC, kk = k % KC;
        const size_t o = (size_t)c * (128 * KC) + (size_t)(kk >> 3) * (16 * 64) + (size_t)(r >> 3) * 64 + (r & 7) * 8 + (kk & 7);
        A_hi[o] = h;
        A_lo[o] = l;
    }
    for (int i = tid; i < ((a.Lt + 7) / 8) * 64; i += 128) {
        // element t of shifted copy r lives at (t / 8) * 64 + r * 8 + t % 8 and holds f[r + t]
        const int blk = i >> 6, r = (i >> 3) & 7, e = i & 7;
        const int src = r + blk * 8 + e;
        const float v = (src < a.Lt ? a.f[src] : 0.f) * a.st;
        const __half h = __float2half_rn(v);
        T_hi[i] = h;
        T_lo[i] = __float2half_rn(v - __half2float(h));
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic stores -> visible to the MMA's async proxy
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem_base = tmem_base_s;

    const uint32_t idesc = (1u << 4) | ((uint32_t)(a.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);  // f16 x f16 -> f32, K-major both
    long long t0 = 0, t1 = 0;
    if (tid == 0) {
        const uint32_t lboA = a.variant ? 128u : 2048u, sboA = a.variant ? 2048u : 128u;
        t0 = clock64();
        for (int rep = 0; rep < a.repeat; ++rep) {
            for (int acc = 0; acc < a.naccum; ++acc) {
                const uint32_t dcol = tmem_base + acc * 256;
                for (int ks = 0; ks < a.K / 16; ++ks) {
                    const int c = ks / (KC / 16), sub = ks % (KC / 16);
                    const uint32_t aoff = c * chunk_bytes + sub * 2 * 2048;
                    const uint32_t boff = ((a.off + ks * 16) / 8) * 128;
                    const uint64_t dAh = make_desc(smem_u32(A_hi) + aoff, lboA, sboA);
                    const uint64_t dAl = make_desc(smem_u32(A_lo) + aoff, lboA, sboA);
                    const uint64_t dTh = make_desc(smem_u32(T_hi) + boff, 128u, 128u);
                    const uint64_t dTl = make_desc(smem_u32(T_lo) + boff, 128u, 128u);
                    mma_f16(dcol, dAh, dTh, idesc, (ks > 0) ? 1u : 0u);
                    if (a.split) {
                        mma_f16(dcol, dAl, dTh, idesc, 1u);
                        mma_f16(dcol, dAh, dTl, idesc, 1u);
                    }
                }
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    mbar_wait(smem_u32(&mbar), 0);
    if (tid == 0) {
        t1 = clock64();
        a.cycles[0] = t1 - t0;
    }
    asm volatile("tcgen05.fence::after_thread_sync;");

    // ---- accumulators: warp w reads TMEM lanes 32 w .. 32 w + 31 (= rows), 16 columns per load ----
    const float inv = 1.0f / (a.sx * a.st);
    for (int acc = 0; acc < a.naccum; ++acc) {
        for (int c0 = 0; c0 < a.N; c0 += 16) {
            uint32_t v[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 256 + c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            const int row = warp * 32 + lane;
            for (int j = 0; j < 16; ++j) a.D[((size_t)acc * 128 + row) * a.N + c0 + j] = __uint_as_float(v[j]) * inv;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
}

static float half_round(float x) { return __half2float(__float2half_rn(x)); }

int main() {
    cudaDeviceProp p;
    CHECK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d\n", p.name, p.major, p.minor);
    CHECK(cudaFuncSetAttribute(umma_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024));
    srand(1);
    struct Case {
        int K, N, off, variant, split, naccum, repeat;
    };
    const Case cases[] = {
        {64, 64, 0, 0, 0, 1, 1},    {64, 176, 8, 0, 0, 1, 1},   {64, 256, 264, 0, 0, 1, 1},
        {64, 16, 16, 0, 0, 1, 1},   {192, 176, 352, 0, 1, 2, 1}, {320, 256, 0, 0, 1, 2, 1},  {320, 176, 0, 0, 1, 2, 40},
        {320, 256, 0, 0, 1, 2, 40}, {320, 256, 0, 0, 0, 2, 40},  {320, 128, 0, 0, 1, 2, 40}, {320, 64, 0, 0, 1, 2, 40},
    };
    for (const Case& c : cases) {
        const int Lt = c.off + c.N + c.K + 16;
        std::vector<float> X((size_t)128 * c.K), f(Lt);
        for (auto& v : X) v = 0.8f * ((float)rand() / RAND_MAX - 0.5f);
        for (int i = 0; i < Lt; ++i) f[i] = 0.04f * std::exp(-0.5f * (float)((i % 97) * (i % 97)) / 400.f) * ((i & 1) ? 1.f : -0.7f);
        float *dX, *df, *dD;
        long long* dc;
        CHECK(cudaMalloc(&dX, X.size() * 4));
        CHECK(cudaMalloc(&df, f.size() * 4));
        CHECK(cudaMalloc(&dD, (size_t)2 * 128 * c.N * 4));
        CHECK(cudaMalloc(&dc, 8));
        CHECK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
        CHECK(cudaMemcpy(df, f.data(), f.size() * 4, cudaMemcpyHostToDevice));
        CHECK(cudaMemset(dD, 0, (size_t)2 * 128 * c.N * 4));
        Args a;
        a.X = dX;
        a.f = df;
        a.D = dD;
        a.cycles = dc;
        a.K = c.K;
        a.N = c.N;
        a.off = c.off;
        a.Lt = Lt;
        a.variant = c.variant;
        a.split = c.split;
        a.naccum = c.naccum;
        a.repeat = c.repeat;
        a.sx = 1024.f;
        a.st = 256.f;
        const size_t smem = (size_t)2 * (c.K / KC) * 128 * KC * 2 + (size_t)2 * ((Lt + 7) / 8) * 128;
        if (smem > 226 * 1024) {
            printf("case skipped: smem %zu\n", smem);
            continue;
        }
        umma_probe<<<1, 128, smem>>>(a);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) {
            printf("K=%d N=%d off=%d variant=%d: kernel failed: %s\n", c.K, c.N, c.off, c.variant, cudaGetErrorString(e));
            return 3;
        }
        std::vector<float> D((size_t)c.naccum * 128 * c.N);
        long long cyc = 0;
        CHECK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
        CHECK(cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost));
        // references: exact (double, unrounded inputs) and fp16-rounded inputs (what a single product computes)
        double err_exact = 0, err_h = 0, ref_max = 0;
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < c.N; ++n) {
                double se = 0, sh = 0;
                for (int k = 0; k < c.K; ++k) {
                    const float x = X[(size_t)r * c.K + k], t = f[n + k + c.off];
                    se += (double)x * t;
                    sh += (double)half_round(x * a.sx) * half_round(t * a.st) / (a.sx * a.st);
                }
                ref_max = std::max(ref_max, std::fabs(se));
                for (int acc = 0; acc < c.naccum; ++acc) {
                    const double d = D[((size_t)acc * 128 + r) * c.N + n];
                    err_exact = std::max(err_exact, std::fabs(d - se));
                    err_h = std::max(err_h, std::fabs(d - sh));
                }
            }
        const long long nmma = (long long)c.repeat * c.naccum * (c.K / 16) * (c.split ? 3 : 1);
        printf("K=%4d N=%3d off=%3d variant=%d split=%d acc=%d rep=%2d | max|ref| %.4f  err vs exact %.3e  err vs fp16-input ref %.3e | %lld cycles, %lld MMAs, %.1f cyc/MMA\n",
               c.K, c.N, c.off, c.variant, c.split, c.naccum, c.repeat, ref_max, err_exact, err_h, cyc, nmma,
               (double)cyc / (double)nmma);
        cudaFree(dX);
        cudaFree(df);
        cudaFree(dD);
        cudaFree(dc);
    }
    return 0;
}
