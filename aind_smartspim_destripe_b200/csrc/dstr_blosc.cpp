// Blosc1 frames (zstd / lz4 streams, byte shuffle) for the Zarr chunk I/O of the tile driver: the
// reference writes blosc-zstd-3-shuffle chunks (/root/reference/code/aind_smartspim_destripe/
// zarr_destriper.py:1066-1074).  Host-only code; libzstd.so.1 / liblz4.so.1 are bound with dlopen at
// first use, so the CUDA library has no link-time dependency on them.  The container format is
// described in aind_smartspim_destripe_b200/blosc1.py (the pure-Python implementation of the same
// format, against which this one is tested).
#include <dlfcn.h>

#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include "../../include/dstr_b200.h"

namespace {

constexpr int kHeader = 16;
constexpr int kMinBuffer = 128;
constexpr uint8_t kShuffle = 0x01, kMemcpyed = 0x02, kBitShuffle = 0x04, kDontSplit = 0x10;

struct Codecs {
    size_t (*zstd_bound)(size_t) = nullptr;
    size_t (*zstd_compress)(void*, size_t, const void*, size_t, int) = nullptr;
    size_t (*zstd_decompress)(void*, size_t, const void*, size_t) = nullptr;
    unsigned (*zstd_is_error)(size_t) = nullptr;
    int (*zstd_max_level)() = nullptr;
    int (*lz4_bound)(int) = nullptr;
    int (*lz4_compress)(const char*, char*, int, int) = nullptr;
    int (*lz4_decompress)(const char*, char*, int, int) = nullptr;
};

Codecs g_codecs;
std::once_flag g_once;

void load_codecs() {
    if (void* z = dlopen("libzstd.so.1", RTLD_NOW | RTLD_GLOBAL)) {
        g_codecs.zstd_bound = (size_t(*)(size_t))dlsym(z, "ZSTD_compressBound");
        g_codecs.zstd_compress = (size_t(*)(void*, size_t, const void*, size_t, int))dlsym(z, "ZSTD_compress");
        g_codecs.zstd_decompress = (size_t(*)(void*, size_t, const void*, size_t))dlsym(z, "ZSTD_decompress");
        g_codecs.zstd_is_error = (unsigned (*)(size_t))dlsym(z, "ZSTD_isError");
        g_codecs.zstd_max_level = (int (*)())dlsym(z, "ZSTD_maxCLevel");
    }
    if (void* l = dlopen("liblz4.so.1", RTLD_NOW | RTLD_GLOBAL)) {
        g_codecs.lz4_bound = (int (*)(int))dlsym(l, "LZ4_compressBound");
        g_codecs.lz4_compress = (int (*)(const char*, char*, int, int))dlsym(l, "LZ4_compress_default");
        g_codecs.lz4_decompress = (int (*)(const char*, char*, int, int))dlsym(l, "LZ4_decompress_safe");
    }
}

const Codecs& codecs() {
    std::call_once(g_once, load_codecs);
    return g_codecs;
}

bool have(int compressor) {
    const Codecs& c = codecs();
    if (compressor == 4) return c.zstd_bound && c.zstd_compress && c.zstd_decompress && c.zstd_is_error && c.zstd_max_level;
    if (compressor == 1) return c.lz4_bound && c.lz4_compress && c.lz4_decompress;
    return false;
}

void put32(uint8_t* p, uint32_t v) { std::memcpy(p, &v, 4); }  // little-endian hosts only (x86-64, aarch64)
uint32_t get32(const uint8_t* p) {
    uint32_t v;
    std::memcpy(&v, p, 4);
    return v;
}

// byte j of element i -> dst[j * n + i]; trailing bytes copied unchanged
void shuffle(const uint8_t* src, uint8_t* dst, size_t len, int ts) {
    const size_t n = len / ts;
    if (ts == 2) {
        uint8_t *lo = dst, *hi = dst + n;
        for (size_t i = 0; i < n; ++i) {
            lo[i] = src[2 * i];
            hi[i] = src[2 * i + 1];
        }
    } else {
        for (int j = 0; j < ts; ++j) {
            uint8_t* d = dst + (size_t)j * n;
            for (size_t i = 0; i < n; ++i) d[i] = src[i * ts + j];
        }
    }
    std::memcpy(dst + n * ts, src + n * ts, len - n * ts);
}

void unshuffle(const uint8_t* src, uint8_t* dst, size_t len, int ts) {
    const size_t n = len / ts;
    if (ts == 2) {
        const uint8_t *lo = src, *hi = src + n;
        for (size_t i = 0; i < n; ++i) {
            dst[2 * i] = lo[i];
            dst[2 * i + 1] = hi[i];
        }
    } else {
        for (int j = 0; j < ts; ++j) {
            const uint8_t* s = src + (size_t)j * n;
            for (size_t i = 0; i < n; ++i) dst[i * ts + j] = s[i];
        }
    }
    std::memcpy(dst + n * ts, src + n * ts, len - n * ts);
}

size_t auto_blocksize(size_t nbytes, int ts, int clevel) {
    static const int kb[10] = {32, 64, 128, 256, 256, 512, 512, 1024, 1024, 2048};
    size_t bs = (size_t)kb[clevel < 0 ? 0 : (clevel > 9 ? 9 : clevel)] * 1024;
    if (bs > nbytes) bs = nbytes;
    if (bs > (size_t)ts) bs -= bs % ts;
    return bs ? bs : 1;
}

int64_t memcpy_frame(const uint8_t* src, uint64_t nbytes, int ts, uint8_t flags, uint32_t bs, uint8_t* dst, uint64_t cap) {
    if (cap < nbytes + kHeader) return DSTR_E_ARG;
    dst[0] = 2;
    dst[1] = 1;
    dst[2] = (uint8_t)(flags | kMemcpyed);
    dst[3] = (uint8_t)ts;
    put32(dst + 4, (uint32_t)nbytes);
    put32(dst + 8, bs);
    put32(dst + 12, (uint32_t)(nbytes + kHeader));
    std::memcpy(dst + kHeader, src, nbytes);
    return (int64_t)(nbytes + kHeader);
}

}  // namespace

extern "C" {

int dstr_blosc_available(int compressor) { return have(compressor) ? 1 : 0; }

int64_t dstr_blosc_compress(const void* src_, uint64_t nbytes, int typesize, int clevel, int do_shuffle, int compressor,
                            uint64_t blocksize, void* dst_, uint64_t dst_capacity) {
    const uint8_t* src = (const uint8_t*)src_;
    uint8_t* dst = (uint8_t*)dst_;
    if ((!src && nbytes) || !dst || nbytes > 0x7fffffffu - kHeader || dst_capacity < nbytes + kHeader) return DSTR_E_ARG;
    if (!have(compressor)) return DSTR_E_UNSUPPORTED;
    const Codecs& c = codecs();
    const int ts = (typesize >= 1 && typesize <= 255) ? typesize : 1;
    const uint8_t flags = (uint8_t)(((do_shuffle && ts > 1) ? kShuffle : 0) | kDontSplit | (compressor << 5));
    size_t bs = blocksize ? (size_t)blocksize : auto_blocksize(nbytes, ts, clevel);
    if (bs > nbytes) bs = nbytes ? nbytes : 1;
    if (clevel == 0 || nbytes < (uint64_t)kMinBuffer) return memcpy_frame(src, nbytes, ts, flags, (uint32_t)bs, dst, dst_capacity);
    const size_t nblocks = (nbytes + bs - 1) / bs;
    size_t pos = kHeader + 4 * nblocks;
    int level = clevel >= 9 ? c.zstd_max_level() : (clevel == 8 ? c.zstd_max_level() - 2 : (2 * clevel - 1 < 1 ? 1 : 2 * clevel - 1));
    std::vector<uint8_t> tmp((flags & kShuffle) ? bs : 0);
    std::vector<uint8_t> cbuf(compressor == 4 ? c.zstd_bound(bs) : (size_t)c.lz4_bound((int)bs));
    bool overflow = pos >= dst_capacity;
    for (size_t b = 0; b < nblocks && !overflow; ++b) {
        const size_t blen = (b + 1) * bs <= nbytes ? bs : nbytes - b * bs;
        const uint8_t* blk = src + b * bs;
        if (flags & kShuffle) {
            shuffle(blk, tmp.data(), blen, ts);
            blk = tmp.data();
        }
        put32(dst + kHeader + 4 * b, (uint32_t)pos);
        // the coder gets a full-size bound buffer (with less room zstd may give up although the result would fit)
        size_t csize = 0;
        if (compressor == 4) {
            const size_t r = c.zstd_compress(cbuf.data(), cbuf.size(), blk, blen, level);
            csize = c.zstd_is_error(r) ? 0 : r;
        } else {
            const int r = c.lz4_compress((const char*)blk, (char*)cbuf.data(), (int)blen, (int)cbuf.size());
            csize = r > 0 ? (size_t)r : 0;
        }
        const bool verbatim = csize == 0 || csize >= blen;  // stored verbatim: csize == raw length
        if (verbatim) csize = blen;
        if (pos + 4 + csize > dst_capacity) {  // cannot beat a plain copy any more
            overflow = true;
            break;
        }
        std::memcpy(dst + pos + 4, verbatim ? blk : cbuf.data(), csize);
        put32(dst + pos, (uint32_t)csize);
        pos += 4 + csize;
    }
    if (overflow || pos >= nbytes + kHeader) return memcpy_frame(src, nbytes, ts, flags, (uint32_t)bs, dst, dst_capacity);
    dst[0] = 2;
    dst[1] = 1;
    dst[2] = flags;
    dst[3] = (uint8_t)ts;
    put32(dst + 4, (uint32_t)nbytes);
    put32(dst + 8, (uint32_t)bs);
    put32(dst + 12, (uint32_t)pos);
    return (int64_t)pos;
}

int64_t dstr_blosc_decompress(const void* frame_, uint64_t frame_bytes, void* dst_, uint64_t dst_capacity) {
    const uint8_t* f = (const uint8_t*)frame_;
    uint8_t* dst = (uint8_t*)dst_;
    if (!f || frame_bytes < (uint64_t)kHeader) return DSTR_E_ARG;
    const uint8_t flags = f[2];
    const int ts = f[3] ? f[3] : 1;
    const uint64_t nbytes = get32(f + 4), bs = get32(f + 8), cbytes = get32(f + 12);
    if (!dst) return (int64_t)nbytes;  // size query
    if (f[0] != 2 || cbytes > frame_bytes || dst_capacity < nbytes || (flags & kBitShuffle)) return DSTR_E_ARG;
    if (flags & kMemcpyed) {
        if (frame_bytes < nbytes + kHeader) return DSTR_E_ARG;
        std::memcpy(dst, f + kHeader, nbytes);
        return (int64_t)nbytes;
    }
    if (nbytes == 0) return 0;
    const int compressor = flags >> 5;
    if (!have(compressor)) return DSTR_E_UNSUPPORTED;
    if (bs == 0) return DSTR_E_ARG;
    const Codecs& c = codecs();
    const uint64_t nblocks = (nbytes + bs - 1) / bs;
    if (kHeader + 4 * nblocks > cbytes) return DSTR_E_ARG;
    const bool shuf = (flags & kShuffle) && ts > 1;
    std::vector<uint8_t> tmp(shuf ? bs : 0);
    for (uint64_t b = 0; b < nblocks; ++b) {
        const uint64_t blen = (b + 1) * bs <= nbytes ? bs : nbytes - b * bs;
        const bool leftover = (b == nblocks - 1) && blen != bs;
        const bool split = !(flags & kDontSplit) && ts > 1 && !leftover && blen % ts == 0;
        const int nstreams = split ? ts : 1;
        const uint64_t slen = blen / nstreams;
        uint64_t pos = get32(f + kHeader + 4 * b);
        uint8_t* target = shuf ? tmp.data() : dst + b * bs;
        for (int s = 0; s < nstreams; ++s) {
            if (pos + 4 > cbytes) return DSTR_E_ARG;
            const uint64_t csize = get32(f + pos);
            pos += 4;
            if (pos + csize > cbytes) return DSTR_E_ARG;
            uint8_t* out = target + (uint64_t)s * slen;
            if (csize == slen) {
                std::memcpy(out, f + pos, csize);
            } else if (compressor == 4) {
                const size_t r = c.zstd_decompress(out, slen, f + pos, csize);
                if (c.zstd_is_error(r) || r != slen) return DSTR_E_ARG;
            } else {
                const int r = c.lz4_decompress((const char*)f + pos, (char*)out, (int)csize, (int)slen);
                if (r < 0 || (uint64_t)r != slen) return DSTR_E_ARG;
            }
            pos += csize;
        }
        if (shuf) unshuffle(tmp.data(), dst + b * bs, blen, ts);
    }
    return (int64_t)nbytes;
}

// PNG row filters (PNG specification, section 9): `scan` holds h rows of (1 filter byte + stride data bytes) as they
// come out of inflate; the reconstructed rows go to `out` (h * stride bytes).  bpp = bytes per complete pixel.
// Used by the TIFF / PNG front-end (destriper.imread, reference readers.py:64-89 reads PNG through imageio).
int dstr_png_unfilter(const uint8_t* scan, int h, int stride, int bpp, uint8_t* out) {
    if (!scan || !out || h <= 0 || stride <= 0 || bpp <= 0) return DSTR_E_ARG;
    for (int y = 0; y < h; ++y) {
        const uint8_t* in = scan + (size_t)y * (stride + 1);
        const int ft = in[0];
        ++in;
        uint8_t* cur = out + (size_t)y * stride;
        const uint8_t* up = y ? cur - stride : nullptr;
        switch (ft) {
            case 0:
                std::memcpy(cur, in, stride);
                break;
            case 1:
                for (int x = 0; x < stride; ++x) cur[x] = (uint8_t)(in[x] + (x >= bpp ? cur[x - bpp] : 0));
                break;
            case 2:
                for (int x = 0; x < stride; ++x) cur[x] = (uint8_t)(in[x] + (up ? up[x] : 0));
                break;
            case 3:
                for (int x = 0; x < stride; ++x) {
                    const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0;
                    cur[x] = (uint8_t)(in[x] + ((a + b) >> 1));
                }
                break;
            case 4:
                for (int x = 0; x < stride; ++x) {
                    const int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
                    const int p = a + b - c;
                    const int pa = p > a ? p - a : a - p, pb = p > b ? p - b : b - p, pc = p > c ? p - c : c - p;
                    const int pr = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
                    cur[x] = (uint8_t)(in[x] + pr);
                }
                break;
            default:
                return DSTR_E_ARG;
        }
    }
    return 0;
}

}  // extern "C"
