// Host side of the packed upload path (see upload.cu): 1-byte-per-base sequence bytes are packed
// to 2 bits per base before they cross PCIe.  Bytes >= 4 (invalid for DNA, and anything a caller
// may have stored) are recorded as (position, value) exceptions so the device reconstructs the
// caller's bytes exactly.  This is transfer encoding only — no result is computed on the host.
//
// Packed layout: output byte q holds input bytes 4q..4q+3 as b0 | b1<<2 | b2<<4 | b3<<6.
#include <immintrin.h>
#include <stddef.h>
#include <stdint.h>

#include <atomic>

namespace dvs {

struct PackExc {
    uint32_t* pos;                 // block-relative positions
    uint8_t* val;                  // original byte values
    std::atomic<uint32_t>* count;  // entries reserved so far (may exceed cap: overflow)
    uint32_t cap;
};

static inline void pack_scalar(const uint8_t* src, size_t n, uint8_t* dst, uint32_t rel0, PackExc& ex) {
    // n is a multiple of 4 except possibly at the very end of the stream
    size_t q = 0;
    for (; q + 4 <= n; q += 4) {
        uint32_t out = 0;
        for (int j = 0; j < 4; ++j) {
            uint8_t b = src[q + j];
            if (b > 3) {
                uint32_t slot = ex.count->fetch_add(1, std::memory_order_relaxed);
                if (slot < ex.cap) {
                    ex.pos[slot] = rel0 + (uint32_t)(q + j);
                    ex.val[slot] = b;
                }
                b = 0;
            }
            out |= (uint32_t)b << (2 * j);
        }
        dst[q >> 2] = (uint8_t)out;
    }
    if (q < n) {
        uint32_t out = 0;
        for (size_t j = 0; q + j < n; ++j) {
            uint8_t b = src[q + j];
            if (b > 3) {
                uint32_t slot = ex.count->fetch_add(1, std::memory_order_relaxed);
                if (slot < ex.cap) {
                    ex.pos[slot] = rel0 + (uint32_t)(q + j);
                    ex.val[slot] = b;
                }
                b = 0;
            }
            out |= (uint32_t)b << (2 * j);
        }
        dst[q >> 2] = (uint8_t)out;
    }
}

__attribute__((target("avx2"))) static void pack_avx2(const uint8_t* src, size_t n, uint8_t* dst, uint32_t rel0,
                                                      PackExc& ex) {
    const __m256i hi6 = _mm256_set1_epi8((char)0xFC);
    const __m256i w1 = _mm256_set1_epi16(0x0401);      // byte pairs: b_even*1 + b_odd*4
    const __m256i w2 = _mm256_set1_epi32(0x00100001);  // 16-bit pairs: t_even*1 + t_odd*16
    const __m256i order = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    const bool aligned = ((uintptr_t)dst & 15) == 0;
    size_t q = 0;
    for (; q + 64 <= n; q += 64) {
        _mm_prefetch((const char*)(src + q + 1024), _MM_HINT_T0);
        const __m256i v0 = _mm256_loadu_si256((const __m256i*)(src + q));
        const __m256i v1 = _mm256_loadu_si256((const __m256i*)(src + q + 32));
        if (!_mm256_testz_si256(_mm256_or_si256(v0, v1), hi6)) {  // some byte > 3: exceptions
            pack_scalar(src + q, 64, dst + (q >> 2), rel0 + (uint32_t)q, ex);
            continue;
        }
        const __m256i u0 = _mm256_madd_epi16(_mm256_maddubs_epi16(v0, w1), w2);  // 8 dwords, one packed byte each
        const __m256i u1 = _mm256_madd_epi16(_mm256_maddubs_epi16(v1, w1), w2);
        const __m256i p16 = _mm256_packus_epi32(u0, u1);   // lane0: u0 d0-3, u1 d0-3 | lane1: u0 d4-7, u1 d4-7
        const __m256i p8 = _mm256_packus_epi16(p16, p16);  // low 8 bytes of each lane carry the data
        const __m256i r = _mm256_permutevar8x32_epi32(p8, order);  // dwords: u0 d0-3, u0 d4-7, u1 d0-3, u1 d4-7
        // non-temporal store: the packed block is only read again by the DMA engine, and a regular
        // store would first read the destination line (write-allocate) on a DRAM-bandwidth-bound loop
        if (aligned)
            _mm_stream_si128((__m128i*)(dst + (q >> 2)), _mm256_castsi256_si128(r));
        else
            _mm_storeu_si128((__m128i*)(dst + (q >> 2)), _mm256_castsi256_si128(r));
    }
    _mm_sfence();
    if (q < n) pack_scalar(src + q, n - q, dst + (q >> 2), rel0 + (uint32_t)q, ex);
}

// packs src[0..n) (n bytes, block-relative offset rel0, rel0 % 4 == 0) into dst[rel0/4 ...]
void pack_bytes(const uint8_t* src, size_t n, uint8_t* dst_block, uint32_t rel0, PackExc& ex) {
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
    uint8_t* dst = dst_block + (rel0 >> 2);
    if (have_avx2)
        pack_avx2(src, n, dst, rel0, ex);
    else
        pack_scalar(src, n, dst, rel0, ex);
}

}  // namespace dvs
