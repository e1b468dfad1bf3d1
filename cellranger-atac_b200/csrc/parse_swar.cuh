// Word-at-a-time field parsers for the fragment decoder (K1).  Host + device so that the bit tricks can be
// unit-tested with a plain C++ compiler (tests/test_swar.py).  All loads are 32-bit aligned words taken
// from the staged text tile; `get4(base, p)` returns the 4 bytes starting at byte offset p (little endian).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SWAR_HD __host__ __device__ __forceinline__
#else
#define SWAR_HD inline
#endif

// 4 text bytes at an arbitrary offset from two aligned word loads
SWAR_HD uint32_t swar_get4(const uint32_t* words, int p) {
    const uint32_t w0 = words[p >> 2], w1 = words[(p >> 2) + 1];
    const int sh = (p & 3) * 8;
#ifdef __CUDA_ARCH__
    return __funnelshift_r(w0, w1, sh);
#else
    return sh ? (w0 >> sh) | (w1 << (32 - sh)) : w0;
#endif
}

// parse the decimal field [p, p + len), 1 <= len <= 10; false on a non-digit or a value above INT32_MAX
SWAR_HD bool swar_parse_uint(const uint32_t* words, int p, int len, int32_t* out) {
    if (len < 1 || len > 10) return false;
    if (len == 1) {                      // the duplicate count is a single digit nearly always
        const uint32_t d = (swar_get4(words, p) & 0xFFu) - 0x30u;
        *out = (int32_t)d;
        return d <= 9u;
    }
    uint64_t head = 0;
    if (len > 8) {                       // leading 1-2 digits one by one, the last 8 digits word-wise
        const uint32_t w = swar_get4(words, p);
        for (int i = 0; i < len - 8; i++) {
            const uint32_t d = ((w >> (8 * i)) & 0xFFu) - 0x30u;
            if (d > 9u) return false;
            head = head * 10 + d;
        }
        p += len - 8;
        len = 8;
    }
    uint64_t v = (uint64_t)swar_get4(words, p) | ((uint64_t)swar_get4(words, p + 4) << 32);
    const int s = 8 * (8 - len);         // right-align: the last digit goes to the top byte, '0' fills the bottom
    if (s) v = (v << s) | (0x3030303030303030ULL >> (64 - s));
    const uint64_t d = v - 0x3030303030303030ULL;
    // every byte of d must be 0..9: no borrow across bytes can hide a bad byte because a byte < '0' sets bit 7
    if (((d | (d + 0x7676767676767676ULL)) & 0x8080808080808080ULL) != 0) return false;
    uint64_t x = d;
    x = ((x & 0x0F0F0F0F0F0F0F0FULL) * 2561ULL) >> 8;
    x = ((x & 0x00FF00FF00FF00FFULL) * 6553601ULL) >> 16;
    x = ((x & 0x0000FFFF0000FFFFULL) * 42949672960001ULL) >> 32;
    const uint64_t val = head * 100000000ULL + x;
    if (val > 2147483647ULL) return false;
    *out = (int32_t)val;
    return true;
}

// 4 barcode characters -> 8 bits (2 bits per base, first base in the top bits; code = (ch >> 1) & 3 so that
// A=0 C=1 T=2 G=3); *ok is cleared when a character is not one of ACGT
SWAR_HD uint32_t swar_pack4(uint32_t w, bool* ok) {
    const uint32_t c = (w >> 1) & 0x03030303u;
    const uint32_t lo = c & 0x01010101u, hi = (c >> 1) & 0x01010101u;
    const uint32_t both = lo & hi, only_hi = hi & ~lo;
    const uint32_t expect = 0x41414141u + (lo << 1) + (both << 2) + only_hi * 0x13u;   // A, C(+2), G(+6), T(+0x13)
    if (expect != w) *ok = false;
    return (c * 0x40100401u) >> 24;
}

// FNV-1a over [p, p + len)
SWAR_HD uint64_t swar_fnv1a(const uint32_t* words, int p, int len) {
    uint64_t h = 1469598103934665603ULL;
    for (int i = 0; i < len; i += 4) {
        uint32_t w = swar_get4(words, p + i);
        const int n = len - i < 4 ? len - i : 4;
        for (int k = 0; k < n; k++) { h ^= (w >> (8 * k)) & 0xFFu; h *= 1099511628211ULL; }
    }
    return h;
}

// Key of a contig name: the bytes themselves (little endian, zero padded) for names of up to 8 bytes, FNV-1a
// with the top bit set for longer ones; and the slot hash of the contig table.
SWAR_HD uint64_t swar_contig_key(const uint32_t* words, int p, int len) {
    if (len <= 8) {
        uint32_t lo = swar_get4(words, p), hi = len > 4 ? swar_get4(words, p + 4) : 0u;
        if (len < 4) lo &= (1u << (8 * len)) - 1u;
        else if (len < 8) hi &= (1u << (8 * (len - 4))) - 1u;
        return (uint64_t)lo | ((uint64_t)hi << 32);
    }
    return swar_fnv1a(words, p, len) | 0x8000000000000000ULL;
}
SWAR_HD uint32_t swar_key_slot(uint64_t key) {
    const uint32_t h = (uint32_t)key * 0x9E3779B1u ^ (uint32_t)(key >> 32) * 0x85EBCA77u;
    return h ^ (h >> 15);
}

// bit i of the result = (byte i of the 16 bytes v0..v3 equals the replicated pattern byte)
SWAR_HD uint32_t swar_eqmask16(uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3, uint32_t pat4) {
#ifdef __CUDA_ARCH__
    // SIMD byte compare: 0xFF per equal byte, one bit kept per byte and gathered by a multiply
    return ((((__vcmpeq4(v0, pat4) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) |
           (((((__vcmpeq4(v1, pat4) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << 4) |
           (((((__vcmpeq4(v2, pat4) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << 8) |
           (((((__vcmpeq4(v3, pat4) & 0x01010101u) * 0x01020408u) >> 24) & 0xFu) << 12);
#endif
    uint32_t m[4] = {v0 ^ pat4, v1 ^ pat4, v2 ^ pat4, v3 ^ pat4};
    uint32_t out = 0;
    for (int k = 0; k < 4; k++) {
        // zero-byte detector: bit 7 of each byte set iff the byte is zero
        uint32_t z = ~(((m[k] & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | m[k] | 0x7F7F7F7Fu);
        out |= (((z >> 7) & 0x01010101u) * 0x01020408u >> 24 & 0xFu) << (4 * k);
    }
    return out;
}
