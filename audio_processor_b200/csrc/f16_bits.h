// f16_bits.h — IEEE binary16 <-> binary32 on raw bits, round-to-nearest-even (host side: filter-bank tables for the
// tensor-core FIR; also used by the TEST-ONLY emulation).  Plain C++, no CUDA types.
#pragma once
#include <cstdint>
#include <cstring>

namespace b2a_f16 {

static inline float f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {                                  // subnormal: renormalise
            int e = -1;
            do { man <<= 1; e++; } while (!(man & 0x400u));
            bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | (man << 13);
    else bits = sign | ((exp + 112u) << 23) | (man << 13);
    float f;
    memcpy(&f, &bits, 4);
    return f;
}

static inline uint16_t f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    const uint32_t ax = x & 0x7fffffffu;
    if (ax >= 0x7f800000u) return (uint16_t)(sign | 0x7c00u | ((ax > 0x7f800000u) ? 0x200u : 0u));   // inf / nan
    if (ax >= 0x477ff000u) return (uint16_t)(sign | 0x7c00u);                                         // rounds to inf
    if (ax < 0x33000001u) return sign;                                                                // < half the smallest subnormal
    int e = (int)(ax >> 23) - 127;
    uint32_t m = (ax & 0x7fffffu) | 0x800000u;          // 24-bit significand
    int shift = (e < -14) ? (13 + (-14 - e)) : 13;      // bits to drop
    uint32_t q = m >> shift, rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (q & 1u))) q++;
    uint32_t h = (e < -14) ? q : (((uint32_t)(e + 15) << 10) + (q - 0x400u));   // carry out of the significand bumps the exponent
    return (uint16_t)(sign | h);
}

}  // namespace b2a_f16
