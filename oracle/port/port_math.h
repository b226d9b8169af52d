// TEST INFRASTRUCTURE — CPU restatement ("port") of the reference's device library.
// Each function cites the reference lines it restates.  Compiled with
// -ffp-contract=off on x86-64 (no FMA): every operation rounds once, in the
// order written, which is the arithmetic the CUDA path reproduces (-fmad=false).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace port
{
struct V2 { float x, y; };
struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };

// component-wise operators, common/math.hpp:27-107 (mutating compound forms: SURVEY.md section 8a)
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 operator*(V3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator*(float s, V3 a) { return {a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }

constexpr float kPi = 3.14159265358979323846f;  // math.hpp:23
constexpr float kFltMax = 3.402823466e+38f;     // math.hpp:25

// math.hpp:109-123
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(V3 a) { return sqrtf(dot(a, a)); }
inline V3 normalize(V3 a) { return a / length(a); }
inline V3 mix(V3 a, V3 b, float t) { return a + (b - a) * t; }
// math.hpp:125-130
inline float luminance(V3 a) { return dot(a, V3{0.1762044f, 0.8129847f, 0.0108109f}); }

// ---- transcendental hooks: mode 0 = glibc float functions (bit-identical to the reference
// built as host C++), mode 1 = correctly rounded via double (what the CUDA "exact" mode computes).
extern int g_math_mode;
inline float m_log(float x) { return g_math_mode ? (float)log((double)x) : logf(x); }
inline float m_exp(float x) { return g_math_mode ? (float)exp((double)x) : expf(x); }
inline float m_pow(float x, float y) { return g_math_mode ? (float)pow((double)x, (double)y) : powf(x, y); }
inline float m_sin(float x) { return g_math_mode ? (float)sin((double)x) : sinf(x); }
inline float m_cos(float x) { return g_math_mode ? (float)cos((double)x) : cosf(x); }

// ---- common/rng.hpp
struct Pcg  // rng.hpp:8-40
{
    uint64_t state, inc;
    Pcg(uint64_t seed, uint64_t sequence)
    {
        state = 0u;
        inc = (sequence << 1u) | 1u;
        next_u32();
        state += seed;
        next_u32();
    }
    uint32_t next_u32()
    {
        const uint64_t old = state;
        state = old * 6364136223846793005ULL + inc;
        const uint32_t xorshifted = (uint32_t)(((old >> 18u) ^ old) >> 27u);
        const uint32_t rot = (uint32_t)(old >> 59u);
        return (xorshifted >> rot) | (xorshifted << ((0u - rot) & 31u));
    }
    float next_f()  // rng.hpp:29-35: 23 mantissa bits in [1,2) minus 1
    {
        const uint32_t bits = (next_u32() >> 9) | 0x3f800000u;
        float v;
        memcpy(&v, &bits, 4);
        return v - 1.0f;
    }
};
inline uint32_t hash_pcg(uint32_t v)  // rng.hpp:43-48
{
    const uint32_t state = v * 747796405u + 2891336453u;
    const uint32_t word = ((state >> ((state >> 28) + 4)) ^ state) * 277803737u;
    return (word >> 22) ^ word;
}
inline uint32_t hash_pcg3(uint32_t x, uint32_t y, uint32_t z) { return hash_pcg(hash_pcg(hash_pcg(x) + y) + z); }
inline uint32_t hash_pcg4(uint32_t x, uint32_t y, uint32_t z, uint32_t w)
{
    return hash_pcg(hash_pcg(hash_pcg(hash_pcg(x) + y) + z) + w);
}
}  // namespace port
