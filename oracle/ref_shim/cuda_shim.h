// TEST INFRASTRUCTURE (oracle/_ref): force-included before the reference's
// *unmodified* examples/NN/*.cu so that g++ can compile them as host C++.
// Nothing here is reference code; it only supplies the CUDA built-ins the
// kernels expect.  No std header may be included
// here: common/types.hpp:13-16 typedefs uint64_t itself under __CUDACC__.
#pragma once
#define __device__
#define __host__
#define __global__
#define __shared__ static thread_local

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct shim_uint3 { unsigned x, y, z; };
extern thread_local shim_uint3 threadIdx, blockIdx, blockDim;

typedef unsigned long size_t;
extern "C" {
float sqrtf(float) noexcept; float cosf(float) noexcept; float sinf(float) noexcept;
float tanf(float) noexcept;  float powf(float, float) noexcept; float expf(float) noexcept;
float logf(float) noexcept;  float fabsf(float) noexcept;
void* memcpy(void*, const void*, size_t) noexcept;
}
// the transcendental calls go through these hooks so that the driver can swap
// glibc's float functions for correctly-rounded ones (same switch as the port)
extern "C" float shim_logf(float); extern "C" float shim_expf(float);
extern "C" float shim_powf(float, float); extern "C" float shim_sinf(float);
extern "C" float shim_cosf(float);
#define logf shim_logf
#define expf shim_expf
#define powf shim_powf
inline float sqrt(float x) { return sqrtf(x); }
inline float cos(float x) { return shim_cosf(x); }
inline float sin(float x) { return shim_sinf(x); }
inline float tan(float x) { return tanf(x); }
inline float log(float x) { return shim_logf(x); }
inline float abs(float x) { return fabsf(x); }
inline float min(float a, float b) { return a < b ? a : b; }
inline float max(float a, float b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }

// 06-10 are compiled with -DNO_VECTOR_OP_OVERLOAD (10_restir_di.cpp:61-63), so
// the vector operators come from outside common/math.hpp; these are the
// component-wise, *mutating* forms the kernels rely on.
inline float3 operator+(const float3& a, const float3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline float3 operator-(const float3& a, const float3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float3 operator*(const float3& a, const float3& b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline float3 operator/(const float3& a, const float3& b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline float3 operator+(const float3& a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline float3 operator+(float s, const float3& a) { return {a.x + s, a.y + s, a.z + s}; }
inline float3 operator*(const float3& a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator*(float s, const float3& a) { return {a.x * s, a.y * s, a.z * s}; }
inline float3 operator/(const float3& a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float3 operator-(const float3& a) { return {-a.x, -a.y, -a.z}; }
inline float3& operator+=(float3& a, const float3& b) { a = a + b; return a; }
inline float3& operator-=(float3& a, const float3& b) { a = a - b; return a; }
inline float3& operator*=(float3& a, const float3& b) { a = a * b; return a; }
inline float3& operator*=(float3& a, float s) { a = a * s; return a; }
inline float3& operator/=(float3& a, float s) { a = a / s; return a; }
inline float2 operator*(float s, const float2& a) { return {s * a.x, s * a.y}; }
inline float4& operator+=(float4& a, const float4& b)
{
    a = {a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w};
    return a;
}
