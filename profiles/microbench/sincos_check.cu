// MEASUREMENT TOOL: is sincosf(phi) bit-identical to (sinf(phi), cosf(phi)) for every argument sample_2d_gaussian can
// produce (reservoir.hpp:89-95: phi = 2 pi rv1, rv1 = k 2^-23, k < 2^23)?  If so, one range reduction serves both.
//   nvcc -gencode arch=compute_100a,code=sm_100a [-fmad=false] -o sincos_check sincos_check.cu && ./sincos_check
#include <cstdio>
#include <cstdint>
__global__ void k(unsigned long long* bad_sin, unsigned long long* bad_cos, unsigned* first)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << 23)) return;
    const float rv1 = __uint_as_float((i << 0) | 0x3f800000u) - 1.0f;  // k 2^-23 exactly, as Pcg::next_f makes it
    const float phi = 2.0f * 3.14159265358979323846f * rv1;
    float s, c;
    sincosf(phi, &s, &c);
    const float s2 = sinf(phi), c2 = cosf(phi);
    if (__float_as_uint(s) != __float_as_uint(s2)) { atomicAdd(bad_sin, 1ull); atomicMin(first, i); }
    if (__float_as_uint(c) != __float_as_uint(c2)) { atomicAdd(bad_cos, 1ull); atomicMin(first, i); }
}
int main()
{
    unsigned long long *d, h[2];
    unsigned *f, hf = 0xffffffffu;
    cudaMalloc(&d, 16); cudaMemset(d, 0, 16);
    cudaMalloc(&f, 4); cudaMemcpy(f, &hf, 4, cudaMemcpyHostToDevice);
    k<<<(1u << 23) / 256, 256>>>(d, d + 1, f);
    cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    cudaMemcpy(&hf, f, 4, cudaMemcpyDeviceToHost);
    printf("sincosf vs sinf/cosf over 2^23 arguments: %llu sin mismatches, %llu cos mismatches, first at k = %u\n", h[0], h[1], hf);
    return 0;
}
