// TEST INFRASTRUCTURE: stand-in for <hiprt/hiprt_device.h>.  HIPRT 2.4.6b6daf9
// is a closed binary (libs/hiprt/hiprt/linux64/libhiprt0200464.so); its
// traversal is replaced here by cpu_closest_hit(), a conservative CPU BVH whose
// leaf test is the reference's own intersect_ray_triangle (common/core.hpp:91).
#pragma once
typedef void* hiprtGeometry;
struct hiprtRay
{
    float3 origin;
    float minT = 0.0f;
    float3 direction;
    float maxT = 3.402823466e+38f;
};
struct hiprtHit
{
    unsigned primID = ~0u;
    float2 uv;
    float3 normal;
    float t = -1.0f;
    bool hasHit() const { return primID != ~0u; }
};
struct hiprtGlobalStackBuffer {};
struct hiprtSharedStackBuffer { unsigned n; void* p; };
struct hiprtGlobalStack
{
    hiprtGlobalStack(hiprtGlobalStackBuffer, hiprtSharedStackBuffer) {}
};
hiprtHit cpu_closest_hit(hiprtGeometry geom, const hiprtRay& ray);
template <class Stack>
struct hiprtGeomTraversalClosestCustomStack
{
    hiprtGeometry g;
    hiprtRay r;
    hiprtGeomTraversalClosestCustomStack(hiprtGeometry geom, const hiprtRay& ray, Stack&) : g(geom), r(ray) {}
    hiprtHit getNextHit() { return cpu_closest_hit(g, r); }
};
