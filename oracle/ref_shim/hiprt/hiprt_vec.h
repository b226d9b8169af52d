// TEST INFRASTRUCTURE: stand-in for <hiprt/hiprt_vec.h> (see cuda_shim.h).
#pragma once
#define hiprtFloat2 float2
#define hiprtFloat3 float3
