// cuda_host_shim.h — TEST INFRASTRUCTURE. Lets g++ compile the device-side cell logic (glnexus_b200/csrc/*.cuh)
// as ordinary host code so that its control flow can be diffed against the oracle on a machine without a GPU.
// Nothing in the product links this; the product path is the CUDA build of the same headers.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)

struct int4 { int x, y, z, w; };
struct int2 { int x, y; };
struct uint2 { unsigned x, y; };
inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
inline int2 make_int2(int x, int y) { return int2{x, y}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct char2 { signed char x, y; };
struct uchar2 { unsigned char x, y; };
inline char2 make_char2(signed char x, signed char y) { return char2{x, y}; }
inline uchar2 make_uchar2(unsigned char x, unsigned char y) { return uchar2{x, y}; }

using std::max;
using std::min;

inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) { auto o = *p; if (v < o) *p = v; return o; }
inline uint32_t atomicOr(uint32_t* p, uint32_t v) { auto o = *p; *p |= v; return o; }
inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { auto o = *p; *p += v; return o; }
inline double __longlong_as_double(long long x) { double d; memcpy(&d, &x, 8); return d; }
inline float __int_as_float(int32_t x) { float f; memcpy(&f, &x, 4); return f; }
inline int32_t __float_as_int(float f) { int32_t x; memcpy(&x, &f, 4); return x; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline int __popc(uint32_t x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T>
inline T __ldg(const T* p) { return *p; }
