// Strict binary64 arithmetic shared by every result-bearing device function.
//
// SB_HOST_SIM: tests/hostsim compiles the predicate / ray headers with g++
// (-ffp-contract=off) to check their LOGIC against the reference on the CPU
// before GPU time is spent.  That build is test-only; the product never uses it.
#pragma once
#ifdef SB_HOST_SIM
#include <cmath>
#include <cstdint>
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return std::sqrt(a); }
#else
#include <cuda_runtime.h>
#include <stdint.h>
#endif

// ---------------------------------------------------------------------------
// Strict IEEE binary64 arithmetic.  The reference is compiled without FMA
// contraction (one rounded op per C operator); nvcc would contract a*b+c into
// DFMA, so every result-bearing operation goes through these explicitly rounded
// intrinsics (and the translation units are built with -fmad=false as a second
// line of defence).
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double xsqrt(double a) { return __dsqrt_rn(a); }

struct d3 {
    double x, y, z;
};

__device__ __forceinline__ d3 d3sub(const d3 &a, const d3 &b)
{
    return {xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)};
}
// (a.x*b.x + a.y*b.y) + a.z*b.z -- the evaluation order of the reference's DOT
// macro (tri_tri_intersect.c:78) and of Vector3::dotProduct (vector3.h:150-153).
__device__ __forceinline__ double d3dot(const d3 &a, const d3 &b)
{
    return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z));
}
// two rounded products and one subtraction per component (tri_tri_intersect.c:73-76)
__device__ __forceinline__ d3 d3cross(const d3 &a, const d3 &b)
{
    return {xsub(xmul(a.y, b.z), xmul(a.z, b.y)),
            xsub(xmul(a.z, b.x), xmul(a.x, b.z)),
            xsub(xmul(a.x, b.y), xmul(a.y, b.x))};
}

