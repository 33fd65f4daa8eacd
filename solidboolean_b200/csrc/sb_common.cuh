// Shared device/host helpers for the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <string.h>
#include "sb_fp64.cuh"

#define SB_WARP 32
#define SB_FULL 0xffffffffu

// ---------------------------------------------------------------------------
// 32-byte record shared by LBVH children, cluster boxes and sorted-triangle
// leaves: conservative float box + a reference.  Exactly one 32-B sector.
struct __align__(32) Rec32 {
    float lox, loy, loz, hix;
    float hiy, hiz;
    int ref; // child: >=0 internal node, <0 ~cluster; leaf: original triangle id
    int aux;
};

struct BoxF {
    float lox, loy, loz, hix, hiy, hiz;
};

struct BoxD {
    double lox, loy, loz, hix, hiy, hiz;
};

// one 256-bit load per record (sm_100: LDG.E.256)
__device__ __forceinline__ Rec32 load_rec(const Rec32 *p)
{
    Rec32 r;
    float fr, fa;
    asm("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=f"(r.lox), "=f"(r.loy), "=f"(r.loz), "=f"(r.hix), "=f"(r.hiy), "=f"(r.hiz), "=f"(fr), "=f"(fa)
        : "l"(p));
    r.ref = __float_as_int(fr);
    r.aux = __float_as_int(fa);
    return r;
}

// coherent (L2) load, for data written earlier in the same kernel by another thread
__device__ __forceinline__ Rec32 load_rec_cg(const Rec32 *p)
{
    Rec32 r;
    float fr, fa;
    asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(r.lox), "=f"(r.loy), "=f"(r.loz), "=f"(r.hix), "=f"(r.hiy), "=f"(r.hiz), "=f"(fr), "=f"(fa)
                 : "l"(p)
                 : "memory");
    r.ref = __float_as_int(fr);
    r.aux = __float_as_int(fa);
    return r;
}

#ifndef SB_REC256
#define SB_REC256 0 // 256-bit record stores measured slightly slower than two 128-bit ones (leaf kernel +5 us)
#endif
__device__ __forceinline__ void store_rec(Rec32 *p, const BoxF &b, int ref, int aux)
{
#if !SB_REC256
    float4 *q = reinterpret_cast<float4 *>(p);
    q[0] = make_float4(b.lox, b.loy, b.loz, b.hix);
    q[1] = make_float4(b.hiy, b.hiz, __int_as_float(ref), __int_as_float(aux));
    return;
#endif
    // one 256-bit store per record (sm_100): a warp writing consecutive records fills whole 128-byte lines
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(b.lox), "f"(b.loy), "f"(b.loz), "f"(b.hix),
                 "f"(b.hiy), "f"(b.hiz), "f"(__int_as_float(ref)), "f"(__int_as_float(aux))
                 : "memory");
}

// closed-interval overlap, written with the same comparisons as
// AxisAlignedBoudingBox::intersectWith (axisalignedboundingbox.h:95-105) so NaN
// behaves identically (any NaN -> no overlap).
__device__ __forceinline__ bool overlap_f(const BoxF &a, float lox, float loy, float loz, float hix, float hiy, float hiz)
{
    return a.lox <= hix && a.hix >= lox && a.loy <= hiy && a.hiy >= loy && a.loz <= hiz && a.hiz >= loz;
}

__device__ __forceinline__ bool overlap_d(const BoxD &a, const BoxD &b)
{
    return a.lox <= b.hix && a.hix >= b.lox && a.loy <= b.hiy && a.hiy >= b.loy && a.loz <= b.hiz && a.hiz >= b.loz;
}

// double box stored as 3 x double2: {lox,loy} {loz,hix} {hiy,hiz}
__device__ __forceinline__ BoxD load_boxd(const double2 *p)
{
    double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
    return {a.x, a.y, b.x, b.y, c.x, c.y};
}

__device__ __forceinline__ void store_boxd(double2 *p, const BoxD &b)
{
    p[0] = make_double2(b.lox, b.loy);
    p[1] = make_double2(b.loz, b.hix);
    p[2] = make_double2(b.hiy, b.hiz);
}

// conservative float enclosure of a double box
__device__ __forceinline__ BoxF enclose(const BoxD &b)
{
    return {__double2float_rd(b.lox), __double2float_rd(b.loy), __double2float_rd(b.loz),
            __double2float_ru(b.hix), __double2float_ru(b.hiy), __double2float_ru(b.hiz)};
}

__device__ __forceinline__ BoxF empty_boxf()
{
    return {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
}

__device__ __forceinline__ void merge_f(BoxF &a, const BoxF &b)
{
    a.lox = fminf(a.lox, b.lox); a.loy = fminf(a.loy, b.loy); a.loz = fminf(a.loz, b.loz);
    a.hix = fmaxf(a.hix, b.hix); a.hiy = fmaxf(a.hiy, b.hiy); a.hiz = fmaxf(a.hiz, b.hiz);
}

__device__ __forceinline__ BoxF shfl_xor_box(const BoxF &b, int mask)
{
    return {__shfl_xor_sync(SB_FULL, b.lox, mask), __shfl_xor_sync(SB_FULL, b.loy, mask),
            __shfl_xor_sync(SB_FULL, b.loz, mask), __shfl_xor_sync(SB_FULL, b.hix, mask),
            __shfl_xor_sync(SB_FULL, b.hiy, mask), __shfl_xor_sync(SB_FULL, b.hiz, mask)};
}

// vertices are stored padded to double4 (one 32-B sector each, two 128-bit loads)
// One 256-bit read-only load (sm_100: LDG.E.256) of a 32-byte aligned record: one instruction, one
// L1 tag lookup per lane where two 128-bit loads cost two -- gathers are bound by those lookups.
__device__ __forceinline__ double4 ldg256(const double4 *p)
{
    double4 r;
    asm("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256(double4 *p, double a, double b, double c, double d)
{
    asm volatile("st.global.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__device__ __forceinline__ d3 load_vertex(const double4 *v, uint32_t i)
{
    const double4 a = ldg256(v + i);
    return {a.x, a.y, a.z};
}

// Per-triangle record of the classifier (sb_build.cu writes it, sb_classify*.cu read it): the unit normal
// (SolidMesh::prepare, src/solidmesh.cpp:47-55) and, in the fourth word, the triangle's three vertex indices
// -- ONE 256-bit gather instead of three index loads followed by three normal loads, and one dependent round
// trip less.  Layout of the word: i0 in 22 bits, i1 - i0 and i2 - i0 as signed 21-bit differences (the corners
// of a triangle are rarely more than a million vertices apart; a batch mesh keeps every job's vertices together).
// A triangle that does not fit (i0 >= 2^22 - 1 or a difference outside +-2^20) gets the all-ones word and the
// kernels read its index triple from `tri`.
#define SB_PACKED_IDX_NONE 0xffffffffffffffffull
__device__ __forceinline__ unsigned long long pack_tri_idx(uint32_t i0, uint32_t i1, uint32_t i2)
{
    const long long d1 = (long long)i1 - (long long)i0, d2 = (long long)i2 - (long long)i0;
    const long long lim = 1ll << 20;
    if (i0 >= (1u << 22) - 1u || d1 < -lim || d1 >= lim || d2 < -lim || d2 >= lim)
        return SB_PACKED_IDX_NONE;
    return (unsigned long long)i0 | ((unsigned long long)(d1 & 0x1fffff) << 22) | ((unsigned long long)(d2 & 0x1fffff) << 43);
}
__device__ __forceinline__ void unpack_tri_idx(unsigned long long w, uint32_t &i0, uint32_t &i1, uint32_t &i2)
{
    i0 = (uint32_t)w & 0x3fffffu;
    const int d1 = ((int)((uint32_t)(w >> 22) << 11)) >> 11; // sign-extend 21 bits
    const int d2 = ((int)((uint32_t)(w >> 43) << 11)) >> 11;
    i1 = (uint32_t)((int)i0 + d1);
    i2 = (uint32_t)((int)i0 + d2);
}

__device__ __forceinline__ uint32_t lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// order-preserving double <-> uint64 map (for atomicMin/Max on doubles)
__device__ __forceinline__ unsigned long long dkey(double x)
{
    unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ __forceinline__ double dkey_inv(unsigned long long k)
{
    unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double d;
    memcpy(&d, &b, sizeof(d));
    return d;
#endif
}
