// Stage 1: everything SolidMesh::prepare() computes (reference src/solidmesh.cpp:
// 42-76), rebuilt for the device:
//   K2  bounds_pad    whole-mesh box (:64-72) + 32-byte padded vertex copy
//   K1  tri_prepare   unit normals (:47-55), Morton key of the centre of the exact
//                     per-triangle box (:57-62)
//   K3  onesweep radix sort of (Morton key, triangle id)
//   K1b leaf_gather   exact boxes (:57-62) and face centroids in Morton order, straight
//                     from the vertices; sorted leaf records + warp-shuffle reduction of every K
//                     consecutive leaves into a cluster box (the bottom log2 K
//                     levels of the bottom-up refit, done in registers); also the
//                     quantised boxes and per-cell counts of the ray grids (sb_grid.cu)
//   K4  tree_build    agglomerative bottom-up LBVH over the clusters: topology
//                     and boxes in ONE pass with an atomic rendezvous per node
// The reference's top-down mean-split tree (axisalignedboundingboxtree.cpp:27-141)
// is not reproduced: the candidate-pair set only depends on the leaf boxes.
#include "sb_internal.h"
#include <algorithm>
#include "sb_radix.cuh"
#include "sb_gridq.cuh"

namespace {

#ifndef SB_LEAF_STAGE
#define SB_LEAF_STAGE 0 // boxes / centroids staged through shared memory for 256-bit stores: measured SLOWER (95 -> 110 us at 1.3M triangles)
#endif
constexpr int K = SB_CLUSTER;
static_assert(K == 1 || K == 2 || K == 4 || K == 8 || K == 16 || K == 32, "cluster size must divide the warp");

// ---- K2 ---------------------------------------------------------------------
__global__ void __launch_bounds__(256) bounds_pad_kernel(const double *__restrict__ xyz, uint32_t nV,
    double4 *__restrict__ vtx, unsigned long long *__restrict__ bounds, float2 *__restrict__ zf)
{
    double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nV; i += gridDim.x * blockDim.x) {
        double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        double2 *o = reinterpret_cast<double2 *>(vtx + i);
        o[0] = make_double2(x, y);
        o[1] = make_double2(z, 0.0);
        if (zf) // multi-GPU planning (sb_shard.cu): z as a float bracket, a table dense enough to stay in cache
            zf[i] = make_float2(__double2float_rd(z), __double2float_ru(z));
        lo[0] = fmin(lo[0], x); hi[0] = fmax(hi[0], x);
        lo[1] = fmin(lo[1], y); hi[1] = fmax(hi[1], y);
        lo[2] = fmin(lo[2], z); hi[2] = fmax(hi[2], z);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(SB_FULL, lo[k], off));
            hi[k] = fmax(hi[k], __shfl_xor_sync(SB_FULL, hi[k], off));
        }
    __shared__ double s_lo[8][3], s_hi[8][3];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0)
        for (int k = 0; k < 3; ++k) {
            s_lo[warp][k] = lo[k];
            s_hi[warp][k] = hi[k];
        }
    __syncthreads();
    if (threadIdx.x < 3) {
        int k = threadIdx.x;
        double l = s_lo[0][k], h = s_hi[0][k];
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
            l = fmin(l, s_lo[w][k]);
            h = fmax(h, s_hi[w][k]);
        }
        atomicMin(&bounds[k], dkey(l));
        atomicMax(&bounds[3 + k], dkey(h));
    }
}

// ---- K1 ---------------------------------------------------------------------
// AxisAlignedBoudingBox::update x3 from the +-DBL_MAX seeds, strict compares
// (src/axisalignedboundingbox.h:31-41)
__device__ __forceinline__ BoxD tri_box(const d3 &a, const d3 &b, const d3 &c)
{
    BoxD bx = {DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};
#define SB_UPD(v)                                   \
    if (v.x > bx.hix) bx.hix = v.x;                 \
    if (v.x < bx.lox) bx.lox = v.x;                 \
    if (v.y > bx.hiy) bx.hiy = v.y;                 \
    if (v.y < bx.loy) bx.loy = v.y;                 \
    if (v.z > bx.hiz) bx.hiz = v.z;                 \
    if (v.z < bx.loz) bx.loz = v.z;
    SB_UPD(a) SB_UPD(b) SB_UPD(c)
#undef SB_UPD
    return bx;
}

__device__ __forceinline__ uint32_t expand10(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

// 10 bits per axis -> 30-bit Morton key = four 8-bit radix passes
__device__ __forceinline__ uint32_t quant10(double c, double lo, double inv)
{
    float f = (float)((c - lo) * inv) * 1024.0f;
    f = fminf(fmaxf(f, 0.0f), 1023.0f); // NaN -> 0
    return (uint32_t)f;
}

#ifndef SB_TRIPREP_MINB
#define SB_TRIPREP_MINB 4 // 64 registers: four CTAs per SM (74 without the bound: three)
#endif
#ifndef SB_TRIPREP_HIST
#define SB_TRIPREP_HIST 1 // the kernel also counts the radix digits of the Morton keys it writes: no separate histogram pass over them
#endif
// Grid-stride over the triangles (with SB_TRIPREP_HIST the grid is one resident wave, so that a CTA's digit counts
// are flushed once for many triangles).
__global__ void __launch_bounds__(256, SB_TRIPREP_MINB) tri_prepare_kernel(const double4 *__restrict__ vtx, const uint32_t *__restrict__ tri,
    uint32_t nT, uint32_t nV, const unsigned long long *__restrict__ bounds, double4 *__restrict__ nrm4,
    uint32_t *__restrict__ mkey, uint32_t *__restrict__ order, int *__restrict__ err,
    unsigned long long *__restrict__ extentSum, const uint16_t *__restrict__ triJob,
    uint32_t *__restrict__ hist /* [passes][256] of the sort to come, or null */, uint32_t *__restrict__ histTicket /* null: one of several
    launches over ranges of the triangles, the histogram is finished by sbradix::hist_scan_kernel */, int beginBit, int passes, uint32_t first)
{
    __shared__ uint32_t s_hist[4 * 256];
    if (hist) {
        for (int k = threadIdx.x; k < passes * 256; k += 256)
            s_hist[k] = 0;
        __syncthreads();
    }
    // the triangles' box extents as 2^-24 fractions of the mesh extent (integers:
    // the sums, and with them the ray-grid resolution, do not depend on atomic order)
    unsigned long long sx = 0, sy = 0, sz = 0;
    const double blx = dkey_inv(bounds[0]), bly = dkey_inv(bounds[1]), blz = dkey_inv(bounds[2]);
    const double ex = dkey_inv(bounds[3]) - blx, ey = dkey_inv(bounds[4]) - bly, ez = dkey_inv(bounds[5]) - blz;
    const double ix = ex > 0 ? 1.0 / ex : 0.0, iy = ey > 0 ? 1.0 / ey : 0.0, iz = ez > 0 ? 1.0 / ez : 0.0;
    for (uint32_t i = first + blockIdx.x * 256 + threadIdx.x; i < nT; i += gridDim.x * 256) { // (nT: end of this launch's range)
    uint32_t i0 = tri[3 * (size_t)i], i1 = tri[3 * (size_t)i + 1], i2 = tri[3 * (size_t)i + 2];
    if (i0 >= nV || i1 >= nV || i2 >= nV) { // reported as SB_ERR_INVALID by the host
        *err = 1;
        i0 = i1 = i2 = 0;
    }
    d3 a = load_vertex(vtx, i0), b = load_vertex(vtx, i1), c = load_vertex(vtx, i2);

    // the triangle's box, here only for the Morton key and the grid resolution (the
    // sorted copy that the queries read is formed by leaf_gather_kernel)
    const BoxD bx = tri_box(a, b, c);

    // Vector3::normal (src/vector3.h:155-176)
    d3 ba = d3sub(b, a), ca = d3sub(c, a);
    d3 cr = d3cross(ba, ca);
    double len2 = xadd(xadd(xmul(cr.x, cr.x), xmul(cr.y, cr.y)), xmul(cr.z, cr.z));
    double len = xsqrt(len2);
    d3 n = {0.0, 0.0, 0.0};
    if (!(fabs(len) <= 2.2204460492503131e-16))
        n = {xdiv(cr.x, len), xdiv(cr.y, len), xdiv(cr.z, len)};
    stg256(nrm4 + i, n.x, n.y, n.z,
        __longlong_as_double((long long)pack_tri_idx(i0, i1, i2)));
    // 30-bit Morton key of the box centre inside the mesh box (ordering only)
    uint32_t qx = quant10(0.5 * (bx.lox + bx.hix), blx, ix);
    uint32_t qy = quant10(0.5 * (bx.loy + bx.hiy), bly, iy);
    uint32_t qz = quant10(0.5 * (bx.loz + bx.hiz), blz, iz);
    const uint32_t morton = (expand10(qx) << 2) | (expand10(qy) << 1) | expand10(qz);
    // batch mesh: the job leads the key (12 bits), 20 Morton bits follow -- every job is one run of the order
    const uint32_t key = triJob ? ((uint32_t)triJob[i] << 20) | (morton >> 10) : morton;
    mkey[i] = key;
    order[i] = i;
    if (hist)
        for (int p = 0; p < passes; ++p)
            atomicAdd(&s_hist[p * 256 + ((key >> (beginBit + 8 * p)) & 255u)], 1u);
    sx += (unsigned int)fmin(fmax((bx.hix - bx.lox) * ix * 16777216.0, 0.0), 16777216.0); // NaN -> 0
    sy += (unsigned int)fmin(fmax((bx.hiy - bx.loy) * iy * 16777216.0, 0.0), 16777216.0);
    sz += (unsigned int)fmin(fmax((bx.hiz - bx.loz) * iz * 16777216.0, 0.0), 16777216.0);
    }
    // mean triangle-box extent per axis (sizes the ray grids): CTA reduction, then
    // three atomics per CTA spread over 32 slots (same-address atomics serialise)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        sx += __shfl_xor_sync(SB_FULL, sx, off);
        sy += __shfl_xor_sync(SB_FULL, sy, off);
        sz += __shfl_xor_sync(SB_FULL, sz, off);
    }
    __shared__ unsigned long long s_ext[8][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        s_ext[warp][0] = sx;
        s_ext[warp][1] = sy;
        s_ext[warp][2] = sz;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        unsigned long long t = 0;
        for (int w = 0; w < 8; ++w)
            t += s_ext[w][threadIdx.x];
        atomicAdd(extentSum + 3 * (blockIdx.x & 31) + threadIdx.x, t);
    }
    if (hist)
        sbradix::hist_flush_and_scan<8>(s_hist, passes, hist, histTicket);
}

// ---- K1b --------------------------------------------------------------------
#ifndef SB_LEAF_MINB
#define SB_LEAF_MINB 4 // 64 registers: four CTAs per SM (the depth slabs took the kernel to 66 and one CTA less: 96 -> 109 us at C3)
#endif
__global__ void __launch_bounds__(256, SB_LEAF_MINB) leaf_gather_kernel(const uint32_t *__restrict__ sortedTri,
    const uint32_t *__restrict__ sortedKey, const double4 *__restrict__ vtx, const uint32_t *__restrict__ tri,
    uint32_t nT, uint32_t nV, uint32_t nTpad, Rec32 *__restrict__ leaf, double2 *__restrict__ sbox, double *__restrict__ scent,
    Rec32 *__restrict__ cbox, uint32_t *__restrict__ ckey, const GridParams *__restrict__ gp, uint4 *__restrict__ qbox,
    uint32_t *__restrict__ gridE, uint32_t *__restrict__ gridBigCount, int gridAxes, const uint16_t *__restrict__ triJob,
    double latPitch)
{
    __shared__ GridParams g;
    if (threadIdx.x < sizeof(GridParams) / 4)
        reinterpret_cast<uint32_t *>(&g)[threadIdx.x] = reinterpret_cast<const uint32_t *>(gp)[threadIdx.x];
    __syncthreads();
    uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nTpad) // nTpad is a multiple of 32: whole warps leave together
        return;
    BoxD bd = {DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};
    BoxF bf = empty_boxf();
    int ref = -1;
    d3 cen = {0.0, 0.0, 0.0};
    if (j < nT) {
        uint32_t t = sortedTri[j];
        uint32_t i0 = __ldg(tri + 3 * (size_t)t), i1 = __ldg(tri + 3 * (size_t)t + 1), i2 = __ldg(tri + 3 * (size_t)t + 2);
        if (i0 >= nV || i1 >= nV || i2 >= nV) // flagged by tri_prepare_kernel, reported by the host
            i0 = i1 = i2 = 0;
        const d3 a = load_vertex(vtx, i0), b = load_vertex(vtx, i1), c = load_vertex(vtx, i2);
        bd = tri_box(a, b, c); // exact box (reference `update` semantics), Morton order
        bf = enclose(bd);
        uint32_t job = 0;
        if (triJob) {
            // batch mesh: the conservative float box moves to the job's lattice position (rounded outwards
            // after the shift); only these floats are moved, the exact box below stays where it is
            job = triJob[t];
            const double ox = latPitch * lattice3(job, 0), oy = latPitch * lattice3(job, 1), oz = latPitch * lattice3(job, 2);
            bf = {__double2float_rd(bd.lox + ox), __double2float_rd(bd.loy + oy), __double2float_rd(bd.loz + oz),
                  __double2float_ru(bd.hix + ox), __double2float_ru(bd.hiy + oy), __double2float_ru(bd.hiz + oz)};
        }
        ref = (int)t;
        // face centroid exactly as decideGroupSide forms its query point:
        // (v0 + v1 + v2) / 3.0  (src/solidboolean.cpp:497-499)
        cen = {xdiv(xadd(xadd(a.x, b.x), c.x), 3.0), xdiv(xadd(xadd(a.y, b.y), c.y), 3.0), xdiv(xadd(xadd(a.z, b.z), c.z), 3.0)};
        // ray grids (sb_grid.cu): the quantised box, and the count pass while it is in registers
        const uint4 q = quantise_box(bd, g, t, job);
        qbox[j] = q;
#ifndef SB_EXP_NOCOUNT
        grid_count_tri(q, g, gridE, gridBigCount, gridAxes);
#endif
    }
#if SB_LEAF_STAGE
    // The 48-byte boxes and 24-byte centroids of a warp are 1536 + 768 contiguous bytes: staged through shared
    // memory and written as 256-bit stores, whole lines at a time (per-lane stores of 16 / 8 bytes at those
    // strides touched 21 sectors per request).  The padding lanes write an empty box and a zero centroid.
    {
        __shared__ __align__(32) double s_stage[8][32 * 9];
        const int lane = threadIdx.x & 31;
        double *sb = s_stage[threadIdx.x >> 5], *sc = sb + 32 * 6;
        sb[6 * lane] = bd.lox; sb[6 * lane + 1] = bd.loy; sb[6 * lane + 2] = bd.loz; // store_boxd order
        sb[6 * lane + 3] = bd.hix; sb[6 * lane + 4] = bd.hiy; sb[6 * lane + 5] = bd.hiz;
        sc[3 * lane] = cen.x; sc[3 * lane + 1] = cen.y; sc[3 * lane + 2] = cen.z;
        __syncwarp();
        const uint32_t j0 = j - lane; // multiple of 32
        double4 *db = reinterpret_cast<double4 *>(sbox + 3 * (size_t)j0), *dc = reinterpret_cast<double4 *>(scent + 3 * (size_t)j0);
        const double4 *s4 = reinterpret_cast<const double4 *>(sb);
        double4 q = s4[lane];
        stg256(db + lane, q.x, q.y, q.z, q.w);
        if (lane < 16) {
            q = s4[32 + lane];
            stg256(db + 32 + lane, q.x, q.y, q.z, q.w);
        } else if (lane < 16 + 24) {
            q = s4[48 + lane - 16]; // centroids: 24 records of 32 bytes, lanes 16..31 take the first 16 ...
            stg256(dc + lane - 16, q.x, q.y, q.z, q.w);
        }
        if (lane < 8) { // ... lanes 0..7 the rest
            q = s4[48 + 16 + lane];
            stg256(dc + 16 + lane, q.x, q.y, q.z, q.w);
        }
    }
#else
    store_boxd(sbox + 3 * (size_t)j, bd);
    if (j < nT) {
        scent[3 * (size_t)j] = cen.x;
        scent[3 * (size_t)j + 1] = cen.y;
        scent[3 * (size_t)j + 2] = cen.z;
    }
#endif
    store_rec(leaf + j, bf, ref, (int)j);
    // segmented (width K) min/max reduction in registers
    BoxF cb = bf;
#pragma unroll
    for (int off = 1; off < K; off <<= 1) {
        BoxF o = shfl_xor_box(cb, off);
        merge_f(cb, o);
    }
    if ((j % K) == 0 && j < nT) {
        uint32_t c = j / K;
        store_rec(cbox + c, cb, ~(int)c, 0);
        ckey[c] = sortedKey[j];
    }
}

// ---- K4 ---------------------------------------------------------------------
// Highest differing bit of the composite keys (Morton << 32 | index) of
// clusters i and i+1; distinct for the two ends of any radix-tree range.
__device__ __forceinline__ int split_level(const uint32_t *__restrict__ ckey, uint32_t i)
{
    uint32_t x = __ldg(ckey + i) ^ __ldg(ckey + i + 1);
    if (x)
        return 63 - __clz(x);
    return 31 - __clz(i ^ (i + 1));
}

__global__ void __launch_bounds__(256) tree_build_kernel(const Rec32 *__restrict__ cbox, const uint32_t *__restrict__ ckey,
    uint32_t M, Rec32 *nodes, int *slot, int *root)
{
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M)
        return;
    if (M == 1) {
        *root = -1; // ~0: the single cluster is the root
        return;
    }
    Rec32 me = load_rec(cbox + c);
    BoxF box = {me.lox, me.loy, me.loz, me.hix, me.hiy, me.hiz};
    int ref = ~(int)c;
    uint32_t l = c, r = c;
    while (true) {
        bool up; // true: parent is node r (we are its left child)
        if (l == 0)
            up = true;
        else if (r == M - 1)
            up = false;
        else
            up = split_level(ckey, r) < split_level(ckey, l - 1);
        uint32_t p = up ? r : l - 1;
        int side = up ? 0 : 1;
        store_rec(nodes + 2 * (size_t)p + side, box, ref, (int)(up ? l : r));
        __threadfence();
        int other = atomicExch(slot + p, (int)(up ? l : r));
        if (other < 0)
            return; // first to arrive: the sibling's thread carries on
        __threadfence();
        Rec32 sib = load_rec_cg(nodes + 2 * (size_t)p + (1 - side));
        BoxF sb = {sib.lox, sib.loy, sib.loz, sib.hix, sib.hiy, sib.hiz};
        merge_f(box, sb);
        if (up)
            r = (uint32_t)other;
        else
            l = (uint32_t)other;
        ref = (int)p;
        if (l == 0 && r == M - 1) {
            *root = (int)p;
            return;
        }
    }
}

// exact triangle boxes in ORIGINAL order (sb_mesh_triangle_boxes; the front end itself
// only reads the Morton-ordered copy)
__global__ void __launch_bounds__(256) tri_boxes_kernel(const double4 *__restrict__ vtx, const uint32_t *__restrict__ tri,
    uint32_t nT, uint32_t nV, double2 *__restrict__ out)
{
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nT)
        return;
    uint32_t i0 = tri[3 * (size_t)i], i1 = tri[3 * (size_t)i + 1], i2 = tri[3 * (size_t)i + 2];
    if (i0 >= nV || i1 >= nV || i2 >= nV)
        i0 = i1 = i2 = 0;
    store_boxd(out + 3 * (size_t)i, tri_box(load_vertex(vtx, i0), load_vertex(vtx, i1), load_vertex(vtx, i2)));
}

// batch upload: job-local vertex indices -> batch-global ones, job of every triangle
__global__ void __launch_bounds__(256) batch_fixup_kernel(uint32_t *__restrict__ tri, uint32_t nT,
    const uint32_t *__restrict__ triStart, const uint32_t *__restrict__ vtxStart, uint32_t nJobs, uint16_t *__restrict__ triJob,
    int *__restrict__ err)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nT)
        return;
    uint32_t lo = 0, hi = nJobs; // last job whose first triangle is <= t
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(triStart + mid) <= t)
            lo = mid;
        else
            hi = mid;
    }
    const uint32_t v0 = __ldg(vtxStart + lo), nv = __ldg(vtxStart + lo + 1) - v0;
    triJob[t] = (uint16_t)lo;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t i = tri[3 * (size_t)t + k];
        if (i >= nv) { // reported as SB_ERR_INVALID by the first build
            *err = 1;
            i = 0;
        }
        tri[3 * (size_t)t + k] = v0 + i;
    }
}

} // namespace

cudaError_t sbk_batch_fixup(cudaStream_t s, uint32_t *tri, uint32_t nT, const uint32_t *triStart, const uint32_t *vtxStart,
    uint32_t nJobs, uint16_t *triJob, int *err, LaunchCounter &lc)
{
    batch_fixup_kernel<<<(nT + 255) / 256, 256, 0, s>>>(tri, nT, triStart, vtxStart, nJobs, triJob, err);
    lc.kernels += 1;
    return cudaGetLastError();
}

cudaError_t sbk_triangle_boxes(cudaStream_t s, const MeshDev &m, double2 *out, LaunchCounter &lc)
{
    if (m.nT == 0)
        return cudaSuccess;
    tri_boxes_kernel<<<(m.nT + 255) / 256, 256, 0, s>>>(m.vtx, m.tri, m.nT, m.nV, out);
    lc.kernels += 1;
    return cudaGetLastError();
}

size_t sbk_radix_workspace_words(size_t n) { return sbradix::Workspace::words(sbradix::tiles_for(n)); }

// whole-mesh box + padded vertex copy (K2)
cudaError_t sbk_bounds_pad(cudaStream_t s, MeshDev &m, int smCount, LaunchCounter &lc, float2 *zf)
{
    // bounds seeds: min slots all-ones, max slots zero (order-encoded doubles)
    cudaMemsetAsync(m.bounds, 0xff, 3 * sizeof(unsigned long long), s);
    cudaMemsetAsync(m.bounds + 3, 0x00, 3 * sizeof(unsigned long long), s);
    int vb = (int)((m.nV + 255) / 256);
    if (vb > smCount * 8)
        vb = smCount * 8;
    if (vb < 1)
        vb = 1;
    bounds_pad_kernel<<<vb, 256, 0, s>>>(m.xyz, m.nV, m.vtx, m.bounds, zf);
    lc.kernels += 1;
    return cudaGetLastError();
}

static uint32_t tri_prepare_blocks(uint32_t count, int smCount, bool fused)
{
    uint32_t blocks = (count + 255) / 256;
    if (fused) {
        static int perSm = 0; // one resident wave
        if (!perSm && (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, tri_prepare_kernel, 256, 0) != cudaSuccess || perSm < 1))
            perSm = 2;
        blocks = std::min<uint32_t>(blocks, (uint32_t)(smCount * perSm));
    }
    return std::max<uint32_t>(blocks, 1);
}

static bool sort_is_fused(const MeshDev &m, int &passes, int &endBit)
{
    endBit = m.triJob ? 32 : 30;
    passes = sbradix::sort_passes(m.nT, m.sortBeginBit, endBit, 8);
    return SB_TRIPREP_HIST && passes > 0 && passes <= 4;
}

// The head of sbk_build_sort in three pieces, for a mesh whose arrays are still arriving from the host (sb_capi.cu
// upload_from_host): clears + bounds + padded vertices once the coordinates are there, the per-triangle kernel over
// each chunk of index triples as it lands, the digit offsets when all of them have been counted.
bool sbk_prep_supported(const MeshDev &m)
{
    int passes, endBit;
    return m.nT && !m.sharedVtx && !m.triJob && sort_is_fused(m, passes, endBit);
}

cudaError_t sbk_prep_begin(cudaStream_t s, MeshDev &m, uint32_t *radixWs, int smCount, LaunchCounter &lc)
{
    int passes, endBit;
    sort_is_fused(m, passes, endBit);
    cudaMemsetAsync(m.root, 0, 8, s); // root + the index-check flag the per-triangle kernel raises
    cudaMemsetAsync(m.extentSum, 0, 96 * sizeof(unsigned long long), s);
    sbradix::Workspace ws;
    ws.mem = radixWs;
    sbradix::sort_clear(s, ws, m.nT, passes);
    return sbk_bounds_pad(s, m, smCount, lc);
}

cudaError_t sbk_prep_triangles(cudaStream_t s, MeshDev &m, uint32_t *radixWs, uint32_t first, uint32_t end, int smCount, LaunchCounter &lc)
{
    if (end <= first)
        return cudaSuccess;
    int passes, endBit;
    sort_is_fused(m, passes, endBit);
    sbradix::Workspace ws;
    ws.mem = radixWs;
    tri_prepare_kernel<<<tri_prepare_blocks(end - first, smCount, true), 256, 0, s>>>(m.vtx, m.tri, end, m.nV, m.bounds, m.nrm4, m.mkey,
        m.order, m.err, m.extentSum, m.triJob, sbradix::sort_hist(ws), nullptr, m.sortBeginBit, passes, first);
    lc.kernels += 1;
    return cudaGetLastError();
}

cudaError_t sbk_prep_end(cudaStream_t s, MeshDev &m, uint32_t *radixWs, LaunchCounter &lc)
{
    int passes, endBit;
    sort_is_fused(m, passes, endBit);
    sbradix::Workspace ws;
    ws.mem = radixWs;
    sbradix::hist_scan_kernel<<<1, 256, 0, s>>>(sbradix::sort_hist(ws), passes);
    lc.kernels += 1;
    return cudaGetLastError();
}

// prepared: sbk_prep_begin / _triangles / _end have run for this geometry -- only the sort passes are left
cudaError_t sbk_build_sort(cudaStream_t s, MeshDev &m, uint32_t *radixWs, int smCount, LaunchCounter &lc, bool prepared)
{
    if (m.nT == 0)
        return cudaSuccess;
    sbradix::Workspace ws;
    ws.mem = radixWs;
    int passes, endBit;
    const bool fused = sort_is_fused(m, passes, endBit);
    if (!prepared) {
        cudaMemsetAsync(m.extentSum, 0, 96 * sizeof(unsigned long long), s);
        if (!m.sharedVtx) { // (a multi-GPU selection reads its parent's padded vertices and bounds)
            cudaError_t eb = sbk_bounds_pad(s, m, smCount, lc);
            if (eb != cudaSuccess)
                return eb;
            lc.kernels -= 1;
        }
        if (fused)
            sbradix::sort_clear(s, ws, m.nT, passes);
        tri_prepare_kernel<<<tri_prepare_blocks(m.nT, smCount, fused), 256, 0, s>>>(m.vtx, m.tri, m.nT, m.nV, m.bounds, m.nrm4, m.mkey, m.order, m.err,
            m.extentSum, m.triJob, fused ? sbradix::sort_hist(ws) : nullptr, sbradix::sort_ticket(ws), m.sortBeginBit, passes, 0u);
        lc.kernels += 2;
    }
    uint32_t *sk = nullptr, *sv = nullptr;
    lc.kernels += sbradix::sort<uint32_t, 8>(s, m.mkey, m.mkeyTmp, m.order, m.orderTmp, m.nT, m.sortBeginBit, endBit, ws, smCount, &sk, &sv,
        fused || prepared);
    m.sortedKey = sk;
    m.sortedTri = sv;
    return cudaGetLastError();
}

cudaError_t sbk_build_leaves(cudaStream_t s, MeshDev &m, LaunchCounter &lc)
{
    if (m.nT == 0)
        return cudaSuccess;
    leaf_gather_kernel<<<(m.nTpad + 255) / 256, 256, 0, s>>>(m.sortedTri, m.sortedKey, m.vtx, m.tri, m.nT, m.nV, m.nTpad,
        m.leaf, m.sbox, m.scent, m.cbox, m.ckey, m.gridParams, m.qbox, m.gridE, m.gridBigCount, m.gridAxes, m.triJob, m.latPitch);
    lc.kernels += 1;
    return cudaGetLastError();
}

// The LBVH topology + inner boxes are only needed when the mesh is the TARGET of a
// traversal (mesh B of sb_intersect); they are built on first such use.
cudaError_t sbk_build_tree(cudaStream_t s, MeshDev &m, LaunchCounter &lc)
{
    if (m.nT == 0)
        return cudaSuccess;
    if (m.M > 1)
        cudaMemsetAsync(m.slot, 0xff, sizeof(int) * (m.M - 1), s);
    tree_build_kernel<<<(m.M + 255) / 256, 256, 0, s>>>(m.cbox, m.ckey, m.M, m.nodes, m.slot, m.root);
    lc.kernels += 1;
    return cudaGetLastError();
}
