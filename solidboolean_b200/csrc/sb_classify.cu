// Stage 4: inside/outside classification.  Replaces SolidBoolean::isPointInMesh
// (reference src/solidboolean.cpp:48-92) as driven by decideGroupSide (:482-510).
//
// ONE kernel; a warp owns 32 neighbouring query points.  Round 1 traces the rays of
// axes 0 and 1 of every point; round 2 traces axis 2 only where those two votes
// disagree -- the third ray cannot change the majority otherwise -- or for every point
// when the per-axis bits are wanted.  Within a round the warp alternates between
// phases so that neither the latency-bound nor the FP64-bound part runs divergent:
//
//  count  one lane per POINT (its two rays one after the other, both cell ranges
//         fetched up front).  Ray box exactly as :53-58; the ray's cell list of the
//         target's axis-projected grid (sb_grid.cu) is filtered with the quantised
//         16-byte references (sb_gridq.cuh: three subtractions and a mask per
//         reference, integer only, slightly over-accepting).
//  fill   a prefix sum over the warp's 64 rays lays their matches out ray by ray in a
//         shared-memory pool; the rays with matches walk their (now cached) lists
//         again and store the triangle ids.
//  eval   one lane per staged (ray, triangle) ENTRY, dense, 32 at a time (chunks end
//         on ray boundaries).  Exact double box test = the reference's candidate set
//         (:55-63), then segment/plane hit + the two edge-normal sign tests
//         (sb_raytri.cuh, bit-exact) and the PositionKey of the hit.  Hits whose key
//         equals that of an earlier hit of the same ray are dropped
//         (std::set<PositionKey>, :64/:85); one ballot then gives every ray the
//         number of its distinct crossings; odd = inside (:89).
//
// A ray with more than 32 matches (stacked layers, silhouettes) gets the whole warp
// for its entries, its distinct keys collected in a list (64 in shared memory, the
// rest in a global scratch that the host sizes and, if ever too small, re-sizes
// exactly).  A ray whose box straddles a cell border or that has more matches than
// the pool holds is traced reference by reference by the warp (big_ray).
// Majority: (float)insideCount / totalCount > 0.5 with totalCount == 3 (:508).
// A target starts with the grids of axes 0 and 1 only; if a vote needs the third ray
// and that grid does not exist yet, the points are listed for a second launch (sb_capi.cu).
#include "sb_classify.cuh"

namespace {

#ifndef SB_CLS_CT
#define SB_CLS_CT 128
#endif
#ifndef SB_CLS_MINB
#define SB_CLS_MINB 6
#endif
#ifndef SB_CLS_UNROLL
#define SB_CLS_UNROLL 8 // references loaded per scan step (<= 8: allocation padding)
#endif
constexpr int CT = SB_CLS_CT; // threads per CTA
constexpr int CW = CT / 32;   // warps per CTA
#ifndef SB_CLS_POOL
#define SB_CLS_POOL 512
#endif
#ifndef SB_CLS_KSM
#define SB_CLS_KSM 32
#endif
constexpr int POOL = SB_CLS_POOL; // staged (ray, triangle) entries per warp
constexpr int KSM = SB_CLS_KSM; // distinct keys of a long ray held in shared memory
constexpr int KS = POOL * 4 / 24; // ... of a big_ray ray (in the then idle pool)

struct __align__(16) WarpStage {
    uint32_t tri[POOL];     // triangle ids, ray by ray
    long long key[32][3];   // PositionKeys of the current chunk's hits
    long long list[KSM][3]; // distinct keys of the long ray being evaluated
    uint8_t owner[POOL];    // entry -> slot * 32 + owner lane
    uint8_t hit[32];
    // per-lane state that lives here rather than in registers (the FP64 evaluation needs
    // them all; what does not fit is spilled to local memory, which misses L1 half of the
    // time at this footprint): the points, and per slot the packed ray + its cell list range
    double px[32], py[32], pz[32];
    uint32_t ray[2][4][32]; // [slot][packed ray x, y, list begin, list end][lane]
};
static_assert(KS * 24 <= POOL * 4, "big_ray keeps its keys in the pool");

// The same as a separately compiled function (its FP64 register pressure stays out of
// the scan loops): bit 0 = hit, bit 1 = candidate; the keys of a hit go to key3[0..2].
#ifndef SB_CLS_EVAL_CALL
#define SB_CLS_EVAL_CALL 0
#endif
#if SB_CLS_EVAL_CALL
__device__ __noinline__
#else
__device__ __forceinline__
#endif
uint32_t eval_call(const Target &T, double px, double py, double pz, int axis, uint32_t f, long long *key3)
{
    long long k0 = 0, k1 = 0, k2 = 0;
    bool isCand;
    const d3 p = {px, py, pz};
    const bool h = eval_entry(T, p, axis, f, k0, k1, k2, isCand);
    if (h) {
        key3[0] = k0;
        key3[1] = k1;
        key3[2] = k2;
    }
    return (h ? 1u : 0u) | (isCand ? 2u : 0u);
}

// A ray with more than one cell or more matches than the pool takes, traced by the
// whole warp: 32 references per step, hits de-duplicated against the ray's key list
// (the first KS keys in the idle pool, the rest in the global scratch).  n = its
// number of quantised matches, 0 if not known.  Returns {distinct hit keys, exact
// candidates seen}.
__device__ __noinline__ uint2 big_ray(const GridParams &g, const Target &T, const Out &o, long long *skeys, int axis, double px,
    double py, double pz, uint32_t job, uint32_t n, int lane)
{
    const d3 p = {px, py, pz};
    const RaySetup rs = ray_setup(g, axis, p, job);
    if (!rs.any)
        return make_uint2(0, 0);
    const uint4 *bigList = T.bigRefs + (size_t)axis * T.bigCap;
    const uint32_t nBig = big_list_length(T, axis);
    const RayQ rqAbs = ray_pack(rs.aU, rs.bU, rs.aV, rs.bV, rs.aA);
    // visit(probe, count) for each reference range of the ray; probe(i, id) = does reference i
    // of the range match (a triangle spanning several of the ray's cells is taken in the
    // first of them only: not if it also covers the previous cell of the ray's range)
    auto ranges = [&](auto &&visit) {
        for (uint32_t cv = rs.cv0; cv <= rs.cv1; ++cv)
            for (uint32_t cu = rs.cu0; cu <= rs.cu1; ++cu) {
                uint32_t i0, i1;
                grid_ray_range(g, T.E, axis, cu, cv, rs.aA, i0, i1);
                const CellRay cq = ray_in_cell(g, axis, rs, cu, cv);
                const bool firstU = cu == rs.cu0, firstV = cv == rs.cv0;
                visit([&](uint32_t i, uint32_t &id) {
                    const uint2 r = __ldg(T.refs + i0 + i);
                    id = cell_ref_id(r);
                    return cell_ref_match(cq, r) && (firstU || !cell_ref_ext_u(r)) && (firstV || !cell_ref_ext_v(r));
                }, i1 - i0);
            }
        visit([&](uint32_t i, uint32_t &id) {
            const uint4 r = __ldg(bigList + i);
            id = r.w;
            return ray_ref_match(rqAbs, r);
        }, nBig);
    };
    if (n == 0) { // count the matches (bounds the number of distinct keys)
        ranges([&](auto &&probe, uint32_t len) {
            for (uint32_t i = lane; i < len; i += 32) {
                uint32_t id;
                n += probe(i, id) ? 1u : 0u;
            }
        });
        n = __reduce_add_sync(SB_FULL, n);
        if (n == 0)
            return make_uint2(0, 0);
    }
    long long *gkeys = nullptr;
    if (n > KS) {
        unsigned long long base = 0;
        if (lane == 0)
            base = atomicAdd(o.bigNeeded, (unsigned long long)(n - KS));
        base = __shfl_sync(SB_FULL, base, 0);
        if (base + (n - KS) > o.bigCap)
            return make_uint2(0, 0); // scratch too small: the host sees bigNeeded > bigCap and repeats the call
        gkeys = o.bigKeys + 3 * base;
    }
    uint32_t count = 0, exact = 0; // distinct hit keys so far (<= n)
    ranges([&](auto &&probe, uint32_t len) {
        for (uint32_t i = 0; i < len; i += 32) {
            bool h = false;
            long long k0 = 0, k1 = 0, k2 = 0;
            uint32_t id = 0;
            if (i + lane < len && probe(i + lane, id)) {
                long long k[3] = {0, 0, 0};
                const uint32_t fl = eval_call(T, px, py, pz, axis, id, k);
                h = fl & 1u;
                exact += fl >> 1;
                k0 = k[0]; k1 = k[1]; k2 = k[2];
            }
            uint32_t hm = __ballot_sync(SB_FULL, h);
            while (hm) {
                const int s = __ffs(hm) - 1;
                hm &= hm - 1;
                const long long x = __shfl_sync(SB_FULL, k0, s), y = __shfl_sync(SB_FULL, k1, s), z = __shfl_sync(SB_FULL, k2, s);
                bool dup = false;
                for (uint32_t t = lane; t < count; t += 32) {
                    if (t < KS)
                        dup |= skeys[3 * t] == x && skeys[3 * t + 1] == y && skeys[3 * t + 2] == z;
                    else
                        dup |= __ldcg(gkeys + 3 * (t - KS)) == x && __ldcg(gkeys + 3 * (t - KS) + 1) == y &&
                               __ldcg(gkeys + 3 * (t - KS) + 2) == z;
                }
                if (!__any_sync(SB_FULL, dup)) {
                    if (lane == 0) {
                        long long *dst = count < KS ? skeys + 3 * count : gkeys + 3 * (count - KS);
                        dst[0] = x;
                        dst[1] = y;
                        dst[2] = z;
                    }
                    ++count;
                    __syncwarp();
                }
            }
        }
    });
    return make_uint2(count, __reduce_add_sync(SB_FULL, exact));
}

// One ray's cell list, references [a, b) of refs (b even, sb_grid.cu), walked by its lane in
// aligned 16-byte PAIRS: number of references whose quantised box the ray matches.
__device__ __forceinline__ uint32_t scan_count(const uint2 *__restrict__ refs, uint32_t qx, uint32_t qy, uint32_t a, uint32_t b)
{
    const CellRay rq = {qx, qy};
    uint32_t m = 0;
    // independent loads in flight; the lists are followed by at least eight more readable
    // references (padding), so the last group stays in bounds
    for (uint32_t i = a & ~1u; i < b; i += SB_CLS_UNROLL) {
        uint4 q[SB_CLS_UNROLL / 2];
#pragma unroll
        for (int k = 0; k < SB_CLS_UNROLL / 2; ++k)
            q[k] = __ldg(reinterpret_cast<const uint4 *>(refs + i) + k);
#pragma unroll
        for (int k = 0; k < SB_CLS_UNROLL / 2; ++k) {
            m += ((i + 2 * k >= a) & (i + 2 * k < b) & cell_ref_match(rq, make_uint2(q[k].x, q[k].y))) ? 1u : 0u;
            m += ((i + 2 * k + 1 < b) & cell_ref_match(rq, make_uint2(q[k].z, q[k].w))) ? 1u : 0u;
        }
    }
    return m;
}

// the same walk, storing the matching triangle ids (and their owner) from `pos` on
__device__ __forceinline__ uint32_t scan_fill(const uint2 *__restrict__ refs, uint32_t qx, uint32_t qy, uint32_t a, uint32_t b,
    uint32_t *tri, uint8_t *owner, uint32_t pos, uint8_t rid)
{
    const CellRay rq = {qx, qy};
    for (uint32_t i = a & ~1u; i < b; i += SB_CLS_UNROLL) {
        uint4 q[SB_CLS_UNROLL / 2];
#pragma unroll
        for (int k = 0; k < SB_CLS_UNROLL / 2; ++k)
            q[k] = __ldg(reinterpret_cast<const uint4 *>(refs + i) + k);
#pragma unroll
        for (int k = 0; k < SB_CLS_UNROLL / 2; ++k) {
            const uint2 r0 = make_uint2(q[k].x, q[k].y), r1 = make_uint2(q[k].z, q[k].w);
            if ((i + 2 * k >= a) & (i + 2 * k < b) & cell_ref_match(rq, r0)) {
                tri[pos] = cell_ref_id(r0);
                owner[pos] = rid;
                ++pos;
            }
            if ((i + 2 * k + 1 < b) & cell_ref_match(rq, r1)) {
                tri[pos] = cell_ref_id(r1);
                owner[pos] = rid;
                ++pos;
            }
        }
    }
    return pos;
}

// the per-axis big list (triangles covering too many cells; usually empty): absolute references
__device__ __forceinline__ uint32_t big_count(const uint4 *__restrict__ big, uint32_t nBig, const RayQ &rq)
{
    uint32_t m = 0;
    for (uint32_t i = 0; i < nBig; ++i)
        m += ray_ref_match(rq, __ldg(big + i)) ? 1u : 0u;
    return m;
}

__device__ __forceinline__ void big_fill(const uint4 *__restrict__ big, uint32_t nBig, const RayQ &rq, uint32_t *tri,
    uint8_t *owner, uint32_t pos, uint8_t rid)
{
    for (uint32_t i = 0; i < nBig; ++i) {
        const uint4 q = __ldg(big + i);
        if (ray_ref_match(rq, q)) {
            tri[pos] = q.w;
            owner[pos] = rid;
            ++pos;
        }
    }
}

__device__ __forceinline__ uint32_t low_mask(uint32_t n) { return n >= 32 ? 0xffffffffu : (1u << n) - 1u; }

// One round for the warp's points: rays along axis0 .. axis0 + nax - 1 (nax = 1 or 2)
// of the lanes that `want` them.  Returns bit s set when the ray along axis0 + s
// crosses an odd number of distinct surface points (:89).
__device__ __forceinline__ uint32_t trace_round(const GridParams &g, const Target &T, const Out &o, WarpStage &W, int axis0, int nax,
    bool want, int lane, uint32_t job, uint32_t &exact)
{
    const d3 p = {W.px[lane], W.py[lane], W.pz[lane]};
    // ---- count ----
    // per slot s (ray along axis0 + s): cell-relative packed ray, cell list range
    uint32_t qx0 = 0, qy0 = 0, a0 = 0, b0 = 0, n0 = 0;
    uint32_t qx1 = 0, qy1 = 0, a1 = 0, b1 = 0, n1 = 0;
    bool legacy0 = false, legacy1 = false;
    if (want) {
        const RaySetup r = ray_setup(g, axis0, p, job);
        // a ray box is a point widened by DBL_EPSILON: it almost always sits in ONE cell;
        // the others go the general way (big_ray)
        legacy0 = r.any && (r.cu0 != r.cu1 || r.cv0 != r.cv1);
        if (r.any && !legacy0) {
            const CellRay cq = ray_in_cell(g, axis0, r, r.cu0, r.cv0);
            qx0 = cq.x; qy0 = cq.y;
            grid_ray_range(g, T.E, axis0, r.cu0, r.cv0, r.aA, a0, b0);
            const uint32_t nBig = big_list_length(T, axis0);
            if (nBig)
                n0 = big_count(T.bigRefs + (size_t)axis0 * T.bigCap, nBig, ray_pack(r.aU, r.bU, r.aV, r.bV, r.aA));
        }
    }
    if (want && nax > 1) {
        const RaySetup r = ray_setup(g, axis0 + 1, p, job);
        legacy1 = r.any && (r.cu0 != r.cu1 || r.cv0 != r.cv1);
        if (r.any && !legacy1) {
            const CellRay cq = ray_in_cell(g, axis0 + 1, r, r.cu0, r.cv0);
            qx1 = cq.x; qy1 = cq.y;
            grid_ray_range(g, T.E, axis0 + 1, r.cu0, r.cv0, r.aA, a1, b1);
            const uint32_t nBig = big_list_length(T, axis0 + 1);
            if (nBig)
                n1 = big_count(T.bigRefs + (size_t)(axis0 + 1) * T.bigCap, nBig, ray_pack(r.aU, r.bU, r.aV, r.bV, r.aA));
        }
    }
    n0 += scan_count(T.refs, qx0, qy0, a0, b0);
    n1 += scan_count(T.refs, qx1, qy1, a1, b1);
    W.ray[0][0][lane] = qx0; W.ray[0][1][lane] = qy0; W.ray[0][2][lane] = a0; W.ray[0][3][lane] = b0;
    W.ray[1][0][lane] = qx1; W.ray[1][1][lane] = qy1; W.ray[1][2][lane] = a1; W.ray[1][3][lane] = b1;
    legacy0 = legacy0 || n0 > o.poolLimit;
    legacy1 = legacy1 || n1 > o.poolLimit;
    const uint32_t ns0 = legacy0 ? 0u : n0, ns1 = legacy1 ? 0u : n1;
    uint32_t incl0 = ns0, incl1 = ns1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t0 = __shfl_up_sync(SB_FULL, incl0, d), t1 = __shfl_up_sync(SB_FULL, incl1, d);
        if (lane >= d) {
            incl0 += t0;
            incl1 += t1;
        }
    }
    // dense order: the entries of the slot-0 rays (by lane), then those of the slot-1 rays
    const uint32_t total0 = __shfl_sync(SB_FULL, incl0, 31);
    const uint32_t total = total0 + __shfl_sync(SB_FULL, incl1, 31);
    const uint32_t off0 = incl0 - ns0, off1 = total0 + incl1 - ns1;
    const uint32_t lelane = lanemask_lt() | (1u << lane);
    uint32_t parity = 0;
    // windows of at most POOL entries (almost always one), ending on ray boundaries
    for (uint32_t Wb = 0; Wb < total;) {
        const uint32_t w0 = (ns0 && off0 >= Wb && off0 + ns0 > Wb + POOL) ? off0 : total;
        const uint32_t w1 = (ns1 && off1 >= Wb && off1 + ns1 > Wb + POOL) ? off1 : total;
        const uint32_t We = __reduce_min_sync(SB_FULL, min(w0, w1));
        // ---- fill ---- (the rays with matches walk their, now cached, lists again)
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const uint32_t nss = s ? ns1 : ns0, offs = s ? off1 : off0;
            if (nss && offs >= Wb && offs < We) {
                const uint32_t fa = W.ray[s][2][lane], fb = W.ray[s][3][lane];
                const uint8_t rid = (uint8_t)(32 * s + lane);
                const uint32_t pos = scan_fill(T.refs, W.ray[s][0][lane], W.ray[s][1][lane], fa, fb, W.tri, W.owner, offs - Wb, rid);
                const uint32_t nBig = big_list_length(T, axis0 + s);
                if (nBig) { // same order as counted: the cell list, then the big list
                    const RaySetup r = ray_setup(g, axis0 + s, p, job);
                    big_fill(T.bigRefs + (size_t)(axis0 + s) * T.bigCap, nBig, ray_pack(r.aU, r.bU, r.aV, r.bV, r.aA), W.tri, W.owner, pos, rid);
                }
            }
        }
        __syncwarp();
        // ---- eval ----
        for (uint32_t S = Wb; S < We;) {
            // the chunk ends where the first ray that does not fit into 32 entries begins
            const uint32_t c0 = (ns0 && off0 >= S && off0 + ns0 > S + 32) ? off0 : We;
            const uint32_t c1 = (ns1 && off1 >= S && off1 + ns1 > S + 32) ? off1 : We;
            const uint32_t Snext = min(__reduce_min_sync(SB_FULL, min(c0, c1)), We);
            if (Snext == S) {
                // a ray with more than 32 entries starts here: the whole warp takes it, 32
                // entries per step, distinct keys collected in W.list / the global scratch
                const uint32_t m0 = __ballot_sync(SB_FULL, ns0 > 32 && off0 == S);
                const uint32_t m1 = __ballot_sync(SB_FULL, ns1 > 32 && off1 == S);
                const int sl = m0 ? 0 : 1;
                const int b = __ffs(m0 ? m0 : m1) - 1;
                const uint32_t nl = __shfl_sync(SB_FULL, sl ? ns1 : ns0, b);
                const d3 pb = {W.px[b], W.py[b], W.pz[b]};
                long long *gl = nullptr;
                bool ok = true;
                if (nl > KSM) {
                    unsigned long long base = 0;
                    if (lane == 0)
                        base = atomicAdd(o.bigNeeded, (unsigned long long)(nl - KSM));
                    base = __shfl_sync(SB_FULL, base, 0);
                    ok = base + (nl - KSM) <= o.bigCap; // else the host repeats the call with a larger scratch
                    gl = o.bigKeys + 3 * base;
                }
                uint32_t count = 0;
                for (uint32_t c = 0; c < nl && ok; c += 32) {
                    bool h = false;
                    long long k0 = 0, k1 = 0, k2 = 0;
                    if (c + lane < nl) {
                        const uint32_t fl = eval_call(T, pb.x, pb.y, pb.z, axis0 + sl, W.tri[S - Wb + c + lane], &W.key[lane][0]);
                        h = fl & 1u;
                        exact += fl >> 1;
                    }
                    W.hit[lane] = h ? 1 : 0;
                    if (h) {
                        k0 = W.key[lane][0];
                        k1 = W.key[lane][1];
                        k2 = W.key[lane][2];
                    }
                    __syncwarp();
                    if (h)
                        for (int q = 0; q < lane; ++q)
                            if (W.hit[q] && W.key[q][0] == k0 && W.key[q][1] == k1 && W.key[q][2] == k2) {
                                h = false;
                                break;
                            }
                    for (uint32_t t = 0; t < count; ++t) {
                        long long x, y, z;
                        if (t < KSM) {
                            x = W.list[t][0]; y = W.list[t][1]; z = W.list[t][2];
                        } else {
                            x = __ldcg(gl + 3 * (t - KSM)); y = __ldcg(gl + 3 * (t - KSM) + 1); z = __ldcg(gl + 3 * (t - KSM) + 2);
                        }
                        h = h && !(x == k0 && y == k1 && z == k2);
                    }
                    const uint32_t sm = __ballot_sync(SB_FULL, h);
                    if (h) {
                        const uint32_t pos = count + __popc(sm & lanemask_lt());
                        long long *dst = pos < KSM ? &W.list[pos][0] : gl + 3 * (pos - KSM);
                        dst[0] = k0;
                        dst[1] = k1;
                        dst[2] = k2;
                    }
                    count += __popc(sm);
                    __syncwarp();
                }
                if (lane == b)
                    parity |= (count & 1u) << sl;
                S += nl;
                continue;
            }
            const bool have = (uint32_t)lane < Snext - S;
            uint32_t rid = lane;
            if (have)
                rid = W.owner[S - Wb + lane];
            const int ow = rid & 31, sl = rid >> 5;
            // first lane of this entry's ray (chunks start on ray boundaries)
            const uint32_t ridPrev = __shfl_up_sync(SB_FULL, rid, 1);
            const uint32_t heads = __ballot_sync(SB_FULL, have && (lane == 0 || rid != ridPrev));
            const int segStart = 31 - __clz(heads & lelane);
            const d3 pp = {W.px[ow], W.py[ow], W.pz[ow]};
            bool h = false;
            long long k0 = 0, k1 = 0, k2 = 0;
            if (have) {
                const uint32_t fl = eval_call(T, pp.x, pp.y, pp.z, axis0 + sl, W.tri[S - Wb + lane], &W.key[lane][0]);
                h = fl & 1u;
                exact += fl >> 1;
                W.hit[lane] = h ? 1 : 0;
                if (h) {
                    k0 = W.key[lane][0];
                    k1 = W.key[lane][1];
                    k2 = W.key[lane][2];
                }
            }
            __syncwarp();
            // std::set<PositionKey>: a hit whose key equals that of an earlier hit of the ray does not count
            if (h)
                for (int q = segStart; q < lane; ++q)
                    if (W.hit[q] && W.key[q][0] == k0 && W.key[q][1] == k1 && W.key[q][2] == k2) {
                        h = false;
                        break;
                    }
            const uint32_t dm = __ballot_sync(SB_FULL, h);
            if (ns0 && off0 >= S && off0 + ns0 <= Snext)
                parity |= __popc((dm >> (off0 - S)) & low_mask(ns0)) & 1u;
            if (ns1 && off1 >= S && off1 + ns1 <= Snext)
                parity |= (__popc((dm >> (off1 - S)) & low_mask(ns1)) & 1u) << 1;
            __syncwarp();
            S = Snext;
        }
        Wb = We;
    }
    // ---- rays that span several cells or overflow the pool ----
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        uint32_t bigMask = __ballot_sync(SB_FULL, s ? legacy1 : legacy0);
        while (bigMask) {
            const int b = __ffs(bigMask) - 1;
            bigMask &= bigMask - 1;
            const uint32_t nb = __shfl_sync(SB_FULL, s ? n1 : n0, b); // 0 for a ray of several cells: not counted yet
            const uint2 d = big_ray(g, T, o, reinterpret_cast<long long *>(W.tri), axis0 + s, W.px[b], W.py[b], W.pz[b],
                __shfl_sync(SB_FULL, job, b), nb, lane);
            if (lane == b)
                parity |= (d.x & 1u) << s;
            if (lane == 0)
                exact += d.y;
            __syncwarp();
        }
    }
    return parity;
}

__global__ void __launch_bounds__(CT, SB_CLS_MINB) classify_kernel(const __grid_constant__ Query q, const __grid_constant__ Target T,
    const __grid_constant__ Out o)
{
    __shared__ GridParams g;
    __shared__ WarpStage s_stage[CW];
    unsigned long long traceStart = 0;
    if (o.trace && threadIdx.x == 0)
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(traceStart));
    if (threadIdx.x < sizeof(GridParams) / 4)
        reinterpret_cast<uint32_t *>(&g)[threadIdx.x] = reinterpret_cast<const uint32_t *>(T.gp)[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    WarpStage &W = s_stage[threadIdx.x >> 5];
    const uint32_t j = q.first + blockIdx.x * CT + threadIdx.x;

    d3 p = {0, 0, 0};
    uint32_t outIndex = 0, local = 0, job = 0;
    bool active = j < q.count;
    if (active) {
        local = q.list ? __ldg(q.list + j) : j;
        const uint32_t idx = q.begin + local;
        if (q.rawTri) {
            p = raw_face_centroid(q, idx);
            outIndex = idx;
        } else if (q.pts) {
            p = {q.pts[3 * (size_t)idx], q.pts[3 * (size_t)idx + 1], q.pts[3 * (size_t)idx + 2]};
            outIndex = idx;
        } else if (idx < q.nT) {
            // faces mode: the centroids ((v0 + v1) + v2) / 3.0 (src/solidboolean.cpp:497-499) were
            // formed at build time and stored in Morton order
            p = {__ldg(q.scent + 3 * (size_t)idx), __ldg(q.scent + 3 * (size_t)idx + 1), __ldg(q.scent + 3 * (size_t)idx + 2)};
            outIndex = __ldg(q.sortedTri + idx);
            if (q.triJob)
                job = q.triJob[outIndex];
            if (q.ownMode && !(p.z >= q.ownLo && (q.ownMode == 2 ? p.z <= q.ownHi : p.z < q.ownHi)))
                active = false; // another rank's face (sb_shard.cu)
            if (q.origFace)
                outIndex = __ldg(q.origFace + outIndex);
        } else {
            active = false; // padding of the sorted order
        }
    }

    W.px[lane] = p.x;
    W.py[lane] = p.y;
    W.pz[lane] = p.z;
    __syncwarp();
    uint32_t votes = 0, exact = 0;
    bool undecided = false, deferred = false;
    // second launch (thirdOnly): votes 0 and 1 of the listed points disagreed, so only round 1
    // (axis 2) runs and its vote is the majority
    for (int round = o.thirdOnly ? 1 : 0; round < 2; ++round) {
        bool want = active;
        if (round == 1 && !o.perAxis && !o.thirdOnly) {
            // lazy majority: the third ray only where the first two disagree
            undecided = active && (((votes >> 1) ^ votes) & 1u);
            want = undecided;
            if (T.naxes < 3) { // no third grid yet: the host has it built and launches again for these points
                deferred = true;
                break;
            }
        }
        if (!__any_sync(SB_FULL, want))
            continue;
        votes |= trace_round(g, T, o, W, 2 * round, 2 - round, want, lane, job, exact) << (2 * round);
    }
    if (active) {
        bool in;
        if (o.perAxis) {
            o.perAxis[3 * (size_t)outIndex] = votes & 1u;
            o.perAxis[3 * (size_t)outIndex + 1] = (votes >> 1) & 1u;
            o.perAxis[3 * (size_t)outIndex + 2] = (votes >> 2) & 1u;
            in = __popc(votes) >= 2; // (float)insideCount / totalCount > 0.5 (:508)
        } else {
            in = (undecided || o.thirdOnly) ? ((votes >> 2) & 1u) != 0 : (votes & 1u) != 0;
        }
        if (!(deferred && undecided))
            o.inside[outIndex] = in ? 1 : 0;
    }
    const uint32_t ex = __reduce_add_sync(SB_FULL, exact);
    const uint32_t um = __ballot_sync(SB_FULL, undecided);
    unsigned int ubase = 0;
    if (lane == 0) {
        if (ex)
            atomicAdd(o.exactCount, (unsigned long long)ex);
        if (um)
            ubase = atomicAdd(o.undecidedCount, (unsigned int)__popc(um));
    }
    if (deferred && um) {
        ubase = __shfl_sync(SB_FULL, ubase, 0);
        if (undecided)
            o.undecidedList[ubase + __popc(um & lanemask_lt())] = local;
    }
    if (o.trace) { // dev instrumentation
        if (lane == 0 && ex)
            atomicAdd(o.trace + 4 * (size_t)blockIdx.x + 3, (unsigned long long)ex);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t1;
            uint32_t smid;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
            o.trace[4 * (size_t)blockIdx.x] = smid;
            o.trace[4 * (size_t)blockIdx.x + 1] = traceStart;
            o.trace[4 * (size_t)blockIdx.x + 2] = t1;
        }
    }
}

} // namespace

uint32_t sbk_classify_blocks(uint32_t points) { return (points + CT - 1) / CT; }

cudaError_t sbk_classify(cudaStream_t s, const MeshDev &target, const ClassifyArgs &a, long long *bigKeys,
    unsigned long long bigCap, unsigned long long *bigNeeded, unsigned long long *exactCount, unsigned int *undecidedCount,
    uint32_t poolLimit, unsigned long long *trace, LaunchCounter &lc)
{
    if (a.end <= a.begin)
        return cudaSuccess;
    const MeshDev *qm = a.queryMesh;
    Query q;
    q.pts = a.pts;
    q.scent = qm ? qm->scent : nullptr;
    q.sortedTri = qm ? qm->sortedTri : nullptr;
    q.nT = qm ? qm->nT : 0;
    q.begin = a.begin;
    q.count = a.list ? a.listCount : a.end - a.begin;
    q.list = a.list;
    q.triJob = qm ? qm->triJob : nullptr;
    q.origFace = qm ? qm->origFace : nullptr;
    q.ownMode = qm && qm->ownFilter ? (qm->ownClosed ? 2 : 1) : 0;
    q.ownLo = qm ? qm->ownLo : 0.0;
    q.ownHi = qm ? qm->ownHi : 0.0;
    q.rawTri = a.rawFaces && qm ? qm->tri : nullptr;
    q.rawXyz = a.rawFaces && qm ? qm->xyz : nullptr;
    q.rawNV = qm ? qm->nV : 0;
    q.first = a.list ? 0u : a.first;
    if (q.count <= q.first)
        return cudaSuccess;
    Target T;
    T.gp = target.gridParams;
    T.E = target.gridE;
    T.refs = target.gridRefs;
    T.refPairs = (target.gridRefCap + 8) / 2;
    T.nT = target.nT;
    T.bigRefs = target.gridBigRefs;
    T.bigCap = target.gridBigCap;
    T.bigN0 = target.gridBigN[0];
    T.bigN1 = target.gridBigN[1];
    T.bigN2 = target.gridBigN[2];
    T.naxes = target.gridAxes;
    T.vtx = target.vtx;
    T.tri = target.tri;
    T.nrm4 = target.nrm4;
    Out o = {};
    o.inside = a.inside;
    o.perAxis = a.perAxis;
    o.bigKeys = bigKeys;
    o.bigCap = bigCap;
    o.bigNeeded = bigNeeded;
    o.exactCount = exactCount;
    o.undecidedCount = undecidedCount;
    o.undecidedList = a.undecidedList;
    o.thirdOnly = a.thirdAxisOnly ? 1 : 0;
    o.trace = trace;
    o.poolLimit = poolLimit ? (poolLimit < (uint32_t)POOL ? poolLimit : (uint32_t)POOL) : (uint32_t)POOL;
    if (trace)
        cudaMemsetAsync(trace, 0, 32 * (size_t)((q.count + CT - 1) / CT), s);
    classify_kernel<<<(q.count - q.first + CT - 1) / CT, CT, 0, s>>>(q, T, o);
    lc.kernels += 1;
    return cudaGetLastError();
}
