// Stage 4, the balanced kernel: inside/outside classification of SolidBoolean::isPointInMesh
// (reference src/solidboolean.cpp:48-92) as driven by decideGroupSide (:482-510), for the
// common case -- every ray box inside ONE cell of the target's ray grid, no per-axis big
// lists.  Same arithmetic as sb_classify.cu (the shared helpers of sb_classify.cuh: ray box,
// quantised filter, exact box test, segment/plane hit, PositionKey); what differs is how the
// work is laid out over the warp.
//
// A warp owns 32 neighbouring query points and, per round, their rays along two axes
// (64 rays).  The cell lists of those rays differ a lot in length (C3: median 0-14
// references, 1 % above 100), so a lane per ray leaves half the lanes idle, and counting
// before filling walks every list twice.  Here the 64 lists are laid end to end as ONE
// virtual list of 16-byte reference pairs (a warp prefix sum over the pair counts) and
// walked once, 32 consecutive pairs per step, whatever ray they belong to:
//
//   owners  per window of 256 pairs: each ray drops its number at the position where its
//           list starts, a byte-wise running maximum spreads it (one SIMD max-scan, no loops)
//   walk    lane i of step k takes pair 32 k + i: its ray's packed box from shared memory,
//           one 128-bit load, two quantised tests (sb_gridq.cuh); matches are compacted with
//           two ballots into the pool -- in list order, so the entries of a ray stay together
//   eval    32 entries per step (chunks end where a ray ends; a ray that continues in the next
//           window is held back): exact box test = the reference's candidate set (:55-63),
//           segment/plane hit and edge-normal signs (sb_raytri.cuh), PositionKey; hits with
//           the key of an earlier hit of the same ray are dropped (std::set, :64/:85); one
//           ballot gives every ray the parity of its distinct crossings (:89)
//
// A ray with more than 32 matches takes whole chunks, its distinct keys collected in a short list.
// Points this layout does not cover -- a ray box that straddles a cell border, a ray with
// more than 32 DISTINCT crossings -- are listed for the general kernel (sb_classify.cu), which the
// host launches over that list afterwards; they are rare (C3: none).
// Majority vote, lazy third ray and the "third grid missing" list work as in sb_classify.cu.
#include "sb_classify.cuh"
#include <algorithm>

namespace {

#ifndef SB_CLS2_CT
#define SB_CLS2_CT 128
#endif
#ifndef SB_CLS2_MINB
#define SB_CLS2_MINB 6
#endif
#ifndef SB_CLS2_OPEN
#define SB_CLS2_OPEN 1         // evaluation steps may end inside a ray (32 of 32 lanes instead of 27)
#endif
#ifndef SB_CLS2_FULLCHUNKS
#define SB_CLS2_FULLCHUNKS 1   // evaluation steps wait for 32 entries (except in a round's last window)
#endif
constexpr int CT2 = SB_CLS2_CT;
constexpr int CW2 = CT2 / 32;
constexpr int WINP = 256;             // pairs per window (one owner byte each)
constexpr int POOL2 = 2 * WINP + 32;  // every reference of a window may match, plus the held-back tail
constexpr int KL = 32;                // distinct keys kept for a ray that spans several chunks

#ifndef SB_CLS2_PREFETCH
#define SB_CLS2_PREFETCH 0 // 1 / 2: a matched triangle's record is prefetched into L1 / L2 when the walk finds it
#endif
__device__ __forceinline__ void prefetch_record(const Target &T, uint32_t id)
{
#if SB_CLS2_PREFETCH == 1
    asm volatile("prefetch.global.L1 [%0];" ::"l"(T.nrm4 + id));
#elif SB_CLS2_PREFETCH == 2
    asm volatile("prefetch.global.L2 [%0];" ::"l"(T.nrm4 + id));
#else
    (void)T; (void)id;
#endif
}

struct __align__(16) Stage2 {
    uint32_t tri[POOL2];    // triangle ids of the matches, ray by ray
    long long key[32][3];   // PositionKeys of the current chunk's hits
    long long list[KL][3];  // distinct keys of the earlier chunks of a ray with more than 32 entries
    uint4 ray[64];          // per ray: packed box x, y; list position of its first pair if that starts with a foreign slot, else ~0
    uint32_t rayBase[64];   // per ray: pair address - list position
    double px[32], py[32], pz[32];
    uint8_t own[WINP];      // window pair -> ray + 1
    uint8_t owner[POOL2];   // entry -> ray (slot * 32 + lane)
    uint8_t hit[32];
    uint8_t vote[64];       // ray -> parity of its distinct crossings | exact candidates << 1
};

__device__ __forceinline__ uint32_t warp_incl_add(uint32_t v, int lane)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(SB_FULL, v, d);
        if (lane >= d)
            v += t;
    }
    return v;
}

// byte-wise inclusive running maximum inside a word (byte 0 first)
__device__ __forceinline__ uint32_t prefix_max_bytes(uint32_t x)
{
    x = __vmaxu4(x, x << 8);
    return __vmaxu4(x, x << 16);
}

// The ray of point p along AXIS: its packed box (rx, ry), where its cell list starts (ra) and how many
// 16-byte pairs it spans (np).  Returns true when the ray box straddles a cell border (general kernel).
template <int AXIS>
__device__ __forceinline__ bool setup_slot(const GridParams &g, const Target &T, const d3 &p, uint32_t job, uint32_t &rx,
    uint32_t &ry, uint32_t &ra, uint32_t &np)
{
    const RaySetup r = ray_setup_finite<AXIS>(g, p, job);
    if (!r.any)
        return false;
    // a ray box is a point widened by DBL_EPSILON: it almost always sits in ONE cell
    if (r.cu0 != r.cu1 || r.cv0 != r.cv1)
        return true;
    const CellRay cq = ray_in_cell(g, AXIS, r, r.cu0, r.cv0);
    uint32_t a, b; // the sub-lists of the depth slabs the ray can reach (sb_gridq.cuh)
    grid_ray_range(g, T.E, AXIS, r.cu0, r.cv0, r.aA, a, b);
#if SB_CLS_GUARDS
    // (a run enqueued before the host knew whether the rebuild fitted its lists, see Target: the walk stays inside the allocation)
    b = min(b, 2u * T.refPairs);
    a = min(a, b);
#endif
    rx = cq.x;
    ry = cq.y;
    ra = a;
    np = b > a ? (b - (a & ~1u)) >> 1 : 0u;
    return false;
}

// One round for the warp's points: rays along axis0 .. axis0 + nax - 1 (nax = 1 or 2) of the
// lanes that `want` them.  Returns bit s set when the ray along axis0 + s crosses an odd
// number of distinct surface points (:89); lanes whose point needs the general kernel get
// their bit in `legacy`.
template <int axis0, int nax>
__device__ __forceinline__ uint32_t trace_round2(const GridParams &g, const Target &T, Stage2 &W, bool want,
    int lane, uint32_t job, uint32_t &exact, uint32_t &legacy)
{
    const d3 p = {W.px[lane], W.py[lane], W.pz[lane]};
    // ---- setup: packed ray, cell list, pairs ----
    uint32_t np[2] = {0, 0}, rx[2] = {0, 0}, ry[2] = {0, 0}, ra[2] = {0, 0};
    bool multi = false;
    if (want) {
        multi |= setup_slot<axis0>(g, T, p, job, rx[0], ry[0], ra[0], np[0]);
        if (nax > 1)
            multi |= setup_slot<(axis0 + 1) % 3>(g, T, p, job, rx[1], ry[1], ra[1], np[1]);
    }
    const uint32_t multiMask = __ballot_sync(SB_FULL, multi);
    legacy |= multiMask;
    if (multi)
        np[0] = np[1] = 0;
    // ---- the virtual list: rays of slot 0 by lane, then those of slot 1 ----
    const uint32_t incl0 = warp_incl_add(np[0], lane);
    const uint32_t total0 = __shfl_sync(SB_FULL, incl0, 31);
    const uint32_t incl1 = warp_incl_add(np[1], lane) + total0;
    const uint32_t total = __shfl_sync(SB_FULL, incl1, 31);
    const uint32_t pref[2] = {incl0 - np[0], incl1 - np[1]};
    W.ray[lane] = make_uint4(rx[0], ry[0], (ra[0] & 1u) ? pref[0] : 0xffffffffu, 0u);
    W.ray[32 + lane] = make_uint4(rx[1], ry[1], (ra[1] & 1u) ? pref[1] : 0xffffffffu, 0u);
    W.rayBase[lane] = (ra[0] >> 1) - pref[0];
    W.rayBase[32 + lane] = (ra[1] >> 1) - pref[1];
    W.vote[lane] = 0;
    W.vote[32 + lane] = 0;
    const uint32_t ltmask = lanemask_lt(), lelane = ltmask | (1u << lane);
    const uint4 *__restrict__ pairs = reinterpret_cast<const uint4 *>(T.refs);
    uint32_t carry = 0; // entries of the ray that straddles the window border, kept at the front of the pool
    uint32_t longRay = 0xffu, longCount = 0; // a ray with more than 32 entries under way, its distinct keys so far
    for (uint32_t w0 = 0; w0 < total; w0 += WINP) {
        const uint32_t wn = min((uint32_t)WINP, total - w0), w1 = w0 + wn;
        // ---- owners ----
        reinterpret_cast<uint2 *>(W.own)[lane] = make_uint2(0u, 0u);
        __syncwarp();
        uint32_t cont = 0xffu; // the ray whose list goes on in the next window
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (np[s]) {
                const uint32_t b = pref[s], e = pref[s] + np[s];
                if (b >= w0 && b < w1)
                    W.own[b - w0] = (uint8_t)(32 * s + lane + 1);
                else if (b < w0 && e > w0)
                    W.own[0] = (uint8_t)(32 * s + lane + 1);
                if (b < w1 && e > w1)
                    cont = 32 * s + lane;
            }
        }
        cont = __reduce_min_sync(SB_FULL, cont); // at most one lane has one
        __syncwarp();
        {
            uint2 v = reinterpret_cast<uint2 *>(W.own)[lane];
            v.x = prefix_max_bytes(v.x);
            v.y = __vmaxu4(prefix_max_bytes(v.y), (v.x >> 24) * 0x01010101u);
            uint32_t m = v.y >> 24; // running maximum over the lanes before this one
#pragma unroll
            for (int d = 1; d < 32; d <<= 1)
                m = max(m, __shfl_up_sync(SB_FULL, m, d));
            m = __shfl_up_sync(SB_FULL, m, 1);
            if (lane == 0)
                m = 0;
            m *= 0x01010101u;
            reinterpret_cast<uint2 *>(W.own)[lane] = make_uint2(__vmaxu4(v.x, m), __vmaxu4(v.y, m));
        }
        __syncwarp();
        // ---- walk ---- (four steps' loads in flight before the first test)
        uint32_t pos = carry;
        for (uint32_t i0 = 0; i0 < wn; i0 += 128) {
            uint32_t rr[4];
            uint4 qq[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t gi = i0 + 32 * k + lane;
                rr[k] = gi < wn ? (uint32_t)W.own[gi] - 1u : 64u;
                qq[k] = make_uint4(0u, 0u, 0u, 0u);
                if (gi < wn)
                    qq[k] = __ldg(pairs + (W.rayBase[rr[k]] + w0 + gi));
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (i0 + 32 * k >= wn)
                    break;
                const uint32_t r = rr[k];
                const bool valid = r < 64u;
                const uint4 R = W.ray[valid ? r : 0u];
                const uint4 q = qq[k];
                const CellRay rq = {R.x, R.y};
                // the first pair of a list that starts on an odd slot begins with another cell's padding
                const bool skip0 = R.z == w0 + i0 + 32 * k + lane;
                const bool m0 = valid && !skip0 && cell_ref_match(rq, make_uint2(q.x, q.y));
                const bool m1 = valid && cell_ref_match(rq, make_uint2(q.z, q.w));
                const uint32_t b0 = __ballot_sync(SB_FULL, m0), b1 = __ballot_sync(SB_FULL, m1);
                uint32_t at = pos + __popc(b0 & ltmask) + __popc(b1 & ltmask);
                if (m0) {
                    const uint32_t id = cell_ref_id(make_uint2(q.x, q.y));
                    W.tri[at] = id;
                    W.owner[at] = (uint8_t)r;
                    prefetch_record(T, id);
                    ++at;
                }
                if (m1) {
                    const uint32_t id = cell_ref_id(make_uint2(q.z, q.w));
                    W.tri[at] = id;
                    W.owner[at] = (uint8_t)r;
                    prefetch_record(T, id);
                }
                pos += __popc(b0) + __popc(b1);
            }
        }
        __syncwarp();
        // ---- eval ----
        uint32_t S = 0;
        carry = 0;
        const bool lastWindow = w0 + WINP >= total;
        while (S < pos) {
            if (SB_CLS2_FULLCHUNKS && !lastWindow && pos - S < 32u) {
                // not enough for a full step: these entries wait at the front of the pool for the next window's
                carry = pos - S;
                break;
            }
#if SB_CLS2_OPEN
            // A step takes the next 32 entries whatever rays they belong to.  The ray of its last lanes may go on behind the
            // step (in the pool, or -- `cont` -- in the next window): such an OPEN ray has no vote yet, its distinct keys are
            // collected in W.list (longRay / longCount) and the step that holds its last entries adds its own to them.  An open
            // ray that turns out to have no further entries is closed by the next step (or at the end of the round).
            const uint32_t e = S + lane;
            const uint32_t avail = min(32u, pos - S);
            const bool have = (uint32_t)lane < avail;
            const uint32_t o = have ? (uint32_t)W.owner[e] : 0x100u;
            const uint32_t onext = S + 32 < pos ? (uint32_t)W.owner[S + 32] : (avail == 32u ? cont : 0xffu);
            const uint32_t rid = have ? o : 0x200u + lane;
            const uint32_t ridPrev = __shfl_up_sync(SB_FULL, rid, 1);
            const uint32_t heads = __ballot_sync(SB_FULL, have && (lane == 0 || rid != ridPrev));
            const uint32_t openMask = __ballot_sync(SB_FULL, have && o == onext); // a suffix of the step (or nothing)
            const uint32_t rid0 = __shfl_sync(SB_FULL, rid, 0);
            if (longRay != 0xffu && rid0 != longRay) {
                if (lane == 0) // the open ray had no further entries: what was collected is its answer
                    W.vote[longRay] = (uint8_t)(longCount & 1u);
                longRay = 0xffu;
            }
            bool h = false, isCand = false;
            long long k0 = 0, k1 = 0, k2 = 0;
            if (have) {
                const int ow = rid & 31;
                const d3 pp = {W.px[ow], W.py[ow], W.pz[ow]};
                h = eval_entry<true>(T, pp, axis0 + (int)(rid >> 5), W.tri[e], k0, k1, k2, isCand);
                if (h) {
                    W.key[lane][0] = k0;
                    W.key[lane][1] = k1;
                    W.key[lane][2] = k2;
                }
            }
            __syncwarp();
            const uint32_t hm = __ballot_sync(SB_FULL, h);
            // std::set<PositionKey>: a hit whose key equals that of an earlier hit of the ray does not count
            const int segStart = 31 - __clz(heads & lelane);
            uint32_t prior = h ? (hm & ltmask & ~((1u << segStart) - 1u)) : 0u;
            while (prior) {
                const int qq = __ffs(prior) - 1;
                prior &= prior - 1;
                if (W.key[qq][0] == k0 && W.key[qq][1] == k1 && W.key[qq][2] == k2) {
                    h = false;
                    break;
                }
            }
            if (h && rid == longRay) // ... nor does one that equals a key of the ray's earlier steps
                for (uint32_t t = 0; t < min(longCount, (uint32_t)KL); ++t)
                    if (W.list[t][0] == k0 && W.list[t][1] == k1 && W.list[t][2] == k2) {
                        h = false;
                        break;
                    }
            const uint32_t dm = __ballot_sync(SB_FULL, h), cm = __ballot_sync(SB_FULL, isCand);
            // closed segments: the vote of their ray
            if (have && ((heads >> lane) & 1u) && !((openMask >> lane) & 1u)) {
                const uint32_t after = heads & ~lelane;
                const uint32_t segEnd = after ? (uint32_t)__ffs(after) - 1u : avail;
                const uint32_t segMask = (segEnd >= 32 ? 0xffffffffu : (1u << segEnd) - 1u) & ~ltmask;
                const uint32_t before = rid == longRay ? longCount : 0u;
                // parity of the distinct crossings | exact candidates of this step (<= 32) << 1
                W.vote[rid] = (uint8_t)(((__popc(dm & segMask) + before) & 1) | (__popc(cm & segMask) << 1));
            }
            if (!(openMask & 1u) && rid0 == longRay)
                longRay = 0xffu; // its last entries were the first segment of this step
            if (openMask) {
                __syncwarp(); // (the list of a ray closed in this step was read above; a ray opening here writes it from slot 0)
                if (longRay != onext) { // a ray opens here
                    longRay = onext;
                    longCount = 0;
                }
                const uint32_t mine = dm & openMask;
                const uint32_t at = longCount + __popc(mine & ltmask);
                if (h && ((openMask >> lane) & 1u) && at < (uint32_t)KL) {
                    W.list[at][0] = k0;
                    W.list[at][1] = k1;
                    W.list[at][2] = k2;
                }
                longCount += __popc(mine);
                if (longCount > (uint32_t)KL) // more distinct crossings than the list holds: the general kernel takes the point
                    legacy |= 1u << (longRay & 31u);
                if ((uint32_t)lane == (longRay & 31u))
                    exact += __popc(cm & openMask);
            }
            __syncwarp();
            S += avail;
#else
            const uint32_t e = S + lane;
            const bool inPool = e < pos;
            const uint32_t o = inPool ? (uint32_t)W.owner[e] : 0x100u;
            const uint32_t onext = S + 32 < pos ? (uint32_t)W.owner[S + 32] : cont;
            // lanes of the ray that goes on behind this chunk (they end the chunk)
            const uint32_t tail = __ballot_sync(SB_FULL, inPool && o == onext);
            uint32_t cnt = min(32u, pos - S) - __popc(tail);
            bool longChunk = false;
            if (cnt == 0) {
                if (S + 32 >= pos) { // the rest of the pool is the straddling ray's: hold it back
                    carry = pos - S;
                    break;
                }
                // 32 entries of ONE ray with more to come: evaluated as they are; the ray's distinct keys
                // are collected in W.list until the chunk that holds its last entries
                longChunk = true;
                cnt = 32;
                if (longRay != onext) {
                    longRay = onext;
                    longCount = 0;
                }
            }
            const bool have = (uint32_t)lane < cnt;
            const uint32_t rid = have ? o : 0x200u + lane;
            const uint32_t ridPrev = __shfl_up_sync(SB_FULL, rid, 1);
            const uint32_t heads = __ballot_sync(SB_FULL, have && (lane == 0 || rid != ridPrev));
            bool h = false, isCand = false;
            long long k0 = 0, k1 = 0, k2 = 0;
            if (have) {
                const int ow = rid & 31;
                const d3 pp = {W.px[ow], W.py[ow], W.pz[ow]};
                h = eval_entry<true>(T, pp, axis0 + (int)(rid >> 5), W.tri[e], k0, k1, k2, isCand);
                if (h) {
                    W.key[lane][0] = k0;
                    W.key[lane][1] = k1;
                    W.key[lane][2] = k2;
                }
            }
            __syncwarp();
            const uint32_t hm = __ballot_sync(SB_FULL, h);
            // std::set<PositionKey>: a hit whose key equals that of an earlier hit of the ray does not count
            const int segStart = 31 - __clz(heads & lelane);
            uint32_t prior = h ? (hm & ltmask & ~((1u << segStart) - 1u)) : 0u;
            while (prior) {
                const int qq = __ffs(prior) - 1;
                prior &= prior - 1;
                if (W.key[qq][0] == k0 && W.key[qq][1] == k1 && W.key[qq][2] == k2) {
                    h = false;
                    break;
                }
            }
            if (h && rid == longRay) // ... nor does one that equals a key of the ray's earlier chunks
                for (uint32_t t = 0; t < min(longCount, (uint32_t)KL); ++t)
                    if (W.list[t][0] == k0 && W.list[t][1] == k1 && W.list[t][2] == k2) {
                        h = false;
                        break;
                    }
            const uint32_t dm = __ballot_sync(SB_FULL, h), cm = __ballot_sync(SB_FULL, isCand);
            if (longChunk) {
                const uint32_t at = longCount + __popc(dm & ltmask);
                if (h && at < (uint32_t)KL) {
                    W.list[at][0] = k0;
                    W.list[at][1] = k1;
                    W.list[at][2] = k2;
                }
                longCount += __popc(dm);
                if (longCount > (uint32_t)KL) // more distinct crossings than the list holds: the general kernel takes the point
                    legacy |= 1u << (longRay & 31u);
                if ((uint32_t)lane == (longRay & 31u))
                    exact += __popc(cm);
            } else {
                if (have && ((heads >> lane) & 1u)) {
                    const uint32_t after = heads & ~lelane;
                    const uint32_t segEnd = after ? (uint32_t)__ffs(after) - 1u : cnt;
                    const uint32_t segMask = (segEnd >= 32 ? 0xffffffffu : (1u << segEnd) - 1u) & ~ltmask;
                    const uint32_t before = rid == longRay ? longCount : 0u;
                    // parity of the distinct crossings | exact candidates of this chunk (<= 32) << 1
                    W.vote[rid] = (uint8_t)(((__popc(dm & segMask) + before) & 1) | (__popc(cm & segMask) << 1));
                }
                if (__shfl_sync(SB_FULL, rid, 0) == longRay)
                    longRay = 0xffu; // its last entries were the first segment of this chunk
            }
            __syncwarp();
            S += cnt;
#endif
        }
        if (carry) { // at most 32 entries (cnt == 0 of a chunk that reaches the end of the pool)
            const bool mv = (uint32_t)lane < carry;
            const uint32_t t = mv ? W.tri[S + lane] : 0u;
            const uint8_t ow = mv ? W.owner[S + lane] : (uint8_t)0;
            __syncwarp();
            if (mv) {
                W.tri[lane] = t;
                W.owner[lane] = ow;
            }
        }
        __syncwarp();
    }
#if SB_CLS2_OPEN
    if (longRay != 0xffu && lane == 0) // still open at the end of the round: no further entries came
        W.vote[longRay] = (uint8_t)(longCount & 1u);
#endif
    __syncwarp();
    const uint32_t v0 = W.vote[lane], v1 = W.vote[32 + lane];
    exact += (v0 >> 1) + (v1 >> 1); // of this lane's own rays: dropped again if the point goes to the general kernel
    return (v0 & 1u) | ((v1 & 1u) << 1);
}

__global__ void __launch_bounds__(CT2, SB_CLS2_MINB) classify2_kernel(const __grid_constant__ Query q, const __grid_constant__ Target T,
    const __grid_constant__ Out o)
{
    __shared__ GridParams g;
    __shared__ Stage2 s_stage[CW2];
    if (threadIdx.x < sizeof(GridParams) / 4)
        reinterpret_cast<uint32_t *>(&g)[threadIdx.x] = reinterpret_cast<const uint32_t *>(T.gp)[threadIdx.x];
    __syncthreads();
    const int lane = threadIdx.x & 31;
    Stage2 &W = s_stage[threadIdx.x >> 5];
    const uint32_t j = q.first + blockIdx.x * CT2 + threadIdx.x;

    d3 p = {0, 0, 0};
    uint32_t outIndex = 0, job = 0;
    bool active = j < q.count;
    if (active) {
        const uint32_t idx = q.begin + j;
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
    // a point with an infinite or NaN coordinate goes to the general kernel: the rounds below form the ray box
    // as lo = p, hi = p + axis vector, which needs |p| < DBL_MAX (ray_box in sb_classify.cuh)
    const bool finite = fabs(p.x) < DBL_MAX && fabs(p.y) < DBL_MAX && fabs(p.z) < DBL_MAX;
    uint32_t votes = 0, exact = 0, legacy = __ballot_sync(SB_FULL, active && !finite);
    bool undecided = false, deferred = false;
    for (int round = 0; round < 2; ++round) {
        bool want = active && !((legacy >> lane) & 1u);
        if (round == 1 && !o.perAxis) {
            // lazy majority: the third ray only where the first two disagree
            undecided = want && (((votes >> 1) ^ votes) & 1u);
            want = undecided;
            if (T.naxes < 3) { // no third grid yet: the host has it built and launches again for these points
                deferred = true;
                break;
            }
        }
        if (!__any_sync(SB_FULL, want))
            continue;
        if (round == 0)
            votes |= trace_round2<0, 2>(g, T, W, want, lane, job, exact, legacy);
        else
            votes |= trace_round2<2, 1>(g, T, W, want, lane, job, exact, legacy) << 2;
    }
    const bool mine = (legacy >> lane) & 1u; // the general kernel classifies this point (and counts its candidates)
    if (mine) {
        undecided = false;
        exact = 0;
    }
    if (active && !mine) {
        bool in;
        if (o.perAxis) {
            o.perAxis[3 * (size_t)outIndex] = votes & 1u;
            o.perAxis[3 * (size_t)outIndex + 1] = (votes >> 1) & 1u;
            o.perAxis[3 * (size_t)outIndex + 2] = (votes >> 2) & 1u;
            in = __popc(votes) >= 2; // (float)insideCount / totalCount > 0.5 (:508)
        } else {
            in = undecided ? ((votes >> 2) & 1u) != 0 : (votes & 1u) != 0;
        }
        if (!(deferred && undecided))
            o.inside[outIndex] = in ? 1 : 0;
    }
    const uint32_t ex = __reduce_add_sync(SB_FULL, exact);
    const uint32_t um = __ballot_sync(SB_FULL, undecided);
    const uint32_t lm = __ballot_sync(SB_FULL, active && mine);
    unsigned int ubase = 0, lbase = 0;
    if (lane == 0) {
        if (ex)
            atomicAdd(o.exactCount, (unsigned long long)ex);
        if (um)
            ubase = atomicAdd(o.undecidedCount, (unsigned int)__popc(um));
        if (lm)
            lbase = atomicAdd(o.legacyCount, (unsigned int)__popc(lm));
    }
    if (deferred && um) {
        ubase = __shfl_sync(SB_FULL, ubase, 0);
        if (undecided)
            o.undecidedList[ubase + __popc(um & lanemask_lt())] = j;
    }
    if (lm) {
        lbase = __shfl_sync(SB_FULL, lbase, 0);
        if (active && mine)
            o.legacyList[lbase + __popc(lm & lanemask_lt())] = j;
    }
}

} // namespace

cudaError_t sbk_classify2(cudaStream_t s, const MeshDev &target, const ClassifyArgs &a, unsigned long long *exactCount,
    unsigned int *undecidedCount, unsigned int *legacyCount, uint32_t *legacyList, LaunchCounter &lc)
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
    q.count = a.end - a.begin;
    q.list = nullptr;
    q.triJob = qm ? qm->triJob : nullptr;
    q.origFace = qm ? qm->origFace : nullptr;
    q.ownMode = qm && qm->ownFilter ? (qm->ownClosed ? 2 : 1) : 0;
    q.ownLo = qm ? qm->ownLo : 0.0;
    q.ownHi = qm ? qm->ownHi : 0.0;
    q.rawTri = a.rawFaces && qm ? qm->tri : nullptr;
    q.rawXyz = a.rawFaces && qm ? qm->xyz : nullptr;
    q.rawNV = qm ? qm->nV : 0;
    q.first = a.first;
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
    T.bigN0 = T.bigN1 = T.bigN2 = 0;
    T.naxes = target.gridAxes;
    T.vtx = target.vtx;
    T.tri = target.tri;
    T.nrm4 = target.nrm4;
    Out o = {};
    o.inside = a.inside;
    o.perAxis = a.perAxis;
    o.exactCount = exactCount;
    o.undecidedCount = undecidedCount;
    o.undecidedList = a.undecidedList;
    o.legacyCount = legacyCount;
    o.legacyList = legacyList;
    classify2_kernel<<<(q.count - q.first + CT2 - 1) / CT2, CT2, 0, s>>>(q, T, o);
    lc.kernels += 1;
    return cudaGetLastError();
}
