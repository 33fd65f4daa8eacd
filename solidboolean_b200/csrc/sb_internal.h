// Internal host-side interface between the C ABI (sb_capi.cu) and the kernel
// translation units.  Not installed; include/solidboolean_b200.h is the public ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "sb_common.cuh"

// The balanced classifier keeps its reads inside the allocations whatever the cell lists hold (what the optimistic
// enqueue of sb_capi.cu's front end relies on); 0 = without those two checks (dev: their cost), no optimistic enqueue.
#ifndef SB_CLS_GUARDS
#define SB_CLS_GUARDS 1
#endif

// Cluster size K of the LBVH (leaves are K Morton-consecutive triangles).
#ifndef SB_CLUSTER
#define SB_CLUSTER 8
#endif

// Axis-projected grids (sb_grid.cu).  One 15-bit quantiser per world axis (sb_gridq.cuh);
// grid a (rays along axis a) bins the two perpendicular dimensions u, v.
struct GridParams {
    double org[3], scl[3]; // q(x) = clamp(floor((x - org) * scl), 0, 32767)
    double lo[3], hi[3];   // mesh bounds
    int shiftU[3], shiftV[3]; // cell = q >> shift
    uint32_t nu[3];        // cells along u
    uint32_t cellBase[3];  // first entry of axis a in E (a multiple of 16)
    uint32_t totalCells;   // entries of all three grids (cells x depth slabs, each grid rounded up to a multiple of 16)
    uint32_t usedCells;    // ... of the grids actually binned (axes 0 .. gridAxes-1): what the clear and the scan cover
    uint32_t slabBits[3];  // log2 of the depth slabs per cell: the list of a cell is ordered by the slab of each
                           // triangle's FAR bound along the ray axis, so a ray starting at depth aA walks only the
                           // sub-lists of the slabs >= slab(aA) -- a suffix of the cell's list (sb_gridq.cuh)
    uint32_t latShift;     // 0: plain mesh.  Batch mesh (sb_batch_upload): quantised coordinates are job-local in the bits
                           // below latShift (10), the job's lattice position (sb_gridq.cuh lattice3) sits above them
    double qmax;           // clamp of the job-local part: 32767 (plain) or 1023 (batch)
};

// Device-resident mesh (all pointers are device pointers).
struct MeshDev {
    uint32_t nV = 0, nT = 0;
    uint32_t nTpad = 0; // nT rounded up to a multiple of 32
    uint32_t M = 0;     // clusters = ceil(nT / K)
    int sortBeginBit = 0; // Morton bits below this are not sorted (fewer radix passes)
    // geometry as uploaded
    double *xyz = nullptr;   // AoS 3*nV
    uint32_t *tri = nullptr; // 3*nT
    // derived (sb_mesh_build)
    double4 *vtx = nullptr;             // padded vertices
    unsigned long long *bounds = nullptr; // 6 order-encoded doubles: min xyz, max xyz
    double4 *nrm4 = nullptr;            // nT, original order: unit normal + packed vertex indices (sb_common.cuh pack_tri_idx)
    double *scent = nullptr;            // 3*nTpad face centroids ((v0+v1)+v2)/3.0 in Morton order (classification queries)
    uint32_t *mkey = nullptr, *mkeyTmp = nullptr;   // Morton keys (sort ping-pong)
    uint32_t *order = nullptr, *orderTmp = nullptr; // triangle ids (sort ping-pong)
    uint32_t *sortedKey = nullptr;      // -> mkey or mkeyTmp after the sort
    uint32_t *sortedTri = nullptr;      // -> order or orderTmp: sorted position -> triangle id
    Rec32 *leaf = nullptr;              // nTpad sorted-triangle records (float box + id)
    double2 *sbox = nullptr;            // 3*nTpad exact boxes in sorted order
    uint4 *qbox = nullptr;              // nTpad quantised boxes (sb_gridq.cuh) + triangle id, sorted order
    Rec32 *cbox = nullptr;              // M cluster boxes
    uint32_t *ckey = nullptr;           // M cluster keys
    Rec32 *nodes = nullptr;             // 2*(M-1) child records
    int *slot = nullptr;                // M-1 rendezvous slots
    int *root = nullptr;                // device scalar
    int *err = nullptr;                 // device scalar: 1 = triangle index out of range
    // ray grids
    unsigned long long *extentSum = nullptr; // 32 x 3 partial sums of triangle-box extents (2^-24 of the mesh extent)
    uint32_t gridCellBits = 0;          // at most 2^bits cells per axis (allocation bound)
    int gridAxes = 2;                   // grids actually binned: axes 0..gridAxes-1 (the third one only once a vote needed it)
    GridParams *gridParams = nullptr;
    uint32_t *gridE = nullptr;          // totalCells + 2: entry e = refs[E[e+1] .. E[e+2]); &gridE[1] is 16-byte aligned
    uint2 *gridRefs = nullptr;          // cell_ref_pack (sb_gridq.cuh): cell-relative quantised box + triangle id
    uint32_t gridRefCap = 0;            // entries allocated behind gridRefs
    uint4 *gridBigRefs = nullptr;       // 3 * gridBigCap: triangles covering too many cells (grid_ref_pack, absolute)
    uint32_t *gridBigCount = nullptr;   // [0..2] counts (count pass), [3..5] fill cursors, [6] total refs
    uint32_t gridBigCap = 0;
    uint32_t gridBigN[3] = {0, 0, 0};   // host copy of the big-list lengths
    // batch mesh (sb_batch_upload): many independent small meshes ("jobs") laid end to end.  Triangle and vertex
    // indices are global; the jobs overlap in space, so every spatial structure keeps them apart by a VIRTUAL
    // translation that never touches the doubles: the job number leads the Morton key, the conservative float
    // boxes of the LBVH are shifted by latPitch x lattice3(job), the quantised coordinates of the ray grids by
    // lattice3(job) << 10.  Exact tests read the real coordinates.
    // multi-GPU selection (sb_shard.cu): an ordinary mesh over the triangles a rank needs, sharing the parent's
    // vertices; origFace maps its triangle numbers back; as a QUERY mesh only the faces whose centroid z lies
    // in [ownLo, ownHi) (ownClosed: <= ownHi) take part -- each face of the parent belongs to one rank
    bool sharedVtx = false;             // vtx / bounds belong to the parent (built there): the build skips them
    const uint32_t *origFace = nullptr; // nT, ascending, or null
    bool ownFilter = false;
    double ownLo = 0.0, ownHi = 0.0;
    bool ownClosed = false;
    const uint16_t *triJob = nullptr;   // nT: job of each triangle (original order), or null
    uint32_t nJobs = 0;
    double latPitch = 0.0;
};

// sb_capi.cu: sets the calling thread's sb_last_error() text
void sbi_set_error(const char *msg);

struct LaunchCounter {
    uint64_t kernels = 0;
};

// sb_build.cu
// sbk_build_sort: bounds, per-triangle data, Morton sort; sbk_build_leaves (after
// sbk_grid_prepare): sorted leaves / boxes / centroids / clusters + quantised boxes and
// the per-cell counts of the ray grids
cudaError_t sbk_build_sort(cudaStream_t s, MeshDev &m, uint32_t *radixWs, int smCount, LaunchCounter &lc, bool prepared = false);
// the head of sbk_build_sort in pieces (bounds + padded vertices / per-triangle kernel over a range / digit offsets), for a
// mesh whose arrays are still arriving from the host
bool sbk_prep_supported(const MeshDev &m);
cudaError_t sbk_prep_begin(cudaStream_t s, MeshDev &m, uint32_t *radixWs, int smCount, LaunchCounter &lc);
cudaError_t sbk_prep_triangles(cudaStream_t s, MeshDev &m, uint32_t *radixWs, uint32_t first, uint32_t end, int smCount, LaunchCounter &lc);
cudaError_t sbk_prep_end(cudaStream_t s, MeshDev &m, uint32_t *radixWs, LaunchCounter &lc);
cudaError_t sbk_build_leaves(cudaStream_t s, MeshDev &m, LaunchCounter &lc);
cudaError_t sbk_build_tree(cudaStream_t s, MeshDev &m, LaunchCounter &lc);
cudaError_t sbk_batch_fixup(cudaStream_t s, uint32_t *tri, uint32_t nT, const uint32_t *triStart, const uint32_t *vtxStart,
    uint32_t nJobs, uint16_t *triJob, int *err, LaunchCounter &lc);
cudaError_t sbk_triangle_boxes(cudaStream_t s, const MeshDev &m, double2 *out /* 3*nT, original order */, LaunchCounter &lc);

// sb_broad.cu -- candidate keys: (((a << bitsB) | b) << 2), code bits zero
cudaError_t sbk_broad_phase(cudaStream_t s, const MeshDev &A, const MeshDev &B, uint32_t groupBegin, uint32_t groupEnd,
    unsigned bitsB, unsigned long long *outKeys, unsigned long long capacity, unsigned long long *outCount,
    int *errFlag, LaunchCounter &lc);

// sb_narrow.cu
cudaError_t sbk_predicate(cudaStream_t s, const MeshDev &A, const MeshDev &B, unsigned long long *keys, uint32_t nPairs,
    unsigned bitsB, unsigned long long *hitKeys, uint32_t *hitSlot, double2 *hitSeg, unsigned int *hitCount,
    uint8_t *flagsA, uint8_t *flagsB, unsigned long long *pathCounts, uint8_t *hitTag /* per slot, or null */, LaunchCounter &lc);
cudaError_t sbk_fp64_peak(cudaStream_t s, int smCount, double *scratch, int iters, bool fma, unsigned long long *flops);
cudaError_t sbk_gather_hits(cudaStream_t s, const unsigned long long *sortedHitKeys, const uint32_t *sortedSlot,
    const double2 *hitSeg, uint32_t nHits, unsigned bitsB, uint32_t *outAB, double2 *outSeg, const uint8_t *hitTag, uint8_t *outTag,
    LaunchCounter &lc);
cudaError_t sbk_decode_candidates(cudaStream_t s, const unsigned long long *keys, uint32_t n, unsigned bitsB,
    uint32_t *outAB, uint8_t *outCode, LaunchCounter &lc);
cudaError_t sbk_tri_tri_batch(cudaStream_t s, const double *tris18, uint32_t n, int32_t *ret, int32_t *coplanar,
    double *seg6, LaunchCounter &lc);
cudaError_t sbk_sort_keys(cudaStream_t s, unsigned long long *keys, unsigned long long *keysTmp, uint32_t *vals,
    uint32_t *valsTmp, size_t n, int beginBit, int endBit, uint32_t *radixWs, int smCount,
    unsigned long long **outKeys, uint32_t **outVals, LaunchCounter &lc);

// sb_halfedge.cu -- uncut triangles + half-edge map (SolidBoolean::addUnintersectedTriangles)
uint32_t sbk_uncut_tiles(uint32_t nT);
cudaError_t sbk_uncut_count(cudaStream_t s, const uint8_t *cut /* nT bytes or null */, uint32_t nT, uint32_t *tileScratch,
    uint32_t *total, LaunchCounter &lc);
// bitsO > 0: ordinal packed into the low bits of the key word (keys-only sort on the bits above), ords unused;
// bitsO == 0: separate ordinal array (key, value sort)
cudaError_t sbk_uncut_emit(cudaStream_t s, const uint8_t *cut, const uint32_t *tri, uint32_t nT, uint32_t nV, int *err,
    const uint32_t *tileScratch, uint32_t vertexOffset, unsigned bitsV, unsigned bitsO, uint32_t *face, uint32_t *tri3,
    unsigned long long *keys, uint32_t *ords, LaunchCounter &lc);
cudaError_t sbk_halfedge_link(cudaStream_t s, const unsigned long long *sortedKeys, const uint32_t *sortedOrds, uint32_t n,
    unsigned bitsV, unsigned bitsO, uint32_t nV, uint32_t *vstart /* nV words of scratch */, uint32_t vertexOffset,
    uint32_t triangleOffset, unsigned long long *refKeys, uint32_t *owner, int32_t *adj, uint32_t *ordOut /* bitsO > 0 */,
    uint32_t *firstRepeat, LaunchCounter &lc);
size_t sbk_uncut_components_scratch(uint32_t nTri, uint32_t nT, bool ordered);
cudaError_t sbk_uncut_components(cudaStream_t s, const int32_t *adj, uint32_t nTri, uint32_t triangleOffset,
    const uint32_t *sortedTri /* Morton order of the mesh, or null */, const uint32_t *face, uint32_t nT, uint32_t *scratch,
    uint32_t *label, uint32_t *count, LaunchCounter &lc);

// sb_flood.cu -- buildFaceGroups over uncut components + retriangulated pieces
size_t sbk_flood_scratch_words(uint32_t nU, uint32_t nP, uint32_t nF);
cudaError_t sbk_flood(cudaStream_t s, const uint32_t *pieces, uint32_t nP, const uint32_t *fences, uint32_t nF,
    const unsigned long long *uKeys, const uint32_t *uOwner, uint32_t nUK, const uint32_t *uLabel, uint32_t nU, uint32_t triangleOffset,
    unsigned keyBits, uint32_t *scratch, int smCount, uint32_t *labelUncut, uint32_t *labelPiece, unsigned int *nGroups, LaunchCounter &lc);

// sb_cuts.cu -- per-triangle intersection contexts (the pair-loop body of SolidBoolean::combine)
size_t sbk_cut_contexts_scratch(size_t nHits); // 4-byte words
cudaError_t sbk_cut_contexts(cudaStream_t s, const uint32_t *hitAB, const double *seg, uint32_t n, int which, unsigned bitsTri,
    uint32_t *scratch, uint32_t *radixWs, int smCount, uint32_t *cutTri /* n */, uint32_t *pointStart /* n + 1 */,
    double *points /* 6 n */, uint32_t *edgeStart /* n + 1 */, uint32_t *edges /* 2 n */, uint32_t *counts /* dev: 3 */,
    LaunchCounter &lc);

// sb_shard.cu -- multi-GPU selection (every pass covers both parents in one launch)
size_t sbk_shard_tiles(uint32_t nT);
size_t sbk_shard_hist_words();
cudaError_t sbk_shard_plan(cudaStream_t s, const MeshDev &A, const MeshDev &B, float2 *const zf[2] /* per vertex, from sbk_bounds_pad */,
    float2 *const zr[2], uint32_t *const tiles[2], uint32_t *hist, int n, double *cuts /* n + 2 */, int smCount, LaunchCounter &lc);
cudaError_t sbk_shard_count(cudaStream_t s, const MeshDev &A, const MeshDev &B, float2 *const zr[2], uint32_t *const tiles[2],
    const double *cuts, int rank, int n, uint32_t *totals, const uint32_t *expect, uint32_t *mismatch, LaunchCounter &lc);
cudaError_t sbk_shard_emit(cudaStream_t s, const MeshDev &A, const MeshDev &B, float2 *const zr[2], uint32_t *const tiles[2],
    const double *cuts, int rank, int n, const uint32_t cap[2], uint32_t *const outTri[2], uint32_t *const outFace[2],
    LaunchCounter &lc);
cudaError_t sbk_shard_remap_hits(cudaStream_t s, uint32_t *ab, uint32_t n, const uint32_t *faceA, const uint32_t *faceB, LaunchCounter &lc);
cudaError_t sbk_bounds_pad(cudaStream_t s, MeshDev &m, int smCount, LaunchCounter &lc, float2 *zf = nullptr /* nV, optional */);

// sb_classify.cu
struct ClassifyArgs {
    // query points: either explicit (pts != null, AoS 3*Q, processed in given order)
    // or the face centroids of `queryMesh` at sorted positions [begin, end)
    const double *pts = nullptr;
    const MeshDev *queryMesh = nullptr;
    uint32_t begin = 0, end = 0; // point range (explicit) or sorted-position range (faces)
    uint8_t *inside = nullptr;   // indexed by point index / original triangle id
    uint8_t *perAxis = nullptr;  // optional, 3 per point (needs all three grids of the target)
    // target with two grids only: points whose two votes disagree are appended (as indices
    // relative to `begin`) to undecidedList; a second launch with list / listCount /
    // thirdAxisOnly set traces their third ray once the target's third grid exists
    uint32_t *undecidedList = nullptr;
    const uint32_t *list = nullptr;
    uint32_t listCount = 0;
    bool thirdAxisOnly = false;
    // raw faces: the query mesh's faces in ORIGINAL order from its uploaded arrays (it need not be built); the launch
    // covers faces [first, end) with begin = 0 -- one launch per upload chunk, lists and counters shared
    bool rawFaces = false;
    uint32_t first = 0;
};
// One launch classifies all points (sb_classify.cu).  Lazy vote when only `inside` is
// wanted: axes 0 and 1 for every point, axis 2 where the two disagree
// (*undecidedCount counts those points); all three axes when perAxis is set.
// bigKeys / bigCap: scratch (3 x bigCap int64) for rays with many quantised matches;
// *bigNeeded (device, zeroed by the caller) receives the entries they asked for -- if
// it exceeds bigCap the results are invalid and the call must be repeated with a
// scratch of that size.  *exactCount accumulates the true candidates (exact box overlap).
cudaError_t sbk_classify(cudaStream_t s, const MeshDev &target, const ClassifyArgs &a, long long *bigKeys,
    unsigned long long bigCap, unsigned long long *bigNeeded, unsigned long long *exactCount, unsigned int *undecidedCount,
    uint32_t poolLimit /* 0 = default; tests lower it to reach the general path */,
    unsigned long long *trace /* dev: 4 words per CTA, or null */, LaunchCounter &lc);
uint32_t sbk_classify_blocks(uint32_t points);
// The balanced single-walk kernel (sb_classify2.cu) for targets without per-axis big lists and
// launches without a point list: same outputs; points it leaves to the general kernel (a ray box
// over several cells, more than 32 matches on a ray) are counted in *legacyCount and listed
// (relative to a.begin) in legacyList, for a second launch of sbk_classify over that list.
cudaError_t sbk_classify2(cudaStream_t s, const MeshDev &target, const ClassifyArgs &a, unsigned long long *exactCount,
    unsigned int *undecidedCount, unsigned int *legacyCount, uint32_t *legacyList, LaunchCounter &lc);

size_t sbk_radix_workspace_words(size_t n);

// sb_grid.cu
size_t sbk_grid_entry_bound(uint32_t gridCellBits);   // entries of E the allocation has to hold (+ 2 closing words)
size_t sbk_grid_scan_status_words(uint32_t gridCellBits);
cudaError_t sbk_grid_prepare(cudaStream_t s, MeshDev &m, uint32_t *scanScratch, float beta, int slabBitsMax, LaunchCounter &lc);
cudaError_t sbk_grid_scan(cudaStream_t s, MeshDev &m, uint32_t *scanScratch, LaunchCounter &lc);
// counts again from the stored quantised boxes, for m.gridAxes axes (a mesh built with two grids gets its third)
cudaError_t sbk_grid_recount(cudaStream_t s, MeshDev &m, uint32_t *scanScratch, LaunchCounter &lc);
cudaError_t sbk_grid_fill(cudaStream_t s, MeshDev &m, LaunchCounter &lc);
