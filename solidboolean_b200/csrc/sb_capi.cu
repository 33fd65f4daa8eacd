// C ABI (include/solidboolean_b200.h): contexts, device memory, stage timing
// and the orchestration of the kernel stages.  No CPU fallback anywhere: every
// compute entry point needs a CUDA device and fails with SB_ERR_CUDA otherwise.
#include "../../include/solidboolean_b200.h"
#include "sb_internal.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

} // namespace

// (for the translation units that sit on top of the ABI: sb_comm.cu)
void sbi_set_error(const char *msg) { snprintf(g_err, sizeof(g_err), "%s", msg ? msg : ""); }

namespace {

#define SB_CUDA(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(SB_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

unsigned bits_for(size_t n)
{
    unsigned b = 1;
    while (((size_t)1 << b) < n)
        ++b;
    return b;
}

constexpr size_t SB_BATCH_MAX_JOBS_HOST = 1024; // = SB_BATCH_MAX_JOBS of sb_gridq.cuh (the 32 x 32 job lattice)

struct DeviceScalars { // device counters of one lane (128-byte slots)
    unsigned long long pairCount;
    unsigned long long stats[2];
    unsigned int hitCount;
    unsigned int overflowCount;
    int err;
    int pad;
    unsigned long long paths[5]; // predicate exit histogram (TriTriPath)
    unsigned int uncutTotal;     // sb_*_uncut: uncut faces
    unsigned int heRepeat;       // ... first refused half-edge insertion (ordinal), UINT_MAX = none
    unsigned int ccCount;        // sb_uncut_components: components
    unsigned int cuts[3];        // sb_isect_contexts: contexts, points, relations
    unsigned int legacyCount;    // classification: points the balanced kernel left to the general one
};
static_assert(sizeof(DeviceScalars) <= 128, "lane slot too small");

} // namespace

struct sb_context {
    std::recursive_mutex mu; // held by every entry point that works on this context (DeviceGuard)
    int device = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr;
    LaunchCounter lc;
    // stage timing
    bool timing = false;
    bool earlyLeaf = false;
    // Destroyed plain meshes are PARKED (device arena, streams, events, captured rebuild graphs and all) and handed out again
    // by the next sb_mesh_upload of the same counts -- a caller that makes a new SolidMesh per operation then pays what
    // sb_mesh_update pays: the copies and a full rebuild, no allocation, no stream creation, no first-build round trip (the
    // reference lists keep their size and are verified against the new geometry, see sb_mesh_update).  At most
    // SB_MESH_CACHE meshes (default 4, 0 = off) and 4 GiB of device memory are kept; sb_context_destroy releases them.
    std::vector<sb_mesh *> parked;
    size_t parkedBytes = 0;
    int parkMax = 4;
    struct Span {
        int stage;
        cudaEvent_t a, b;
    };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> freeEvents;
    float acc[SB_STAGE_COUNT] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaEvent_t t0 = nullptr;      // timing reference (recorded at reset)
    cudaEvent_t orderEvent = nullptr; // orders per-mesh streams behind the context stream
    // scratch
    uint32_t *radixWs = nullptr;
    size_t radixWsWords = 0;
    DeviceScalars *dScalars = nullptr;
    DeviceScalars *hScalars = nullptr; // pinned mirror
    uint32_t *hPool = nullptr;         // pinned: 256 slots of 4 words for per-mesh count read-backs
    unsigned hPoolNext = 0;
    uint8_t *classifyOut = nullptr;    // inside flags (+ per-axis) of the last classify call
    size_t classifyOutBytes = 0;
    uint32_t *overflowList = nullptr;
    uint32_t overflowCap = 0;
    uint64_t lastRays = 0, lastCands = 0;
    // "lanes": stream + device/host counter block.  Lane 0 is the context stream; lanes 1
    // and 2 let the two classification directions of sb_front_end overlap the
    // intersection stages.
    struct Lane {
        cudaStream_t stream = nullptr;
        DeviceScalars *d = nullptr, *h = nullptr;
        cudaEvent_t done = nullptr;
        unsigned long long bigHint = 0; // scratch entries the many-layer rays of the last call needed
    } lanes[3];
    uint32_t *scanScratch = nullptr; // grid-build scan status words
    size_t scanScratchWords = 0;
    void *feFlags = nullptr;         // sb_front_end_host: device copy of the two flag arrays (grow-only)
    size_t feFlagsBytes = 0;
    uint64_t uploadSeq = 0;          // counts the host uploads (which of two fresh meshes arrives last)
    bool earlyPrep = false;          // SB_EARLY_PREP=1: the head of a mesh's build (bounds, padded vertices, per-triangle kernel) runs beside its
                                     // upload, chunk by chunk.  Measured at C3: 1.87 against 1.89 ms per host-buffer step (the 35 us it hides are
                                     // partly paid back in stream waits); with the stage events of sb_context_enable_timing on the streams the first
                                     // mesh's build is then held back until the second mesh's upload is over (2.13 ms) -- off
    size_t earlyPrepMin = 1u << 16;  // ... meshes below this many triangles always do (SB_EARLY_PREP_MIN; tests lower it)
    bool optimisticVerify = true;    // SB_OPTIMISTIC=0: a front end waits for the rebuilds' reference counts before it enqueues anything
    bool deferVerify = false;        // ... set while such a front end enqueues its work (mesh_finish then leaves the check alone)
    uint64_t optimisticRedone = 0;   // front ends that had to be repeated because a rebuild did not fit its lists
    bool streamClassify = false;     // SB_STREAM_CLASSIFY=1: chunk-by-chunk classification of a just-uploaded mesh (measured at
                                     // C3: 1.97 against 1.98 ms per host-buffer step -- the GPU is busy either way; off)
    float gridBeta = 1.0f;           // ray-grid cell size / mean triangle-box extent (SB_GRID_BETA)
    int gridSlabBits = 2;            // at most 2^this depth slabs per ray-grid cell (SB_GRID_SLABS = the bits; 0: one list per cell)
    int sortBeginBit = -1;           // lowest Morton bit that is sorted (SB_SORT_BEGIN_BIT); -1 = by mesh size
    uint32_t classifyPoolLimit = 0;  // SB_CLASSIFY_POOL_LIMIT: rays with more matches take the general path (tests)
    std::vector<void *> shardPinned;  // pinned read-back blocks of destroyed shards, reused (cudaMallocHost is slow)
    uint64_t shardUndecided = 0;     // points of a multi-GPU selection whose first two votes disagreed (need the whole target)
    bool classifyBalanced = true;    // SB_CLASSIFY_V2=0: every launch uses the general kernel (sb_classify.cu)
    bool useGraphs = true;           // SB_GRAPHS=0: rebuilds enqueue their kernels one by one
    size_t grid3EagerBelow = 65536;  // SB_GRID3_EAGER_BELOW: meshes with fewer triangles get their third ray grid right away
                                     // (launch-bound sizes: binning it costs nothing, building it later costs host round trips)
};

struct sb_mesh {
    sb_context *ctx = nullptr;
    MeshDev d;
    void *arena = nullptr;
    void *gridArena = nullptr; // references, sized after the count pass
    size_t gridArenaBytes = 0;
    bool built = false;
    // builds run on the mesh's own stream so that independent meshes overlap;
    // consumers on the context stream wait for `ready`
    cudaStream_t stream = nullptr;
    cudaEvent_t ready = nullptr;     // everything built (incl. the ray grids)
    cudaStream_t treeStream = nullptr; // the (latency-bound) LBVH kernel runs here, beside the grid scan / fill
    cudaEvent_t leavesDone = nullptr;  // leaf kernel finished (orders treeStream behind the mesh stream)
    // sb_mesh_update put new geometry into a mesh whose reference list was sized for the old one: the next build
    // fills within the old capacity (bounds-guarded) and sends the new counts to the host; the first use of the
    // mesh reads them (mesh_finish) -- new big-list lengths, and a proper first build if the capacity was exceeded
    bool geomChanged = false, verifyPending = false;
    cudaEvent_t verifyEv = nullptr;
    size_t arenaBytes = 0;           // size of `arena` (accounting of the parked meshes)
    cudaEvent_t leafReady = nullptr; // sorted leaves / boxes / centroids: all a QUERY mesh needs,
                                     // recorded before the grids (and the LBVH) are built
    uint32_t *radixWs = nullptr;     // in the arena
    uint32_t *scanScratch = nullptr; // in the arena
    uint32_t *hCounts = nullptr;     // pinned: [0] total refs, [1..3] big-list lengths
    uint32_t *hErr = nullptr;        // pinned: index-validation flag read back with the counts
    bool gridSized = false;          // reference list already sized by an earlier build
    bool gridPending = false;        // first build: counts on their way to the host, list not yet sized / filled
                                     // (mesh_finish completes it at the first use, so that the host does not
                                     // wait for one mesh before it has enqueued the next one's build)
    bool grid3Wanted = false;        // a vote (or a per-axis query) needed the third grid: builds include it from now on
    // A REbuild (same immutable geometry, reference list already sized) is a fixed sequence of
    // ~20 launches and memsets on two streams: captured once, replayed as one CUDA graph.
    cudaGraphExec_t buildGraph = nullptr;  // sort .. leaves (then leafReady is recorded)
    cudaGraphExec_t gridGraph = nullptr;   // {grid scan -> fill} beside {LBVH}
    unsigned graphSig = 0;           // what the capture depended on: grids, LBVH wanted, sorted bits
    uint64_t graphKernels = 0;       // kernels in the graph (launch accounting)
    std::vector<uint32_t> jobTriStart, jobVtxStart; // batch mesh (sb_batch_upload): n_jobs + 1 each, else empty
    uint32_t *dJobStart = nullptr;   // ... and their device copy: [triangle starts][vertex starts] (in the arena)
    bool treeBuilt = false;          // LBVH topology built (lazily, on first use as a traversal target)
    bool treeWanted = false;         // the mesh has been a traversal target: rebuilds include the LBVH
    // Geometry that has only just been sent from host memory (sb_mesh_upload / sb_mesh_update): the index triples go up
    // in chunks with an event behind each, and the first front end after it classifies this mesh's faces in their
    // ORIGINAL order chunk by chunk as they arrive (needs the other mesh's grids, not this mesh's build) -- that
    // direction then runs beside the rest of the transfer and this mesh's build instead of after them.
    static constexpr int MAX_CHUNKS = 4;
    cudaEvent_t chunkEv[MAX_CHUNKS] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t chunkEnd[MAX_CHUNKS] = {0, 0, 0, 0};
    int nChunks = 0;
    // ... and the head of this mesh's build runs while they arrive, on treeStream: bounds + padded vertices behind the
    // coordinates, the per-triangle kernel (normals, Morton keys, digit counts) behind each chunk of index triples
    // (sb_build.cu sbk_prep_*); the next sb_mesh_build waits for prepEv and goes on from the sort passes
    bool prepared = false;
    cudaEvent_t xyzEv = nullptr, prepEv = nullptr;
    bool fresh = false;              // uploaded from the host since the last front end
    uint64_t uploadSeq = 0;          // order of the uploads within the context
};

struct sb_isect {
    sb_context *ctx = nullptr;
    const sb_mesh *A = nullptr, *B = nullptr;
    unsigned bitsA = 1, bitsB = 1;
    size_t nCand = 0, nHit = 0;
    unsigned long long *candKeys = nullptr; // final (sorted) candidate keys
    void *candAlloc[2] = {nullptr, nullptr};
    uint32_t *hitAB = nullptr;
    double2 *hitSeg = nullptr;
    uint8_t *hitTag = nullptr;     // nHit: which edges the segment end points lie on (sb_tritri.cuh tt_segment_tag)
    uint8_t *flagsA = nullptr, *flagsB = nullptr;
    unsigned long long paths[5] = {0, 0, 0, 0, 0}; // predicate exit histogram
    std::vector<void *> owned; // stream-ordered allocations to release
    bool candSorted = false;   // candidates are ordered lazily, when somebody asks for them
    bool noSort = false;
};

// Result of sb_*_uncut (sb_halfedge.cu): all arrays device-resident, sized for the nTri uncut faces.
struct sb_uncut {
    sb_context *ctx = nullptr;
    const sb_mesh *mesh = nullptr;
    uint32_t nTri = 0;            // uncut faces
    uint32_t vertexOffset = 0, triangleOffset = 0;
    uint32_t repeatOrd = 0xffffffffu; // ordinal (3 * rank + edge) of the first refused insertion
    uint32_t *face = nullptr;     // nTri
    uint32_t *tri3 = nullptr;     // 3 nTri shifted index triples
    unsigned long long *keys = nullptr; // 3 nTri reference-format keys, ascending
    uint32_t *owner = nullptr;    // 3 nTri
    uint32_t *ords = nullptr;     // 3 nTri insertion ordinal of each sorted entry
    int32_t *adj = nullptr;       // 3 nTri, indexed by ordinal
    uint32_t *label = nullptr;    // nTri component labels (sb_uncut_components, on first request)
    size_t nComponents = 0;
    std::vector<void *> owned;
};

// Result of sb_isect_contexts (sb_cuts.cu), device-resident.
struct sb_cuts {
    sb_context *ctx = nullptr;
    size_t nCtx = 0, nPoints = 0, nEdges = 0;
    uint32_t *tri = nullptr, *pointStart = nullptr, *edgeStart = nullptr, *edges = nullptr;
    double *points = nullptr;
    std::vector<void *> owned;
};

namespace {

// Every entry point starts with one of these: the context's device becomes current and -- the reference
// allows concurrent SolidBoolean objects over shared const SolidMesh from several threads (SURVEY 8b,
// "Threading") -- the context is locked for the duration of the call.  A context owns ONE stream, one block
// of pinned read-back scalars, one radix workspace: calls on the same context from several host threads are
// serialised here (recursive: entry points call each other); different contexts run side by side.
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    std::recursive_mutex *mu = nullptr;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess)
            ok = true;
    }
    explicit DeviceGuard(const sb_context *c) : mu(&const_cast<sb_context *>(c)->mu)
    {
        mu->lock();
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(c->device) == cudaSuccess)
            ok = true;
    }
    ~DeviceGuard()
    {
        if (prev >= 0)
            cudaSetDevice(prev);
        if (mu)
            mu->unlock();
    }
};

struct StageTimer {
    sb_context *c;
    sb_context::Span span;
    bool on;
    cudaStream_t st;
    StageTimer(sb_context *ctx, int stage, cudaStream_t stream = nullptr) : c(ctx), on(ctx->timing), st(stream ? stream : ctx->stream)
    {
        if (!on)
            return;
        span.stage = stage;
        span.a = grab();
        span.b = grab();
        cudaEventRecord(span.a, st);
    }
    ~StageTimer()
    {
        if (!on)
            return;
        cudaEventRecord(span.b, st);
        c->spans.push_back(span);
    }
    cudaEvent_t grab()
    {
        if (!c->freeEvents.empty()) {
            cudaEvent_t e = c->freeEvents.back();
            c->freeEvents.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};

int ensure_radix_ws(sb_context *c, size_t n)
{
    size_t words = sbk_radix_workspace_words(n);
    if (words <= c->radixWsWords)
        return SB_OK;
    if (c->radixWs) {
        SB_CUDA(cudaStreamSynchronize(c->stream));
        SB_CUDA(cudaFree(c->radixWs));
        c->radixWs = nullptr;
        c->radixWsWords = 0;
    }
    words = words + words / 4;
    SB_CUDA(cudaMalloc(&c->radixWs, words * sizeof(uint32_t)));
    c->radixWsWords = words;
    return SB_OK;
}

int ensure_classify_out(sb_context *c, size_t bytes, uint32_t overflowCap)
{
    if (bytes > c->classifyOutBytes) {
        if (c->classifyOut) {
            SB_CUDA(cudaStreamSynchronize(c->stream));
            SB_CUDA(cudaFree(c->classifyOut));
            c->classifyOut = nullptr;
        }
        SB_CUDA(cudaMalloc(&c->classifyOut, bytes));
        c->classifyOutBytes = bytes;
    }
    if (overflowCap > c->overflowCap) {
        if (c->overflowList) {
            SB_CUDA(cudaStreamSynchronize(c->stream));
            SB_CUDA(cudaFree(c->overflowList));
            c->overflowList = nullptr;
        }
        SB_CUDA(cudaMalloc(&c->overflowList, sizeof(uint32_t) * overflowCap));
        c->overflowCap = overflowCap;
    }
    return SB_OK;
}

// make the context stream wait for the mesh's last build
inline void use_mesh(sb_context *c, const sb_mesh *m)
{
    if (m->ready)
        cudaStreamWaitEvent(c->stream, m->ready, 0);
}
// ... or only for the part of it the intersection reads (leaves, boxes, LBVH)
inline void use_mesh_leaves(sb_context *c, const sb_mesh *m)
{
    if (m->leafReady)
        cudaStreamWaitEvent(c->stream, m->leafReady, 0);
}

// make the mesh stream wait for everything enqueued so far on the context stream
inline void order_after_context(sb_context *c, const sb_mesh *m)
{
    cudaEventRecord(c->orderEvent, c->stream);
    cudaStreamWaitEvent(m->stream, c->orderEvent, 0);
}

template <typename T>
int alloc_async(sb_context *c, T **p, size_t count, std::vector<void *> *owned)
{
    void *q = nullptr;
    SB_CUDA(cudaMallocAsync(&q, std::max<size_t>(count, 1) * sizeof(T), c->stream));
    *p = static_cast<T *>(q);
    if (owned)
        owned->push_back(q);
    return SB_OK;
}

// releases a stream-ordered scratch allocation on every way out of a function
struct ScratchGuard {
    void *p;
    cudaStream_t s;
    ~ScratchGuard()
    {
        if (p)
            cudaFreeAsync(p, s);
    }
};

int mesh_alloc(sb_context *ctx, size_t nV, size_t nT, sb_mesh **out, size_t nJobs = 0, const sb_mesh *vertexParent = nullptr)
{
    if (nV >= (1ull << 31) || nT >= (1ull << SB_MAX_TRIANGLE_BITS))
        return fail(SB_ERR_INVALID, "mesh too large: %zu vertices, %zu triangles", nV, nT);
    sb_mesh *m = new (std::nothrow) sb_mesh;
    if (!m)
        return fail(SB_ERR_NOMEM, "out of host memory");
    m->ctx = ctx;
    m->grid3Wanted = nT < ctx->grid3EagerBelow;
    MeshDev &d = m->d;
    d.nV = (uint32_t)nV;
    d.nT = (uint32_t)nT;
    d.nTpad = (uint32_t)((nT + 31) / 32 * 32);
    d.M = (uint32_t)((nT + SB_CLUSTER - 1) / SB_CLUSTER);
    size_t nI = d.M > 1 ? d.M - 1 : 0;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += align256(std::max<size_t>(bytes, 16));
        return o;
    };
    // (a multi-GPU selection reads the vertices, padded vertices and bounds of its parent)
    size_t oXyz = take(vertexParent ? 0 : 24 * nV), oTri = take(12 * nT), oVtx = take(vertexParent ? 0 : 32 * nV), oBounds = take(48);
    size_t oNormal = take(32 * nT), oScent = take(24 * (size_t)((nT + 31) / 32 * 32));
    size_t oKey = take(4 * nT), oKeyT = take(4 * nT), oOrd = take(4 * nT), oOrdT = take(4 * nT);
    size_t oLeaf = take(32 * (size_t)d.nTpad), oSbox = take(48 * (size_t)d.nTpad), oQbox = take(16 * (size_t)d.nTpad);
    size_t oCbox = take(32 * (size_t)d.M), oCkey = take(4 * (size_t)d.M + 4);
    size_t oNodes = take(64 * nI), oSlot = take(4 * nI), oRoot = take(8);
    {
        // allocation bound: at most ~2 nT cells per axis (the device picks the
        // actual resolution from the mean triangle extent, sb_grid.cu)
        double lg = nT ? std::log2((double)nT) + 1.0 : 0.0;
        int bits = (int)std::floor(lg + 0.5);
        if (nJobs) // batch mesh: every job has its own block of cells on the 32 x 32 lattice, most of them sparsely used
            bits = std::max(bits + 2, 12);
        d.gridCellBits = (uint32_t)std::max(0, std::min(bits, 26));
    }
    size_t oTriJob = take(nJobs ? 2 * nT : 0), oJobStart = take(nJobs ? 8 * (nJobs + 1) : 0);
    size_t oRadix = take(4 * sbk_radix_workspace_words(nT));
    size_t oScan = take(4 * sbk_grid_scan_status_words(d.gridCellBits));
    size_t oGridP = take(sizeof(GridParams)), oGridE = take(4 * (sbk_grid_entry_bound(d.gridCellBits) + 8)), oGridBig = take(32 + 96 * 8);
    // stream-ordered allocation: the pool keeps the block cached between calls
    m->arenaBytes = off;
    cudaError_t e = cudaMallocAsync(&m->arena, off, ctx->stream);
    if (e != cudaSuccess) {
        delete m;
        return fail(e == cudaErrorMemoryAllocation ? SB_ERR_NOMEM : SB_ERR_CUDA, "cudaMallocAsync(%zu): %s", off,
            cudaGetErrorString(e));
    }
    char *b = static_cast<char *>(m->arena);
    d.xyz = (double *)(b + oXyz);
    d.tri = (uint32_t *)(b + oTri);
    d.vtx = (double4 *)(b + oVtx);
    d.bounds = (unsigned long long *)(b + oBounds);
    d.nrm4 = (double4 *)(b + oNormal);
    d.scent = (double *)(b + oScent);
    d.mkey = (uint32_t *)(b + oKey);
    d.mkeyTmp = (uint32_t *)(b + oKeyT);
    d.order = (uint32_t *)(b + oOrd);
    d.orderTmp = (uint32_t *)(b + oOrdT);
    d.leaf = (Rec32 *)(b + oLeaf);
    d.sbox = (double2 *)(b + oSbox);
    d.qbox = (uint4 *)(b + oQbox);
    d.cbox = (Rec32 *)(b + oCbox);
    d.ckey = (uint32_t *)(b + oCkey);
    d.nodes = (Rec32 *)(b + oNodes);
    d.slot = (int *)(b + oSlot);
    d.root = (int *)(b + oRoot);
    d.err = d.root + 1;
    d.gridParams = (GridParams *)(b + oGridP);
    d.gridE = (uint32_t *)(b + oGridE) + 3; // &gridE[1] is 16-byte aligned: the scan and the clear work on E + 1 with 128-bit accesses
    d.gridBigCount = (uint32_t *)(b + oGridBig);
    d.extentSum = (unsigned long long *)(d.gridBigCount + 8);
    if (vertexParent) {
        d.xyz = vertexParent->d.xyz;
        d.vtx = vertexParent->d.vtx;
        d.bounds = vertexParent->d.bounds;
        d.sharedVtx = true;
        m->grid3Wanted = false; // its third ray would need the whole parent (sb_shard.cu)
    }
    if (nJobs) {
        d.triJob = (const uint16_t *)(b + oTriJob);
        d.nJobs = (uint32_t)nJobs;
        m->dJobStart = (uint32_t *)(b + oJobStart);
    }
    m->radixWs = (uint32_t *)(b + oRadix);
    m->scanScratch = (uint32_t *)(b + oScan);
    // builds are on the critical path of everything that follows: highest priority, so
    // that work started early on the context / lane streams only fills their gaps
    int prioLow = 0, prioHigh = 0;
    cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh);
    if (cudaStreamCreateWithPriority(&m->stream, cudaStreamNonBlocking, prioHigh) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithPriority(&m->treeStream, cudaStreamNonBlocking, prioHigh) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->leavesDone, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->verifyEv, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&m->leafReady, cudaEventDisableTiming) != cudaSuccess) {
        cudaFreeAsync(m->arena, ctx->stream);
        delete m;
        return fail(SB_ERR_CUDA, "stream/event creation failed");
    }
    m->hCounts = ctx->hPool + 8 * (ctx->hPoolNext++ % 256);
    m->hErr = m->hCounts + 4;
    order_after_context(ctx, m); // the arena was allocated in context-stream order
    *out = m;
    return SB_OK;
}

} // namespace

extern "C" {

const char *sb_last_error(void) { return g_err; }

const char *sb_version(void) { return "solidboolean_b200 0.1 sm_100a"; }

int sb_context_create(int device, sb_context **out)
{
    if (!out)
        return fail(SB_ERR_INVALID, "out is null");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(SB_ERR_CUDA, "no CUDA device available (%s); this library has no CPU fallback",
            e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= count)
        return fail(SB_ERR_INVALID, "device %d out of range [0,%d)", device, count);
    DeviceGuard g(device);
    if (!g.ok)
        return fail(SB_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    sb_context *c = new (std::nothrow) sb_context;
    if (!c)
        return fail(SB_ERR_NOMEM, "out of host memory");
    c->device = device;
    // (everything below that can fail runs inside `init`: a failure releases what was created so far)
    auto init = [&]() -> int {
    cudaDeviceProp prop;
    SB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->smCount = prop.multiProcessorCount;
    // Priorities: the intersection stages on the context stream are short and the host waits
    // for their counts twice, the two classification launches are long and saturate every
    // SM: with equal priorities the broad phase only got its CTAs once the classifiers had
    // dispatched all of theirs, and the whole intersection ran as a tail after them.
    int prioLow = 0, prioHigh = 0;
    cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh);
    SB_CUDA(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prioHigh));
    SB_CUDA(cudaMalloc(&c->dScalars, 512));
    SB_CUDA(cudaMemset(c->dScalars, 0, 512)); // the counter blocks are read back whole
    SB_CUDA(cudaMallocHost(&c->hScalars, 512));
    SB_CUDA(cudaMallocHost(&c->hPool, 256 * 32));
    for (int l = 0; l < 3; ++l) {
        c->lanes[l].d = reinterpret_cast<DeviceScalars *>(reinterpret_cast<char *>(c->dScalars) + 128 * l);
        c->lanes[l].h = reinterpret_cast<DeviceScalars *>(reinterpret_cast<char *>(c->hScalars) + 128 * l);
        if (l == 0)
            c->lanes[l].stream = c->stream;
        else
            SB_CUDA(cudaStreamCreateWithPriority(&c->lanes[l].stream, cudaStreamNonBlocking, prioLow));
        SB_CUDA(cudaEventCreateWithFlags(&c->lanes[l].done, cudaEventDisableTiming));
    }
    SB_CUDA(cudaEventCreate(&c->t0));
    SB_CUDA(cudaEventCreateWithFlags(&c->orderEvent, cudaEventDisableTiming));
    SB_CUDA(cudaEventRecord(c->t0, c->stream));
    return SB_OK;
    };
    if (int r = init()) {
        char msg[512];
        snprintf(msg, sizeof(msg), "%s", g_err);
        sb_context_destroy(c); // tolerates the members that were never created
        cudaGetLastError();
        snprintf(g_err, sizeof(g_err), "%s", msg);
        return r;
    }
    if (const char *e = getenv("SB_SORT_BEGIN_BIT"))
        c->sortBeginBit = std::max(0, std::min(atoi(e), 24));
    if (const char *e = getenv("SB_GRAPHS"))
        c->useGraphs = atoi(e) != 0;
    if (const char *e = getenv("SB_MESH_CACHE"))
        c->parkMax = std::max(0, std::min(atoi(e), 64));
    if (const char *e = getenv("SB_EARLY_LEAF")) // dev: a mesh's face queries may start once its sorted centroids exist
        c->earlyLeaf = atoi(e) != 0;
    if (const char *e = getenv("SB_GRID3_EAGER_BELOW"))
        c->grid3EagerBelow = (size_t)std::max(0ll, atoll(e));
    if (const char *e = getenv("SB_CLASSIFY_POOL_LIMIT"))
        c->classifyPoolLimit = (uint32_t)std::max(0, atoi(e));
    if (const char *e = getenv("SB_CLASSIFY_V2"))
        c->classifyBalanced = atoi(e) != 0;
    if (const char *e = getenv("SB_EARLY_PREP"))
        c->earlyPrep = atoi(e) != 0;
    if (const char *e = getenv("SB_EARLY_PREP_MIN"))
        c->earlyPrepMin = (size_t)std::max(1ll, atoll(e));
    if (const char *e = getenv("SB_OPTIMISTIC"))
        c->optimisticVerify = atoi(e) != 0;
    if (const char *e = getenv("SB_STREAM_CLASSIFY"))
        c->streamClassify = atoi(e) != 0;
    if (const char *e = getenv("SB_GRID_SLABS"))
        c->gridSlabBits = std::max(0, std::min(atoi(e), 4));
    if (const char *e = getenv("SB_GRID_BETA")) {
        float b = (float)atof(e);
        if (b > 0.01f && b < 100.0f)
            c->gridBeta = b;
    }
    // keep stream-ordered allocations cached in the pool between calls
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    *out = c;
    return SB_OK;
}

static void mesh_destroy_now(sb_mesh *m);

void sb_context_destroy(sb_context *c)
{
    if (!c)
        return;
    DeviceGuard g(c->device); // (not the locking form: the mutex goes away with the context)
    for (sb_mesh *m : c->parked)
        mesh_destroy_now(m);
    c->parked.clear();
    cudaStreamSynchronize(c->stream);
    for (auto &s : c->spans) {
        cudaEventDestroy(s.a);
        cudaEventDestroy(s.b);
    }
    for (auto e : c->freeEvents)
        cudaEventDestroy(e);
    cudaFree(c->radixWs);
    cudaFree(c->dScalars);
    cudaFreeHost(c->hScalars);
    cudaFreeHost(c->hPool);
    for (void *q : c->shardPinned)
        cudaFreeHost(q);
    for (int l = 0; l < 3; ++l) {
        if (l && c->lanes[l].stream) {
            cudaStreamSynchronize(c->lanes[l].stream);
            cudaStreamDestroy(c->lanes[l].stream);
        }
        if (c->lanes[l].done)
            cudaEventDestroy(c->lanes[l].done);
    }
    cudaEventDestroy(c->t0);
    cudaEventDestroy(c->orderEvent);
    cudaFree(c->classifyOut);
    cudaFree(c->overflowList);
    cudaFree(c->scanScratch);
    cudaFree(c->feFlags);
    cudaStreamDestroy(c->stream);
    delete c;
}

int sb_context_synchronize(sb_context *c)
{
    if (!c)
        return fail(SB_ERR_INVALID, "context is null");
    DeviceGuard g(c);
    SB_CUDA(cudaStreamSynchronize(c->stream));
    return SB_OK;
}

void *sb_context_stream(sb_context *c) { return c ? (void *)c->stream : nullptr; }
int sb_context_device(const sb_context *c) { return c ? c->device : -1; }

int sb_context_enable_timing(sb_context *c, int enable)
{
    if (!c)
        return fail(SB_ERR_INVALID, "context is null");
    c->timing = enable != 0;
    return SB_OK;
}

// Spans of one stage may overlap (meshes build concurrently on their own streams):
// a stage's time is the length of the UNION of its spans.
static int drain_spans(sb_context *c)
{
    SB_CUDA(cudaDeviceSynchronize());
    std::vector<std::pair<float, float>> iv[SB_STAGE_COUNT];
    for (auto &s : c->spans) {
        float a = 0, b = 0;
        if (cudaEventElapsedTime(&a, c->t0, s.a) == cudaSuccess && cudaEventElapsedTime(&b, c->t0, s.b) == cudaSuccess) {
            iv[s.stage].push_back({a, b});
            if (getenv("SB_DEBUG_SPANS")) // dev: the timeline of the stage spans since the last reset
                fprintf(stderr, "[sb] span stage %d: %.4f .. %.4f ms\n", s.stage, a, b);
        }
        c->freeEvents.push_back(s.a);
        c->freeEvents.push_back(s.b);
    }
    c->spans.clear();
    for (int st = 0; st < SB_STAGE_COUNT; ++st) {
        std::sort(iv[st].begin(), iv[st].end());
        float curA = 0, curB = -1;
        for (auto &p : iv[st]) {
            if (curB < curA || p.first > curB) {
                if (curB > curA)
                    c->acc[st] += curB - curA;
                curA = p.first;
                curB = p.second;
            } else {
                curB = std::max(curB, p.second);
            }
        }
        if (curB > curA)
            c->acc[st] += curB - curA;
    }
    return SB_OK;
}

int sb_context_reset_timing(sb_context *c)
{
    if (!c)
        return fail(SB_ERR_INVALID, "context is null");
    DeviceGuard g(c);
    int r = drain_spans(c);
    if (r)
        return r;
    for (float &a : c->acc)
        a = 0;
    c->lc.kernels = 0;
    SB_CUDA(cudaEventRecord(c->t0, c->stream));
    return SB_OK;
}

int sb_context_get_timing(sb_context *c, float *ms, uint64_t *kernel_launches)
{
    if (!c)
        return fail(SB_ERR_INVALID, "context is null");
    DeviceGuard g(c);
    int r = drain_spans(c);
    if (r)
        return r;
    if (ms)
        for (int i = 0; i < SB_STAGE_COUNT; ++i)
            ms[i] = c->acc[i];
    if (kernel_launches)
        *kernel_launches = c->lc.kernels;
    return SB_OK;
}

int sb_context_classify_stats(sb_context *c, uint64_t *rays, uint64_t *candidates)
{
    if (!c)
        return fail(SB_ERR_INVALID, "context is null");
    if (rays)
        *rays = c->lastRays;
    if (candidates)
        *candidates = c->lastCands;
    return SB_OK;
}

// ---- mesh ---------------------------------------------------------------------

// how many Morton bits the build sorts (a function of the mesh size and the context's settings)
static void decide_sort_bits(const sb_context *c, sb_mesh *m)
{
    if (c->sortBeginBit >= 0) {
        m->d.sortBeginBit = c->sortBeginBit;
    } else if (m->d.triJob) {
        m->d.sortBeginBit = 8; // batch: 12 job bits + the 12 leading Morton bits = three passes
    } else {
        // The order only has to be spatially coherent: 8 or more Morton cells per triangle are
        // plenty (ties keep their input order), so a 1M-triangle mesh sorts 24 of the 30 bits
        // -- three 8-bit passes instead of four -- and a 5K-triangle mesh two.
        int want = 3;
        while (want < 30 && ((size_t)1 << (want - 3)) < m->d.nT)
            ++want;
        const int passes = std::min(4, (want + 7) / 8);
        m->d.sortBeginBit = std::max(0, 30 - 8 * passes);
    }
}

// host -> device copies of a mesh's two arrays on its stream: coordinates first, then the index triples in up to
// MAX_CHUNKS pieces with an event behind each (see sb_mesh::chunkEv)
static cudaError_t upload_from_host(sb_mesh *m, const void *xyz, const void *tri)
{
    cudaError_t e = cudaSuccess;
    sb_context *c = m->ctx;
    const size_t nV = m->d.nV, nT = m->d.nT;
    m->nChunks = 0;
    m->fresh = false;
    m->prepared = false;
    decide_sort_bits(c, m);
    // the head of the build beside the transfer (both arrays new, a plain mesh of some size)
    bool prep = c->earlyPrep && xyz && tri && nV && nT >= c->earlyPrepMin && sbk_prep_supported(m->d);
    auto ev = [&](cudaEvent_t *p) { return *p ? cudaSuccess : cudaEventCreateWithFlags(p, cudaEventDisableTiming); };
    if (prep) {
        e = ev(&m->xyzEv);
        if (e == cudaSuccess) e = ev(&m->prepEv);
        // treeStream behind whatever the mesh stream still does with the buffers the preparation writes
        if (e == cudaSuccess) e = cudaEventRecord(m->prepEv, m->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(m->treeStream, m->prepEv, 0);
    }
    if (e == cudaSuccess && xyz && nV)
        e = cudaMemcpyAsync(m->d.xyz, xyz, 24 * nV, cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess && prep) {
        e = cudaEventRecord(m->xyzEv, m->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(m->treeStream, m->xyzEv, 0);
        if (e == cudaSuccess) {
            StageTimer t(c, SB_STAGE_BUILD, m->treeStream);
            e = sbk_prep_begin(m->treeStream, m->d, m->radixWs, c->smCount, c->lc);
        }
    }
    if (e != cudaSuccess || !tri || !nT)
        return e;
    int n = nT >= (1u << 18) ? sb_mesh::MAX_CHUNKS : nT >= (1u << 16) ? 2 : 1;
    if (prep && c->earlyPrepMin < (1u << 16) && nT >= 2048)
        n = sb_mesh::MAX_CHUNKS; // (tests: ranged launches on small meshes too)
    const size_t per = (((nT + n - 1) / n) + 1023) / 1024 * 1024;
    size_t done = 0;
    for (int k = 0; k < n && done < nT && e == cudaSuccess; ++k) {
        const size_t end = std::min(nT, done + per);
        e = cudaMemcpyAsync(m->d.tri + 3 * done, static_cast<const uint32_t *>(tri) + 3 * done, 12 * (end - done),
            cudaMemcpyHostToDevice, m->stream);
        if (e == cudaSuccess) e = ev(&m->chunkEv[k]);
        if (e == cudaSuccess)
            e = cudaEventRecord(m->chunkEv[k], m->stream);
        if (e == cudaSuccess && prep) {
            e = cudaStreamWaitEvent(m->treeStream, m->chunkEv[k], 0);
            if (e == cudaSuccess) {
                StageTimer t(c, SB_STAGE_BUILD, m->treeStream);
                e = sbk_prep_triangles(m->treeStream, m->d, m->radixWs, (uint32_t)done, (uint32_t)end, c->smCount, c->lc);
            }
        }
        m->chunkEnd[k] = (uint32_t)end;
        m->nChunks = k + 1;
        done = end;
    }
    if (e == cudaSuccess && prep) {
        e = sbk_prep_end(m->treeStream, m->d, m->radixWs, c->lc);
        if (e == cudaSuccess) e = cudaEventRecord(m->prepEv, m->treeStream);
        m->prepared = e == cudaSuccess;
    }
    if (e == cudaSuccess && xyz) { // (both arrays new: what the faces' centroids are made of is on its way)
        m->fresh = true;
        m->uploadSeq = ++m->ctx->uploadSeq;
    }
    return e;
}

int sb_mesh_upload(sb_context *ctx, const double *xyz, size_t nV, const uint32_t *tri, size_t nT, sb_mesh **out)
{
    if (!ctx || !out)
        return fail(SB_ERR_INVALID, "null context or out");
    *out = nullptr;
    if ((nV && !xyz) || (nT && !tri))
        return fail(SB_ERR_INVALID, "null geometry pointer");
    if (nT && !nV)
        return fail(SB_ERR_INVALID, "triangles without vertices");
    DeviceGuard g(ctx);
    sb_mesh *m = nullptr;
    for (size_t k = 0; k < ctx->parked.size(); ++k)
        if (ctx->parked[k]->d.nV == nV && ctx->parked[k]->d.nT == nT) {
            // a parked mesh of the same counts: new bytes into it, exactly what sb_mesh_update does
            m = ctx->parked[k];
            ctx->parked.erase(ctx->parked.begin() + (std::ptrdiff_t)k);
            ctx->parkedBytes -= m->arenaBytes + m->gridArenaBytes;
            m->built = false;
            m->gridPending = false;
            m->verifyPending = false;
            m->geomChanged = true;
            order_after_context(ctx, m);
            break;
        }
    if (!m) {
        int r = mesh_alloc(ctx, nV, nT, &m);
        if (r)
            return r;
    }
    cudaError_t e = upload_from_host(m, xyz, tri);
    if (e == cudaSuccess)
        e = cudaEventRecord(m->ready, m->stream);
    if (e == cudaSuccess)
        e = cudaEventRecord(m->leafReady, m->stream);
    if (e != cudaSuccess) {
        mesh_destroy_now(m);
        return fail(SB_ERR_CUDA, "upload: %s", cudaGetErrorString(e));
    }
    *out = m;
    return SB_OK;
}

// geometry that is already on the device (e.g. received from another GPU): copied into the mesh on its stream,
// ordered after everything the caller has enqueued on the context stream so far
int sb_mesh_upload_device(sb_context *ctx, const void *d_xyz, size_t nV, const void *d_tri, size_t nT, sb_mesh **out)
{
    if (!ctx || !out)
        return fail(SB_ERR_INVALID, "null context or out");
    *out = nullptr;
    if ((nV && !d_xyz) || (nT && !d_tri))
        return fail(SB_ERR_INVALID, "null geometry pointer");
    if (nT && !nV)
        return fail(SB_ERR_INVALID, "triangles without vertices");
    DeviceGuard g(ctx);
    sb_mesh *m = nullptr;
    int r = mesh_alloc(ctx, nV, nT, &m); // (the mesh stream waits for the context stream here)
    if (r)
        return r;
    cudaError_t e = cudaSuccess;
    if (nV)
        e = cudaMemcpyAsync(m->d.xyz, d_xyz, 24 * nV, cudaMemcpyDeviceToDevice, m->stream);
    if (e == cudaSuccess && nT)
        e = cudaMemcpyAsync(m->d.tri, d_tri, 12 * nT, cudaMemcpyDeviceToDevice, m->stream);
    if (e == cudaSuccess)
        e = cudaEventRecord(m->ready, m->stream);
    if (e == cudaSuccess)
        e = cudaEventRecord(m->leafReady, m->stream);
    if (e != cudaSuccess) {
        sb_mesh_destroy(m);
        return fail(SB_ERR_CUDA, "device upload: %s", cudaGetErrorString(e));
    }
    *out = m;
    return SB_OK;
}

// new coordinates / index triples of the SAME sizes into an existing mesh (a deforming mesh, or the next frame of
// a stream of same-sized inputs): everything built from the old geometry is stale until the next sb_mesh_build
int sb_mesh_update(sb_mesh *m, const void *xyz, const void *tri, int on_device)
{
    if (!m)
        return fail(SB_ERR_INVALID, "mesh is null");
    if (m->d.triJob || m->d.sharedVtx)
        return fail(SB_ERR_INVALID, "only plain meshes can be updated");
    sb_context *c = m->ctx;
    DeviceGuard g(c);
    order_after_context(c, m);
    if (on_device) {
        if (xyz && m->d.nV)
            SB_CUDA(cudaMemcpyAsync(m->d.xyz, xyz, 24 * (size_t)m->d.nV, cudaMemcpyDeviceToDevice, m->stream));
        if (tri && m->d.nT)
            SB_CUDA(cudaMemcpyAsync(m->d.tri, tri, 12 * (size_t)m->d.nT, cudaMemcpyDeviceToDevice, m->stream));
        m->fresh = false;
    } else {
        SB_CUDA(upload_from_host(m, xyz, tri));
    }
    m->built = false;
    m->gridPending = false;
    m->geomChanged = true; // the next build reports its reference counts (mesh_finish checks them against the capacity)
    SB_CUDA(cudaEventRecord(m->ready, m->stream));
    SB_CUDA(cudaEventRecord(m->leafReady, m->stream));
    return SB_OK;
}

// ---- batch meshes ---------------------------------------------------------------------
int sb_batch_upload(sb_context *ctx, size_t n_jobs, const double *xyz, const size_t *vertex_start, const uint32_t *tri,
    const size_t *triangle_start, double lattice_pitch, sb_mesh **out)
{
    if (!ctx || !out)
        return fail(SB_ERR_INVALID, "null context or out");
    *out = nullptr;
    if (!n_jobs || n_jobs > SB_BATCH_MAX_JOBS_HOST)
        return fail(SB_ERR_INVALID, "a batch holds 1 .. %u jobs", (unsigned)SB_BATCH_MAX_JOBS_HOST);
    if (!vertex_start || !triangle_start || vertex_start[0] || triangle_start[0])
        return fail(SB_ERR_INVALID, "job start arrays must begin with 0");
    for (size_t j = 0; j < n_jobs; ++j)
        if (vertex_start[j + 1] < vertex_start[j] || triangle_start[j + 1] < triangle_start[j])
            return fail(SB_ERR_INVALID, "job start arrays must not decrease (job %zu)", j);
    const size_t nV = vertex_start[n_jobs], nT = triangle_start[n_jobs];
    if ((nV && !xyz) || (nT && !tri))
        return fail(SB_ERR_INVALID, "null geometry pointer");
    if (nT && !nV)
        return fail(SB_ERR_INVALID, "triangles without vertices");
    if (!(lattice_pitch > 0.0) || !(lattice_pitch < 1.0e30))
        return fail(SB_ERR_INVALID, "lattice_pitch must be a positive number (>= 4 x the largest |coordinate| of both batches)");
    DeviceGuard g(ctx);
    sb_mesh *m = nullptr;
    int r = mesh_alloc(ctx, nV, nT, &m, n_jobs);
    if (r)
        return r;
    m->d.latPitch = lattice_pitch;
    m->jobTriStart.resize(n_jobs + 1);
    m->jobVtxStart.resize(n_jobs + 1);
    for (size_t j = 0; j <= n_jobs; ++j) {
        m->jobTriStart[j] = (uint32_t)triangle_start[j];
        m->jobVtxStart[j] = (uint32_t)vertex_start[j];
    }
    cudaError_t e = cudaSuccess;
    if (nV)
        e = cudaMemcpyAsync(m->d.xyz, xyz, 24 * nV, cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess && nT)
        e = cudaMemcpyAsync(m->d.tri, tri, 12 * nT, cudaMemcpyHostToDevice, m->stream);
    // (the vectors outlive the copies: the stream is synchronised before the mesh can be destroyed)
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(m->dJobStart, m->jobTriStart.data(), 4 * (n_jobs + 1), cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(m->dJobStart + n_jobs + 1, m->jobVtxStart.data(), 4 * (n_jobs + 1), cudaMemcpyHostToDevice, m->stream);
    if (e == cudaSuccess)
        e = cudaMemsetAsync(m->d.root, 0, 8, m->stream); // clears the error flag the fix-up may raise
    if (e == cudaSuccess && nT)
        e = sbk_batch_fixup(m->stream, m->d.tri, (uint32_t)nT, m->dJobStart, m->dJobStart + n_jobs + 1, (uint32_t)n_jobs,
            const_cast<uint16_t *>(m->d.triJob), m->d.err, ctx->lc);
    if (e == cudaSuccess)
        e = cudaEventRecord(m->ready, m->stream);
    if (e == cudaSuccess)
        e = cudaEventRecord(m->leafReady, m->stream);
    if (e != cudaSuccess) {
        sb_mesh_destroy(m);
        return fail(SB_ERR_CUDA, "batch upload: %s", cudaGetErrorString(e));
    }
    *out = m;
    return SB_OK;
}

int sb_batch_info(const sb_mesh *m, size_t *n_jobs, size_t *vertex_start, size_t *triangle_start)
{
    if (!m)
        return fail(SB_ERR_INVALID, "mesh is null");
    const size_t J = m->d.nJobs;
    if (n_jobs)
        *n_jobs = J;
    for (size_t j = 0; J && j <= J; ++j) {
        if (vertex_start)
            vertex_start[j] = m->jobVtxStart[j];
        if (triangle_start)
            triangle_start[j] = m->jobTriStart[j];
    }
    return SB_OK;
}

// device error flag of a build -> message (bit 0: index check; bit 1: batch lattice pitch)
static int mesh_error(const sb_mesh *m, int err)
{
    if (err & 2)
        return fail(SB_ERR_INVALID, "lattice_pitch %g is too small: it must be at least 4 x the largest |coordinate| of the batch",
            m->d.latPitch);
    if (m->d.triJob)
        return fail(SB_ERR_INVALID, "triangle index out of range (>= the vertex count of its job)");
    return fail(SB_ERR_INVALID, "triangle index out of range (>= %u vertices)", m->d.nV);
}

// First build (or first build with a third grid): the reference list is sized from the
// counts (16-byte read-back on the mesh's stream), then filled.
static int grid_counts_readback(sb_mesh *m)
{
    cudaStream_t st = m->stream;
    uint32_t *h = m->hCounts;
    SB_CUDA(cudaMemcpyAsync(h, m->d.gridBigCount + 6, 4, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpyAsync(h + 1, m->d.gridBigCount, 12, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpyAsync(m->hErr, m->d.err, sizeof(int), cudaMemcpyDeviceToHost, st));
    return SB_OK;
}

static int grid_size_and_fill(sb_mesh *m, bool readback = true)
{
    sb_context *c = m->ctx;
    cudaStream_t st = m->stream;
    uint32_t *h = m->hCounts;
    if (readback) {
        int r = grid_counts_readback(m);
        if (r)
            return r;
    }
    int *hErr = reinterpret_cast<int *>(m->hErr);
    SB_CUDA(cudaStreamSynchronize(st));
    if (*hErr)
        return mesh_error(m, *hErr);
    size_t nRefs = h[0];
    uint32_t bigMax = std::max(h[1], std::max(h[2], h[3]));
    for (int k = 0; k < 3; ++k)
        m->d.gridBigN[k] = h[1 + k];
    // [8-byte references: nRefs + 8 (the classifier reads whole groups)][16-byte big lists: 3 x bigMax]
    const size_t refBytes = align256(8 * (std::max<size_t>(nRefs, 1) + 8));
    size_t bytes = refBytes + 16 * 3 * (size_t)std::max<uint32_t>(bigMax, 1);
    if (m->gridArena)
        cudaFreeAsync(m->gridArena, st);
    if (m->buildGraph) { // the captured rebuild has the old lists' addresses and capacities in its kernel arguments
        cudaGraphExecDestroy(m->buildGraph);
        cudaGraphExecDestroy(m->gridGraph);
        m->buildGraph = m->gridGraph = nullptr;
    }
    SB_CUDA(cudaMallocAsync(&m->gridArena, bytes, st));
    m->gridArenaBytes = bytes;
    m->d.gridRefs = static_cast<uint2 *>(m->gridArena);
    m->d.gridRefCap = (uint32_t)std::max<size_t>(nRefs, 1);
    m->d.gridBigRefs = reinterpret_cast<uint4 *>(static_cast<char *>(m->gridArena) + refBytes);
    m->d.gridBigCap = std::max<uint32_t>(bigMax, 1);
    m->gridSized = true;
    StageTimer t(c, SB_STAGE_BUILD, st);
    SB_CUDA(sbk_grid_fill(st, m->d, c->lc));
    return SB_OK;
}

// A rebuild after sb_mesh_update: did the new geometry fit the reference lists sized for the old one?  Waits for the
// counts the build sent to the host.  *changed: the lists overflowed (the mesh is then built again with lists of the
// right size) or the per-axis big lists have other lengths than the host believed -- whatever was enqueued against
// the mesh's grids before this check must be repeated.
static int mesh_finish(const sb_mesh *mc);
static int mesh_verify(sb_mesh *m, bool *changed)
{
    if (changed)
        *changed = false;
    if (!m || !m->verifyPending)
        return SB_OK;
    m->verifyPending = false;
    DeviceGuard g(m->ctx);
    SB_CUDA(cudaEventSynchronize(m->verifyEv));
    if (*reinterpret_cast<int *>(m->hErr))
        return mesh_error(m, *reinterpret_cast<int *>(m->hErr));
    const uint32_t *h = m->hCounts;
    if (h[0] > m->d.gridRefCap || std::max(h[1], std::max(h[2], h[3])) > m->d.gridBigCap) {
        m->gridSized = false; // no: a first build sizes them again (counts on their way, finished by mesh_finish)
        if (changed)
            *changed = true;
        int rb = sb_mesh_build(m);
        if (rb)
            return rb;
    } else {
        for (int k = 0; k < 3; ++k) {
            if (changed && m->d.gridBigN[k] != h[1 + k])
                *changed = true;
            m->d.gridBigN[k] = h[1 + k];
        }
    }
    return SB_OK;
}

// the same question without side effects (the counts are waited for): would mesh_verify report a change?
static bool mesh_verify_would_change(const sb_mesh *m)
{
    if (!m || !m->verifyPending)
        return false;
    DeviceGuard g(m->ctx);
    if (cudaEventSynchronize(m->verifyEv) != cudaSuccess || *reinterpret_cast<const int *>(m->hErr))
        return true;
    const uint32_t *h = m->hCounts;
    if (h[0] > m->d.gridRefCap || std::max(h[1], std::max(h[2], h[3])) > m->d.gridBigCap)
        return true;
    for (int k = 0; k < 3; ++k)
        if (m->d.gridBigN[k] != h[1 + k])
            return true;
    return false;
}

// Completes a first build whose reference list is still to be sized and filled.
static int mesh_finish(const sb_mesh *mc)
{
    sb_mesh *m = const_cast<sb_mesh *>(mc);
    if (m && m->verifyPending && !m->ctx->deferVerify) {
        int rv = mesh_verify(m, nullptr);
        if (rv)
            return rv;
    }
    if (!m || !m->gridPending)
        return SB_OK;
    m->gridPending = false;
    DeviceGuard g(m->ctx);
    int r = grid_size_and_fill(m, false);
    if (r) {
        m->built = false;
        return r;
    }
    cudaStream_t st = m->stream;
    if (m->treeBuilt && m->d.nT)
        SB_CUDA(cudaStreamWaitEvent(st, m->leavesDone, 0)); // `ready` covers the LBVH too
    SB_CUDA(cudaEventRecord(m->ready, st));
    return SB_OK;
}

// The third ray grid (rays along z) is only binned once something needs it: a point
// whose first two votes disagree, or a per-axis query.  The mesh's grids are then built
// again from the stored quantised boxes with all three axes (and every later
// sb_mesh_build includes the third one).  Nothing may be reading the grids meanwhile.
static int ensure_grid3(const sb_mesh *mc)
{
    sb_mesh *m = const_cast<sb_mesh *>(mc);
    if (m->d.gridAxes == 3 || !m->d.nT)
        return SB_OK;
    {
        int rf = mesh_finish(m);
        if (rf)
            return rf;
    }
    sb_context *c = m->ctx;
    for (int l = 0; l < 3; ++l)
        SB_CUDA(cudaStreamSynchronize(c->lanes[l].stream));
    cudaStream_t st = m->stream;
    m->d.gridAxes = 3;
    m->grid3Wanted = true;
    m->gridSized = false;
    {
        StageTimer t(c, SB_STAGE_BUILD, st);
        SB_CUDA(sbk_grid_recount(st, m->d, m->scanScratch, c->lc));
        SB_CUDA(sbk_grid_scan(st, m->d, m->scanScratch, c->lc));
    }
    int r = grid_size_and_fill(m);
    if (r)
        return r;
    SB_CUDA(cudaEventRecord(m->ready, st));
    SB_CUDA(cudaStreamSynchronize(st));
    return SB_OK;
}

int sb_mesh_build(sb_mesh *m)
{
    if (!m)
        return fail(SB_ERR_INVALID, "mesh is null");
    sb_context *c = m->ctx;
    DeviceGuard g(c);
    cudaStream_t st = m->stream;
    m->gridPending = false; // an unfinished first build is simply redone
    decide_sort_bits(c, m);
    const bool prepared = m->prepared; // the head of the build has run beside the upload (upload_from_host)
    m->prepared = false;
    if (m->d.sharedVtx)
        m->grid3Wanted = false;
    if ((m->d.gridAxes == 3) != m->grid3Wanted)
        m->gridSized = false; // the reference list was sized for another number of grids
    m->d.gridAxes = m->grid3Wanted ? 3 : 2;
    // after whatever the context stream still does with this mesh's buffers
    order_after_context(c, m);
    if (c->useGraphs && m->d.nT && m->gridSized) {
        const unsigned sig = (unsigned)m->d.gridAxes | (m->treeWanted ? 4u : 0u) | ((unsigned)m->d.sortBeginBit << 3) | (prepared ? 0x10000u : 0u);
        if (m->buildGraph && m->graphSig != sig) {
            cudaGraphExecDestroy(m->buildGraph);
            cudaGraphExecDestroy(m->gridGraph);
            m->buildGraph = m->gridGraph = nullptr;
        }
        if (!m->buildGraph) {
            const uint64_t k0 = c->lc.kernels;
            // two graphs, so that the event the classification queries of OTHER streams wait
            // for (sorted centroids ready) can be recorded between them
            auto capture = [&](cudaGraphExec_t *exec, auto &&enqueue) -> int {
                SB_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeRelaxed));
                cudaError_t e = enqueue();
                cudaGraph_t graph = nullptr;
                cudaError_t e2 = cudaStreamEndCapture(st, &graph);
                if (e != cudaSuccess || e2 != cudaSuccess || !graph) {
                    if (graph)
                        cudaGraphDestroy(graph);
                    return fail(SB_ERR_CUDA, "build graph capture: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
                }
                e = cudaGraphInstantiate(exec, graph, 0);
                cudaGraphDestroy(graph);
                if (e != cudaSuccess) {
                    *exec = nullptr;
                    return fail(SB_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
                }
                return SB_OK;
            };
            int r = capture(&m->buildGraph, [&]() {
                cudaError_t e = cudaSuccess;
                if (!prepared) // (prepared: cleared before the per-triangle kernel raised its index-check flag)
                    e = cudaMemsetAsync(m->d.root, 0, m->d.triJob ? 4 : 8, st); // (batch: the upload's index check stays)
                if (e == cudaSuccess) e = sbk_build_sort(st, m->d, m->radixWs, c->smCount, c->lc, prepared);
                if (e == cudaSuccess) e = sbk_grid_prepare(st, m->d, m->scanScratch, c->gridBeta, c->gridSlabBits, c->lc);
                if (e == cudaSuccess) e = sbk_build_leaves(st, m->d, c->lc);
                return e;
            });
            if (!r)
                r = capture(&m->gridGraph, [&]() {
                    cudaError_t e = cudaSuccess;
                    if (m->treeWanted) { // fork: the LBVH beside the grid scan / fill
                        e = cudaEventRecord(m->leavesDone, st);
                        if (e == cudaSuccess) e = cudaStreamWaitEvent(m->treeStream, m->leavesDone, 0);
                        if (e == cudaSuccess) e = sbk_build_tree(m->treeStream, m->d, c->lc);
                        if (e == cudaSuccess) e = cudaEventRecord(m->leavesDone, m->treeStream);
                    }
                    if (e == cudaSuccess) e = sbk_grid_scan(st, m->d, m->scanScratch, c->lc);
                    if (e == cudaSuccess) e = sbk_grid_fill(st, m->d, c->lc);
                    if (e == cudaSuccess && m->treeWanted)
                        e = cudaStreamWaitEvent(st, m->leavesDone, 0); // join
                    return e;
                });
            m->graphKernels = c->lc.kernels - k0;
            c->lc.kernels = k0;
            if (r) {
                if (m->buildGraph)
                    cudaGraphExecDestroy(m->buildGraph);
                m->buildGraph = m->gridGraph = nullptr;
                return r;
            }
            m->graphSig = sig;
        }
        if (prepared)
            SB_CUDA(cudaStreamWaitEvent(st, m->prepEv, 0));
        {
            StageTimer t(c, SB_STAGE_BUILD, st);
            SB_CUDA(cudaGraphLaunch(m->buildGraph, st));
            if (c->earlyLeaf)
                SB_CUDA(cudaEventRecord(m->leafReady, st));
            SB_CUDA(cudaGraphLaunch(m->gridGraph, st));
        }
        if (m->geomChanged) { // new geometry in a list sized for the old one: counts to the host, checked at first use
            int rv = grid_counts_readback(m);
            if (rv)
                return rv;
            SB_CUDA(cudaEventRecord(m->verifyEv, st));
            m->verifyPending = true;
            m->geomChanged = false;
        }
        // Recorded at the END of a rebuild on purpose: letting the other mesh's face queries
        // start as soon as the sorted centroids exist (between the two graphs) was measured
        // slower -- the two long classification launches then no longer run side by side and
        // the later one ends with its tail alone (1.32 -> 1.37 ms/step at 1M+1M).
        if (!c->earlyLeaf)
            SB_CUDA(cudaEventRecord(m->leafReady, st));
        c->lc.kernels += m->graphKernels;
        m->treeBuilt = m->treeWanted;
        SB_CUDA(cudaEventRecord(m->ready, st));
        m->built = true;
        return SB_OK;
    }
    if (prepared)
        SB_CUDA(cudaStreamWaitEvent(st, m->prepEv, 0));
    {
        StageTimer t(c, SB_STAGE_BUILD, st);
        if (!prepared)
            SB_CUDA(cudaMemsetAsync(m->d.root, 0, m->d.triJob ? 4 : 8, st));
        SB_CUDA(sbk_build_sort(st, m->d, m->radixWs, c->smCount, c->lc, prepared));
        SB_CUDA(sbk_grid_prepare(st, m->d, m->scanScratch, c->gridBeta, c->gridSlabBits, c->lc));
        SB_CUDA(sbk_build_leaves(st, m->d, c->lc)); // also counts the grid cells
        m->treeBuilt = false;
        // classification queries of this mesh's faces can start here (sorted centroids)
        SB_CUDA(cudaEventRecord(m->leafReady, st));
        if (m->treeWanted) {
            // known traversal target: the LBVH right away, on its own stream -- one
            // latency-bound kernel (an atomic rendezvous per level) next to the
            // bandwidth-bound grid scan / fill of the same mesh; `ready` covers it
            SB_CUDA(cudaEventRecord(m->leavesDone, st));
            SB_CUDA(cudaStreamWaitEvent(m->treeStream, m->leavesDone, 0));
            {
                StageTimer tt(c, SB_STAGE_BUILD, m->treeStream);
                SB_CUDA(sbk_build_tree(m->treeStream, m->d, c->lc));
            }
            SB_CUDA(cudaEventRecord(m->leavesDone, m->treeStream));
            m->treeBuilt = true;
        }
        SB_CUDA(sbk_grid_scan(st, m->d, m->scanScratch, c->lc));
        if (m->d.nT && m->gridSized) {
            // Rebuild of the same (immutable) geometry: every step above is
            // deterministic, so the reference count equals the one the list was
            // sized for -- no host round trip (the fill is bounds-guarded anyway).
            SB_CUDA(sbk_grid_fill(st, m->d, c->lc));
            if (m->geomChanged) {
                int rv = grid_counts_readback(m);
                if (rv)
                    return rv;
                SB_CUDA(cudaEventRecord(m->verifyEv, st));
                m->verifyPending = true;
            }
        }
    }
    m->geomChanged = false;
    if (m->d.nT && !m->gridSized) {
        // first build: the list has to be sized from the counts.  They are sent to the host
        // here; the rest (size, fill, `ready`) happens at the first use of the mesh
        int r = grid_counts_readback(m);
        if (r)
            return r;
        m->gridPending = true;
        m->built = true;
        return SB_OK;
    }
    if (m->treeBuilt && m->d.nT)
        SB_CUDA(cudaStreamWaitEvent(st, m->leavesDone, 0)); // `ready` covers the LBVH too
    SB_CUDA(cudaEventRecord(m->ready, st));
    m->built = true;
    return SB_OK;
}

// LBVH of a traversal target, built on the mesh's stream on first use
static int ensure_tree(const sb_mesh *mc)
{
    sb_mesh *m = const_cast<sb_mesh *>(mc);
    if (m->treeBuilt)
        return SB_OK;
    {
        int rf = mesh_finish(m);
        if (rf)
            return rf;
    }
    sb_context *c = m->ctx;
    {
        StageTimer t(c, SB_STAGE_BUILD, m->stream);
        SB_CUDA(sbk_build_tree(m->stream, m->d, c->lc));
    }
    SB_CUDA(cudaEventRecord(m->ready, m->stream));
    SB_CUDA(cudaEventRecord(m->leafReady, m->stream));
    m->treeBuilt = true;
    m->treeWanted = true;
    return SB_OK;
}

static int mesh_check(sb_mesh *m)
{
    {
        int rf = mesh_finish(m);
        if (rf)
            return rf;
    }
    sb_context *c = m->ctx;
    int err = 0;
    use_mesh(c, m);
    SB_CUDA(cudaMemcpyAsync(&c->hScalars->err, m->d.err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    err = c->hScalars->err;
    if (err)
        return mesh_error(m, err);
    return SB_OK;
}

int sb_mesh_create(sb_context *ctx, const double *xyz, size_t nV, const uint32_t *tri, size_t nT, sb_mesh **out)
{
    int r = sb_mesh_upload(ctx, xyz, nV, tri, nT, out);
    if (r)
        return r;
    DeviceGuard g(ctx);
    r = sb_mesh_build(*out);
    if (!r)
        r = mesh_check(*out);
    if (r) {
        sb_mesh_destroy(*out);
        *out = nullptr;
    }
    return r;
}

static void mesh_destroy_now(sb_mesh *m)
{
    if (m->stream) {
        // free in mesh-stream order, after the context stream is done with the buffers
        order_after_context(m->ctx, m);
        cudaFreeAsync(m->arena, m->stream);
        if (m->gridArena)
            cudaFreeAsync(m->gridArena, m->stream);
        cudaStreamDestroy(m->stream); // deferred by the runtime until the stream drains
        if (m->treeStream)
            cudaStreamDestroy(m->treeStream);
    } else {
        cudaFreeAsync(m->arena, m->ctx->stream);
    }
    if (m->ready)
        cudaEventDestroy(m->ready);
    if (m->leafReady)
        cudaEventDestroy(m->leafReady);
    if (m->leavesDone)
        cudaEventDestroy(m->leavesDone);
    if (m->verifyEv)
        cudaEventDestroy(m->verifyEv);
    for (cudaEvent_t ev : m->chunkEv)
        if (ev)
            cudaEventDestroy(ev);
    if (m->xyzEv)
        cudaEventDestroy(m->xyzEv);
    if (m->prepEv)
        cudaEventDestroy(m->prepEv);
    if (m->buildGraph)
        cudaGraphExecDestroy(m->buildGraph);
    if (m->gridGraph)
        cudaGraphExecDestroy(m->gridGraph);
    delete m;
}

void sb_mesh_destroy(sb_mesh *m)
{
    if (!m)
        return;
    sb_context *c = m->ctx;
    DeviceGuard g(c);
    const size_t bytes = m->arenaBytes + m->gridArenaBytes;
    if (m->stream && !m->d.triJob && !m->d.sharedVtx && !m->d.origFace && (int)c->parked.size() < c->parkMax &&
        c->parkedBytes + bytes <= ((size_t)4 << 30)) {
        c->parked.push_back(m); // see sb_context::parked
        c->parkedBytes += bytes;
        return;
    }
    mesh_destroy_now(m);
}

size_t sb_mesh_num_triangles(const sb_mesh *m) { return m ? m->d.nT : 0; }
size_t sb_mesh_num_vertices(const sb_mesh *m) { return m ? m->d.nV : 0; }

static int mesh_download(const sb_mesh *m, void *dst, const void *src, size_t bytes)
{
    if (!m || !dst)
        return fail(SB_ERR_INVALID, "null mesh or output");
    if (!m->built)
        return fail(SB_ERR_INVALID, "mesh not built");
    {
        int rf = mesh_finish(m);
        if (rf)
            return rf;
    }
    DeviceGuard g(m->ctx);
    use_mesh(m->ctx, m);
    if (bytes)
        SB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, m->ctx->stream));
    SB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return SB_OK;
}

// The device keeps the normals in 32-byte records (normal + packed vertex indices, sb_common.cuh): a strided copy drops the fourth word.
int sb_mesh_normals(const sb_mesh *m, double *out)
{
    if (!m || !out)
        return fail(SB_ERR_INVALID, "null mesh or output");
    if (!m->built)
        return fail(SB_ERR_INVALID, "mesh not built");
    {
        int rf = mesh_finish(m);
        if (rf)
            return rf;
    }
    DeviceGuard g(m->ctx);
    use_mesh(m->ctx, m);
    if (m->d.nT)
        SB_CUDA(cudaMemcpy2DAsync(out, 24, m->d.nrm4, 32, 24, m->d.nT, cudaMemcpyDeviceToHost, m->ctx->stream));
    SB_CUDA(cudaStreamSynchronize(m->ctx->stream));
    return SB_OK;
}
// The front end only keeps the Morton-ordered boxes; the original-order copy is formed on request.
int sb_mesh_triangle_boxes(const sb_mesh *m, double *out)
{
    if (!m || !out)
        return fail(SB_ERR_INVALID, "null mesh or output");
    if (!m->built)
        return fail(SB_ERR_INVALID, "mesh not built");
    if (!m->d.nT)
        return SB_OK;
    {
        int rf = mesh_finish(m);
        if (rf)
            return rf;
    }
    sb_context *c = m->ctx;
    DeviceGuard g(c);
    use_mesh(c, m);
    double2 *tmp = nullptr;
    int r = alloc_async(c, &tmp, 3 * (size_t)m->d.nT, nullptr);
    if (r)
        return r;
    SB_CUDA(sbk_triangle_boxes(c->stream, m->d, tmp, c->lc));
    SB_CUDA(cudaMemcpyAsync(out, tmp, 48 * (size_t)m->d.nT, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFreeAsync(tmp, c->stream);
    return SB_OK;
}
int sb_mesh_order(const sb_mesh *m, uint32_t *out) { return mesh_download(m, out, m ? m->d.sortedTri : nullptr, m ? 4 * (size_t)m->d.nT : 0); }

int sb_mesh_bounds(const sb_mesh *m, double *out6)
{
    unsigned long long enc[6];
    int r = mesh_download(m, enc, m ? m->d.bounds : nullptr, sizeof(enc));
    if (r)
        return r;
    for (int i = 0; i < 6; ++i)
        out6[i] = dkey_inv(enc[i]);
    return SB_OK;
}

int sb_mesh_bvh_info(const sb_mesh *m, sb_bvh_info *out)
{
    if (!m || !out || !m->built)
        return fail(SB_ERR_INVALID, "null or unbuilt mesh");
    {
        DeviceGuard g(m->ctx);
        int rt = ensure_tree(m);
        if (rt)
            return rt;
    }
    int root = 0;
    int r = mesh_download(m, &root, m ? m->d.root : nullptr, sizeof(int));
    if (r)
        return r;
    out->cluster_size = SB_CLUSTER;
    out->num_clusters = m->d.M;
    out->num_internal = m->d.M > 1 ? m->d.M - 1 : 0;
    out->root = root;
    return SB_OK;
}

int sb_mesh_bvh_nodes(const sb_mesh *m, void *out)
{
    if (m && m->built) {
        DeviceGuard g(m->ctx);
        int rt = ensure_tree(m);
        if (rt)
            return rt;
    }
    size_t nI = m && m->d.M > 1 ? m->d.M - 1 : 0;
    return mesh_download(m, out, m ? m->d.nodes : nullptr, 64 * nI);
}

int sb_mesh_bvh_leaves(const sb_mesh *m, void *out, size_t *padded)
{
    if (padded)
        *padded = m ? m->d.nTpad : 0;
    return mesh_download(m, out, m ? m->d.leaf : nullptr, m ? 32 * (size_t)m->d.nTpad : 0);
}

int sb_mesh_grid_info(const sb_mesh *m, sb_grid_info *out)
{
    if (!m || !out)
        return fail(SB_ERR_INVALID, "null mesh or output");
    memset(out, 0, sizeof(*out));
    if (!m->d.nT)
        return SB_OK;
    {
        int rf = mesh_finish(m);
        if (rf)
            return rf;
    }
    GridParams g;
    int r = mesh_download(m, &g, m->d.gridParams, sizeof(g));
    if (r)
        return r;
    uint32_t cnt[8];
    unsigned long long ext[96];
    r = mesh_download(m, cnt, m->d.gridBigCount, sizeof(cnt));
    if (!r)
        r = mesh_download(m, ext, m->d.extentSum, sizeof(ext));
    if (r)
        return r;
    for (int a = 0; a < 3; ++a) {
        out->nu[a] = g.nu[a];
        out->nv[a] = 1u << (15 - g.shiftV[a]);
        out->big[a] = cnt[a];
        unsigned long long sum = 0;
        for (int k = 0; k < 32; ++k)
            sum += ext[3 * k + a];
        out->mean_extent[a] = (float)((double)sum / 16777216.0 * (g.hi[a] - g.lo[a]) / (double)m->d.nT);
    }
    out->total_cells = g.totalCells;
    out->total_refs = cnt[6];
    return SB_OK;
}

// batch meshes only meet batch meshes of the same shape (job j of one against job j of the other)
static int batch_pair_check(const sb_mesh *A, const sb_mesh *B)
{
    if (!A->d.triJob && !B->d.triJob)
        return SB_OK;
    if (!A->d.triJob || !B->d.triJob)
        return fail(SB_ERR_INVALID, "a batch mesh can only be combined with another batch mesh");
    if (A->d.nJobs != B->d.nJobs)
        return fail(SB_ERR_INVALID, "batch meshes with different job counts (%u, %u)", A->d.nJobs, B->d.nJobs);
    if (A->d.latPitch != B->d.latPitch)
        return fail(SB_ERR_INVALID, "batch meshes with different lattice pitches (%g, %g)", A->d.latPitch, B->d.latPitch);
    return SB_OK;
}

// ---- intersection -----------------------------------------------------------------

void sb_isect_destroy(sb_isect *x)
{
    if (!x)
        return;
    DeviceGuard g(x->ctx);
    for (void *p : x->owned)
        cudaFreeAsync(p, x->ctx->stream);
    delete x;
}

int sb_intersect_range(const sb_mesh *A, const sb_mesh *B, size_t begin, size_t end, unsigned flags, sb_isect **out)
{
    if (!A || !B || !out)
        return fail(SB_ERR_INVALID, "null mesh or out");
    *out = nullptr;
    if (A->ctx != B->ctx)
        return fail(SB_ERR_INVALID, "meshes belong to different contexts");
    if (!A->built || !B->built)
        return fail(SB_ERR_INVALID, "mesh not built");
    {
        int rf = mesh_finish(A);
        if (!rf)
            rf = mesh_finish(B);
        if (rf)
            return rf;
    }
    if (end > A->d.nT)
        end = A->d.nT;
    if (begin > end)
        begin = end;
    if (begin % 32)
        return fail(SB_ERR_INVALID, "range begin must be a multiple of 32");
    {
        int rb = batch_pair_check(A, B);
        if (rb)
            return rb;
    }
    sb_context *c = A->ctx;
    DeviceGuard g(c);
    {
        int rt = ensure_tree(B);
        if (rt)
            return rt;
    }
    use_mesh_leaves(c, A); // the query side reads the sorted leaves only: may overlap A's grid build
    use_mesh(c, B);        // the target's LBVH is finished together with its grids
    sb_isect *x = new (std::nothrow) sb_isect;
    if (!x)
        return fail(SB_ERR_NOMEM, "out of host memory");
    x->ctx = c;
    x->A = A;
    x->B = B;
    x->bitsA = bits_for(A->d.nT);
    x->bitsB = bits_for(B->d.nT);
    int r = SB_OK;
    auto bail = [&](int code) {
        sb_isect_destroy(x);
        return code;
    };
#define SB_TRY(expr)          \
    do {                      \
        r = (expr);           \
        if (r)                \
            return bail(r);   \
    } while (0)
#define SB_CUDA_X(expr)                                                                                         \
    do {                                                                                                        \
        cudaError_t e_ = (expr);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            return bail(fail(SB_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)

    SB_TRY(alloc_async(c, &x->flagsA, A->d.nT, &x->owned));
    SB_TRY(alloc_async(c, &x->flagsB, B->d.nT, &x->owned));
    SB_CUDA_X(cudaMemsetAsync(x->flagsA, 0, std::max<size_t>(A->d.nT, 1), c->stream));
    SB_CUDA_X(cudaMemsetAsync(x->flagsB, 0, std::max<size_t>(B->d.nT, 1), c->stream));

    // ---- broad phase (retry once with the exact size if the guess was small) ----
    uint32_t gBegin = (uint32_t)(begin / 32), gEnd = (uint32_t)((end + 31) / 32);
    size_t cap = std::max<size_t>(1 << 16, 4 * ((end - begin) + (size_t)B->d.nT));
    unsigned long long *keys = nullptr;
    size_t nCand = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        SB_TRY(alloc_async(c, &keys, cap, &x->owned)); // owned from the start: an error below releases it with the isect
        {
            StageTimer t(c, SB_STAGE_BROAD);
            SB_CUDA_X(cudaMemsetAsync(c->dScalars, 0, sizeof(DeviceScalars), c->stream));
            SB_CUDA_X(sbk_broad_phase(c->stream, A->d, B->d, gBegin, gEnd, x->bitsB, keys, cap, &c->dScalars->pairCount,
                &c->dScalars->err, c->lc));
        }
        SB_CUDA_X(cudaMemcpyAsync(c->hScalars, c->dScalars, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA_X(cudaStreamSynchronize(c->stream));
        nCand = (size_t)c->hScalars->pairCount;
        if (nCand <= cap)
            break;
        x->owned.pop_back();
        cudaFreeAsync(keys, c->stream);
        keys = nullptr;
        if (attempt == 1)
            return bail(fail(SB_ERR_CAPACITY, "candidate buffer overflow after retry (%zu > %zu)", nCand, cap));
        cap = nCand;
    }
    if (nCand >= (1ull << 32))
        return bail(fail(SB_ERR_CAPACITY, "too many candidate pairs (%zu)", nCand));
    x->nCand = nCand;
    x->candKeys = keys;

    // ---- narrow phase ----
    unsigned long long *hitKeys = nullptr, *hitKeysTmp = nullptr, *keysTmp = nullptr;
    uint32_t *hitSlot = nullptr, *hitSlotTmp = nullptr;
    double2 *hitSegRaw = nullptr;
    uint8_t *hitTagRaw = nullptr;
    if (nCand) {
        SB_TRY(alloc_async(c, &hitTagRaw, nCand, &x->owned));
        SB_TRY(alloc_async(c, &hitKeys, nCand, &x->owned));
        SB_TRY(alloc_async(c, &hitSlot, nCand, &x->owned));
        SB_TRY(alloc_async(c, &hitSegRaw, 3 * nCand, &x->owned));
        {
            StageTimer t(c, SB_STAGE_PREDICATE);
            SB_CUDA_X(sbk_predicate(c->stream, A->d, B->d, keys, (uint32_t)nCand, x->bitsB, hitKeys, hitSlot, hitSegRaw,
                &c->dScalars->hitCount, x->flagsA, x->flagsB, c->dScalars->paths, hitTagRaw, c->lc));
        }
        SB_CUDA_X(cudaMemcpyAsync(c->hScalars, c->dScalars, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA_X(cudaStreamSynchronize(c->stream));
        x->nHit = c->hScalars->hitCount;
        for (int k = 0; k < 5; ++k)
            x->paths[k] = c->hScalars->paths[k];
    }
    if (nCand) {
        StageTimer t(c, SB_STAGE_NARROW);
        const bool doSort = !(flags & SB_ISECT_NO_SORT);
        unsigned long long *sortedHitKeys = hitKeys;
        uint32_t *sortedSlot = hitSlot;
        x->noSort = !doSort;
        (void)keysTmp;
        if (doSort) {
            SB_TRY(ensure_radix_ws(c, std::max<size_t>(x->nHit, 1)));
            if (x->nHit > 1) {
                SB_TRY(alloc_async(c, &hitKeysTmp, x->nHit, &x->owned));
                SB_TRY(alloc_async(c, &hitSlotTmp, x->nHit, &x->owned));
                SB_CUDA_X(sbk_sort_keys(c->stream, hitKeys, hitKeysTmp, hitSlot, hitSlotTmp, x->nHit, 0,
                    (int)(x->bitsA + x->bitsB), c->radixWs, c->smCount, &sortedHitKeys, &sortedSlot, c->lc));
            }
        }
        if (x->nHit) {
            SB_TRY(alloc_async(c, &x->hitAB, 2 * x->nHit, &x->owned));
            SB_TRY(alloc_async(c, &x->hitSeg, 3 * x->nHit, &x->owned));
            SB_TRY(alloc_async(c, &x->hitTag, x->nHit, &x->owned));
            SB_CUDA_X(sbk_gather_hits(c->stream, sortedHitKeys, sortedSlot, hitSegRaw, (uint32_t)x->nHit, x->bitsB, x->hitAB,
                x->hitSeg, hitTagRaw, x->hitTag, c->lc));
        }
    }
#undef SB_TRY
#undef SB_CUDA_X
    *out = x;
    return SB_OK;
}

int sb_intersect(const sb_mesh *A, const sb_mesh *B, unsigned flags, sb_isect **out)
{
    return sb_intersect_range(A, B, 0, A ? A->d.nT : 0, flags, out);
}

int sb_isect_path_counts(const sb_isect *x, uint64_t *out5)
{
    if (!x || !out5)
        return fail(SB_ERR_INVALID, "null argument");
    for (int k = 0; k < 5; ++k)
        out5[k] = x->paths[k];
    return SB_OK;
}

int sb_fp64_peak(sb_context *c, double *nofma_gflops, double *fma_gflops)
{
    if (!c)
        return fail(SB_ERR_INVALID, "context is null");
    DeviceGuard g(c);
    double *scratch = nullptr;
    SB_CUDA(cudaMalloc(&scratch, 64));
    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0));
    SB_CUDA(cudaEventCreate(&e1));
    double results[2] = {0, 0};
    for (int fma = 0; fma < 2; ++fma) {
        unsigned long long flops = 0;
        sbk_fp64_peak(c->stream, c->smCount, scratch, 2000, fma != 0, &flops); // warm-up
        double best = 0;
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0, c->stream);
            sbk_fp64_peak(c->stream, c->smCount, scratch, 20000, fma != 0, &flops);
            cudaEventRecord(e1, c->stream);
            SB_CUDA(cudaEventSynchronize(e1));
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms > 0)
                best = std::max(best, (double)flops / (ms * 1e-3) / 1e9);
        }
        results[fma] = best;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(scratch);
    if (nofma_gflops)
        *nofma_gflops = results[0];
    if (fma_gflops)
        *fma_gflops = results[1];
    return SB_OK;
}

int sb_isect_counts(const sb_isect *x, size_t *nCand, size_t *nHit)
{
    if (!x)
        return fail(SB_ERR_INVALID, "isect is null");
    if (nCand)
        *nCand = x->nCand;
    if (nHit)
        *nHit = x->nHit;
    return SB_OK;
}

// The full candidate list is only consumed by tests / callers that ask for it:
// ordering it by (a, b) is deferred to the first request.
static int ensure_candidates_sorted(sb_isect *x)
{
    if (x->candSorted || x->noSort || x->nCand < 2) {
        x->candSorted = true;
        return SB_OK;
    }
    sb_context *c = x->ctx;
    int r = ensure_radix_ws(c, x->nCand);
    if (r)
        return r;
    unsigned long long *tmp = nullptr, *sorted = nullptr;
    r = alloc_async(c, &tmp, x->nCand, &x->owned);
    if (r)
        return r;
    StageTimer t(c, SB_STAGE_NARROW);
    SB_CUDA(sbk_sort_keys(c->stream, x->candKeys, tmp, nullptr, nullptr, x->nCand, 2, 2 + (int)(x->bitsA + x->bitsB),
        c->radixWs, c->smCount, &sorted, nullptr, c->lc));
    x->candKeys = sorted;
    x->candSorted = true;
    return SB_OK;
}

int sb_isect_candidates(const sb_isect *xc, uint32_t *ab, uint8_t *code)
{
    sb_isect *x = const_cast<sb_isect *>(xc);
    if (!x || !ab)
        return fail(SB_ERR_INVALID, "null isect or output");
    sb_context *c = x->ctx;
    DeviceGuard g(c);
    if (!x->nCand)
        return SB_OK;
    {
        int rs = ensure_candidates_sorted(x);
        if (rs)
            return rs;
    }
    uint32_t *dAB = nullptr;
    uint8_t *dCode = nullptr;
    int r = alloc_async(c, &dAB, 2 * x->nCand, nullptr);
    if (r)
        return r;
    r = alloc_async(c, &dCode, x->nCand, nullptr);
    if (r)
        return r;
    SB_CUDA(sbk_decode_candidates(c->stream, x->candKeys, (uint32_t)x->nCand, x->bitsB, dAB, dCode, c->lc));
    SB_CUDA(cudaMemcpyAsync(ab, dAB, 8 * x->nCand, cudaMemcpyDeviceToHost, c->stream));
    if (code)
        SB_CUDA(cudaMemcpyAsync(code, dCode, x->nCand, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    cudaFreeAsync(dAB, c->stream);
    cudaFreeAsync(dCode, c->stream);
    return SB_OK;
}

int sb_isect_hits(const sb_isect *x, uint32_t *ab, double *seg)
{
    if (!x)
        return fail(SB_ERR_INVALID, "isect is null");
    sb_context *c = x->ctx;
    DeviceGuard g(c);
    if (!x->nHit)
        return SB_OK;
    if (ab)
        SB_CUDA(cudaMemcpyAsync(ab, x->hitAB, 8 * x->nHit, cudaMemcpyDeviceToHost, c->stream));
    if (seg)
        SB_CUDA(cudaMemcpyAsync(seg, x->hitSeg, 48 * x->nHit, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    return SB_OK;
}

int sb_isect_hit_edges(const sb_isect *x, uint8_t *tags)
{
    if (!x || !tags)
        return fail(SB_ERR_INVALID, "null argument");
    sb_context *c = x->ctx;
    DeviceGuard g(c);
    if (!x->nHit)
        return SB_OK;
    if (!x->hitTag)
        return fail(SB_ERR_INVALID, "this intersection carries no edge tags");
    SB_CUDA(cudaMemcpyAsync(tags, x->hitTag, x->nHit, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    return SB_OK;
}

int sb_batch_job_ranges(const sb_isect *x, size_t *hit_start)
{
    if (!x || !hit_start)
        return fail(SB_ERR_INVALID, "null argument");
    const sb_mesh *A = x->A;
    if (!A || !A->d.triJob)
        return fail(SB_ERR_INVALID, "not the intersection of two batch meshes");
    if (x->noSort)
        return fail(SB_ERR_INVALID, "hits were left unsorted (SB_ISECT_NO_SORT)");
    const size_t J = A->d.nJobs, H = x->nHit;
    std::vector<uint32_t> ab(2 * std::max<size_t>(H, 1));
    if (H) {
        DeviceGuard g(x->ctx);
        SB_CUDA(cudaMemcpyAsync(ab.data(), x->hitAB, 8 * H, cudaMemcpyDeviceToHost, x->ctx->stream));
        SB_CUDA(cudaStreamSynchronize(x->ctx->stream));
    }
    // the hits ascend in (a, b) and a job's triangles are one index range: one sweep
    size_t k = 0;
    for (size_t j = 0; j <= J; ++j) {
        while (k < H && ab[2 * k] < A->jobTriStart[j])
            ++k;
        hit_start[j] = k;
    }
    hit_start[J] = H;
    return SB_OK;
}

int sb_isect_face_flags(const sb_isect *x, uint8_t *flagsA, uint8_t *flagsB)
{
    if (!x)
        return fail(SB_ERR_INVALID, "isect is null");
    sb_context *c = x->ctx;
    DeviceGuard g(c);
    if (flagsA && x->A->d.nT)
        SB_CUDA(cudaMemcpyAsync(flagsA, x->flagsA, x->A->d.nT, cudaMemcpyDeviceToHost, c->stream));
    if (flagsB && x->B->d.nT)
        SB_CUDA(cudaMemcpyAsync(flagsB, x->flagsB, x->B->d.nT, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    return SB_OK;
}

int sb_isect_device_ptrs(const sb_isect *xc, void **cand_keys, unsigned *bits_b, void **hit_ab, void **hit_seg,
    void **flagsA, void **flagsB)
{
    sb_isect *x = const_cast<sb_isect *>(xc);
    if (!x)
        return fail(SB_ERR_INVALID, "isect is null");
    if (cand_keys) {
        DeviceGuard g(x->ctx);
        int rs = ensure_candidates_sorted(x);
        if (rs)
            return rs;
    }
    if (cand_keys) *cand_keys = x->candKeys;
    if (bits_b) *bits_b = x->bitsB;
    if (hit_ab) *hit_ab = x->hitAB;
    if (hit_seg) *hit_seg = x->hitSeg;
    if (flagsA) *flagsA = x->flagsA;
    if (flagsB) *flagsB = x->flagsB;
    return SB_OK;
}

// ---- uncut triangles + half-edge map (sb_halfedge.cu) ---------------------------------

void sb_uncut_destroy(sb_uncut *u)
{
    if (!u)
        return;
    DeviceGuard g(u->ctx);
    for (void *p : u->owned)
        cudaFreeAsync(p, u->ctx->stream);
    delete u;
}

// dCut: device flags (nT bytes, 8-byte aligned) or null; everything on the context stream
static int uncut_run(const sb_mesh *mesh, const uint8_t *dCut, size_t vertexOffset, size_t triangleOffset, sb_uncut **out)
{
    sb_context *c = mesh->ctx;
    const MeshDev &d = mesh->d;
    if (vertexOffset + d.nV > 0x100000000ull || triangleOffset + d.nT > 0x7fffffffull)
        return fail(SB_ERR_INVALID, "vertex / triangle offset out of the 32-bit index range");
    sb_uncut *u = new (std::nothrow) sb_uncut;
    if (!u)
        return fail(SB_ERR_NOMEM, "out of host memory");
    u->ctx = c;
    u->mesh = mesh;
    u->vertexOffset = (uint32_t)vertexOffset;
    u->triangleOffset = (uint32_t)triangleOffset;
    int r = SB_OK;
    auto bail = [&](int code) {
        sb_uncut_destroy(u);
        return code;
    };
#define SB_TRY(expr)          \
    do {                      \
        r = (expr);           \
        if (r)                \
            return bail(r);   \
    } while (0)
#define SB_CUDA_X(expr)                                                                                         \
    do {                                                                                                        \
        cudaError_t e_ = (expr);                                                                                \
        if (e_ != cudaSuccess)                                                                                  \
            return bail(fail(SB_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)
    StageTimer timer(c, SB_STAGE_HALFEDGE);
    uint32_t *tileScratch = nullptr;
    SB_TRY(alloc_async(c, &tileScratch, sbk_uncut_tiles(d.nT), &u->owned));
    SB_CUDA_X(cudaMemsetAsync(&c->dScalars->heRepeat, 0xff, sizeof(unsigned int), c->stream));
    SB_CUDA_X(cudaMemsetAsync(&c->dScalars->err, 0, sizeof(int), c->stream));
    SB_CUDA_X(sbk_uncut_count(c->stream, dCut, d.nT, tileScratch, &c->dScalars->uncutTotal, c->lc));
    // the number of uncut faces sizes everything that follows (one 128-byte read-back)
    SB_CUDA_X(cudaMemcpyAsync(c->hScalars, c->dScalars, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA_X(cudaStreamSynchronize(c->stream));
    u->nTri = c->hScalars->uncutTotal;
    const size_t n = 3 * (size_t)u->nTri;
    if (n) {
        const unsigned bitsV = bits_for(d.nV); // keys are sorted on the mesh's own vertex ids
        // one 8-byte record per half-edge when key and ordinal fit 64 bits together (they do up to
        // ~2 M vertices x 4 M uncut triangles): keys-only sort, a third less traffic per pass
        unsigned bitsO = bits_for(n);
        if (2 * bitsV + bitsO > 64 || std::getenv("SB_HE_PAIR_SORT"))
            bitsO = 0;
        unsigned long long *k0 = nullptr, *k1 = nullptr, *sk = nullptr;
        uint32_t *o0 = nullptr, *o1 = nullptr, *so = nullptr, *vstart = nullptr;
        SB_TRY(alloc_async(c, &u->face, u->nTri, &u->owned));
        SB_TRY(alloc_async(c, &u->tri3, n, &u->owned));
        SB_TRY(alloc_async(c, &k0, n, &u->owned));
        SB_TRY(alloc_async(c, &k1, n, &u->owned));
        SB_TRY(alloc_async(c, &o0, n, &u->owned)); // sort values, or the unpacked ordinals of the sorted entries
        if (!bitsO)
            SB_TRY(alloc_async(c, &o1, n, &u->owned));
        SB_TRY(alloc_async(c, &u->owner, n, &u->owned));
        SB_TRY(alloc_async(c, &u->adj, n, &u->owned));
        SB_TRY(alloc_async(c, &vstart, d.nV, &u->owned));
        SB_TRY(ensure_radix_ws(c, n));
        SB_CUDA_X(sbk_uncut_emit(c->stream, dCut, d.tri, d.nT, d.nV, &c->dScalars->err, tileScratch, u->vertexOffset, bitsV, bitsO,
            u->face, u->tri3, k0, o0, c->lc));
        if (bitsO)
            SB_CUDA_X(sbk_sort_keys(c->stream, k0, k1, nullptr, nullptr, n, (int)bitsO, (int)(bitsO + 2 * bitsV), c->radixWs,
                c->smCount, &sk, nullptr, c->lc));
        else
            SB_CUDA_X(sbk_sort_keys(c->stream, k0, k1, o0, o1, n, 0, (int)(2 * bitsV), c->radixWs, c->smCount, &sk, &so, c->lc));
        // the reference-format keys go to whichever key buffer the sort left free
        u->keys = sk == k0 ? k1 : k0;
        u->ords = bitsO ? o0 : so;
        SB_CUDA_X(sbk_halfedge_link(c->stream, sk, so, (uint32_t)n, bitsV, bitsO, d.nV, vstart, u->vertexOffset, u->triangleOffset,
            u->keys, u->owner, u->adj, bitsO ? o0 : nullptr, &c->dScalars->heRepeat, c->lc));
        SB_CUDA_X(cudaMemcpyAsync(c->hScalars, c->dScalars, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA_X(cudaStreamSynchronize(c->stream));
        u->repeatOrd = c->hScalars->heRepeat;
        if (c->hScalars->err)
            return bail(fail(SB_ERR_INVALID, "triangle index out of range"));
    }
#undef SB_TRY
#undef SB_CUDA_X
    *out = u;
    return SB_OK;
}

int sb_isect_uncut(const sb_isect *x, int which, size_t vertex_offset, size_t triangle_offset, sb_uncut **out)
{
    if (!x || !out || (which != 0 && which != 1))
        return fail(SB_ERR_INVALID, "null isect / out, or which not 0 / 1");
    *out = nullptr;
    DeviceGuard g(x->ctx);
    return uncut_run(which == 0 ? x->A : x->B, which == 0 ? x->flagsA : x->flagsB, vertex_offset, triangle_offset, out);
}

int sb_mesh_uncut(const sb_mesh *mesh, const uint8_t *cut_flags, size_t vertex_offset, size_t triangle_offset, sb_uncut **out)
{
    if (!mesh || !out)
        return fail(SB_ERR_INVALID, "null mesh or out");
    *out = nullptr;
    sb_context *c = mesh->ctx;
    DeviceGuard g(c);
    use_mesh(c, mesh); // the geometry upload ran on the mesh stream
    uint8_t *dCut = nullptr;
    if (cut_flags && mesh->d.nT) {
        int r = alloc_async(c, &dCut, mesh->d.nT, nullptr);
        if (r)
            return r;
        cudaError_t ec = cudaMemcpyAsync(dCut, cut_flags, mesh->d.nT, cudaMemcpyHostToDevice, c->stream);
        if (ec != cudaSuccess) {
            cudaFreeAsync(dCut, c->stream);
            return fail(SB_ERR_CUDA, "cudaMemcpyAsync(cut flags): %s", cudaGetErrorString(ec));
        }
    }
    int r = uncut_run(mesh, dCut, vertex_offset, triangle_offset, out);
    if (dCut)
        cudaFreeAsync(dCut, c->stream);
    return r;
}

// what the reference leaves behind: everything, or the state at the refused insertion
static inline size_t uncut_tri_count(const sb_uncut *u) { return u->repeatOrd == 0xffffffffu ? u->nTri : u->repeatOrd / 3 + 1; }

int sb_uncut_counts(const sb_uncut *u, size_t *n_triangles, size_t *n_half_edges, int *ok)
{
    if (!u)
        return fail(SB_ERR_INVALID, "uncut is null");
    const bool good = u->repeatOrd == 0xffffffffu;
    if (n_triangles)
        *n_triangles = uncut_tri_count(u);
    if (n_half_edges)
        *n_half_edges = good ? 3 * (size_t)u->nTri : u->repeatOrd;
    if (ok)
        *ok = good ? 1 : 0;
    return SB_OK;
}

int sb_uncut_triangles(const sb_uncut *u, uint32_t *face, uint32_t *tri3)
{
    if (!u)
        return fail(SB_ERR_INVALID, "uncut is null");
    sb_context *c = u->ctx;
    DeviceGuard g(c);
    const size_t n = uncut_tri_count(u);
    if (!n)
        return SB_OK;
    if (face)
        SB_CUDA(cudaMemcpyAsync(face, u->face, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    if (tri3)
        SB_CUDA(cudaMemcpyAsync(tri3, u->tri3, 12 * n, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    return SB_OK;
}

// Rare path (the reference returned false): the arrays hold all uncut faces; keep the entries
// inserted before the refused one.  Host-side filter of the downloaded arrays, order preserved.
static int uncut_truncated(const sb_uncut *u, std::vector<unsigned long long> &keys, std::vector<uint32_t> &owner)
{
    sb_context *c = u->ctx;
    const size_t n = 3 * (size_t)u->nTri;
    std::vector<unsigned long long> k(n);
    std::vector<uint32_t> o(n), ord(n);
    SB_CUDA(cudaMemcpyAsync(k.data(), u->keys, 8 * n, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaMemcpyAsync(o.data(), u->owner, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaMemcpyAsync(ord.data(), u->ords, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    keys.clear();
    owner.clear();
    for (size_t i = 0; i < n; ++i)
        if (ord[i] < u->repeatOrd) {
            keys.push_back(k[i]);
            owner.push_back(o[i]);
        }
    return SB_OK;
}

int sb_uncut_half_edges(const sb_uncut *u, uint64_t *keys, uint32_t *owner)
{
    if (!u)
        return fail(SB_ERR_INVALID, "uncut is null");
    sb_context *c = u->ctx;
    DeviceGuard g(c);
    const size_t n = 3 * (size_t)u->nTri;
    if (!n)
        return SB_OK;
    if (u->repeatOrd != 0xffffffffu) {
        std::vector<unsigned long long> k;
        std::vector<uint32_t> o;
        int r = uncut_truncated(u, k, o);
        if (r)
            return r;
        if (keys && !k.empty())
            memcpy(keys, k.data(), 8 * k.size());
        if (owner && !o.empty())
            memcpy(owner, o.data(), 4 * o.size());
        return SB_OK;
    }
    if (keys)
        SB_CUDA(cudaMemcpyAsync(keys, u->keys, 8 * n, cudaMemcpyDeviceToHost, c->stream));
    if (owner)
        SB_CUDA(cudaMemcpyAsync(owner, u->owner, 4 * n, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    return SB_OK;
}

int sb_uncut_adjacency(const sb_uncut *u, int32_t *adj3)
{
    if (!u || !adj3)
        return fail(SB_ERR_INVALID, "null uncut or output");
    sb_context *c = u->ctx;
    DeviceGuard g(c);
    if (!u->nTri)
        return SB_OK;
    if (u->repeatOrd != 0xffffffffu) {
        // look the opposite half-edges up in the truncated map, as buildFaceGroups would
        std::vector<unsigned long long> k;
        std::vector<uint32_t> o;
        int r = uncut_truncated(u, k, o);
        if (r)
            return r;
        const size_t nt = uncut_tri_count(u);
        std::vector<uint32_t> t(3 * nt);
        SB_CUDA(cudaMemcpyAsync(t.data(), u->tri3, 12 * nt, cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA(cudaStreamSynchronize(c->stream));
        for (size_t j = 0; j < nt; ++j)
            for (int e = 0; e < 3; ++e) {
                unsigned long long want = ((unsigned long long)t[3 * j + (e + 1) % 3] << 32) | t[3 * j + e];
                auto it = std::lower_bound(k.begin(), k.end(), want);
                adj3[3 * j + e] = (it != k.end() && *it == want) ? (int32_t)o[it - k.begin()] : -1;
            }
        return SB_OK;
    }
    SB_CUDA(cudaMemcpyAsync(adj3, u->adj, 12 * (size_t)u->nTri, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    return SB_OK;
}

int sb_uncut_components(const sb_uncut *uc, uint32_t *label, size_t *n_components)
{
    sb_uncut *u = const_cast<sb_uncut *>(uc);
    if (!u)
        return fail(SB_ERR_INVALID, "uncut is null");
    sb_context *c = u->ctx;
    DeviceGuard g(c);
    if (n_components)
        *n_components = 0;
    if (!u->nTri)
        return SB_OK;
    if (u->repeatOrd != 0xffffffffu) {
        // Rare path (the reference returned false): the truncated map is not symmetric, so the
        // groups depend on the order of the reference's flood -- repeat it on the host:
        // ascending seeds, forward lookups only (src/solidboolean.cpp:205-238).
        const size_t nt = uncut_tri_count(u);
        std::vector<int32_t> adj(3 * nt);
        int r = sb_uncut_adjacency(u, adj.data());
        if (r)
            return r;
        std::vector<uint32_t> lab(nt, 0xffffffffu), stack;
        size_t comps = 0;
        for (size_t s = 0; s < nt; ++s) {
            if (lab[s] != 0xffffffffu)
                continue;
            ++comps;
            lab[s] = u->triangleOffset + (uint32_t)s;
            stack.assign(1, (uint32_t)s);
            while (!stack.empty()) {
                uint32_t t = stack.back();
                stack.pop_back();
                for (int k = 0; k < 3; ++k) {
                    int32_t o = adj[3 * (size_t)t + k];
                    if (o < 0)
                        continue;
                    uint32_t j = (uint32_t)o - u->triangleOffset;
                    if (j < nt && lab[j] == 0xffffffffu) {
                        lab[j] = u->triangleOffset + (uint32_t)s;
                        stack.push_back(j);
                    }
                }
            }
        }
        if (label)
            memcpy(label, lab.data(), 4 * nt);
        if (n_components)
            *n_components = comps;
        return SB_OK;
    }
    if (!u->label) {
        StageTimer timer(c, SB_STAGE_HALFEDGE);
        uint32_t *parent = nullptr;
        // union-find in the mesh's Morton order when it has one (spatially coherent tiles)
        const MeshDev &d = u->mesh->d;
        const bool ordered = u->mesh->built && d.sortedTri && !std::getenv("SB_CC_FACE_ORDER");
        if (ordered)
            use_mesh_leaves(c, u->mesh);
        int r = alloc_async(c, &u->label, u->nTri, &u->owned);
        if (!r)
            r = alloc_async(c, &parent, sbk_uncut_components_scratch(u->nTri, d.nT, ordered), nullptr);
        if (r)
            return r;
        cudaError_t e = cudaMemsetAsync(&c->dScalars->ccCount, 0, sizeof(unsigned int), c->stream);
        if (e == cudaSuccess)
            e = sbk_uncut_components(c->stream, u->adj, u->nTri, u->triangleOffset, ordered ? d.sortedTri : nullptr, u->face,
                d.nT, parent, u->label, &c->dScalars->ccCount, c->lc);
        cudaFreeAsync(parent, c->stream); // stream-ordered: released once the kernels above are through
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(c->hScalars, c->dScalars, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess)
            e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) {
            u->label = nullptr; // (stays in `owned`) computed again by the next request
            return fail(SB_ERR_CUDA, "components: %s", cudaGetErrorString(e));
        }
        u->nComponents = c->hScalars->ccCount;
    }
    if (label) {
        SB_CUDA(cudaMemcpyAsync(label, u->label, 4 * (size_t)u->nTri, cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA(cudaStreamSynchronize(c->stream));
    }
    if (n_components)
        *n_components = u->nComponents;
    return SB_OK;
}

// the flood of buildFaceGroups over the uncut triangles AND the retriangulated pieces (sb_flood.cu)
int sb_uncut_face_groups(const sb_uncut *uc, const uint32_t *pieces, size_t n_pieces, const uint32_t *fences, size_t n_fences,
    uint32_t *label_uncut, uint32_t *label_piece, size_t *n_groups)
{
    sb_uncut *u = const_cast<sb_uncut *>(uc);
    if (!u || (n_pieces && (!pieces || !label_piece)) || (n_fences && !fences))
        return fail(SB_ERR_INVALID, "null argument");
    if (n_pieces >= (1u << 28) || n_fences >= (1u << 28))
        return fail(SB_ERR_INVALID, "too many pieces / fences");
    sb_context *c = u->ctx;
    DeviceGuard g(c);
    if (n_groups)
        *n_groups = 0;
    if (u->repeatOrd != 0xffffffffu)
        return fail(SB_ERR_INVALID, "the half-edge map was truncated by a repeated half-edge: the groups then depend on the order of the "
                                    "reference's flood (host)");
    const uint32_t nU = u->nTri, nP = (uint32_t)n_pieces, nF = (uint32_t)n_fences;
    if (nU && !label_uncut)
        return fail(SB_ERR_INVALID, "label_uncut is null");
    if (nU + (size_t)nP == 0)
        return SB_OK;
    if (nU) { // the uncut triangles' own components first (kept on the device)
        int r = sb_uncut_components(u, nullptr, nullptr);
        if (r)
            return r;
    }
    StageTimer timer(c, SB_STAGE_HALFEDGE);
    uint32_t maxV = 0;
    for (size_t i = 0; i < 3 * n_pieces; ++i)
        maxV = std::max(maxV, pieces[i]);
    for (size_t i = 0; i < 2 * n_fences; ++i)
        maxV = std::max(maxV, fences[i]);
    unsigned keyBits = 1;
    while (keyBits < 32 && (maxV >> keyBits))
        ++keyBits;
    uint32_t *dPieces = nullptr, *dFences = nullptr, *scratch = nullptr, *dLabU = nullptr, *dLabP = nullptr;
    std::vector<void *> tmp;
    auto bail = [&](int rc) {
        for (void *q : tmp)
            cudaFreeAsync(q, c->stream);
        return rc;
    };
#define SB_TRY_F(expr)             \
    do {                           \
        int r_ = (expr);           \
        if (r_)                    \
            return bail(r_);       \
    } while (0)
#define SB_CUDA_F(expr)                                                                          \
    do {                                                                                         \
        cudaError_t e_ = (expr);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            return bail(fail(SB_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(e_)));           \
    } while (0)
    SB_TRY_F(alloc_async(c, &dPieces, 3 * n_pieces, &tmp));
    SB_TRY_F(alloc_async(c, &dFences, 2 * n_fences, &tmp));
    SB_TRY_F(alloc_async(c, &scratch, sbk_flood_scratch_words(nU, nP, nF), &tmp));
    SB_TRY_F(alloc_async(c, &dLabU, nU, &tmp));
    SB_TRY_F(alloc_async(c, &dLabP, nP, &tmp));
    if (nP)
        SB_CUDA_F(cudaMemcpyAsync(dPieces, pieces, 12 * n_pieces, cudaMemcpyHostToDevice, c->stream));
    if (nF)
        SB_CUDA_F(cudaMemcpyAsync(dFences, fences, 8 * n_fences, cudaMemcpyHostToDevice, c->stream));
    SB_CUDA_F(sbk_flood(c->stream, dPieces, nP, dFences, nF, u->keys, u->owner, 3 * nU, u->label, nU, u->triangleOffset, keyBits, scratch,
        c->smCount, dLabU, dLabP, &c->dScalars->ccCount, c->lc));
    SB_CUDA_F(cudaMemcpyAsync(&c->hScalars->ccCount, &c->dScalars->ccCount, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    if (nU)
        SB_CUDA_F(cudaMemcpyAsync(label_uncut, dLabU, 4 * (size_t)nU, cudaMemcpyDeviceToHost, c->stream));
    if (nP)
        SB_CUDA_F(cudaMemcpyAsync(label_piece, dLabP, 4 * (size_t)nP, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA_F(cudaStreamSynchronize(c->stream));
#undef SB_TRY_F
#undef SB_CUDA_F
    if (n_groups)
        *n_groups = c->hScalars->ccCount;
    return bail(SB_OK);
}

int sb_uncut_device_ptrs(const sb_uncut *u, void **face, void **tri3, void **keys, void **owner, void **adj3)
{
    if (!u)
        return fail(SB_ERR_INVALID, "uncut is null");
    if (face) *face = u->face;
    if (tri3) *tri3 = u->tri3;
    if (keys) *keys = u->keys;
    if (owner) *owner = u->owner;
    if (adj3) *adj3 = u->adj;
    return SB_OK;
}

// ---- per-triangle intersection contexts (sb_cuts.cu) ------------------------------------

void sb_cuts_destroy(sb_cuts *k)
{
    if (!k)
        return;
    DeviceGuard g(k->ctx);
    for (void *p : k->owned)
        cudaFreeAsync(p, k->ctx->stream);
    delete k;
}

int sb_isect_contexts(const sb_isect *x, int which, sb_cuts **out)
{
    if (!x || !out || (which != 0 && which != 1))
        return fail(SB_ERR_INVALID, "null isect / out, or which not 0 / 1");
    *out = nullptr;
    if (x->noSort)
        return fail(SB_ERR_INVALID, "contexts need the hits in ascending order (SB_ISECT_NO_SORT was set)");
    sb_context *c = x->ctx;
    DeviceGuard g(c);
    sb_cuts *k = new (std::nothrow) sb_cuts;
    if (!k)
        return fail(SB_ERR_NOMEM, "out of host memory");
    k->ctx = c;
    const size_t n = x->nHit;
    if (n == 0) {
        *out = k;
        return SB_OK;
    }
    int r = SB_OK;
    uint32_t *scratch = nullptr;
    auto bail = [&](int code) {
        if (scratch)
            cudaFreeAsync(scratch, c->stream);
        sb_cuts_destroy(k);
        return code;
    };
    StageTimer timer(c, SB_STAGE_CONTEXTS);
    if ((r = alloc_async(c, &k->tri, n, &k->owned)) || (r = alloc_async(c, &k->pointStart, n + 1, &k->owned)) ||
        (r = alloc_async(c, &k->edgeStart, n + 1, &k->owned)) || (r = alloc_async(c, &k->points, 6 * n, &k->owned)) ||
        (r = alloc_async(c, &k->edges, 2 * n, &k->owned)) || (r = alloc_async(c, &scratch, sbk_cut_contexts_scratch(n), nullptr)) ||
        (r = ensure_radix_ws(c, n)))
        return bail(r);
    cudaError_t e = sbk_cut_contexts(c->stream, x->hitAB, reinterpret_cast<const double *>(x->hitSeg), (uint32_t)n, which,
        which == 0 ? x->bitsA : x->bitsB, scratch, c->radixWs, c->smCount, k->tri, k->pointStart, k->points, k->edgeStart, k->edges,
        c->dScalars->cuts, c->lc);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(c->hScalars, c->dScalars, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess)
        e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess)
        return bail(fail(SB_ERR_CUDA, "contexts: %s", cudaGetErrorString(e)));
    cudaFreeAsync(scratch, c->stream);
    k->nCtx = c->hScalars->cuts[0];
    k->nPoints = c->hScalars->cuts[1];
    k->nEdges = c->hScalars->cuts[2];
    *out = k;
    return SB_OK;
}

int sb_cuts_counts(const sb_cuts *k, size_t *n_contexts, size_t *n_points, size_t *n_relations)
{
    if (!k)
        return fail(SB_ERR_INVALID, "cuts is null");
    if (n_contexts) *n_contexts = k->nCtx;
    if (n_points) *n_points = k->nPoints;
    if (n_relations) *n_relations = k->nEdges;
    return SB_OK;
}

int sb_cuts_fetch(const sb_cuts *k, uint32_t *tri, uint32_t *point_start, double *points, uint32_t *relation_start,
    uint32_t *relations)
{
    if (!k)
        return fail(SB_ERR_INVALID, "cuts is null");
    sb_context *c = k->ctx;
    DeviceGuard g(c);
    if (!k->nCtx) {
        if (point_start) point_start[0] = 0;
        if (relation_start) relation_start[0] = 0;
        return SB_OK;
    }
    if (tri)
        SB_CUDA(cudaMemcpyAsync(tri, k->tri, 4 * k->nCtx, cudaMemcpyDeviceToHost, c->stream));
    if (point_start)
        SB_CUDA(cudaMemcpyAsync(point_start, k->pointStart, 4 * (k->nCtx + 1), cudaMemcpyDeviceToHost, c->stream));
    if (points && k->nPoints)
        SB_CUDA(cudaMemcpyAsync(points, k->points, 24 * k->nPoints, cudaMemcpyDeviceToHost, c->stream));
    if (relation_start)
        SB_CUDA(cudaMemcpyAsync(relation_start, k->edgeStart, 4 * (k->nCtx + 1), cudaMemcpyDeviceToHost, c->stream));
    if (relations && k->nEdges)
        SB_CUDA(cudaMemcpyAsync(relations, k->edges, 8 * k->nEdges, cudaMemcpyDeviceToHost, c->stream));
    SB_CUDA(cudaStreamSynchronize(c->stream));
    return SB_OK;
}

int sb_cuts_device_ptrs(const sb_cuts *k, void **tri, void **point_start, void **points, void **relation_start, void **relations)
{
    if (!k)
        return fail(SB_ERR_INVALID, "cuts is null");
    if (tri) *tri = k->tri;
    if (point_start) *point_start = k->pointStart;
    if (points) *points = k->points;
    if (relation_start) *relation_start = k->edgeStart;
    if (relations) *relations = k->edges;
    return SB_OK;
}

int sb_isect_pack_device(const sb_isect *x, void *d_record, size_t cap)
{
    if (!x || !d_record)
        return fail(SB_ERR_INVALID, "null isect or record");
    sb_context *c = x->ctx;
    DeviceGuard g(c);
    char *rec = static_cast<char *>(d_record);
    const unsigned long long header[2] = {x->nCand, x->nHit};
    // 16 bytes from the stack: copied into the driver's staging at enqueue time
    SB_CUDA(cudaMemcpyAsync(rec, header, sizeof(header), cudaMemcpyHostToDevice, c->stream));
    const size_t n = std::min(x->nHit, cap);
    if (n) {
        SB_CUDA(cudaMemcpyAsync(rec + 16, x->hitAB, 8 * n, cudaMemcpyDeviceToDevice, c->stream));
        SB_CUDA(cudaMemcpyAsync(rec + 16 + 8 * cap, x->hitSeg, 48 * n, cudaMemcpyDeviceToDevice, c->stream));
    }
    return SB_OK;
}

int sb_tri_tri_batch(sb_context *c, const double *tris18, size_t n, int32_t *ret, int32_t *coplanar, double *seg6)
{
    if (!c || (n && (!tris18 || !ret || !coplanar || !seg6)))
        return fail(SB_ERR_INVALID, "null argument");
    if (n >= (1ull << 31))
        return fail(SB_ERR_INVALID, "batch too large");
    if (!n)
        return SB_OK;
    DeviceGuard g(c);
    double *dT = nullptr, *dSeg = nullptr;
    int32_t *dRet = nullptr, *dCop = nullptr;
    std::vector<void *> owned;
    int r = alloc_async(c, &dT, 18 * n, &owned);
    if (!r) r = alloc_async(c, &dSeg, 6 * n, &owned);
    if (!r) r = alloc_async(c, &dRet, n, &owned);
    if (!r) r = alloc_async(c, &dCop, n, &owned);
    if (!r) {
        cudaError_t e = cudaMemcpyAsync(dT, tris18, 144 * n, cudaMemcpyHostToDevice, c->stream);
        if (e == cudaSuccess) {
            StageTimer t(c, SB_STAGE_NARROW);
            e = sbk_tri_tri_batch(c->stream, dT, (uint32_t)n, dRet, dCop, dSeg, c->lc);
        }
        if (e == cudaSuccess) e = cudaMemcpyAsync(ret, dRet, 4 * n, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(coplanar, dCop, 4 * n, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(seg6, dSeg, 48 * n, cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess)
            r = fail(SB_ERR_CUDA, "sb_tri_tri_batch: %s", cudaGetErrorString(e));
    }
    for (void *p : owned)
        cudaFreeAsync(p, c->stream);
    return r;
}

// ---- classification --------------------------------------------------------------

// One classification on a lane: launch enqueues the kernel and the counter read-back;
// finish waits for the lane, repeats the call once with the exact size if the scratch of
// the many-layer rays was too small, and -- when the target had only its first two ray
// grids and some points' two votes disagree -- has the third grid built and traces the
// third ray of just those points in a second launch.
struct ClassifyJob {
    void *scratch = nullptr;      // [many-layer keys: 24 * cap][undecided list: 4 * points][legacy list: 4 * points]
    unsigned long long cap = 0;
    uint32_t points = 0;
    int firstAxes = 2;
    bool launched = false;
    bool second = false;          // this is the third-axis launch over the undecided list
    void *firstScratch = nullptr; // keeps that list alive during the second launch
    size_t firstKeyBytes = 0;     // offset of the list in firstScratch
    uint32_t undecided = 0;
    bool balanced = false;        // the first launch used the balanced kernel (sb_classify2.cu) ...
    bool legacyPass = false;      // ... and this is the general kernel's launch over the points it left (same scratch,
    uint32_t legacy = 0;          //     counters carried on)
    bool noBalanced = false;      // after a failed legacy pass: everything through the general kernel
    const sb_mesh *rawQuery = nullptr; // raw faces mode: the query mesh whose upload chunks pace the first launch
    int launches = 0;             // classify_launch calls of this job (follow-up passes and retries included)
};

// can the balanced kernel take this launch?  (no per-axis big lists on the grids it may trace)
static bool classify_balanced_ok(const sb_context *c, const sb_mesh *target, const ClassifyArgs &a)
{
    if (!c->classifyBalanced || c->classifyPoolLimit || a.list || a.thirdAxisOnly)
        return false;
    const int axes = a.perAxis ? 3 : target->d.gridAxes;
    for (int k = 0; k < axes; ++k)
        if (target->d.gridBigN[k])
            return false;
    return true;
}

// rawQuery: the query mesh's faces in original order, one launch per upload chunk behind that chunk's event
// (a0.rawFaces; first launch of a job only -- the follow-up launches go over lists of the same numbering)
static int classify_launch(sb_context *c, sb_context::Lane &lane, const sb_mesh *target, const ClassifyArgs &a0,
    ClassifyJob &job, unsigned long long forceCap = 0, const sb_mesh *rawQuery = nullptr)
{
    ClassifyArgs a = a0;
    if (rawQuery)
        job.rawQuery = rawQuery;
    rawQuery = job.rawQuery;
    if (a.perAxis) {
        int r = ensure_grid3(target); // all three rays of every point
        if (r)
            return r;
        cudaStreamWaitEvent(lane.stream, target->ready, 0);
    }
    if (!job.second && !job.legacyPass) {
        job.points = a.end - a.begin;
        job.firstAxes = a.perAxis ? 3 : 2; // without the per-axis bits the vote is lazy (third ray on demand)
    }
    const size_t listBytes = align256(4 * (size_t)std::max<uint32_t>(job.points, 1));
    if (!job.legacyPass) {
        job.cap = forceCap ? forceCap : std::max<unsigned long long>(1 << 16, lane.bigHint + lane.bigHint / 4);
        const size_t keyBytes = align256(24 * (size_t)job.cap);
        SB_CUDA(cudaMallocAsync(&job.scratch, keyBytes + 2 * listBytes, lane.stream));
        SB_CUDA(cudaMemsetAsync(lane.d, 0, sizeof(DeviceScalars), lane.stream));
        if (!job.second)
            job.firstKeyBytes = keyBytes;
    }
    char *const firstBase = static_cast<char *>(job.second ? job.firstScratch : job.scratch);
    uint32_t *const undecidedList = reinterpret_cast<uint32_t *>(firstBase + job.firstKeyBytes);
    uint32_t *const legacyList = reinterpret_cast<uint32_t *>(firstBase + job.firstKeyBytes + listBytes);
    if (job.second) {
        a.list = undecidedList;
        a.listCount = job.undecided;
        a.thirdAxisOnly = true;
    } else {
        a.undecidedList = undecidedList;
        if (job.legacyPass) {
            a.list = legacyList;
            a.listCount = job.legacy;
        }
    }
    job.balanced = !job.second && !job.legacyPass && !job.noBalanced && classify_balanced_ok(c, target, a);
    unsigned long long *trace = nullptr;
    const char *traceFile = job.balanced ? nullptr : getenv("SB_CLASSIFY_TRACE"); // dev: per-CTA timeline of the general kernel's launches
    const uint32_t traceBlocks = sbk_classify_blocks(job.second ? job.undecided : job.legacyPass ? job.legacy : job.points);
    if (traceFile)
        SB_CUDA(cudaMalloc(&trace, 32 * (size_t)traceBlocks));
    const int nLaunch = rawQuery && !a.list && rawQuery->nChunks > 0 ? rawQuery->nChunks : 1;
    for (int k = 0; k < nLaunch; ++k) {
        if (nLaunch > 1 || (rawQuery && !a.list && rawQuery->nChunks == 1)) {
            // faces [chunkEnd[k-1], chunkEnd[k]) once their index triples are on the device
            SB_CUDA(cudaStreamWaitEvent(lane.stream, rawQuery->chunkEv[k], 0));
            a.first = k ? rawQuery->chunkEnd[k - 1] : 0u;
            a.end = rawQuery->chunkEnd[k];
        }
        StageTimer t(c, SB_STAGE_CLASSIFY, lane.stream);
        if (job.balanced)
            SB_CUDA(sbk_classify2(lane.stream, target->d, a, &lane.d->stats[0], &lane.d->overflowCount, &lane.d->legacyCount,
                legacyList, c->lc));
        else
            SB_CUDA(sbk_classify(lane.stream, target->d, a, static_cast<long long *>(job.scratch), job.cap, &lane.d->stats[1],
                &lane.d->stats[0], &lane.d->overflowCount, c->classifyPoolLimit, trace, c->lc));
    }
    if (trace) {
        std::vector<unsigned long long> h(4 * (size_t)traceBlocks);
        SB_CUDA(cudaStreamSynchronize(lane.stream));
        SB_CUDA(cudaMemcpy(h.data(), trace, 32 * (size_t)traceBlocks, cudaMemcpyDeviceToHost));
        cudaFree(trace);
        if (FILE *f = fopen(traceFile, "ab")) {
            unsigned long long hdr[4] = {0xffffffffffffffffull, traceBlocks, job.points, 0};
            fwrite(hdr, 8, 4, f);
            fwrite(h.data(), 8, h.size(), f);
            fclose(f);
        }
    }
    SB_CUDA(cudaMemcpyAsync(lane.h, lane.d, sizeof(DeviceScalars), cudaMemcpyDeviceToHost, lane.stream));
    job.launched = true;
    job.launches += 1;
    return SB_OK;
}

static int classify_finish(sb_context *c, sb_context::Lane &lane, const sb_mesh *target, const ClassifyArgs &a,
    ClassifyJob &job, bool accumulateStats)
{
    if (!accumulateStats)
        c->lastRays = c->lastCands = 0;
    auto release = [&]() {
        if (job.scratch)
            cudaFreeAsync(job.scratch, lane.stream);
        if (job.firstScratch)
            cudaFreeAsync(job.firstScratch, lane.stream);
        job.scratch = job.firstScratch = nullptr;
    };
    for (int attempt = 0;;) {
        SB_CUDA(cudaStreamSynchronize(lane.stream));
        const unsigned long long needed = lane.h->stats[1]; // scratch entries the many-layer rays asked for
        const uint32_t undecided = lane.h->overflowCount;
        lane.bigHint = std::max(lane.bigHint, needed);
        if (getenv("SB_DEBUG"))
            fprintf(stderr, "[sb] classify%s%s: points %u axes %d undecided %u exact candidates %llu big-ray entries %llu (cap %llu) grids %d left to the general kernel %u\n",
                job.second ? " (third axis)" : job.legacyPass ? " (general kernel over the listed points)" : "",
                job.balanced ? " [balanced kernel]" : "", job.second ? job.undecided : job.legacyPass ? job.legacy : job.points,
                job.firstAxes, undecided, (unsigned long long)lane.h->stats[0], needed, job.cap, target->d.gridAxes,
                lane.h->legacyCount);
        if (needed > job.cap) { // scratch too small: repeat with the exact size
            if (job.legacyPass) {
                // the lists live in the same allocation: start over, everything through the general kernel
                release();
                const int launchesSoFar = job.launches;
                job = ClassifyJob();
                job.launches = launchesSoFar;
                job.noBalanced = true;
                attempt = 0;
                int r = classify_launch(c, lane, target, a, job, needed);
                if (r)
                    return r;
                continue;
            }
            cudaFreeAsync(job.scratch, lane.stream);
            job.scratch = nullptr;
            if (attempt++) {
                release();
                return fail(SB_ERR_CAPACITY, "many-layer ray scratch overflow after retry (%llu > %llu)", needed, job.cap);
            }
            int r = classify_launch(c, lane, target, a, job, needed);
            if (r) {
                release();
                return r;
            }
            continue;
        }
        if (job.balanced && lane.h->legacyCount) {
            // points the balanced kernel does not cover (ray box over several cells, more than 32 matches on
            // a ray): the general kernel classifies them, counters and the undecided list carried on
            job.legacy = lane.h->legacyCount;
            job.legacyPass = true;
            int r = classify_launch(c, lane, target, a, job);
            if (r) {
                release();
                return r;
            }
            continue;
        }
        c->lastCands += lane.h->stats[0]; // exact candidates
        if (!job.second) {
            c->lastRays += (unsigned long long)job.firstAxes * job.points + (job.firstAxes == 2 ? undecided : 0);
            if (job.firstAxes == 2 && undecided && target->d.sharedVtx) {
                // a multi-GPU selection holds the part of the target the first two rays can meet; the third
                // ray needs all of it: the caller (sb_shard_front_end) classifies against the whole meshes
                c->shardUndecided += undecided;
            } else if (job.firstAxes == 2 && undecided && target->d.gridAxes == 2) {
                // the kernel could not trace the third ray (no third grid yet): it listed the points
                job.firstScratch = job.scratch;
                job.scratch = nullptr;
                job.undecided = undecided;
                job.second = true;
                job.legacyPass = false;
                attempt = 0;
                int r = ensure_grid3(target);
                if (!r) {
                    cudaStreamWaitEvent(lane.stream, target->ready, 0);
                    r = classify_launch(c, lane, target, a, job);
                }
                if (r) {
                    release();
                    return r;
                }
                continue;
            }
        }
        release();
        return SB_OK;
    }
}

static int classify_run(sb_context *c, const sb_mesh *target, ClassifyArgs a)
{
    ClassifyJob job;
    int r = classify_launch(c, c->lanes[0], target, a, job);
    if (r)
        return r;
    return classify_finish(c, c->lanes[0], target, a, job, false);
}

int sb_classify(const sb_mesh *target, const double *pts, size_t Q, uint8_t *inside, uint8_t *per_axis)
{
    if (!target || (Q && (!pts || !inside)))
        return fail(SB_ERR_INVALID, "null argument");
    if (!target->built)
        return fail(SB_ERR_INVALID, "mesh not built");
    if (target->d.triJob)
        return fail(SB_ERR_INVALID, "explicit query points cannot be classified against a batch mesh (they belong to no job)");
    {
        int rf = mesh_finish(target);
        if (rf)
            return rf;
    }
    if (Q >= (1ull << 31))
        return fail(SB_ERR_INVALID, "too many points");
    if (!Q)
        return SB_OK;
    sb_context *c = target->ctx;
    DeviceGuard g(c);
    use_mesh(c, target);
    int r = ensure_classify_out(c, 4 * Q, 0);
    if (r)
        return r;
    double *dPts = nullptr;
    r = alloc_async(c, &dPts, 3 * Q, nullptr);
    if (r)
        return r;
    ScratchGuard ptsGuard{dPts, c->stream};
    SB_CUDA(cudaMemcpyAsync(dPts, pts, 24 * Q, cudaMemcpyHostToDevice, c->stream));
    ClassifyArgs a;
    a.pts = dPts;
    a.begin = 0;
    a.end = (uint32_t)Q;
    a.inside = c->classifyOut;
    a.perAxis = per_axis ? c->classifyOut + Q : nullptr; // without it the vote is lazy (third ray on demand)
    if (target->d.nT == 0) {
        SB_CUDA(cudaMemsetAsync(c->classifyOut, 0, 4 * Q, c->stream));
        SB_CUDA(cudaStreamSynchronize(c->stream));
    } else {
        r = classify_run(c, target, a);
    }
    if (!r) {
        SB_CUDA(cudaMemcpyAsync(inside, a.inside, Q, cudaMemcpyDeviceToHost, c->stream));
        if (per_axis)
            SB_CUDA(cudaMemcpyAsync(per_axis, a.perAxis, 3 * Q, cudaMemcpyDeviceToHost, c->stream));
        SB_CUDA(cudaStreamSynchronize(c->stream));
    }
    return r;
}

static int classify_faces_impl(const sb_mesh *query, const sb_mesh *target, size_t begin, size_t end, bool wantPerAxis,
    uint8_t *externalInside, uint8_t **dInside, uint8_t **dPerAxis)
{
    if (!query || !target)
        return fail(SB_ERR_INVALID, "null mesh");
    if (query->ctx != target->ctx)
        return fail(SB_ERR_INVALID, "meshes belong to different contexts");
    if (!query->built || !target->built)
        return fail(SB_ERR_INVALID, "mesh not built");
    {
        int rb = batch_pair_check(query, target);
        if (rb)
            return rb;
    }
    {
        int rf = mesh_finish(query);
        if (!rf)
            rf = mesh_finish(target);
        if (rf)
            return rf;
    }
    sb_context *c = query->ctx;
    use_mesh(c, query);
    use_mesh(c, target);
    size_t n = query->d.nT;
    if (end > n)
        end = n;
    if (begin > end)
        begin = end;
    if (begin % 32)
        return fail(SB_ERR_INVALID, "range begin must be a multiple of 32");
    int r = ensure_classify_out(c, 4 * std::max<size_t>(n, 1), 0);
    if (r)
        return r;
    ClassifyArgs a;
    a.queryMesh = &query->d;
    a.begin = (uint32_t)begin;
    a.end = (uint32_t)end;
    a.inside = externalInside ? externalInside : c->classifyOut;
    a.perAxis = wantPerAxis ? c->classifyOut + n : nullptr;
    *dInside = a.inside;
    *dPerAxis = a.perAxis;
    if (target->d.nT == 0) {
        SB_CUDA(cudaMemsetAsync(c->classifyOut, 0, 4 * std::max<size_t>(n, 1), c->stream));
        if (externalInside && n)
            SB_CUDA(cudaMemsetAsync(externalInside, 0, n, c->stream));
        return SB_OK;
    }
    return classify_run(c, target, a);
}

int sb_classify_faces(const sb_mesh *query, const sb_mesh *target, uint8_t *inside, uint8_t *per_axis)
{
    if (!inside)
        return fail(SB_ERR_INVALID, "inside is null");
    if (!query)
        return fail(SB_ERR_INVALID, "null mesh");
    DeviceGuard g(query->ctx);
    uint8_t *dIn = nullptr, *dAx = nullptr;
    int r = classify_faces_impl(query, target, 0, query->d.nT, per_axis != nullptr, nullptr, &dIn, &dAx);
    if (r)
        return r;
    sb_context *c = query->ctx;
    size_t n = query->d.nT;
    if (n) {
        SB_CUDA(cudaMemcpyAsync(inside, dIn, n, cudaMemcpyDeviceToHost, c->stream));
        if (per_axis)
            SB_CUDA(cudaMemcpyAsync(per_axis, dAx, 3 * n, cudaMemcpyDeviceToHost, c->stream));
    }
    SB_CUDA(cudaStreamSynchronize(c->stream));
    return SB_OK;
}

int sb_classify_faces_device(const sb_mesh *query, const sb_mesh *target, size_t begin, size_t end, void *d_inside)
{
    if (!query || !d_inside)
        return fail(SB_ERR_INVALID, "null mesh or output");
    DeviceGuard g(query->ctx);
    uint8_t *dIn = nullptr, *dAx = nullptr;
    // the slow path for overflowing rays needs the counters on the host, so this
    // call synchronises once after the kernel (a 40-byte read-back)
    return classify_faces_impl(query, target, begin, end, false, static_cast<uint8_t *>(d_inside), &dIn, &dAx);
}

// host outputs of a front end (sb_front_end_host): every result is sent on its way as soon as the stream that produces it
// is done -- A's flags while B's classification still runs, the hit list while both do
struct FrontEndHostOut {
    uint8_t *insideA = nullptr, *insideB = nullptr;
    uint32_t *hitAB = nullptr;
    double *hitSeg = nullptr;
    size_t hitCap = 0;
};

static int front_end_core(const sb_mesh *A, const sb_mesh *B, size_t aBegin, size_t aEnd, size_t bBegin, size_t bEnd,
    unsigned flags, sb_isect **out, void *d_insideA, void *d_insideB, const FrontEndHostOut *host)
{
    if (!A || !B || !out || !d_insideA || !d_insideB)
        return fail(SB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (A->ctx != B->ctx)
        return fail(SB_ERR_INVALID, "meshes belong to different contexts");
    if (!A->built || !B->built)
        return fail(SB_ERR_INVALID, "mesh not built");
    {
        int rb = batch_pair_check(A, B);
        if (rb)
            return rb;
    }
    sb_context *c = A->ctx;
    DeviceGuard g(c);
    aEnd = std::min<size_t>(aEnd, A->d.nT);
    bEnd = std::min<size_t>(bEnd, B->d.nT);
    aBegin = std::min(aBegin, aEnd);
    bBegin = std::min(bBegin, bEnd);
    if (aBegin % 32 || bBegin % 32)
        return fail(SB_ERR_INVALID, "range begin must be a multiple of 32");
    // A mesh whose bytes have only just been sent from the host, later than the other mesh's: its faces are
    // classified in their original order, chunk by chunk as the index triples arrive -- the launch waits for the
    // TARGET's grids and the upload events only, not for this mesh's own build (nor does the host: the target is
    // finished first, the streamed direction enqueued, and only then the just-uploaded mesh's build waited for)
    auto streamed = [&](const sb_mesh *q, const sb_mesh *t, size_t begin, size_t end) {
        return c->streamClassify && q->fresh && q->nChunks > 0 && q->uploadSeq > t->uploadSeq && begin == 0 && end == q->d.nT &&
               end > 0 && !q->d.triJob && !q->d.origFace && !q->d.ownFilter && !q->d.sharedVtx && !t->d.sharedVtx && t->d.nT;
    };
    const bool rawA = streamed(A, B, aBegin, aEnd), rawB = !rawA && streamed(B, A, bBegin, bEnd);
    const_cast<sb_mesh *>(A)->fresh = false;
    const_cast<sb_mesh *>(B)->fresh = false;
    // Optimistic enqueue.  After sb_mesh_update + sb_mesh_build the host does not know yet whether the rebuild's
    // references fitted the lists sized for the old geometry (the counts arrive when the build is done).  Waiting for
    // them here would leave the GPU idle between the build and the first kernel of the front end (wake-up + a dozen
    // launches: ~50 us of a 2 ms step); instead everything is enqueued behind the builds right away -- the balanced
    // classifier keeps its reads inside the allocations whatever the lists hold (sb_classify.cuh Target) -- and the
    // counts are checked once they are there, before anything is handed out: had they changed (lists overflowed, big
    // lists appeared), the run is thrown away and repeated the careful way.
    auto plain = [](const sb_mesh *m) {
        return !m->d.triJob && !m->d.origFace && !m->d.sharedVtx && !m->gridPending && m->gridSized &&
               !m->d.gridBigN[0] && !m->d.gridBigN[1] && !m->d.gridBigN[2];
    };
    const bool optimistic = SB_CLS_GUARDS && c->optimisticVerify && (A->verifyPending || B->verifyPending) && plain(A) && plain(B) && !rawA && !rawB &&
                            c->classifyBalanced && !c->classifyPoolLimit && A->d.nT && B->d.nT;
    struct DeferGuard {
        sb_context *c;
        ~DeferGuard() { c->deferVerify = false; }
    } deferGuard{c};
    c->deferVerify = optimistic;
    // the two classification directions run on their own lanes, behind whatever the
    // context stream has enqueued so far and behind the builds they read
    cudaEventRecord(c->orderEvent, c->stream);
    ClassifyArgs qa, qb;
    qa.queryMesh = &A->d; qa.begin = (uint32_t)aBegin; qa.end = (uint32_t)aEnd; qa.inside = static_cast<uint8_t *>(d_insideA);
    qb.queryMesh = &B->d; qb.begin = (uint32_t)bBegin; qb.end = (uint32_t)bEnd; qb.inside = static_cast<uint8_t *>(d_insideB);
    qa.rawFaces = rawA;
    qb.rawFaces = rawB;
    ClassifyJob ja, jb;
    int r = SB_OK;
    auto is_pinned = [](const void *p) {
        cudaPointerAttributes at;
        if (!p || cudaPointerGetAttributes(&at, p) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return at.type == cudaMemoryTypeHost;
    };
    const bool earlyA = host && is_pinned(host->insideA), earlyB = host && is_pinned(host->insideB);
    bool finished[2] = {false, false}; // mesh_finish done for A / B
    auto finish = [&](int which) {
        if (finished[which])
            return SB_OK;
        finished[which] = true;
        return mesh_finish(which ? B : A);
    };
    for (int step = 0; step < 2 && !r; ++step) {
        const int l = rawB ? 2 - step : 1 + step; // (a streamed direction is enqueued first)
        sb_context::Lane &lane = c->lanes[l];
        const sb_mesh *target = l == 1 ? B : A;
        const sb_mesh *query = l == 1 ? A : B;
        const bool raw = l == 1 ? rawA : rawB;
        r = finish(l == 1 ? 1 : 0); // the target's grids (first build: sized and filled here; rebuild: capacity verified)
        if (!r && !raw)
            r = finish(l == 1 ? 0 : 1);
        if (r)
            break;
        cudaStreamWaitEvent(lane.stream, c->orderEvent, 0);
        // queries need the other mesh's sorted centroids only (raw: nothing of the query's build); the target needs its grids
        cudaStreamWaitEvent(lane.stream, target->ready, 0);
        if (!raw)
            cudaStreamWaitEvent(lane.stream, query->leafReady, 0);
        const ClassifyArgs &q = l == 1 ? qa : qb;
        if (q.end > q.begin && target->d.nT)
            r = classify_launch(c, lane, target, q, l == 1 ? ja : jb, 0, raw ? query : nullptr);
        else if (q.end > q.begin && !q.queryMesh->origFace) // empty target: nothing is inside it
            cudaMemsetAsync(q.inside, 0, q.queryMesh->nT, lane.stream); // (a multi-GPU selection writes its own faces only;
                                                                        //  the caller's array starts out cleared)
        // host outputs: the flags follow their kernel at once (no host round trip in between); a job that turns out to
        // need further launches (points left to the general kernel, third rays) sends them again when it is done
        // (pinned host memory only: a copy to pageable memory would block this thread until the kernel is done)
        if (!r && host && (l == 1 ? host->insideA : host->insideB) && query->d.nT && (l == 1 ? earlyA : earlyB))
            cudaMemcpyAsync(l == 1 ? host->insideA : host->insideB, q.inside, query->d.nT, cudaMemcpyDeviceToHost, lane.stream);
    }
    if (!r)
        r = finish(0);
    if (!r)
        r = finish(1);
    // broad + narrow phase on the context stream meanwhile
    int ri = r ? r : sb_intersect_range(A, B, aBegin, aEnd, flags, out);
    c->deferVerify = false;
    if (optimistic) {
        // (the intersection has synchronised the context stream behind both builds: the counts are on the host)
        if (mesh_verify_would_change(A) || mesh_verify_would_change(B)) {
            c->optimisticRedone += 1;
            for (int l = 1; l <= 2; ++l) {
                ClassifyJob &job = l == 1 ? ja : jb;
                cudaStreamSynchronize(c->lanes[l].stream);
                if (job.scratch)
                    cudaFreeAsync(job.scratch, c->lanes[l].stream);
                if (job.firstScratch)
                    cudaFreeAsync(job.firstScratch, c->lanes[l].stream);
                job.scratch = job.firstScratch = nullptr;
            }
            if (*out) {
                sb_isect_destroy(*out);
                *out = nullptr;
            }
            // the careful way: the checks first (rebuilding what did not fit), then the same call again
            int rv = mesh_finish(A);
            if (!rv)
                rv = mesh_finish(B);
            if (rv)
                return rv;
            return front_end_core(A, B, aBegin, aEnd, bBegin, bEnd, flags, out, d_insideA, d_insideB, host);
        }
        int rv = mesh_verify(const_cast<sb_mesh *>(A), nullptr);
        if (!rv)
            rv = mesh_verify(const_cast<sb_mesh *>(B), nullptr);
        if (rv && !r)
            r = rv;
    }
    bool hitsSent = false;
    if (!ri && host && host->hitAB && host->hitSeg && *out && (*out)->nHit && (*out)->nHit <= host->hitCap) {
        cudaMemcpyAsync(host->hitAB, (*out)->hitAB, 8 * (*out)->nHit, cudaMemcpyDeviceToHost, c->stream);
        cudaMemcpyAsync(host->hitSeg, (*out)->hitSeg, 48 * (*out)->nHit, cudaMemcpyDeviceToHost, c->stream);
        hitsSent = true;
    }
    c->lastRays = c->lastCands = 0;
    bool firstStats = true;
    if (ja.launched) {
        int rf = classify_finish(c, c->lanes[1], B, qa, ja, !firstStats);
        firstStats = false;
        if (!r) r = rf;
    }
    if (host && host->insideA && A->d.nT && !r && (ja.launches > 1 || !earlyA))
        cudaMemcpyAsync(host->insideA, d_insideA, A->d.nT, cudaMemcpyDeviceToHost, c->lanes[1].stream);
    if (jb.launched) {
        int rf = classify_finish(c, c->lanes[2], A, qb, jb, !firstStats);
        firstStats = false;
        if (!r) r = rf;
    }
    if (host && host->insideB && B->d.nT && !r && (jb.launches > 1 || !earlyB))
        cudaMemcpyAsync(host->insideB, d_insideB, B->d.nT, cudaMemcpyDeviceToHost, c->lanes[2].stream);
    cudaStreamSynchronize(c->lanes[1].stream);
    cudaStreamSynchronize(c->lanes[2].stream);
    if (hitsSent && cudaStreamSynchronize(c->stream) != cudaSuccess && !r)
        r = fail(SB_ERR_CUDA, "front end: copying the hit list to the host failed");
    if (!r)
        r = ri;
    if (r && *out) {
        sb_isect_destroy(*out);
        *out = nullptr;
    }
    return r;
}

int sb_front_end_range(const sb_mesh *A, const sb_mesh *B, size_t aBegin, size_t aEnd, size_t bBegin, size_t bEnd,
    unsigned flags, sb_isect **out, void *d_insideA, void *d_insideB)
{
    return front_end_core(A, B, aBegin, aEnd, bBegin, bEnd, flags, out, d_insideA, d_insideB, nullptr);
}

int sb_front_end(const sb_mesh *A, const sb_mesh *B, unsigned flags, sb_isect **out, void *d_insideA, void *d_insideB)
{
    return sb_front_end_range(A, B, 0, A ? A->d.nT : 0, 0, B ? B->d.nT : 0, flags, out, d_insideA, d_insideB);
}

int sb_front_end_host(const sb_mesh *A, const sb_mesh *B, unsigned flags, sb_isect **out, uint8_t *insideA, uint8_t *insideB,
    uint32_t *hit_ab, double *hit_seg, size_t hit_capacity)
{
    if (!A || !B || !out || !insideA || !insideB)
        return fail(SB_ERR_INVALID, "null argument");
    if (A->ctx != B->ctx)
        return fail(SB_ERR_INVALID, "meshes belong to different contexts");
    sb_context *c = A->ctx;
    DeviceGuard g(c);
    // the per-face flags live in a buffer the context keeps (grow-only)
    const size_t nA = A->d.nT, nB = B->d.nT, offB = align256(std::max<size_t>(nA, 1)), need = offB + std::max<size_t>(nB, 1);
    if (need > c->feFlagsBytes) {
        if (c->feFlags) {
            SB_CUDA(cudaDeviceSynchronize());
            SB_CUDA(cudaFree(c->feFlags));
            c->feFlags = nullptr;
            c->feFlagsBytes = 0;
        }
        SB_CUDA(cudaMalloc(&c->feFlags, need + need / 4));
        c->feFlagsBytes = need + need / 4;
    }
    FrontEndHostOut h;
    h.insideA = insideA;
    h.insideB = insideB;
    h.hitAB = hit_ab;
    h.hitSeg = hit_seg;
    h.hitCap = hit_ab && hit_seg ? hit_capacity : 0;
    uint8_t *dA = static_cast<uint8_t *>(c->feFlags), *dB = dA + offB;
    return front_end_core(A, B, 0, nA, 0, nB, flags, out, dA, dB, &h);
}

// ---- multi-GPU shards (sb_shard.cu) --------------------------------------------------------

struct sb_shard {
    sb_context *ctx = nullptr;
    const sb_mesh *parent[2] = {nullptr, nullptr};
    int rank = 0, n = 1;
    void *arena = nullptr;
    float2 *zr[2] = {nullptr, nullptr};  // per triangle: box z range (floats rounded outwards)
    float2 *zf[2] = {nullptr, nullptr};  // per vertex: z rounded down / up
    uint32_t *tiles[2] = {nullptr, nullptr};
    uint32_t *hist = nullptr;   // histogram + tallest box
    double *cuts = nullptr;     // n + 2
    uint32_t *totals = nullptr; // [selection sizes: 2][mismatch flag]
    void *hPinned = nullptr;    // [cuts: n + 2 doubles][totals: 2 words][mismatch]
    sb_mesh *sub[2] = {nullptr, nullptr};
    uint32_t *face[2] = {nullptr, nullptr};
    cudaEvent_t padded[2] = {nullptr, nullptr};
    // what the last completed call found: a repeated call over the same geometry sizes its launches from it
    // and lets the device confirm (no host round trip in the middle of the step)
    bool planValid = false;
    std::vector<double> lastCuts;
    uint32_t lastTotals[2] = {0, 0};
    uint64_t fallbacks = 0, replans = 0;
};

int sb_shard_create(const sb_mesh *A, const sb_mesh *B, int rank, int n_ranks, sb_shard **out)
{
    if (!A || !B || !out)
        return fail(SB_ERR_INVALID, "null argument");
    *out = nullptr;
    if (A->ctx != B->ctx)
        return fail(SB_ERR_INVALID, "meshes belong to different contexts");
    if (n_ranks < 1 || n_ranks > 1000 || rank < 0 || rank >= n_ranks)
        return fail(SB_ERR_INVALID, "rank %d of %d", rank, n_ranks);
    if (A->d.triJob || B->d.triJob || A->d.sharedVtx || B->d.sharedVtx)
        return fail(SB_ERR_INVALID, "shards are made of plain meshes");
    sb_context *c = A->ctx;
    DeviceGuard g(c);
    sb_shard *s = new (std::nothrow) sb_shard;
    if (!s)
        return fail(SB_ERR_NOMEM, "out of host memory");
    s->ctx = c;
    s->parent[0] = A;
    s->parent[1] = B;
    s->rank = rank;
    s->n = n_ranks;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += align256(std::max<size_t>(bytes, 16));
        return o;
    };
    size_t oZ[2], oT[2], oF[2];
    for (int k = 0; k < 2; ++k) {
        oF[k] = take(8 * (size_t)s->parent[k]->d.nV);
        oZ[k] = take(8 * (size_t)s->parent[k]->d.nT);
        oT[k] = take(4 * (sbk_shard_tiles(s->parent[k]->d.nT) + 1));
    }
    const size_t oHist = take(4 * sbk_shard_hist_words() + 8), oCuts = take(8 * ((size_t)n_ranks + 2)), oTot = take(16);
    const size_t pinnedBytes = 8 * (1000 + 2) + 16; // (n_ranks <= 1000: one size, so that the blocks can be reused)
    cudaError_t e = cudaMallocAsync(&s->arena, off, c->stream);
    if (e == cudaSuccess) {
        if (!c->shardPinned.empty()) {
            s->hPinned = c->shardPinned.back();
            c->shardPinned.pop_back();
        } else {
            e = cudaMallocHost(&s->hPinned, pinnedBytes);
        }
    }
    for (int k = 0; k < 2 && e == cudaSuccess; ++k)
        e = cudaEventCreateWithFlags(&s->padded[k], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        if (s->arena)
            cudaFreeAsync(s->arena, c->stream);
        if (s->hPinned)
            cudaFreeHost(s->hPinned);
        for (int k = 0; k < 2; ++k)
            if (s->padded[k])
                cudaEventDestroy(s->padded[k]);
        delete s;
        return fail(e == cudaErrorMemoryAllocation ? SB_ERR_NOMEM : SB_ERR_CUDA, "shard scratch: %s", cudaGetErrorString(e));
    }
    memset(s->hPinned, 0, pinnedBytes);
    char *b = static_cast<char *>(s->arena);
    for (int k = 0; k < 2; ++k) {
        s->zr[k] = (float2 *)(b + oZ[k]);
        s->zf[k] = (float2 *)(b + oF[k]);
        s->tiles[k] = (uint32_t *)(b + oT[k]);
    }
    s->hist = (uint32_t *)(b + oHist);
    s->cuts = (double *)(b + oCuts);
    s->totals = (uint32_t *)(b + oTot);
    *out = s;
    return SB_OK;
}

void sb_shard_destroy(sb_shard *s)
{
    if (!s)
        return;
    DeviceGuard g(s->ctx);
    for (int k = 0; k < 2; ++k) {
        if (s->sub[k])
            sb_mesh_destroy(s->sub[k]);
        if (s->face[k])
            cudaFreeAsync(s->face[k], s->ctx->stream);
    }
    cudaFreeAsync(s->arena, s->ctx->stream);
    cudaStreamSynchronize(s->ctx->stream); // the read-backs into the pinned block are done
    s->ctx->shardPinned.push_back(s->hPinned);
    for (int k = 0; k < 2; ++k)
        cudaEventDestroy(s->padded[k]);
    delete s;
}

int sb_shard_info(const sb_shard *s, size_t *selected_a, size_t *selected_b, double *z_lo, double *z_hi, uint64_t *fallbacks)
{
    if (!s)
        return fail(SB_ERR_INVALID, "shard is null");
    if (selected_a)
        *selected_a = s->sub[0] ? s->sub[0]->d.nT : 0;
    if (selected_b)
        *selected_b = s->sub[1] ? s->sub[1]->d.nT : 0;
    if (z_lo)
        *z_lo = s->planValid ? s->lastCuts[s->rank] : 0.0;
    if (z_hi)
        *z_hi = s->planValid ? s->lastCuts[s->rank + 1] : 0.0;
    if (fallbacks)
        *fallbacks = s->fallbacks;
    return SB_OK;
}

// one attempt: speculative = sizes and slab borders of the last call, confirmed by the device afterwards
static int shard_attempt(sb_shard *s, bool speculative, unsigned flags, sb_isect **out, void *d_insideA, void *d_insideB, bool *confirmed)
{
    sb_context *c = s->ctx;
    cudaStream_t st = c->stream;
    const int n = s->n, rank = s->rank;
    double *hCuts = static_cast<double *>(s->hPinned);
    uint32_t *hTot = reinterpret_cast<uint32_t *>(hCuts + n + 2);
    const MeshDev &A = s->parent[0]->d, &B = s->parent[1]->d;
    *confirmed = true;
    // ---- plan: padded vertices + bounds of the parents (each on its own stream), z ranges, slab borders, sizes ----
    {
        for (int k = 0; k < 2; ++k) {
            sb_mesh *p = const_cast<sb_mesh *>(s->parent[k]);
            order_after_context(c, p);
            {
                StageTimer t(c, SB_STAGE_SHARD, p->stream);
                SB_CUDA(sbk_bounds_pad(p->stream, p->d, c->smCount, c->lc, s->zf[k]));
            }
            SB_CUDA(cudaEventRecord(s->padded[k], p->stream));
        }
        for (int k = 0; k < 2; ++k)
            SB_CUDA(cudaStreamWaitEvent(st, s->padded[k], 0));
        StageTimer t(c, SB_STAGE_SHARD, st);
        SB_CUDA(sbk_shard_plan(st, A, B, s->zf, s->zr, s->tiles, s->hist, n, s->cuts, c->smCount, c->lc));
        SB_CUDA(cudaMemsetAsync(s->totals, 0, 12, st));
        SB_CUDA(sbk_shard_count(st, A, B, s->zr, s->tiles, s->cuts, rank, n, s->totals, speculative ? s->lastTotals : nullptr,
            s->totals + 2, c->lc));
    }
    SB_CUDA(cudaMemcpyAsync(hCuts, s->cuts, 8 * ((size_t)n + 2), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaMemcpyAsync(hTot, s->totals, 12, cudaMemcpyDeviceToHost, st));
    if (!speculative) {
        SB_CUDA(cudaStreamSynchronize(st));
        s->lastCuts.assign(hCuts, hCuts + n + 2);
        s->lastTotals[0] = hTot[0];
        s->lastTotals[1] = hTot[1];
        s->planValid = true;
    }
    // ---- the two selections: ordinary meshes over the chosen triangles ----
    uint32_t *outTri[2], *outFace[2];
    for (int k = 0; k < 2; ++k) {
        const sb_mesh *p = s->parent[k];
        const uint32_t want = s->lastTotals[k];
        if (!s->sub[k] || s->sub[k]->d.nT != want) {
            if (s->sub[k])
                sb_mesh_destroy(s->sub[k]);
            s->sub[k] = nullptr;
            if (s->face[k])
                cudaFreeAsync(s->face[k], st);
            s->face[k] = nullptr;
            int r = mesh_alloc(c, p->d.nV, want, &s->sub[k], 0, p);
            if (r)
                return r;
            r = alloc_async(c, &s->face[k], std::max<uint32_t>(want, 1), nullptr);
            if (r)
                return r;
        }
        sb_mesh *m = s->sub[k];
        m->d.origFace = s->face[k];
        m->d.ownFilter = true;
        m->d.ownLo = s->lastCuts[rank];
        m->d.ownHi = s->lastCuts[rank + 1];
        m->d.ownClosed = rank == n - 1;
        outTri[k] = m->d.tri;
        outFace[k] = s->face[k];
    }
    {
        StageTimer t(c, SB_STAGE_SHARD, st);
        SB_CUDA(sbk_shard_emit(st, A, B, s->zr, s->tiles, s->cuts, rank, n, s->lastTotals, outTri, outFace, c->lc));
    }
    for (int k = 0; k < 2; ++k) {
        int r = sb_mesh_build(s->sub[k]);
        if (r)
            return r;
    }
    // ---- front end over the selections; flags land at the parents' triangle ids ----
    c->shardUndecided = 0;
    int r = sb_front_end_range(s->sub[0], s->sub[1], 0, s->sub[0]->d.nT, 0, s->sub[1]->d.nT, flags, out, d_insideA, d_insideB);
    if (r)
        return r;
    sb_isect *x = *out;
    if (speculative) {
        // the front end has waited for the context stream since the copies above were enqueued
        SB_CUDA(cudaStreamSynchronize(st));
        if (hTot[2] || memcmp(hCuts, s->lastCuts.data(), 8 * ((size_t)n + 2)) != 0) {
            *confirmed = false; // the geometry changed under the same handles: plan again, for real
            sb_isect_destroy(x);
            *out = nullptr;
            return SB_OK;
        }
    }
    if (x->nHit)
        SB_CUDA(sbk_shard_remap_hits(st, x->hitAB, (uint32_t)x->nHit, s->face[0], s->face[1], c->lc));
    return SB_OK;
}

int sb_shard_front_end(sb_shard *s, unsigned flags, sb_isect **out, void *d_insideA, void *d_insideB)
{
    if (!s || !out || !d_insideA || !d_insideB)
        return fail(SB_ERR_INVALID, "null argument");
    *out = nullptr;
    sb_context *c = s->ctx;
    DeviceGuard g(c);
    const int n = s->n, rank = s->rank;
    bool confirmed = true;
    const bool speculate = s->planValid && s->sub[0] && s->sub[1] && !getenv("SB_SHARD_NO_SPECULATION");
    int r = shard_attempt(s, speculate, flags, out, d_insideA, d_insideB, &confirmed);
    if (!r && !confirmed) {
        ++s->replans;
        // (flags the discarded attempt wrote for faces that are not this rank's any more are the caller's to clear:
        //  the arrays are zeroed per step; a changed plan under unchanged handles means the caller rewrote the geometry)
        r = shard_attempt(s, false, flags, out, d_insideA, d_insideB, &confirmed);
    }
    if (r)
        return r;
    sb_isect *x = *out;
    if (c->shardUndecided) {
        // some points' first two votes disagree: their third ray (along z) leaves the slab.  Rare (C3: none):
        // this rank's faces are classified again against the WHOLE meshes, built here on demand.
        ++s->fallbacks;
        sb_mesh *P[2] = {const_cast<sb_mesh *>(s->parent[0]), const_cast<sb_mesh *>(s->parent[1])};
        for (int k = 0; k < 2 && !r; ++k) {
            P[k]->d.ownFilter = false;
            r = sb_mesh_build(P[k]);
        }
        for (int k = 0; k < 2 && !r; ++k) {
            P[k]->d.ownFilter = true;
            P[k]->d.ownLo = s->lastCuts[rank];
            P[k]->d.ownHi = s->lastCuts[rank + 1];
            P[k]->d.ownClosed = rank == n - 1;
            r = sb_classify_faces_device(P[k], P[1 - k], 0, P[k]->d.nT, k == 0 ? d_insideA : d_insideB);
            P[k]->d.ownFilter = false;
        }
        if (r) {
            sb_isect_destroy(x);
            *out = nullptr;
            return r;
        }
    }
    return SB_OK;
}

} // extern "C"
