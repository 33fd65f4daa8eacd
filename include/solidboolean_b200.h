/*
 * solidboolean_b200 -- C ABI of the B200 (sm_100a) intersection front end.
 *
 * This is the drop-in boundary for the data-parallel path of
 * huxingyi/solidboolean.  The reference has no FFI of its own: its
 * boundary is the SolidMesh / SolidBoolean C++ classes.  Each entry point below
 * names the reference code it replaces (paths relative to the reference tree);
 * INTEGRATION.md shows the three call sites a maintainer patches.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns 0 on success or an
 *     SB_ERR_* code, and sb_last_error() returns a thread-local message;
 *   - there is NO CPU fallback: without a CUDA device every compute call fails
 *     with SB_ERR_CUDA;
 *   - vertices are AoS double[3] (the layout of Vector3, src/vector3.h:306),
 *     triangles are uint32[3] (the reference's size_t triples, narrowed);
 *   - all output index pairs are ORIGINAL triangle ids, sorted by (a, b);
 *   - a context owns one CUDA stream; objects of one context must not be used
 *     from two host threads at once, different contexts are independent
 *     (the reference allows concurrent SolidBooleans over shared const meshes).
 */
#ifndef SOLIDBOOLEAN_B200_H
#define SOLIDBOOLEAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SB_OK 0
#define SB_ERR_INVALID 1  /* bad argument (null pointer, index out of range ...) */
#define SB_ERR_CUDA 2     /* CUDA runtime error / no device */
#define SB_ERR_CAPACITY 3 /* an internal bound was exceeded (see message) */
#define SB_ERR_NOMEM 4

typedef struct sb_context sb_context;
typedef struct sb_mesh sb_mesh;
typedef struct sb_isect sb_isect;
typedef struct sb_uncut sb_uncut;
typedef struct sb_cuts sb_cuts;

/* Thread-local message of the last failing call ("" if none). */
const char *sb_last_error(void);
/* Library version string, e.g. "solidboolean_b200 0.1 sm_100a". */
const char *sb_version(void);

/* ---- context -------------------------------------------------------------- */
int sb_context_create(int device, sb_context **out);
void sb_context_destroy(sb_context *ctx);
/* Block until everything enqueued on the context's stream has finished. */
int sb_context_synchronize(sb_context *ctx);
/* The context's cudaStream_t (as void*), for callers that enqueue NCCL or their
 * own kernels behind this library's work. */
void *sb_context_stream(sb_context *ctx);
int sb_context_device(const sb_context *ctx);

/* ---- mesh: replaces SolidMesh::prepare(), src/solidmesh.cpp:42-76 ----------
 * normals (:47-55), per-triangle boxes (:57-62), whole-mesh box (:64-72) and
 * the AxisAlignedBoudingBoxTree build (:74-75, axisalignedboundingboxtree.cpp:
 * 27-141), the latter as a Morton-code LBVH over clusters of triangles. */

/* A mesh holds at most 2^SB_MAX_TRIANGLE_BITS - 1 triangles (the ray-grid references keep
 * the triangle id in 25 bits) and 2^31 - 1 vertices; larger inputs fail with SB_ERR_INVALID. */
#define SB_MAX_TRIANGLE_BITS 25

/* Host buffers in, copy + build enqueued; returns after the build completed. */
int sb_mesh_create(sb_context *ctx, const double *xyz, size_t nV,
                   const uint32_t *tri, size_t nT, sb_mesh **out);
/* Split form used for device-resident timing: upload copies the geometry to the
 * device (no build), build (re)builds every derived array from the resident
 * geometry.  build is asynchronous (each mesh has its own stream); an invalid mesh
 * (triangle index out of range) is reported by build or, at the latest, by the first
 * call that uses the mesh. */
int sb_mesh_upload(sb_context *ctx, const double *xyz, size_t nV,
                   const uint32_t *tri, size_t nT, sb_mesh **out);
int sb_mesh_build(sb_mesh *mesh);
void sb_mesh_destroy(sb_mesh *mesh);

size_t sb_mesh_num_triangles(const sb_mesh *mesh);
size_t sb_mesh_num_vertices(const sb_mesh *mesh);
/* SolidMesh::triangleNormals(): 3 doubles per triangle (src/solidmesh.cpp:47-55,
 * Vector3::normal src/vector3.h:155-176), bit-exact. */
int sb_mesh_normals(const sb_mesh *mesh, double *out3nT);
/* SolidMesh::triangleAxisAlignedBoundingBoxes(): lower xyz, upper xyz per
 * triangle (src/axisalignedboundingbox.h:31-41), bit-exact. */
int sb_mesh_triangle_boxes(const sb_mesh *mesh, double *out6nT);
/* Whole-mesh box over all vertices: lower xyz, upper xyz. */
int sb_mesh_bounds(const sb_mesh *mesh, double *out6);
/* Morton order: sorted position -> original triangle id (nT entries). */
int sb_mesh_order(const sb_mesh *mesh, uint32_t *outnT);

/* LBVH introspection for invariant tests.  Nodes are records of 8 x 32-bit:
 * float lo[3], hi[3]; int32 ref; int32 aux.  Internal node p = records 2p, 2p+1
 * (its two children); ref >= 0 is an internal node index, ref < 0 is ~cluster.
 * Cluster c covers sorted triangles [c*K, min((c+1)*K, nT)). */
typedef struct sb_bvh_info {
    uint32_t cluster_size; /* K */
    uint32_t num_clusters; /* M */
    uint32_t num_internal; /* M - 1 (0 when M <= 1) */
    int32_t root;          /* internal node index, or ~0 (= -1) when M == 1 */
} sb_bvh_info;
int sb_mesh_bvh_info(const sb_mesh *mesh, sb_bvh_info *out);
int sb_mesh_bvh_nodes(const sb_mesh *mesh, void *out_records /* 2*num_internal*32 B */);
int sb_mesh_bvh_leaves(const sb_mesh *mesh, void *out_records /* padded nT * 32 B */,
                       size_t *padded_count);

/* Ray-grid introspection (sb_grid.cu): per axis a (rays along a) the grid has
 * nu[a] x nv[a] cells over the two perpendicular dimensions. */
typedef struct sb_grid_info {
    uint32_t nu[3], nv[3];
    uint32_t total_cells;
    uint32_t total_refs;    /* 8-byte references over the grids built so far (the third on demand) */
    uint32_t big[3];        /* triangles kept on the per-axis "big" list */
    float mean_extent[3];   /* mean triangle-box extent per world axis */
} sb_grid_info;
int sb_mesh_grid_info(const sb_mesh *mesh, sb_grid_info *out);

/* ---- intersection: replaces SolidBoolean::searchPotentialIntersectedPairs
 * (src/solidboolean.cpp:94-101 -> axisalignedboundingboxtree.h:54-95) and the
 * predicate loop (src/solidboolean.cpp:315-320 -> intersectTwoFaces :103-122 ->
 * tri_tri_intersection_test_3d, thirdparty/GuigueDevillers03/
 * tri_tri_intersect.c:395-472). */

#define SB_ISECT_DEFAULT 0u
#define SB_ISECT_NO_SORT 1u /* leave candidates/hits in emission order */

int sb_intersect(const sb_mesh *A, const sb_mesh *B, unsigned flags, sb_isect **out);
/* Same, restricted to A's triangles at Morton-sorted positions [begin, end)
 * (begin a multiple of 32) -- the multi-GPU shard of SURVEY 8e. */
int sb_intersect_range(const sb_mesh *A, const sb_mesh *B, size_t begin, size_t end,
                       unsigned flags, sb_isect **out);
void sb_isect_destroy(sb_isect *isect);

int sb_isect_counts(const sb_isect *isect, size_t *nCand, size_t *nHit);
/* Candidate pairs (triangle boxes overlap, closed intervals in double).
 * ab: 2*nCand uint32.  code (optional): per pair, bit0 = predicate return
 * value, bit1 = *coplanar. */
int sb_isect_candidates(const sb_isect *isect, uint32_t *ab, uint8_t *code);
/* Intersecting, non-coplanar pairs (what intersectTwoFaces accepts) and their
 * segments: seg = source xyz, target xyz per hit. */
int sb_isect_hits(const sb_isect *isect, uint32_t *ab, double *seg);
/* WHICH EDGE each end point of a hit's segment lies on, carried out of the predicate -- the branch of
 * CONSTRUCT_INTERSECTION taken (thirdparty/GuigueDevillers03/tri_tri_intersect.c:285-356) names the two edges, the
 * permutations of :443-471 / :360-385 map them back to the caller's vertex order.  One byte per hit, in the order of
 * sb_isect_hits: bits 0-1 = edge of the SOURCE point (edge k joins vertices k and (k + 1) mod 3 of its triangle), bit 2 =
 * the edge belongs to B's triangle (else A's); bits 4-5 / bit 6 the same for the TARGET point; bit 7 is always set.
 * With it the retriangulation of a cut face attaches polyline ends to the face's boundary without the absolute-epsilon
 * collinearity test of src/vector2.h:212-215 (src/retriangulator.cpp:97-105). */
int sb_isect_hit_edges(const sb_isect *isect, uint8_t *tags /* n_hit */);
/* Per-face "intersected" flags (m_firstIntersectedFaces / m_secondIntersectedFaces,
 * src/solidboolean.cpp:318-319): nT(A) and nT(B) bytes. */
int sb_isect_face_flags(const sb_isect *isect, uint8_t *flagsA, uint8_t *flagsB);
/* Device pointers of the result arrays (valid until sb_isect_destroy), for
 * gathering over NCCL without a host round trip.  Any out pointer may be NULL.
 * cand_keys: nCand uint64 = (((a << bits_b) | b) << 2) | code;
 * hit_ab: 2*nHit uint32; hit_seg: 6*nHit double. */
int sb_isect_device_ptrs(const sb_isect *isect, void **cand_keys, unsigned *bits_b,
                         void **hit_ab, void **hit_seg, void **flagsA, void **flagsB);

/* The per-rank record of the multi-GPU exchange, written into a caller-owned DEVICE buffer
 * of 16 + 56 * cap bytes on the context stream, no host round trip:
 *   uint64 nCand, uint64 nHit | uint32 hit pairs [cap][2] | double segments [cap][6]
 * (the first min(nHit, cap) hits; the rest of the buffer is left as it is). */
int sb_isect_pack_device(const sb_isect *isect, void *d_record, size_t cap);

/* How the candidate pairs left the predicate (flop accounting): out5 =
 * {rejected at plane(T2) test, rejected at plane(T1) test, coplanar 2-D test,
 *  rejected at the interval test, segment constructed}. */
int sb_isect_path_counts(const sb_isect *isect, uint64_t *out5);
/* FP64 issue-rate microbenchmark on the context's device, GFLOP/s: separate
 * DMUL+DADD (what the predicate may use: no FMA contraction) and DFMA (counted as 2). */
int sb_fp64_peak(sb_context *ctx, double *nofma_gflops, double *fma_gflops);

/* Raw predicate on explicit triangles: 18 doubles per pair (p1 q1 r1 p2 q2 r2).
 * ret/coplanar as tri_tri_intersection_test_3d; seg = 6 doubles per pair, left
 * zero where the predicate does not write them. */
int sb_tri_tri_batch(sb_context *ctx, const double *tris18, size_t n,
                     int32_t *ret, int32_t *coplanar, double *seg6);

/* ---- per-triangle intersection contexts: replaces the body of the pair loop of
 * SolidBoolean::combine (src/solidboolean.cpp:296-339; addIntersectedPoint :305-310, the neighbour
 * bookkeeping :321-339) -- SURVEY 8f row 1.
 * One context per triangle of mesh `which` (0 = first, 1 = second) that takes part in an
 * intersecting pair: its points = the segment end points de-duplicated by PositionKey
 * (src/positionkey.cpp:32-54), the FIRST position seen with a key kept, in first-seen order; its
 * relations = the undirected pairs of point numbers (3 + index, as the reference numbers them
 * behind the triangle's own corners) that a segment joins, coinciding ends dropped.  "Seen" refers
 * to the pair list in ascending (first, second) order, i.e. the hits as sb_isect_hits returns them
 * (the reference's own list is in the order of its tree traversal; its loop body is the same).
 * Contexts come in ascending triangle id; the relations of a context in ascending (low, high).
 * An sb_isect made by sb_intersect_range / sb_front_end_range holds one shard's pairs: its contexts
 * cover those pairs only (the second mesh's triangles may then appear in several shards). */
int sb_isect_contexts(const sb_isect *isect, int which, sb_cuts **out);
void sb_cuts_destroy(sb_cuts *cuts);
int sb_cuts_counts(const sb_cuts *cuts, size_t *n_contexts, size_t *n_points, size_t *n_relations);
/* tri[n_contexts]; point_start / relation_start [n_contexts + 1] (CSR); points 3 doubles each;
 * relations 2 uint32 each.  Any pointer may be NULL. */
int sb_cuts_fetch(const sb_cuts *cuts, uint32_t *tri, uint32_t *point_start, double *points,
                  uint32_t *relation_start, uint32_t *relations);
/* Device pointers of the same arrays (valid until sb_cuts_destroy). */
int sb_cuts_device_ptrs(const sb_cuts *cuts, void **tri, void **point_start, void **points,
                        void **relation_start, void **relations);

/* ---- uncut triangles + half-edge map: replaces SolidBoolean::addUnintersectedTriangles
 * (src/solidboolean.cpp:250-286, called at :411-421) and answers the half-edge lookups of
 * buildFaceGroups (:205-224) -- SURVEY 8f row 2.
 * Every face of the mesh that is not cut becomes new triangle `triangle_offset + rank` (rank =
 * its position among the uncut faces, original order = the reference's m_newTriangles order)
 * with vertex ids shifted by vertex_offset (= m_newVertices.size() before the call, :253); its
 * half-edges are keyed (first << 32) | second (makeHalfEdgeKey, src/solidboolean.h:75-78).
 * vertex_offset + nV must fit 32 bits.
 * Error behaviour as the reference: a repeated half-edge does not fail the call; *ok = 0 and the
 * object holds exactly what the reference leaves behind when it returns false -- the triangles up
 * to and including the offending one, the half-edges inserted before the refused insertion. */

/* Cut faces = the faces the intersection touched (m_firstIntersectedFaces for which = 0,
 * m_secondIntersectedFaces for which = 1, src/solidboolean.cpp:318-319); no host round trip of
 * the flags. */
int sb_isect_uncut(const sb_isect *isect, int which, size_t vertex_offset, size_t triangle_offset,
                   sb_uncut **out);
/* Same with an explicit per-face flag array (host, nT bytes, non-zero = cut; NULL = keep all). */
int sb_mesh_uncut(const sb_mesh *mesh, const uint8_t *cut_flags, size_t vertex_offset,
                  size_t triangle_offset, sb_uncut **out);
void sb_uncut_destroy(sb_uncut *u);
/* n_triangles new triangles, n_half_edges map entries, *ok = the reference's return value. */
int sb_uncut_counts(const sb_uncut *u, size_t *n_triangles, size_t *n_half_edges, int *ok);
/* face: original face id per new triangle; tri3 (optional): its shifted index triple. */
int sb_uncut_triangles(const sb_uncut *u, uint32_t *face, uint32_t *tri3);
/* The map, sorted by key: keys[n_half_edges], owner[n_half_edges] (new triangle index). */
int sb_uncut_half_edges(const sb_uncut *u, uint64_t *keys, uint32_t *owner);
/* adj3[3 j + k] = owner of the half-edge opposite to edge k (vertices k, k+1 mod 3) of new
 * triangle j, or -1 if the map has none (the neighbour is a cut face or the mesh is open). */
int sb_uncut_adjacency(const sb_uncut *u, int32_t *adj3);
/* Face groups of the uncut triangles -- the flood fill of SolidBoolean::buildFaceGroups
 * (src/solidboolean.cpp:167-239) where no intersection loop fences it (:229-238), as connected
 * components of the adjacency above (lock-free union-find on the device): label[j] = lowest new
 * triangle index of j's component, i.e. the triangle the reference's loop over ascending indices
 * opens that group with; *n_components = number of groups.  SURVEY 8f row 3.  label may be NULL. */
int sb_uncut_components(const sb_uncut *u, uint32_t *label, size_t *n_components);
/* The whole flood of buildFaceGroups (src/solidboolean.cpp:167-239) for one mesh's side of the result: the uncut
 * triangles of `u` AND the retriangulated pieces the host produced (pieces: 3 x n_pieces vertex ids in the numbering of
 * the result vertex array, i.e. with u's vertex offset applied to original vertices).  fences: 2 x n_fences vertex ids,
 * the edges of the intersection loops (:176-203): a fill does not cross them in either direction.  Triangles joined through
 * opposite half-edges that are not fences form a group (connected components on the device, sb_flood.cu).  NODES are
 * numbered: uncut triangle j -> j (0 .. n_triangles-1, the order of sb_uncut_triangles), piece i -> n_triangles + i;
 * label_uncut[j] / label_piece[i] = the lowest node of the triangle's group; *n_groups = number of groups.
 * Where two loop seeds of the reference reach the same region its queue order splits the region between two groups on the
 * same side of every loop; a group here is the union of those (same triangles selected by every operation).
 * Not offered (SB_ERR_INVALID) when the half-edge map was truncated by a repeated half-edge.  SURVEY 8f row 3. */
int sb_uncut_face_groups(const sb_uncut *u, const uint32_t *pieces, size_t n_pieces, const uint32_t *fences, size_t n_fences,
                         uint32_t *label_uncut, uint32_t *label_piece, size_t *n_groups);
/* Device pointers of the same arrays (valid until sb_uncut_destroy) for device-side consumers;
 * any out pointer may be NULL.  After a repeated half-edge (*ok == 0) they hold the untruncated
 * arrays of ALL uncut faces. */
int sb_uncut_device_ptrs(const sb_uncut *u, void **face, void **tri3, void **keys, void **owner,
                         void **adj3);

/* ---- classification: replaces SolidBoolean::isPointInMesh
 * (src/solidboolean.cpp:48-92) as driven by decideGroupSide (:482-510): three
 * rays (g_testAxisList :31-35) per point, PositionKey de-duplication
 * (src/positionkey.cpp:32-37), odd/even, majority of three. */

/* inside: Q bytes.  per_axis (optional): 3 bytes per point. */
int sb_classify(const sb_mesh *target, const double *pts, size_t Q,
                uint8_t *inside, uint8_t *per_axis);
/* Query points = face centroids ((v0+v1)+v2)/3.0 of `query`'s triangles
 * (src/solidboolean.cpp:497-499), computed on the device. Outputs are indexed
 * by original triangle id of `query`. */
int sb_classify_faces(const sb_mesh *query, const sb_mesh *target,
                      uint8_t *inside, uint8_t *per_axis);
/* Device-resident form: classify query triangles at Morton-sorted positions
 * [begin, end) (begin a multiple of 32) -- the multi-GPU shard of SURVEY 8e.
 * d_inside is a caller-owned DEVICE buffer of nT(query) bytes indexed by
 * original triangle id; entries outside the range are left untouched.  No host
 * copy of the flags is made (one 40-byte counter read-back only). */
int sb_classify_faces_device(const sb_mesh *query, const sb_mesh *target,
                             size_t begin, size_t end, void *d_inside);

/* sb_mesh_upload for geometry that already lives on this device (same layouts; e.g. gathered from the
 * other ranks over NVLink): one device-to-device copy, ordered after the work enqueued on the context
 * stream so far. */
int sb_mesh_upload_device(sb_context *ctx, const void *d_xyz, size_t nV, const void *d_tri, size_t nT,
                          sb_mesh **out);

/* New coordinates (and / or index triples) of the SAME counts into an existing mesh -- the next frame of a
 * deforming mesh, or the next of a stream of same-sized inputs; NULL leaves that array as it is.  on_device:
 * the pointers are device pointers.  What was built from the old geometry is stale until sb_mesh_build;
 * a shard bound to the mesh stays valid (its next sb_shard_front_end plans again if the slabs moved). */
int sb_mesh_update(sb_mesh *m, const void *xyz, const void *tri, int on_device);

/* ---- multi-GPU shards (north_star: "mesh A's query triangles shard naturally") ---------
 * One process per GPU.  Every rank uploads (or receives) the geometry of both meshes -- sb_mesh_upload,
 * no build -- and makes a shard for its rank.  sb_shard_front_end then does the rank's share of one
 * front end: the faces of A and of B are dealt to the ranks by the z of their centroid (n equally
 * populated slabs, borders computed on the device, identical on every rank); the rank selects from both
 * meshes the triangles its queries can meet (the two lazy test rays run along x and y, so a slab is
 * closed under them), builds LBVH + ray grids over that selection only, and runs broad phase, predicate
 * and both classifications for its own faces.  Results use the triangle ids of the uploaded meshes:
 * the isect holds the rank's hit pairs (sorted) and segments -- sb_isect_counts / sb_isect_hits /
 * sb_isect_pack_device; candidates, contexts and uncut triangles are not offered on a shard's isect --
 * and d_insideA / d_insideB (caller-owned device arrays of nT(A) / nT(B) bytes) receive the flags of the
 * rank's own faces, the other bytes are left alone.  Summing the flag arrays and concatenating the hit
 * lists of all ranks gives the single-GPU result (the union over ranks is exact: every face belongs to
 * one rank).  A point whose first two votes disagree needs its third ray, along z, through the whole
 * target: the rank then builds the whole meshes and classifies its faces again (rare; counted). */
typedef struct sb_shard sb_shard;
int sb_shard_create(const sb_mesh *A, const sb_mesh *B, int rank, int n_ranks, sb_shard **out);
int sb_shard_front_end(sb_shard *s, unsigned flags, sb_isect **out, void *d_insideA, void *d_insideB);
int sb_shard_info(const sb_shard *s, size_t *selected_a, size_t *selected_b, double *z_lo, double *z_hi,
                  uint64_t *fallbacks);
void sb_shard_destroy(sb_shard *s);

/* ---- several GPUs from ONE host process (`sb_comm_create` + the multi-GPU front end) ------
 * What a C++ caller of SolidBoolean::combine() (src/solidboolean.cpp:288) needs to reach GPUs 1..N-1:
 * no MPI, no torch.distributed.  A comm owns one context per rank (devices[r], or device r when
 * devices is NULL; the same device may appear more than once), and per call one host thread per rank.
 * sb_comm_set_meshes hands both meshes to every rank (host buffers, same layouts as sb_mesh_upload; a
 * second call with the same counts only replaces the bytes).  sb_comm_front_end runs sb_shard_front_end
 * on every rank -- selection, build of the selected triangles, broad phase, predicate, both
 * classifications, each for the rank's own slab of faces -- and gathers: the per-face flags of all ranks
 * are OR-ed on rank 0's GPU (peer copies; every face belongs to one rank) and copied into insideA / insideB
 * (nT(A) / nT(B) bytes, either may be NULL); the hit pairs and segments of all ranks, merged in (a, b) order,
 * are then available from sb_comm_hits.  n_cand / n_hit = the totals over the ranks = the single-GPU counts.
 * replaces: src/solidboolean.cpp:94-122 (search + predicate loop) and :482-510 (decideGroupSide) across GPUs. */
typedef struct sb_comm sb_comm;
typedef struct sb_comm_rank_info {
    int device;
    size_t selected_a, selected_b; /* triangles of A / B the rank selected and built */
    double z_lo, z_hi;             /* its slab of face centroids */
    size_t candidates, hits;       /* its share of the pairs */
    uint64_t fallbacks;            /* front ends that needed the whole meshes (third ray) */
} sb_comm_rank_info;
int sb_comm_create(int n_ranks, const int *devices, sb_comm **out);
int sb_comm_size(const sb_comm *c);
int sb_comm_set_meshes(sb_comm *c, const double *xyzA, size_t nVA, const uint32_t *triA, size_t nTA,
                       const double *xyzB, size_t nVB, const uint32_t *triB, size_t nTB);
int sb_comm_front_end(sb_comm *c, unsigned flags, size_t *n_cand, size_t *n_hit, uint8_t *insideA, uint8_t *insideB);
int sb_comm_hits(const sb_comm *c, uint32_t *ab /* 2 n_hit */, double *seg /* 6 n_hit */);
int sb_comm_rank(const sb_comm *c, int rank, sb_comm_rank_info *out);
void sb_comm_destroy(sb_comm *c);

/* ---- batches of small booleans (BASELINE configs[4]: 1,000 x 5K-triangle jobs; SURVEY 8e "C5") --------
 * The reference runs one SolidBoolean per call (test/main.cpp:92-103); a 5K-triangle mesh cannot fill a
 * B200.  A BATCH MESH holds the meshes of n_jobs independent jobs end to end and goes through the same
 * entry points as a single mesh -- sb_mesh_build, sb_intersect, sb_classify_faces(_device), sb_front_end:
 * ONE sort, ONE leaf / grid build, ONE front-end launch sequence for all jobs.  Job j of batch A only ever
 * meets job j of batch B.  The jobs may (and usually do) overlap in space: inside the library the job
 * number leads the Morton key and moves the conservative float boxes / quantised grid coordinates onto a
 * 32 x 32 x 32 lattice; every exact test reads the real coordinates, so results are those of n_jobs
 * separate calls.
 *   xyz            AoS vertices of all jobs, job after job; vertex_start[n_jobs + 1] (first entry 0)
 *   tri            index triples of all jobs, JOB-LOCAL vertex indices; triangle_start[n_jobs + 1]
 *   lattice_pitch  >= 4 x the largest |coordinate| over BOTH batches of a pair (a power of two is a
 *                  good choice); both batches must be made with the same value.  Checked at build time.
 * Results are batch-global: triangle a of job j has index triangle_start[j] + a in hit / candidate pairs
 * and in the per-face flag arrays.  sb_batch_job_ranges splits the (sorted) hit list by job.
 * At most 1024 jobs per batch; explicit-point classification (sb_classify) is not offered for batches. */
int sb_batch_upload(sb_context *ctx, size_t n_jobs, const double *xyz, const size_t *vertex_start,
                    const uint32_t *tri, const size_t *triangle_start, double lattice_pitch, sb_mesh **out);
int sb_batch_info(const sb_mesh *m, size_t *n_jobs, size_t *vertex_start /* n_jobs + 1, or NULL */,
                  size_t *triangle_start /* n_jobs + 1, or NULL */);
/* hits of job j = entries [hit_start[j], hit_start[j + 1]) of sb_isect_hits; hit_start holds n_jobs + 1 */
int sb_batch_job_ranges(const sb_isect *x, size_t *hit_start);

/* ---- the whole front end of one boolean in one call ------------------------------
 * sb_intersect(A, B) on the context stream, overlapped with the classification of
 * A's faces against B and of B's faces against A on two internal streams (what
 * SolidBoolean::combine() needs from the GPU: src/solidboolean.cpp:292, 315-320 and
 * the isPointInMesh calls of :468-472 applied to every face).  d_insideA / d_insideB:
 * caller-owned DEVICE buffers of nT(A) / nT(B) bytes indexed by original triangle id.
 * The _range form restricts A's triangles (for the intersection and as queries) and
 * B's query triangles to Morton-sorted position ranges -- the multi-GPU shard. */
int sb_front_end(const sb_mesh *A, const sb_mesh *B, unsigned flags, sb_isect **out,
                 void *d_insideA, void *d_insideB);
int sb_front_end_range(const sb_mesh *A, const sb_mesh *B, size_t a_begin, size_t a_end,
                       size_t b_begin, size_t b_end, unsigned flags, sb_isect **out,
                       void *d_insideA, void *d_insideB);
/* The same with HOST outputs -- what a caller of SolidBoolean::combine() (src/solidboolean.cpp:288-349, :468-510)
 * consumes on the CPU: insideA / insideB receive nT(A) / nT(B) flag bytes, hit_ab / hit_seg (optional, both or
 * neither) the hit list as sb_isect_hits returns it when it has at most hit_capacity entries (otherwise they are
 * left untouched and sb_isect_hits fetches it; sb_isect_counts tells).  Every result is copied as soon as the
 * stream that produces it is done (A's flags while B's classification still runs, the hits while both do), so
 * the device-to-host traffic hides behind the kernels; pinned host memory keeps the copies asynchronous.
 * All buffers are complete when the call returns. */
int sb_front_end_host(const sb_mesh *A, const sb_mesh *B, unsigned flags, sb_isect **out,
                      uint8_t *insideA, uint8_t *insideB, uint32_t *hit_ab /* 2 x hit_capacity */,
                      double *hit_seg /* 6 x hit_capacity */, size_t hit_capacity);

/* ---- instrumentation --------------------------------------------------------
 * Stage timing with CUDA events on the context stream.  Stages accumulate the
 * device time of the kernels they enclose since the last reset. */
enum {
    SB_STAGE_BUILD = 0,    /* bounds, boxes, normals, Morton, sort, LBVH */
    SB_STAGE_BROAD = 1,    /* traversal + pair emission */
    SB_STAGE_NARROW = 2,   /* hit / candidate sorts + gather */
    SB_STAGE_CLASSIFY = 3, /* ray classification */
    SB_STAGE_PREDICATE = 4,/* the tri/tri predicate kernel alone (FP64 roofline) */
    SB_STAGE_HALFEDGE = 5, /* uncut-triangle compaction, half-edge sort, adjacency, face groups */
    SB_STAGE_CONTEXTS = 6, /* per-triangle intersection contexts */
    SB_STAGE_SHARD = 7,    /* multi-GPU: padded vertices / bounds of the parents, slab plan, selection */
    SB_STAGE_COUNT = 8
};
int sb_context_enable_timing(sb_context *ctx, int enable);
int sb_context_reset_timing(sb_context *ctx);
/* Synchronises the stream, then writes SB_STAGE_COUNT floats (ms) and the
 * number of kernel launches issued by this library since the last reset. */
int sb_context_get_timing(sb_context *ctx, float *ms, uint64_t *kernel_launches);
/* Work counters of the last classification on this context: rays traced and
 * ray/triangle candidates evaluated (roofline accounting). */
int sb_context_classify_stats(sb_context *ctx, uint64_t *rays, uint64_t *candidates);

#ifdef __cplusplus
}
#endif
#endif /* SOLIDBOOLEAN_B200_H */
