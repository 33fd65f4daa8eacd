// Multi-GPU behind the C ABI (include/solidboolean_b200.h, sb_comm_*): ONE host
// process drives several GPUs, so that a C++ caller of SolidBoolean::combine() (reference
// src/solidboolean.cpp:288) reaches all of them without MPI / torch.distributed.  It is written on
// top of the public single-GPU entry points: one context, one pair of uploaded meshes and one
// shard (sb_shard_*) per rank, one host thread per rank for the duration of a call.
//
// Exchange step (the only communication): every rank's per-face flag bytes travel to rank 0's GPU
// with cudaMemcpyPeerAsync (NVLink where the GPUs are peers) and are OR-ed there by one small
// kernel -- each face is owned by exactly one rank, the others leave its byte zero; hit pairs and
// segments (a few thousand records) are read back per rank and merged on the host in (a, b) order,
// which is the order of the single-GPU result.
#include "../../include/solidboolean_b200.h"
#include "sb_internal.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

struct sb_comm {
    int n = 0;
    std::vector<int> dev;
    std::vector<sb_context *> ctx;
    std::vector<sb_mesh *> A, B;
    std::vector<sb_shard *> shard;
    std::vector<uint8_t *> dFlags; // per rank, on its device: nTA + nTB bytes
    uint8_t *dGather = nullptr;    // rank 0's device: (n - 1) x (nTA + nTB) bytes
    size_t flagBytes = 0;          // nTA + nTB rounded up to 16
    size_t nVA = 0, nTA = 0, nVB = 0, nTB = 0;
    bool haveMeshes = false;
    // result of the last front end
    size_t nCand = 0, nHit = 0;
    std::vector<uint32_t> hitAB;
    std::vector<double> hitSeg;
    std::vector<sb_comm_rank_info> info;
};

namespace {

__global__ void __launch_bounds__(256) or_bytes_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16, int parts, size_t stride16)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n16)
        return;
    uint4 a = dst[i];
    for (int p = 0; p < parts; ++p) {
        const uint4 b = src[p * stride16 + i];
        a.x |= b.x; a.y |= b.y; a.z |= b.z; a.w |= b.w;
    }
    dst[i] = a;
}

struct RankResult {
    int rc = SB_OK;
    std::string msg;
};

// runs fn(rank) on one host thread per rank (rank 0 on the calling thread); the first failure wins
template <typename F>
int for_each_rank(sb_comm *c, F &&fn)
{
    std::vector<RankResult> res(c->n);
    auto body = [&](int r) {
        cudaSetDevice(c->dev[r]);
        res[r].rc = fn(r);
        if (res[r].rc != SB_OK)
            res[r].msg = sb_last_error();
    };
    std::vector<std::thread> th;
    for (int r = 1; r < c->n; ++r)
        th.emplace_back(body, r);
    body(0);
    for (auto &t : th)
        t.join();
    for (int r = 0; r < c->n; ++r)
        if (res[r].rc != SB_OK) {
            char buf[600];
            snprintf(buf, sizeof(buf), "rank %d (device %d): %s", r, c->dev[r], res[r].msg.c_str());
            sbi_set_error(buf);
            return res[r].rc;
        }
    return SB_OK;
}

int cuda_fail(cudaError_t e, const char *what)
{
    char buf[256];
    snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    sbi_set_error(buf);
    return SB_ERR_CUDA;
}
#define COMM_CUDA(expr)                      \
    do {                                     \
        cudaError_t e_ = (expr);             \
        if (e_ != cudaSuccess)               \
            return cuda_fail(e_, #expr);     \
    } while (0)

void drop_meshes(sb_comm *c)
{
    for (int r = 0; r < c->n; ++r) {
        cudaSetDevice(c->dev[r]);
        if (c->shard[r]) sb_shard_destroy(c->shard[r]);
        if (c->A[r]) sb_mesh_destroy(c->A[r]);
        if (c->B[r]) sb_mesh_destroy(c->B[r]);
        if (c->dFlags[r]) cudaFree(c->dFlags[r]);
        c->shard[r] = nullptr;
        c->A[r] = c->B[r] = nullptr;
        c->dFlags[r] = nullptr;
    }
    if (c->dGather) {
        cudaSetDevice(c->dev[0]);
        cudaFree(c->dGather);
        c->dGather = nullptr;
    }
    c->haveMeshes = false;
}

} // namespace

extern "C" {

int sb_comm_create(int n_ranks, const int *devices, sb_comm **out)
{
    if (!out || n_ranks < 1 || n_ranks > 64) {
        sbi_set_error("sb_comm_create: bad arguments (1 <= n_ranks <= 64, out != NULL)");
        return SB_ERR_INVALID;
    }
    *out = nullptr;
    int prev = -1;
    cudaGetDevice(&prev);
    sb_comm *c = new (std::nothrow) sb_comm;
    if (!c) {
        sbi_set_error("out of host memory");
        return SB_ERR_NOMEM;
    }
    c->n = n_ranks;
    c->dev.resize(n_ranks);
    c->ctx.assign(n_ranks, nullptr);
    c->A.assign(n_ranks, nullptr);
    c->B.assign(n_ranks, nullptr);
    c->shard.assign(n_ranks, nullptr);
    c->dFlags.assign(n_ranks, nullptr);
    c->info.resize(n_ranks);
    for (int r = 0; r < n_ranks; ++r)
        c->dev[r] = devices ? devices[r] : r;
    for (int r = 0; r < n_ranks; ++r) {
        int rc = sb_context_create(c->dev[r], &c->ctx[r]); // fails loudly without a device: no CPU fallback
        if (rc != SB_OK) {
            const std::string msg = sb_last_error();
            sb_comm_destroy(c);
            sbi_set_error(msg.c_str());
            if (prev >= 0) cudaSetDevice(prev);
            return rc;
        }
    }
    // peers where the hardware allows it (NVLink / NVSwitch): the flag gather then goes GPU to GPU
    for (int r = 1; r < n_ranks; ++r)
        if (c->dev[r] != c->dev[0]) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, c->dev[0], c->dev[r]) == cudaSuccess && can) {
                cudaSetDevice(c->dev[0]);
                cudaError_t e = cudaDeviceEnablePeerAccess(c->dev[r], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    cudaGetLastError();
                else if (e == cudaErrorPeerAccessAlreadyEnabled)
                    cudaGetLastError();
            }
        }
    if (prev >= 0) cudaSetDevice(prev);
    *out = c;
    return SB_OK;
}

int sb_comm_size(const sb_comm *c) { return c ? c->n : 0; }

void sb_comm_destroy(sb_comm *c)
{
    if (!c)
        return;
    int prev = -1;
    cudaGetDevice(&prev);
    drop_meshes(c);
    for (int r = 0; r < c->n; ++r)
        if (c->ctx[r]) {
            cudaSetDevice(c->dev[r]);
            sb_context_destroy(c->ctx[r]);
        }
    if (prev >= 0) cudaSetDevice(prev);
    delete c;
}

int sb_comm_set_meshes(sb_comm *c, const double *xyzA, size_t nVA, const uint32_t *triA, size_t nTA, const double *xyzB, size_t nVB,
    const uint32_t *triB, size_t nTB)
{
    if (!c || (nVA && !xyzA) || (nTA && !triA) || (nVB && !xyzB) || (nTB && !triB)) {
        sbi_set_error("sb_comm_set_meshes: null argument");
        return SB_ERR_INVALID;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    const bool same = c->haveMeshes && nVA == c->nVA && nTA == c->nTA && nVB == c->nVB && nTB == c->nTB;
    int rc;
    if (same) {
        // the next frame of the same sizes: new bytes into the existing meshes, the shards stay bound
        rc = for_each_rank(c, [&](int r) {
            int e = sb_mesh_update(c->A[r], xyzA, triA, 0);
            return e != SB_OK ? e : sb_mesh_update(c->B[r], xyzB, triB, 0);
        });
    } else {
        drop_meshes(c);
        c->nVA = nVA; c->nTA = nTA; c->nVB = nVB; c->nTB = nTB;
        c->flagBytes = (nTA + nTB + 15) / 16 * 16;
        rc = for_each_rank(c, [&](int r) {
            int e = sb_mesh_upload(c->ctx[r], xyzA, nVA, triA, nTA, &c->A[r]);
            if (e == SB_OK) e = sb_mesh_upload(c->ctx[r], xyzB, nVB, triB, nTB, &c->B[r]);
            if (e == SB_OK) e = sb_shard_create(c->A[r], c->B[r], r, c->n, &c->shard[r]);
            if (e != SB_OK)
                return e;
            COMM_CUDA(cudaMalloc(&c->dFlags[r], std::max<size_t>(c->flagBytes, 16)));
            if (r == 0 && c->n > 1)
                COMM_CUDA(cudaMalloc(&c->dGather, std::max<size_t>(c->flagBytes, 16) * (size_t)(c->n - 1)));
            return SB_OK;
        });
        c->haveMeshes = rc == SB_OK;
        if (rc != SB_OK) {
            const std::string msg = sb_last_error();
            drop_meshes(c);
            sbi_set_error(msg.c_str());
        }
    }
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int sb_comm_front_end(sb_comm *c, unsigned flags, size_t *n_cand, size_t *n_hit, uint8_t *insideA, uint8_t *insideB)
{
    if (!c || !c->haveMeshes) {
        sbi_set_error("sb_comm_front_end: no meshes (sb_comm_set_meshes first)");
        return SB_ERR_INVALID;
    }
    int prev = -1;
    cudaGetDevice(&prev);
    const size_t nTA = c->nTA, nTB = c->nTB, fb = std::max<size_t>(c->flagBytes, 16);
    std::vector<std::vector<uint32_t>> ab(c->n);
    std::vector<std::vector<double>> seg(c->n);
    std::vector<size_t> cand(c->n, 0);
    int rc = for_each_rank(c, [&](int r) {
        COMM_CUDA(cudaMemset(c->dFlags[r], 0, fb));
        sb_isect *x = nullptr;
        int e = sb_shard_front_end(c->shard[r], flags, &x, c->dFlags[r], c->dFlags[r] + nTA);
        if (e != SB_OK)
            return e;
        size_t nc = 0, nh = 0;
        e = sb_isect_counts(x, &nc, &nh);
        if (e == SB_OK) {
            cand[r] = nc;
            ab[r].resize(2 * nh);
            seg[r].resize(6 * nh);
            if (nh)
                e = sb_isect_hits(x, ab[r].data(), seg[r].data());
        }
        sb_isect_destroy(x);
        if (e != SB_OK)
            return e;
        sb_comm_rank_info &ri = c->info[r];
        ri.device = c->dev[r];
        uint64_t fbk = 0;
        e = sb_shard_info(c->shard[r], &ri.selected_a, &ri.selected_b, &ri.z_lo, &ri.z_hi, &fbk);
        ri.fallbacks = fbk;
        ri.candidates = nc;
        ri.hits = nh;
        if (e != SB_OK)
            return e;
        COMM_CUDA(cudaDeviceSynchronize());
        // the exchange: this rank's flag bytes to rank 0's GPU
        if (r > 0)
            COMM_CUDA(cudaMemcpyPeer(c->dGather + (size_t)(r - 1) * fb, c->dev[0], c->dFlags[r], c->dev[r], fb));
        return SB_OK;
    });
    if (rc == SB_OK) {
        cudaSetDevice(c->dev[0]);
        rc = [&]() -> int {
            if (c->n > 1) {
                const size_t n16 = fb / 16;
                or_bytes_kernel<<<(unsigned)((n16 + 255) / 256), 256>>>(reinterpret_cast<uint4 *>(c->dFlags[0]),
                    reinterpret_cast<const uint4 *>(c->dGather), n16, c->n - 1, n16);
                COMM_CUDA(cudaGetLastError());
            }
            if (insideA && nTA)
                COMM_CUDA(cudaMemcpy(insideA, c->dFlags[0], nTA, cudaMemcpyDeviceToHost));
            if (insideB && nTB)
                COMM_CUDA(cudaMemcpy(insideB, c->dFlags[0] + nTA, nTB, cudaMemcpyDeviceToHost));
            COMM_CUDA(cudaDeviceSynchronize());
            return SB_OK;
        }();
    }
    if (rc == SB_OK) {
        // hit lists: every rank's is sorted by (a, b) and the ranks own disjoint pairs -> merge
        size_t H = 0;
        c->nCand = 0;
        for (int r = 0; r < c->n; ++r) {
            H += ab[r].size() / 2;
            c->nCand += cand[r];
        }
        std::vector<std::pair<unsigned long long, std::pair<int, uint32_t>>> keys;
        keys.reserve(H);
        for (int r = 0; r < c->n; ++r)
            for (size_t i = 0; i < ab[r].size() / 2; ++i)
                keys.push_back({((unsigned long long)ab[r][2 * i] << 32) | ab[r][2 * i + 1], {r, (uint32_t)i}});
        std::sort(keys.begin(), keys.end());
        c->nHit = H;
        c->hitAB.resize(2 * H);
        c->hitSeg.resize(6 * H);
        for (size_t k = 0; k < H; ++k) {
            const int r = keys[k].second.first;
            const uint32_t i = keys[k].second.second;
            c->hitAB[2 * k] = ab[r][2 * i];
            c->hitAB[2 * k + 1] = ab[r][2 * i + 1];
            memcpy(&c->hitSeg[6 * k], &seg[r][6 * (size_t)i], 48);
        }
        if (n_cand) *n_cand = c->nCand;
        if (n_hit) *n_hit = c->nHit;
    }
    if (prev >= 0) cudaSetDevice(prev);
    return rc;
}

int sb_comm_hits(const sb_comm *c, uint32_t *ab, double *seg)
{
    if (!c) {
        sbi_set_error("sb_comm_hits: null comm");
        return SB_ERR_INVALID;
    }
    if (ab && c->nHit)
        memcpy(ab, c->hitAB.data(), sizeof(uint32_t) * 2 * c->nHit);
    if (seg && c->nHit)
        memcpy(seg, c->hitSeg.data(), sizeof(double) * 6 * c->nHit);
    return SB_OK;
}

int sb_comm_rank(const sb_comm *c, int rank, sb_comm_rank_info *out)
{
    if (!c || !out || rank < 0 || rank >= c->n) {
        sbi_set_error("sb_comm_rank: bad arguments");
        return SB_ERR_INVALID;
    }
    *out = c->info[rank];
    return SB_OK;
}

} // extern "C"
