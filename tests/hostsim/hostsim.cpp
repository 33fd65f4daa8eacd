// TEST ONLY: compiles the device predicate / ray headers for the host
// (-DSB_HOST_SIM, -ffp-contract=off) so their logic can be compared with the
// reference on the CPU.  Never part of the product library.
#include "../../solidboolean_b200/csrc/sb_tritri.cuh"
#include "../../solidboolean_b200/csrc/sb_raytri.cuh"
#include <cstring>

extern "C" {

void hs_tri_tri_batch(const double *tris, size_t n, int32_t *ret, int32_t *coplanar, double *seg)
{
    for (size_t i = 0; i < n; ++i) {
        const double *v = tris + 18 * i;
        d3 p1 = {v[0], v[1], v[2]}, q1 = {v[3], v[4], v[5]}, r1 = {v[6], v[7], v[8]};
        d3 p2 = {v[9], v[10], v[11]}, q2 = {v[12], v[13], v[14]}, r2 = {v[15], v[16], v[17]};
        int cop = 0;
        d3 s = {0, 0, 0}, t = {0, 0, 0};
        ret[i] = tri_tri_intersection(p1, q1, r1, p2, q2, r2, cop, s, t);
        coplanar[i] = cop;
        seg[6 * i + 0] = s.x; seg[6 * i + 1] = s.y; seg[6 * i + 2] = s.z;
        seg[6 * i + 3] = t.x; seg[6 * i + 4] = t.y; seg[6 * i + 5] = t.z;
    }
}

// one ray against one triangle: rec = p(3) t0(3) t1(3) t2(3); out hit flag + key
void hs_ray_tri_batch(const double *rec, const int32_t *axis, size_t n, uint8_t *hitFlag, long long *keys, int filtered)
{
    for (size_t i = 0; i < n; ++i) {
        const double *v = rec + 12 * i;
        d3 p = {v[0], v[1], v[2]}, t0 = {v[3], v[4], v[5]}, t1 = {v[6], v[7], v[8]}, t2 = {v[9], v[10], v[11]};
        d3 nrm = tri_normal(t0, t1, t2);
        d3 end = ray_end(p, axis[i]);
        d3 hit = {0, 0, 0};
        bool h = filtered ? ray_tri_hit_filtered(p, end, t0, t1, t2, nrm, hit) : ray_tri_hit(p, end, t0, t1, t2, nrm, hit);
        hitFlag[i] = h ? 1 : 0;
        keys[3 * i] = h ? position_key(hit.x) : 0;
        keys[3 * i + 1] = h ? position_key(hit.y) : 0;
        keys[3 * i + 2] = h ? position_key(hit.z) : 0;
    }
}
}
