// fg_pairs.cuh -- fast O(N^2) pair loops of the tile kernel for fp32, uniform agents, N >= 32 (sm_100a).
//
// profiles/r01b_tile243_noobs: the scalar loops of k_step cost 10.3 (contact filter), 10.3 (collision
// filter) and 13 (symmetric Hausdorff) warp instructions per ordered pair and the kernel is ISSUE-bound
// (80 % issue-active, 87 M warp instructions per 1024 envs of 243 agents).  Here:
//   * partner data is kept as structure-of-arrays in shared memory (x[], y[], |p|^2[]), padded to a
//     multiple of 32 with neutral entries, and read four partners at a time with one broadcast LDS.128
//     per array;
//   * arithmetic uses Blackwell's packed fp32x2 instructions (FFMA2 / FADD2 / FMUL2: two partners per
//     issue slot) and the 3-input FMNMX3;
//   * the two FILTERS (contact cut-off on the old positions, core.py:304-312; reward collision on the
//     new ones, formation_hd_env.py:119-121) use |p_i - p_j|^2 - |p_i|^2 = |p_j|^2 - 2 p_i.p_j (two FMAs
//     per pair) against a threshold widened by a rounding margin, and flag GROUPS of four partners
//     (NaN-propagating minimum of the group, one compare).  They only select candidates: every flagged
//     group is re-tested by the caller with the reference's exact arithmetic, in ascending partner
//     order, so results do not depend on the filter;
//   * the Hausdorff minima (formation_hd_env.py:64-66) keep the exact (a-b)^2 form -- the reward needs
//     1e-5 absolute and the expanded form loses that when the formation error is small.
#pragma once
#include "fg_math.cuh"

namespace fg {

typedef unsigned long long u64;

__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 d; asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi)); return d;
}
__device__ __forceinline__ void unpk2(u64 v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
    u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
    u64 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
    u64 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {          // drops NaNs like fminf
    float d; asm("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ float fmin3_nan(float a, float b, float c) {      // NaN if any input is NaN
    float d; asm("min.NaN.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d;
}
__device__ __forceinline__ float fmin2_nan(float a, float b) {
    float d; asm("min.NaN.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b)); return d;
}

// Rounding margin of the expanded-form filters.  With n_i = |p_i|^2 the computed
// q_j = fma(-2x_i, x_j, fma(-2y_i, y_j, n_j)) differs from |p_i-p_j|^2 - n_i by at most
// 2^-24 * 8 * (n_i + n_j) (five roundings on terms bounded by 2(n_i + n_j)); 2^-20 leaves a factor 2.
__device__ __forceinline__ float filter_margin(float ni, float nmax) { return 9.5367431640625e-07f * (ni + nmax); }

// Candidate filter.  R: records of one env, one per group of four partners: {x[4], y[4], n[4]} (48 bytes; pads hold
// x = y = 0, n = +inf).  Bit g of the result is set when some partner j in [4g, 4g+4) may satisfy
// |p_i - p_j|^2 < limit (or is NaN).  thr = (limit - n_i) + filter_margin(n_i, max_j n_j).
// `groups` is a multiple of 8 and at most 64.
__device__ __forceinline__ u64 filter_groups(const ulonglong2* __restrict__ R, int groups, float px, float py,
                                             float thr) {
    const u64 ax = pk2(-2.f * px, -2.f * px), ay = pk2(-2.f * py, -2.f * py);
    u64 mask = 0;
    for (int c = 0; c < groups; c += 8) {
        unsigned m8 = 0;
        const ulonglong2* r = R + 3 * c;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const ulonglong2 x = r[3 * g], y = r[3 * g + 1], n = r[3 * g + 2];
            const u64 q01 = ffma2(ax, x.x, ffma2(ay, y.x, n.x));
            const u64 q23 = ffma2(ax, x.y, ffma2(ay, y.y, n.y));
            float q0, q1, q2, q3;
            unpk2(q01, q0, q1); unpk2(q23, q2, q3);
            const float m = fmin2_nan(fmin3_nan(q0, q1, q2), q3);
            m8 |= (!(m >= thr)) ? (1u << g) : 0u;                 // !(>=): NaN groups stay candidates
        }
        mask |= (u64)m8 << c;
    }
    return mask;
}

// Reward pass on the NEW state: collision-candidate filter on the centred positions (same scheme as
// above; |C_i - C_j| = |p_i - p_j| up to the rounding the margin covers; FILTER = false skips it) fused with
// the two Hausdorff minima in exact (a-b)^2 arithmetic:
//   rowmin = min_j |C_i - S_j|^2,   colmin = min_j |C_j - S_i|^2      (formation_hd_env.py:64-66)
// R: records {cx[4], cy[4], |c|^2[4], sx[4], sy[4]} (80 bytes) of the centred new positions and the centred ideal
// shape.  Pads: C pads x = y = 1e18, n = +inf;  S pads x = y = 1e18  (never a minimum, never a candidate).
template <bool FILTER>
__device__ __forceinline__ u64 reward_pass(const ulonglong2* __restrict__ R, int groups, float cx, float cy,
                                           float sx, float sy, float thr, float* rowmin_out, float* colmin_out) {
    const u64 ax = pk2(-2.f * cx, -2.f * cx), ay = pk2(-2.f * cy, -2.f * cy);
    const u64 ncx = pk2(-cx, -cx), ncy = pk2(-cy, -cy);            // S_j - C_i
    const u64 nsx = pk2(-sx, -sx), nsy = pk2(-sy, -sy);            // C_j - S_i
    float rowmin = INFINITY, colmin = INFINITY;
    u64 mask = 0;
    for (int c = 0; c < groups; c += 4) {
        unsigned m4 = 0;
        const ulonglong2* r = R + 5 * c;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const ulonglong2 x = r[5 * g], y = r[5 * g + 1];
            const ulonglong2 s = r[5 * g + 3], u = r[5 * g + 4];
            float a0, a1, a2, a3;
            // column minimum: partners C_j against my shape point S_i
            u64 dx = fadd2(x.x, nsx), dy = fadd2(y.x, nsy);
            u64 d01 = ffma2(dy, dy, fmul2(dx, dx));
            dx = fadd2(x.y, nsx); dy = fadd2(y.y, nsy);
            u64 d23 = ffma2(dy, dy, fmul2(dx, dx));
            unpk2(d01, a0, a1); unpk2(d23, a2, a3);
            colmin = fmin3(fmin3(colmin, a0, a1), a2, a3);
            // row minimum: partners S_j against my centred position C_i
            dx = fadd2(s.x, ncx); dy = fadd2(u.x, ncy);
            d01 = ffma2(dy, dy, fmul2(dx, dx));
            dx = fadd2(s.y, ncx); dy = fadd2(u.y, ncy);
            d23 = ffma2(dy, dy, fmul2(dx, dx));
            unpk2(d01, a0, a1); unpk2(d23, a2, a3);
            rowmin = fmin3(fmin3(rowmin, a0, a1), a2, a3);
            if (FILTER) {
                const ulonglong2 n = r[5 * g + 2];
                const u64 q01 = ffma2(ax, x.x, ffma2(ay, y.x, n.x));
                const u64 q23 = ffma2(ax, x.y, ffma2(ay, y.y, n.y));
                unpk2(q01, a0, a1); unpk2(q23, a2, a3);
                const float m = fmin2_nan(fmin3_nan(a0, a1, a2), a3);
                m4 |= (!(m >= thr)) ? (1u << g) : 0u;
            }
        }
        if (FILTER) mask |= (u64)m4 << c;
    }
    *rowmin_out = rowmin; *colmin_out = colmin;
    return mask;
}

// ---- hashed cell lists (one per env, in shared memory) ------------------------------------------------
// The two FILTERS above cost 3.2 + 2.5 of the ~12 warp instructions per ordered pair of the large-N step and feed
// the same FMA pipe as the Hausdorff minima.  A uniform grid with cell edge H = 2 x search radius replaces them:
// every agent is appended to the bucket of its cell (atomicExch on the bucket head -> singly linked chain of
// 16-byte nodes {x, y, next}), and a query walks the chains of the 2 x 2 cells its search disc can touch (own
// cell plus the neighbour on the side of the cell the agent sits in), testing every entry with the reference's
// exact arithmetic on the spot.  Buckets are hashed (positions are unbounded), so foreign entries are only
// rejected work.  A partner closer than the radius lies in one of the four cells because the cell index is a
// monotone function of the coordinate and H/2 carries a 2^-9 margin over the radius (index rounding error
// < 2^-12 cell units for |x| < 100).  Envs with a farther or non-finite agent are flagged and take the callers'
// exhaustive fall-back.
__device__ __forceinline__ unsigned cell_bucket(int cx, int cy, int shift) {
    return ((unsigned)cx * 0x9E3779B1u + (unsigned)cy * 0x85EBCA77u) >> shift;
}

struct CellPos { int cx, cy, ox, oy; };

__device__ __forceinline__ CellPos cell_insert(float x, float y, float inv_h, int shift, int* head, float4* node, int i,
                                               int* far_flag) {
    const bool ok = fabsf(x) < 100.f && fabsf(y) < 100.f;
    if (!ok) *far_flag = 1;
    const float fx = ok ? x * inv_h : 0.f, fy = ok ? y * inv_h : 0.f;
    const float flx = floorf(fx), fly = floorf(fy);
    CellPos c;
    c.cx = (int)flx; c.cy = (int)fly;
    c.ox = (fx - flx >= 0.5f) ? 1 : -1; c.oy = (fy - fly >= 0.5f) ? 1 : -1;
    const int prev = atomicExch(&head[cell_bucket(c.cx, c.cy, shift)], i);
    node[i] = make_float4(x, y, __int_as_float(prev), 0.f);
    return c;
}

// Contact cut-off (core.py:304-312): bit g set when a partner j != i in [4g, 4g+4) has !(|p_i - p_j|^2 >= cut2).
__device__ __forceinline__ u64 cell_near_groups(const int* __restrict__ head, const float4* __restrict__ node, int shift,
                                                CellPos c, float px, float py, float cut2, int i) {
    unsigned lo = 0, hi = 0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        int j = head[cell_bucket(c.cx + ((v & 1) ? c.ox : 0), c.cy + ((v & 2) ? c.oy : 0), shift)];
        while (j >= 0) {
            const float4 q = node[j];
            const float dx = px - q.x, dy = py - q.y;
            if (j != i && !(dx * dx + dy * dy >= cut2)) {
                const unsigned bit = 1u << ((j >> 2) & 31);
                if (j & 128) hi |= bit; else lo |= bit;
            }
            j = __float_as_int(q.z);
        }
    }
    return ((u64)hi << 32) | lo;
}

// Reward collisions (formation_hd_env.py:119-121): #{j != i : norm(p_j - p_i) < rthr}, sqrt-then-compare (Q18)
// behind a guarded squared test.  Two of the four cells may share a bucket, whose chain is then walked twice: an
// entry counts only in the walk of the cell it actually lies in.
__device__ __forceinline__ int cell_count_collisions(const int* __restrict__ head, const float4* __restrict__ node,
                                                     int shift, float inv_h, CellPos c, float px, float py,
                                                     float thr2_hi, float thr, int i) {
    int col = 0;
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int vx = c.cx + ((v & 1) ? c.ox : 0), vy = c.cy + ((v & 2) ? c.oy : 0);
        int j = head[cell_bucket(vx, vy, shift)];
        while (j >= 0) {
            const float4 q = node[j];
            const float dx = q.x - px, dy = q.y - py;
            if (j != i && dx * dx + dy * dy < thr2_hi) {
                if ((int)floorf(q.x * inv_h) == vx && (int)floorf(q.y * inv_h) == vy &&
                    sqrtf(dx * dx + dy * dy) < thr) ++col;
            }
            j = __float_as_int(q.z);
        }
    }
    return col;
}

}  // namespace fg
