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

// Candidate filter.  X, Y, Nn: partner arrays of one env (16-byte aligned, `groups`*4 entries, pads
// hold x = y = 0, n = +inf).  Bit g of the result is set when some partner j in [4g, 4g+4) may satisfy
// |p_i - p_j|^2 < limit (or is NaN).  thr = (limit - n_i) + filter_margin(n_i, max_j n_j).
// `groups` is a multiple of 8 and at most 64.
__device__ __forceinline__ u64 filter_groups(const float* __restrict__ X, const float* __restrict__ Y,
                                             const float* __restrict__ Nn, int groups, float px, float py,
                                             float thr) {
    const ulonglong2* X4 = reinterpret_cast<const ulonglong2*>(X);
    const ulonglong2* Y4 = reinterpret_cast<const ulonglong2*>(Y);
    const ulonglong2* N4 = reinterpret_cast<const ulonglong2*>(Nn);
    const u64 ax = pk2(-2.f * px, -2.f * px), ay = pk2(-2.f * py, -2.f * py);
    u64 mask = 0;
    for (int c = 0; c < groups; c += 8) {
        unsigned m8 = 0;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const ulonglong2 x = X4[c + g], y = Y4[c + g], n = N4[c + g];
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
// above; |C_i - C_j| = |p_i - p_j| up to the rounding the margin covers) fused with the two Hausdorff
// minima in exact (a-b)^2 arithmetic:
//   rowmin = min_j |C_i - S_j|^2,   colmin = min_j |C_j - S_i|^2      (formation_hd_env.py:64-66)
// CX, CY, NC: centred new positions and their squared norms; SX, SY: centred ideal shape.  Pads:
// C pads x = y = 1e18, n = +inf;  S pads x = y = 1e18  (never a minimum, never a candidate).
__device__ __forceinline__ u64 reward_pass(const float* __restrict__ CX, const float* __restrict__ CY,
                                           const float* __restrict__ NC, const float* __restrict__ SX,
                                           const float* __restrict__ SY, int groups, float cx, float cy,
                                           float sx, float sy, float thr, bool want_filter,
                                           float* rowmin_out, float* colmin_out) {
    const ulonglong2* CX4 = reinterpret_cast<const ulonglong2*>(CX);
    const ulonglong2* CY4 = reinterpret_cast<const ulonglong2*>(CY);
    const ulonglong2* NC4 = reinterpret_cast<const ulonglong2*>(NC);
    const ulonglong2* SX4 = reinterpret_cast<const ulonglong2*>(SX);
    const ulonglong2* SY4 = reinterpret_cast<const ulonglong2*>(SY);
    const u64 ax = pk2(-2.f * cx, -2.f * cx), ay = pk2(-2.f * cy, -2.f * cy);
    const u64 ncx = pk2(-cx, -cx), ncy = pk2(-cy, -cy);            // S_j - C_i
    const u64 nsx = pk2(-sx, -sx), nsy = pk2(-sy, -sy);            // C_j - S_i
    float rowmin = INFINITY, colmin = INFINITY;
    u64 mask = 0;
    for (int c = 0; c < groups; c += 4) {
        unsigned m4 = 0;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            const ulonglong2 x = CX4[c + g], y = CY4[c + g];
            const ulonglong2 s = SX4[c + g], u = SY4[c + g];
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
            if (want_filter) {
                const ulonglong2 n = NC4[c + g];
                const u64 q01 = ffma2(ax, x.x, ffma2(ay, y.x, n.x));
                const u64 q23 = ffma2(ax, x.y, ffma2(ay, y.y, n.y));
                unpk2(q01, a0, a1); unpk2(q23, a2, a3);
                const float m = fmin2_nan(fmin3_nan(a0, a1, a2), a3);
                m4 |= (!(m >= thr)) ? (1u << g) : 0u;
            }
        }
        mask |= (u64)m4 << c;
    }
    *rowmin_out = rowmin; *colmin_out = colmin;
    return mask;
}

}  // namespace fg
