// fg_policy.cuh -- the reference's hand-written controller on the device (sm_100a):
// ezpolicy (formation_gym/__init__.py:19-47) expanded hierarchically by get_action_BFS
// (formation_gym/__init__.py:49-98), for every env of a batch in one launch.
//
// The reference walks per-agent observation lists on the host; every value it slices out of an
// observation is pos[b] - pos[a], ideal_shape or ideal_vel (formation_hd_env.py:52-59; p_vel is read but
// unused), so the kernel works from the env STATE and never touches the [E,N,6N] observation tensor:
// 16N + 8 bytes read and 8N written per env instead of 24N^2 read.
//
// Mapping: one CTA owns EPC = floor(256 / N) envs, thread <-> (local env, agent).  The BFS tree has
// `levels` = log_n(N) layers; at the layer with groups of M agents, the first agent of each of the n
// subgroups of a group (the "leader", :61) builds the n-agent layer observation from subgroup centroids
// (:65-74), runs ezpolicy, scales by the layer number (:78-79) and hands the result down as the subgroup's
// target velocity (:84-97) through shared memory; the last layer writes the actions (:81-83).
#pragma once
#include "fg_math.cuh"

namespace fg {

constexpr int kPolicyMaxFan = 8;        // agents per layer n <= 8
constexpr int kPolicyMaxLevels = 8;

template <typename T> struct PArgs {
    typedef typename Ops<T>::R2 R2;
    const R2* pos; const R2* shape; const R2* ivel; R2* act;
    int E, N, n, levels, EPC;
    uint32_t magic_n;
    T mult[kPolicyMaxLevels];           // log(M)/log(n) per layer, top first (:78)
    // per layer (top first): group size M, subgroup size nxt = M / n, leaders per env N / nxt, and the fastdiv
    // magics of the three divisors (runtime integer divisions were 25 % of the kernel's instructions)
    int lev_M[kPolicyMaxLevels], lev_nxt[kPolicyMaxLevels], lev_nlead[kPolicyMaxLevels];
    uint32_t mg_M[kPolicyMaxLevels], mg_nxt[kPolicyMaxLevels], mg_nlead[kPolicyMaxLevels];
};

// formation_gym/__init__.py:19-47 on the sliced inputs: others [n-1] (other_pos), tgt [n] (ideal_shape
// before re-centring), tvel (ideal_vel).
// NF > 0: compile-time fan-out (loops unrolled, arrays in registers); NF == 0: runtime n <= 8.
template <typename T, int NF>
__device__ __forceinline__ typename Ops<T>::R2 ezpolicy_dev(const typename Ops<T>::R2* others,
                                                            const typename Ops<T>::R2* tgt,
                                                            typename Ops<T>::R2 tvel, int n_rt) {
    typedef Ops<T> O;
    typedef typename O::R2 R2;
    constexpr int CAP = NF > 0 ? NF : kPolicyMaxFan;
    const int n = NF > 0 ? NF : n_rt;
    R2 ideal[CAP], cur[CAP];
    T d[CAP];
    // ideal_shape - mean (:28); current_shape = [other_pos, (0,0)] - mean (:31-33): sums in row order
    T sx = 0, sy = 0, cx = 0, cy = 0;
#pragma unroll
    for (int k = 0; k < CAP; ++k) {
        if (k < n) {
            sx = O::add(sx, tgt[k].x); sy = O::add(sy, tgt[k].y);
            const R2 c = (k < n - 1) ? others[k] : O::make((T)0, (T)0);
            cur[k] = c;
            cx = O::add(cx, c.x); cy = O::add(cy, c.y);
        }
    }
    sx = O::div_count(sx, n); sy = O::div_count(sy, n); cx = O::div_count(cx, n); cy = O::div_count(cy, n);
#pragma unroll
    for (int k = 0; k < CAP; ++k) {
        if (k < n) {
            ideal[k] = O::make(O::sub(tgt[k].x, sx), O::sub(tgt[k].y, sy));
            cur[k] = O::make(O::sub(cur[k].x, cx), O::sub(cur[k].y, cy));
        }
    }
    const R2 self = cur[n - 1];
    // Distances from me to every landmark (:35).  np.argsort's small-n path is an insertion sort (stable), so
    // landmark k has rank = #{j : d[j] < d[k] or (d[j] == d[k] and j < k)}; walking the ranks in order is the
    // reference's `for idx in sort_mark_idx` without an index array.
#pragma unroll
    for (int k = 0; k < CAP; ++k)
        if (k < n) d[k] = O::lenkey(O::sub(self.x, ideal[k].x), O::sub(self.y, ideal[k].y));
    // rank of every landmark in the stable ascending order of d (computed once, not once per candidate)
    int rank[CAP];
#pragma unroll
    for (int k = 0; k < CAP; ++k) {
        int rk = 0;
        if (k < n) {
#pragma unroll
            for (int j = 0; j < CAP; ++j)
                if (j < n) rk += (d[j] < d[k] || (d[j] == d[k] && j < k)) ? 1 : 0;
        }
        rank[k] = rk;
    }
    R2 act = O::make((T)0, (T)0);
    bool found = false;
#pragma unroll
    for (int r = 0; r < CAP; ++r) {                                           // :36-40
        if (r < n && !found) {
            R2 tg = ideal[0];
#pragma unroll
            for (int k = 1; k < CAP; ++k) if (k < n && rank[k] == r) tg = ideal[k];
            int closest = 0;
            T best = O::lenkey(O::sub(cur[0].x, tg.x), O::sub(cur[0].y, tg.y));
#pragma unroll
            for (int m = 1; m < CAP; ++m) {
                if (m < n) {
                    const T dm = O::lenkey(O::sub(cur[m].x, tg.x), O::sub(cur[m].y, tg.y));
                    if (dm < best) { best = dm; closest = m; }                // np.argmin: first minimum
                }
            }
            if (closest == n - 1 || r == n - 1) {
                T ax = O::mul((T)0.5, O::sub(tg.x, self.x));
                T ay = O::mul((T)0.5, O::sub(tg.y, self.y));
                act = O::make(fmin(fmax(ax, (T)-1), (T)1), fmin(fmax(ay, (T)-1), (T)1));   // np.clip
                found = true;
            }
        }
    }
    // done = ||ideal_shape - current_shape||_F < 0.01 (:42)
    T fro = 0;
#pragma unroll
    for (int k = 0; k < CAP; ++k) {
        if (k < n) {
            const T ex = O::sub(ideal[k].x, cur[k].x), ey = O::sub(ideal[k].y, cur[k].y);
            fro = O::add(fro, O::add(O::mul(ex, ex), O::mul(ey, ey)));
        }
    }
    const bool done = O::sqrt_(fro) < (T)0.01;
    const T g = done ? (T)1 : (T)0.3;                                         // :43-46
    return O::make(O::add(act.x, done ? tvel.x : O::mul(tvel.x, g)), O::add(act.y, done ? tvel.y : O::mul(tvel.y, g)));
}

template <int B, int E_> struct CPow { static constexpr int v = B * CPow<B, E_ - 1>::v; };
template <int B> struct CPow<B, 0> { static constexpr int v = 1; };

// One node of the BFS tree (formation_gym/__init__.py:61-79): leader `i` of a subgroup of `nxt` agents inside the
// group that starts at agent `gb` (subgroup index `si` within the group) builds the n-agent layer observation from the
// subgroup centroids, runs ezpolicy on it with the group's target velocity `tv` and scales by the layer number.
// P, S: the env's positions and ideal shape (shared memory).  NXT_ > 0: compile-time subgroup size (loops unrolled).
template <typename T, int NF, int NXT_>
__device__ __forceinline__ typename Ops<T>::R2 policy_node(const typename Ops<T>::R2* P, const typename Ops<T>::R2* S,
                                                           int i, int gb, int si, int n_rt, int nxt_rt,
                                                           typename Ops<T>::R2 tv, T mult) {
    typedef Ops<T> O;
    typedef typename O::R2 R2;
    constexpr int CAP = NF > 0 ? NF : kPolicyMaxFan;
    constexpr bool CT = NXT_ > 0;
    const int n = NF > 0 ? NF : n_rt;
    const int nxt = CT ? NXT_ : nxt_rt;
    const R2 pi = P[i];
    R2 cur[CAP], tgt[CAP], others[CAP];
#pragma unroll
    for (int k = 0; k < CAP; ++k) {
        if (k >= n) continue;
        // centroid of subgroup k in my frame (:65-66) and of its target points (:70-71): np.mean sums
        // the rows in order and divides by the count
        T cx = 0, cy = 0, tx = 0, ty = 0;
        const int b0 = gb + k * nxt;
#pragma unroll CT ? 9 : 1
        for (int b = b0; b < b0 + nxt; ++b) {
            const R2 q2 = P[b];
            const T rx = (b == i) ? (T)0 : O::sub(q2.x, pi.x);            // own slot is the inserted (0,0)
            const T ry = (b == i) ? (T)0 : O::sub(q2.y, pi.y);
            cx = O::add(cx, rx); cy = O::add(cy, ry);
            const R2 sb = S[b];
            tx = O::add(tx, sb.x); ty = O::add(ty, sb.y);
        }
        cur[k] = O::make(O::div_count(cx, nxt), O::div_count(cy, nxt));
        tgt[k] = O::make(O::div_count(tx, nxt), O::div_count(ty, nxt));
    }
    R2 own = cur[0];                                                      // :67-68
#pragma unroll
    for (int k = 1; k < CAP; ++k) if (k < n && k == si) own = cur[k];
#pragma unroll
    for (int k = 0; k < CAP - 1; ++k) {                                   // np.delete(cur - cur[si], si, 0)
        if (k < n - 1) {
            const R2 c = (k < si) ? cur[k] : cur[k + 1];
            others[k] = O::make(O::sub(c.x, own.x), O::sub(c.y, own.y));
        }
    }
    R2 out = ezpolicy_dev<T, NF>(others, tgt, tv, n);                     // :76-79
    return O::make(O::mul(out.x, mult), O::mul(out.y, mult));
}

// The chain of nodes above agent `i`, top layer first, for a compile-time tree N = NF^LV: at every layer the agent
// evaluates the node of ITS subgroup's leader (the agents of a subgroup compute the same value redundantly instead of
// one of them computing it and handing it down -- no exchange, no divergence), and the last layer's node is its own
// action (:81-97).  Used by the warp-autonomous step kernel (fg_warp.cuh, POL), where lane <-> agent.
template <typename T, int NF, int LV, int L = 0>
__device__ __forceinline__ typename Ops<T>::R2 policy_chain(const typename Ops<T>::R2* P, const typename Ops<T>::R2* S,
                                                            int i, typename Ops<T>::R2 tv, const T* mult) {
    constexpr int M = CPow<NF, LV - L>::v, NXT = M / NF;
    const int il = (i / NXT) * NXT;                                       // the leader of my subgroup in this layer (:61)
    const int gb = (il / M) * M, si = (il - gb) / NXT;
    const typename Ops<T>::R2 out = policy_node<T, NF, NXT>(P, S, il, gb, si, NF, NXT, tv, mult[L]);
    if constexpr (L + 1 < LV) return policy_chain<T, NF, LV, L + 1>(P, S, i, out, mult);
    else return out;
}

// One BFS layer for the envs of a CTA.  M, NXT > 0: compile-time group / subgroup sizes (loops unrolled, divisions by
// constants); M == 0: run-time sizes from the argument block with fastdiv magics.
template <typename T, int NF, int N_, int M_, int NXT_>
__device__ __forceinline__ void policy_layer(const PArgs<T>& a, int l, int nvalid, const typename Ops<T>::R2* s_p,
                                             const typename Ops<T>::R2* s_s, const typename Ops<T>::R2* s_tv0,
                                             typename Ops<T>::R2* s_tv1) {
    typedef Ops<T> O;
    typedef typename O::R2 R2;
    constexpr bool CT = M_ > 0;
    const int N = CT ? N_ : a.N, n = NF > 0 ? NF : a.n;
    const int M = CT ? M_ : a.lev_M[l], nxt = CT ? NXT_ : a.lev_nxt[l];
    const int nlead = CT ? N_ / (NXT_ > 0 ? NXT_ : 1) : a.lev_nlead[l];       // leaders per env in this layer (:61)
    const int t = threadIdx.x;
    // leaders are COMPACTED onto consecutive threads (upper layers have few of them: N / nxt per env), so
    // a layer costs ceil(envs * leaders / 32) warps instead of every warp of the CTA at 1/nxt efficiency
    for (int q = t; q < nvalid * nlead; q += 256) {
        const int qe = CT ? q / nlead : (int)fastdiv((uint32_t)q, a.mg_nlead[l]);
        const int i = (q - qe * nlead) * nxt;
        const R2* P = s_p + qe * N;
        const R2* S = s_s + qe * N;
        const int gb = CT ? (i / M) * M : (int)fastdiv((uint32_t)i, a.mg_M[l]) * M;      // first agent of my group
        const int si = CT ? (i - gb) / nxt : (int)fastdiv((uint32_t)(i - gb), a.mg_nxt[l]);   // my subgroup within the group
        const R2 out = policy_node<T, NF, CT ? NXT_ : 0>(P, S, i, gb, si, n, nxt, s_tv0[qe * N + i], a.mult[l]);
        if (nxt == 1) a.act[((size_t)blockIdx.x * a.EPC + qe) * N + i] = out;              // :81-83
        else for (int b = 0; b < nxt; ++b) s_tv1[qe * N + i + b] = out;       // tar_vel of my subgroup (:84-97)
    }
}

// LV > 0: the tree shape is a compile-time constant (N = NF^LV; every layer's loops unroll and its index divisions are
// by constants -- the run-time version spent a quarter of its instructions on them and on loop control); LV == 0: any
// N = n^levels from the argument block.
// (forcing 7 or 8 resident CTAs per SM -- 36 / 32 registers instead of 40 -- changes nothing: 23.2 us per 131072 x 9)
template <typename T, int NF, int LV = 0>
__global__ void __launch_bounds__(256) k_policy_bfs(const __grid_constant__ PArgs<T> a) {
    typedef Ops<T> O;
    typedef typename O::R2 R2;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NC = LV > 0 ? CPow<NF, (LV > 0 ? LV : 0)>::v : 0;
    const int N = LV > 0 ? NC : a.N, nA = a.EPC * N;
    R2* s_p = reinterpret_cast<R2*>(smem_raw);
    R2* s_s = s_p + nA;
    R2* s_tv0 = s_s + nA;                 // target velocity of the group each agent is in (ping-pong)
    R2* s_tv1 = s_tv0 + nA;
    const int t = threadIdx.x;
    const int le = LV > 0 ? t / NC : (int)fastdiv((uint32_t)t, a.magic_n);
    const int i = t - le * N;
    const int e = blockIdx.x * a.EPC + le;
    const bool active = t < nA && e < a.E;
    const size_t g = (size_t)e * N + i;
    if (active) {
        s_p[t] = a.pos[g];
        s_s[t] = a.shape[g];
        s_tv0[t] = a.ivel[e];                                                 // top layer: the env's ideal_vel (:74)
    }
    __syncthreads();
    const int nvalid = min(a.EPC, a.E - blockIdx.x * a.EPC);
    if constexpr (LV > 0) {
        policy_layer<T, NF, NC, CPow<NF, LV>::v, CPow<NF, LV - 1>::v>(a, 0, nvalid, s_p, s_s, s_tv0, s_tv1);
        if constexpr (LV >= 2) {
            __syncthreads();
            policy_layer<T, NF, NC, CPow<NF, LV - 1>::v, CPow<NF, LV - 2>::v>(a, 1, nvalid, s_p, s_s, s_tv1, s_tv0);
        }
        if constexpr (LV >= 3) {
            __syncthreads();
            policy_layer<T, NF, NC, CPow<NF, LV - 2>::v, CPow<NF, LV - 3>::v>(a, 2, nvalid, s_p, s_s, s_tv0, s_tv1);
        }
        if constexpr (LV >= 4) {
            __syncthreads();
            policy_layer<T, NF, NC, CPow<NF, LV - 3>::v, CPow<NF, LV - 4>::v>(a, 3, nvalid, s_p, s_s, s_tv1, s_tv0);
        }
        if constexpr (LV >= 5) {
            __syncthreads();
            policy_layer<T, NF, NC, CPow<NF, LV - 4>::v, CPow<NF, LV - 5>::v>(a, 4, nvalid, s_p, s_s, s_tv0, s_tv1);
        }
        static_assert(LV <= 5, "instantiate more layers");
    } else {
        for (int l = 0; l < a.levels; ++l) {
            policy_layer<T, NF, 0, 0, 0>(a, l, nvalid, s_p, s_s, s_tv0, s_tv1);
            __syncthreads();
            R2* tmp = s_tv0; s_tv0 = s_tv1; s_tv1 = tmp;
        }
    }
}

}  // namespace fg
