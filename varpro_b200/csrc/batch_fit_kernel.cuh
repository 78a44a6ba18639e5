// batch_fit_kernel.cuh -- K4: independent-batch evaluator + fit (BASELINE config 3).
//
// P independent problems share the model structure, the independent variable x and the weights,
// but each has its OWN observations y_p, nonlinear parameters alpha_p and linear coefficients c_p:
// what a user of the reference writes as a loop over SingleRhs problems,
//   for p in 0..P { LevMarSolver::default().fit(SeparableProblemBuilder::new(model_p).observations(y_p).build()) }
// (src/problem/builder.rs:116-324, src/solvers/levmar/mod.rs:238-254 per problem).
//
// One CTA keeps G problems in flight (see below) and takes the next one from a global counter whenever
// a fit ends (fits need different numbers of evaluations): y_p is read from HBM ONCE per fit (then L2);
// per evaluation of a problem the compute warps of the CTA
//   1. regenerate Phi_w(alpha_p), D(alpha_p) from x into a shared-memory working matrix (m x (n+p))
//      -- materialising Phi for 65 536 problems would take 6.4 GB --
//   2. run n Householder steps on [Phi_w | D | y] (y carried along in registers). EVERY ROW OF THE
//      WORKING SYSTEM BELONGS TO ONE THREAD, so the only thing that crosses threads is the reduction
//      of the reflector's dot products: the sweep that applies reflector J also accumulates the dots
//      reflector J + 1 needs, a finished row is saved and then zeroed in place and the padding rows
//      hold zeros, so the sweeps carry no row predicates at all,
//   3. the last sweep accumulates everything the LM step needs over the rows >= n of the rotated
//      system:  ||r||^2 = sum y~_i^2,  u_e = D~_e . y~,  M_ef = D~_e . D~_f   and the saved top rows give
//      c = R1^-1 y~[0:n];  then  g_k = -sum_{e in k} c_j(e) u_e,  H_kl = sum M_ef c_j(e) c_j(f)
//      (the S = 1 case of the formulas in stream_kernel.cuh; no explicit Q or E is formed),
//   4. a lane of the LM warp advances the lmder state machine (lm_step.cuh) in shared memory.
// n + 1 block reductions per evaluation (warp level: recursive halving, ~K instead of 5 K shuffles
// for K values; the reflector scalars are formed once, by warp 0, between the two barriers).
// Bound: fp64 ALU / exp and block-reduction latency, not HBM (32 KB of y per fit against ~15
// evaluations of ~0.5 MFLOP each).
#pragma once

#include <type_traits>
#include <utility>

#include "device_common.cuh"
#include "lm_step.cuh"
#include "panel_kernel_hh.cuh"

namespace vp {

constexpr unsigned long long BATCH_SPIN_TIMEOUT_NS = 4000000000ull; // a lost partner must not hang the GPU
__host__ __device__ constexpr int pow2_ceil(int k) { int p = 1; while (p < k) p <<= 1; return p; }
__host__ __device__ constexpr int log2_int(int k) { int l = 0; while ((1 << l) < k) ++l; return l; }

template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F &&f, std::integer_sequence<int, Is...>)
{
    (f(std::integral_constant<int, Is>{}), ...);
}
// f(integral_constant<int, 0>) ... f(integral_constant<int, N - 1>): a loop whose index is a constant expression
template <int N, class F>
__device__ __forceinline__ void static_for(F &&f)
{
    static_for_impl(static_cast<F &&>(f), std::make_integer_sequence<int, N>{});
}

// ---- named barriers for the compute warps (the LM warps of the CTA never join them) -----------------------
__device__ __forceinline__ void bar_sync_n(const int id, const int count)
{
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ int bar_or_n(const int id, const int count, const int pred)
{
    int r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.s32 p, %3, 0;\n\t"
        "bar.red.or.pred q, %1, %2, p;\n\t"
        "selp.s32 %0, 1, 0, q;\n\t}"
        : "=r"(r)
        : "r"(id), "r"(count), "r"(pred)
        : "memory");
    return r;
}
// Hand-off flags in shared memory: atomic accesses (what compute-sanitizer's racecheck recognises as
// synchronisation), a block-level fence before the publishing store and after the observing load.
__device__ __forceinline__ int ld_flag(int *p) { return atomicOr(p, 0); }
__device__ __forceinline__ void st_flag(int *p, const int v)
{
    __threadfence_block();
    atomicExch(p, v);
}

// Warp-level sum of KP (a power of two <= 32) values per lane by recursive halving: at the level with lane
// offset o a lane keeps one half of its values and hands the other half to its partner. Afterwards every lane
// holds the warp total of value index (lane >> (5 - log2 KP)). KP - 1 + (5 - log2 KP) 64-bit shuffles instead
// of 5 KP.
template <int KP>
__device__ __forceinline__ double warp_sum_scatter(double (&v)[KP])
{
    static_assert(KP >= 1 && KP <= 32 && (KP & (KP - 1)) == 0, "power of two");
    const int lane = threadIdx.x & 31;
    int o = 16;
#pragma unroll
    for (int h = KP / 2; h >= 1; h >>= 1, o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < h; ++i) {
            const double keep = up ? v[i + h] : v[i];
            const double send = up ? v[i] : v[i + h];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    double t = v[0];
#pragma unroll
    for (int oo = 16 / KP; oo > 0; oo >>= 1) t += __shfl_xor_sync(0xffffffffu, t, oo);
    return t;
}

// First half of a block sum of K values over the NTHREADS compute threads (named barrier 1): per-warp partials
// to buf[warp * KP + k] and ONE barrier, which also
// ORs `flags` over the CTA (the return value). Second half: warp 0, lane k < K, adds the NW partials.
template <int K, int NTHREADS>
__device__ __forceinline__ int block_sum_begin(const double *s, double *buf, int flags)
{
    constexpr int KP = pow2_ceil(K), SH = 5 - log2_int(KP);
    double v[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) v[k] = k < K ? s[k] : 0.0;
    const double t = warp_sum_scatter<KP>(v);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ((lane & ((1 << SH) - 1)) == 0) buf[warp * KP + (lane >> SH)] = t;
    return bar_or_n(1, NTHREADS, flags);
}
template <int K, int NW>
__device__ __forceinline__ double block_sum_total(const double *buf, int lane)
{
    constexpr int KP = pow2_ceil(K);
    double t = 0.0;
    if (lane < KP) {
#pragma unroll
        for (int w = 0; w < NW; ++w) t += buf[w * KP + lane];
    }
    return t;
}

struct BatchArgs {
    ModelDesc md;
    const double *x;      // m
    const double *w;      // m or nullptr
    const double *Y;      // ld x P observations (unweighted), column p = problem p
    long long ld;
    long long P;
    double svd_eps;
    LmConfig cfg;
    const double *alpha0; // q x P
    double *alpha_out;    // q x P
    double *C_out;        // n x P
    double *obj_out;      // P: 0.5 * ||r_w||^2
    int *term_out;        // P: Termination
    int *nfev_out;        // P
    unsigned long long *next; // work counter (zeroed by the host)
    unsigned int *error;      // set to 1 when a hand-off between the compute warps and the LM warp timed out
    ExpTable expc;            // vp_exp_table(): the exp coefficients as direct constant-bank operands
    unsigned long long *dbg;  // optional per-CTA accumulators: [0] evaluation ns, [1] LM ns, [2] ns waiting for LM, [3] evaluations, [5] basis ns, [6] sweep 0 ns, [7] LM steps
};

// What one evaluation leaves behind for the LM phase (per problem slot, shared memory).
template <int N, int P, int KMAX>
struct BatchTail {
    double tv[KMAX];          // ||r||^2, u_e, M_ef (upper triangle)
    double top[N][N + P + 1]; // rows 0..n-1 of the rotated system: R, Q^T D, Q^T y
    double rdiag[N];
    int dropped, bad;
};

// The per-function part of the model descriptor, in shared memory (run-time indexing of kernel parameters would
// put the whole descriptor on the stack).
struct BatchModelS {
    int kind[VP_MAX_N], npar[VP_MAX_N], p0[VP_MAX_N], p1[VP_MAX_N];
    int e_basis[VP_MAX_P];
    double scale[VP_MAX_N];
};

// G problems are in flight per CTA ("slots", two groups of G / 2) and the CTA is WARP-SPECIALISED: THREADS compute
// threads evaluate the slots round-robin with all their lanes; ONE extra warp runs the lmder steps, lane l serving
// slot l of the current group, the lanes in lockstep (SIMT: the warp pays for the slowest step, the others cost
// nothing). The serial LM step of a q = 3 problem is ~3 500 dependent instructions = 30 us, against 11.7 us for an
// evaluation at m = 4096; while the LM warp steps the slots of one group the compute warps evaluate the other
// group, so the step latency disappears behind G / 2 evaluations. Hand-off through shared-memory sequence
// numbers per group: ev_seq = rounds evaluated, lm_seq = rounds stepped (DEAD when the group has no problem left).
// The compute warps only wait when a group's steps are slower than the other group's evaluations (small m) or in
// the drain at the end of the batch. No CTA-wide barrier after the prologue: the compute warps synchronise on
// named barrier 1, and the reduction buffers alternate with the evaluation parity so that warp 0 can finish the
// tail of one evaluation while the other warps are already generating the next basis.
// Working matrix: NPV columns of MP = RPT * THREADS doubles (dynamic shared memory); the rows >= m are zero.
// (The block has THREADS + 32 threads; registers are allocated in units of 4 warps, hence 96 per thread.)
template <int N, int P, int RPT, int THREADS, int G = 16>
__global__ void __launch_bounds__(THREADS + 32, 1)
batch_fit_kernel(const BatchArgs a)
{
    constexpr int NPV = N + P;
    constexpr int NW = THREADS / 32;
    constexpr int MP = RPT * THREADS;
    constexpr int NTAIL = 1 + P + P * (P + 1) / 2; // ||r||^2, u_e, M_ef (upper)
    constexpr int KMAX = (NTAIL > NPV + 1) ? NTAIL : NPV + 1;
    constexpr int KRED = NW * pow2_ceil(KMAX);
    constexpr int HUGE_HI = 0x5ff00000; // high word of 2^512 = 1.34e154: an entry this large overflows the column norm (cf. RANK_HUGE_ENTRY)
    constexpr int DEAD = 0x7fffffff;
    constexpr int GH = G / 2; // slots per group
    static_assert(G % 2 == 0 && GH <= 32, "one lane of the LM warp per slot of a group");
    static_assert(KMAX <= 32, "one lane of warp 0 per reduced value");
    extern __shared__ __align__(16) double colm[]; // NPV columns of MP doubles
    __shared__ double red2[2][KRED];               // per-warp partial sums of the running block reduction (by evaluation parity)
    __shared__ double bc[NPV + 2];                 // reflector scalars from warp 0: v_JJ, tau_k, tau_y
    __shared__ double rowpre[NPV + 1];             // row J of the working system before reflector J is applied
    __shared__ BatchModelS ms;
    __shared__ LmState st_s[G];
    __shared__ BatchTail<N, P, KMAX> tail_s[G];
    __shared__ double coef_acc[G][N];
    __shared__ long long prob_s[G]; // problem in the slot
    __shared__ int ev_seq[2], lm_seq[2], live_s[G];
    __shared__ int exhausted_s, abort_s, go_s[2]; // go_s: verdict of thread 0's bounded wait, alternating by group (a slow reader of one check never sees the next one's)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int m = a.md.m, q = a.md.q;

    if (tid < G) { live_s[tid] = 0; prob_s[tid] = -1; }
    if (tid < 2) { ev_seq[tid] = 0; lm_seq[tid] = 0; }
    if (tid == 0) {
        exhausted_s = 0; abort_s = 0; go_s[0] = go_s[1] = 1;
#pragma unroll
        for (int j = 0; j < VP_MAX_N; ++j) {
            ms.kind[j] = a.md.kind[j]; ms.npar[j] = a.md.npar[j]; ms.p0[j] = a.md.pidx[j][0]; ms.p1[j] = a.md.pidx[j][1];
            ms.scale[j] = a.md.scale[j];
        }
#pragma unroll
        for (int e = 0; e < VP_MAX_P; ++e) ms.e_basis[e] = a.md.e_basis[e];
    }
    __syncthreads(); // the only CTA-wide barrier

    // =============================== the LM warp =======================================================
    if (tid >= THREADS) {
        unsigned long long t_lm = 0, n_lm = 0; // (dbg, lane 0)
        const bool mine = lane < GH;
        // take a problem from the global counter into slot g, or mark the slot dead
        auto refill = [&](const int g) {
            long long pnew = -1;
            if (!ld_flag(&exhausted_s)) {
                pnew = (long long)atomicAdd(a.next, 1ull);
                if (pnew >= a.P) { pnew = -1; atomicExch(&exhausted_s, 1); }
            }
            if (pnew >= 0) {
                double x0[VP_MAX_Q];
                for (int k = 0; k < VP_MAX_Q; ++k) x0[k] = k < q ? a.alpha0[(size_t)pnew * q + k] : 0.0;
                lm_init(st_s[g], q, x0);
                prob_s[g] = pnew;
            }
            live_s[g] = pnew >= 0;
        };
        unsigned alive[2];
#pragma unroll
        for (int grp = 0; grp < 2; ++grp) {
            if (mine) refill(grp * GH + lane);
            alive[grp] = __ballot_sync(0xffffffffu, mine && live_s[grp * GH + lane]);
            if (lane == 0) st_flag(&lm_seq[grp], alive[grp] ? 1 : DEAD);
        }
        for (int round = 1; alive[0] | alive[1]; ++round) {
#pragma unroll 1
            for (int grp = 0; grp < 2; ++grp) {
                if (!alive[grp]) continue; // warp-uniform
                int lost = 0;
                if (lane == 0) {
                    const unsigned long long ts = global_timer_ns();
                    while (ld_flag(&ev_seq[grp]) < round) {
                        __nanosleep(200);
                        if (ld_flag(&abort_s) || global_timer_ns() - ts > BATCH_SPIN_TIMEOUT_NS) { lost = 1; break; }
                    }
                }
                if (__shfl_sync(0xffffffffu, lost, 0)) { // the compute warps are gone (bounded spin: never hang the GPU)
                    if (lane == 0) { *a.error = 1u; atomicExch(&abort_s, 1); }
                    return;
                }
                __threadfence_block();
                unsigned long long t0 = 0;
                if (a.dbg && lane == 0) t0 = global_timer_ns();
                const int g = grp * GH + lane;
                if (mine && live_s[g]) {
                    const BatchTail<N, P, KMAX> &tl = tail_s[g];
                    // inner solve c = R1^-1 (Q^T y) under the rank policy (rank_policy.cuh): triangular inverse, cheap
                    // full-rank test, and -- rarely -- the SVD of R1 with the truncated directions moved into the residual
                    double Rm[N * N], Ri[N * N], coef[N];
#pragma unroll
                    for (int c = 0; c < N; ++c)
#pragma unroll
                        for (int i = 0; i < N; ++i) { Rm[c * N + i] = (i < c) ? tl.top[i][c] : ((i == c) ? tl.rdiag[c] : 0.0); Ri[c * N + i] = 0.0; }
#pragma unroll
                    for (int c = 0; c < N; ++c)
#pragma unroll
                        for (int i = N - 1; i >= 0; --i) {
                            if (i > c) continue;
                            double sacc = (i == c) ? 1.0 : 0.0;
#pragma unroll
                            for (int k = 0; k < N; ++k)
                                if (k > i && k <= c) sacc -= Rm[k * N + i] * Ri[c * N + k];
                            Ri[c * N + i] = sacc / Rm[i * N + i];
                        }
                    double rn2_extra = 0.0;
                    if (rank_surely_full(N, Rm, N, Ri, N, a.svd_eps)) {
#pragma unroll
                        for (int i = 0; i < N; ++i) {
                            double sacc = 0.0;
#pragma unroll
                            for (int c = i; c < N; ++c) sacc += Ri[c * N + i] * tl.top[c][NPV];
                            coef[i] = sacc;
                        }
                    } else {
                        SmallSvd sv;
                        rank_policy_svd(N, Rm, N, a.svd_eps, &sv);
                        double bb[N], b2 = 0.0, bb2 = 0.0;
                        for (int c = 0; c < N; ++c) {
                            double sacc = 0.0;
                            for (int k = 0; k < N; ++k) sacc += sv.Urot[c * N + k] * tl.top[k][NPV];
                            bb[c] = sacc;
                            bb2 += sacc * sacc;
                            b2 += tl.top[c][NPV] * tl.top[c][NPV];
                        }
                        for (int i = 0; i < N; ++i) {
                            double sacc = 0.0;
                            for (int c = 0; c < N; ++c) sacc += sv.RinvEff[c * N + i] * bb[c];
                            coef[i] = sacc;
                        }
                        rn2_extra = fmax(b2 - bb2, 0.0); // the truncated components of Q^T y stay in the residual
                    }
                    LmEval ev;
                    ev.rnorm2 = tl.tv[0] + rn2_extra;
                    const int resid_ok = !tl.bad && isfinite(tl.tv[0] + rn2_extra);
                    int finite = 1; // derivatives
                    for (int k = 0; k < VP_LM_MAXQ; ++k) ev.g[k] = 0.0;
                    for (int k = 0; k < VP_LM_MAXQ * VP_LM_MAXQ; ++k) ev.H[k] = 0.0;
                    double Mm[P > 0 ? P : 1][P > 0 ? P : 1];
                    {
                        int t = 1 + P;
#pragma unroll
                        for (int e = 0; e < P; ++e)
#pragma unroll
                            for (int f2 = e; f2 < P; ++f2) { Mm[e][f2] = tl.tv[t]; Mm[f2][e] = tl.tv[t]; ++t; }
                    }
#pragma unroll
                    for (int e = 0; e < P; ++e) {
                        double ce = 0.0;
#pragma unroll
                        for (int r = 0; r < N; ++r) ce = (a.md.e_basis[e] == r) ? coef[r] : ce;
                        const int ke = a.md.e_param[e];
                        ev.g[ke] -= ce * tl.tv[1 + e];
#pragma unroll
                        for (int f2 = 0; f2 < P; ++f2) {
                            double cf = 0.0;
#pragma unroll
                            for (int r = 0; r < N; ++r) cf = (a.md.e_basis[f2] == r) ? coef[r] : cf;
                            ev.H[a.md.e_param[f2] * q + ke] += Mm[e][f2] * ce * cf;
                        }
                    }
                    for (int k = 0; k < q; ++k) finite = finite && isfinite(ev.g[k]);
                    for (int k = 0; k < q * q; ++k) finite = finite && isfinite(ev.H[k]);
                    ev.finite = (resid_ok ? VP_EVAL_RESIDUAL_OK : 0) | (finite ? VP_EVAL_DERIVS_OK : 0);
                    LmState &st = st_s[g];
                    const bool more = lm_advance(st, a.cfg, ev);
                    if (st.last_accepted) {
#pragma unroll
                        for (int r = 0; r < N; ++r) coef_acc[g][r] = coef[r];
                    }
                    if (!more) { // results of this problem; the slot takes the next problem (or dies)
                        const long long prob = prob_s[g];
                        for (int k = 0; k < q; ++k) a.alpha_out[(size_t)prob * q + k] = st.x[k];
                        for (int r = 0; r < N; ++r) a.C_out[(size_t)prob * N + r] = coef_acc[g][r];
                        a.obj_out[prob] = 0.5 * st.fnorm * st.fnorm;
                        a.term_out[prob] = st.termination;
                        a.nfev_out[prob] = st.nfev;
                        refill(g);
                    }
                }
                __syncwarp();
                alive[grp] = __ballot_sync(0xffffffffu, mine && live_s[g]);
                if (lane == 0) {
                    if (a.dbg) { t_lm += global_timer_ns() - t0; n_lm++; }
                    st_flag(&lm_seq[grp], alive[grp] ? round + 1 : DEAD);
                }
            }
        }
        if (a.dbg && lane == 0) {
            a.dbg[(size_t)blockIdx.x * 8 + 1] = t_lm;
            a.dbg[(size_t)blockIdx.x * 8 + 7] = n_lm;
        }
        return;
    }

    // =============================== compute warps =====================================================
    unsigned long long t_eval = 0, t_wait = 0, n_evals = 0, t_basis = 0, t_sw0 = 0, te = 0, tb = 0; // (thread 0, dbg only)
    // the thread's rows of x and w (shared by all problems)
    double xi[RPT], wi[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        const int i = tid + r * THREADS;
        const bool in = i < m;
        xi[r] = in ? a.x[i] : 0.0;
        wi[r] = in ? (a.w ? a.w[i] : 1.0) : 0.0;
    }
    int parity = 0;
    for (int round = 1;; ++round) {
        int nlive = 0;
#pragma unroll 1
        for (int g = 0; g < G; ++g) {
            const int grp = g >= GH;
            if (g == 0 || g == GH) { // the group's previous LM steps (or its first fill) must be complete
                if (tid == 0) { // (one thread decides, so that a timeout leaves the barriers of the others balanced)
                    const unsigned long long ts = global_timer_ns();
                    int ok = 1;
                    while (ld_flag(&lm_seq[grp]) < round)
                        if (ld_flag(&abort_s) || global_timer_ns() - ts > BATCH_SPIN_TIMEOUT_NS) { ok = 0; break; }
                    go_s[grp] = ok;
                    if (!ok) { *a.error = 1u; atomicExch(&abort_s, 1); }
                    if (a.dbg) t_wait += global_timer_ns() - ts;
                }
                bar_sync_n(1, THREADS);
                if (!go_s[grp]) return; // uniform
            }
            if (live_s[g]) { // uniform: written before lm_seq
            ++nlive;
            parity ^= 1;
            double *const red = red2[parity];
            const long long prob = prob_s[g];
            BatchTail<N, P, KMAX> &tl = tail_s[g];
            if (a.dbg && tid == 0) te = global_timer_ns();
            const double *alpha_s = st_s[g].x_trial;
            // y_p (weighted like builder.rs:307): issued first, consumed after the basis evaluation
            double yt[RPT];
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int i = tid + r * THREADS;
                yt[r] = (i < m) ? wi[r] * __ldcg(&a.Y[(size_t)prob * a.ld + i]) : 0.0;
            }
            // 1. Phi_w, D into the working matrix (rolled over basis functions, rows unrolled). The rows
            //    r < RPT / 2 are always inside the problem (the host picks the smallest RPT covering m); the others
            //    may be padding, which is stored as zero whatever the basis function returns at x = 0.
            int bad_local = 0;
            {
                int e = 0;
#pragma unroll 1
                for (int j = 0; j < N; ++j) {
                    const int kind = ms.kind[j], np = ms.npar[j];
                    const double a0 = np > 0 ? alpha_s[ms.p0[j]] : 0.0;
                    const double a1 = np > 1 ? alpha_s[ms.p1[j]] : 0.0;
                    double *cv = colm + (size_t)j * MP + tid, *ca = colm + (size_t)(N + e) * MP + tid, *cb = ca + MP;
                    int mx = 0; // max over the thread's rows of the high word of |Phi_w[i][j]| (NaN / Inf are the largest)
                    if (kind == VP_BASIS_EXP_DECAY) {
                        const double inv0 = 1.0 / a0; // one division per basis function, see basis_eval_all
                        constexpr int RB = RPT < 4 ? RPT : 4; // rows per block: RB exp chains in flight per thread
                        double done = 0.0;                    // (scheduling only: a block starts after the previous one is
                                                              //  finished -- 8 chains do not fit in 96 registers)
#pragma unroll
                        for (int r0 = 0; r0 < RPT; r0 += RB) {
                            double t[RB], arg[RB], ex[RB];
#pragma unroll
                            for (int v = 0; v < RB; ++v) {
                                double xr = xi[r0 + v];
                                if (r0 > 0) asm volatile("" : "+d"(xr) : "d"(done));
                                t[v] = xr * inv0;
                                arg[v] = -t[v];
                            }
                            vp_exp_vec<RB>(arg, ex, a.expc.c);
#pragma unroll
                            for (int v = 0; v < RB; ++v) {
                                const int r = r0 + v;
                                double pv = wi[r] * ex[v], pa = wi[r] * (ex[v] * t[v] * inv0);
                                if (RPT == 1 || r >= RPT / 2) {
                                    const bool in = tid + r * THREADS < m;
                                    pv = in ? pv : 0.0; pa = in ? pa : 0.0;
                                }
                                mx = max(mx, __double2hiint(pv) & 0x7fffffff);
                                cv[r * THREADS] = pv;
                                ca[r * THREADS] = pa;
                                done = pa;
                            }
                        }
                    } else {
                        const double scale = ms.scale[j];
#pragma unroll
                        for (int r = 0; r < RPT; ++r) {
                            double v, da = 0.0, db = 0.0;
                            if (kind == VP_BASIS_EXP_RATE_COS) {
                                const double ex = vp_exp(-a0 * xi[r]);
                                double sn, cs;
                                sincos(a1 * xi[r], &sn, &cs);
                                v = ex * cs; da = -xi[r] * (ex * cs); db = -xi[r] * ex * sn;
                            } else if (kind == VP_BASIS_SIN_PHASE) {
                                double sn, cs;
                                sincos(a0 * xi[r] + a1, &sn, &cs);
                                v = sn; da = xi[r] * cs; db = cs;
                            } else {
                                v = kind == VP_BASIS_CONSTANT ? 1.0 : (kind == VP_BASIS_LINEAR_X ? scale * xi[r] : nan(""));
                            }
                            const bool in = tid + r * THREADS < m;
                            const double pv = in ? wi[r] * v : 0.0;
                            mx = max(mx, __double2hiint(pv) & 0x7fffffff);
                            cv[r * THREADS] = pv;
                            if (np > 0) ca[r * THREADS] = in ? wi[r] * da : 0.0;
                            if (np > 1) cb[r * THREADS] = in ? wi[r] * db : 0.0;
                        }
                    }
                    bad_local |= (mx >= 0x7ff00000 ? 1 : 0) | (mx >= HUGE_HI ? (2 << j) : 0); // flag word of rank_policy.cuh
                    e += np;
                }
            }

            if (a.dbg && tid == 0) { tb = global_timer_ns(); t_basis += tb - te; }
            // 2. sweep 0: sigma_0 and the dots of column 0 with every column and y. The barrier of the reduction
            //    also ORs the flag words; overflowing basis columns (rare) are zeroed together with their derivative
            //    columns -- every thread in its own rows -- and the sweep is repeated.
            double s[KMAX];
            int bad = 0;
            {
                int flags = bad_local;
                bool redo;
                do {
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) s[k] = 0.0;
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        const double *rowp = colm + tid + r * THREADS;
                        const double aj = rowp[0];
#pragma unroll
                        for (int k = 0; k < NPV; ++k) s[k] = fma(aj, rowp[(size_t)k * MP], s[k]);
                        s[NPV] = fma(aj, yt[r], s[NPV]);
                    }
                    if (tid == 0) { // row 0 before its reflector is applied (thread 0 owns it, r = 0)
#pragma unroll
                        for (int k = 0; k < NPV; ++k) rowpre[k] = colm[(size_t)k * MP];
                        rowpre[NPV] = yt[0];
                    }
                    redo = false;
                    if (block_sum_begin<NPV + 1, THREADS>(s, red, flags)) {
                        bad = 0; // (rare) the whole flag word, bit by bit
                        for (int b = 0; b < 1 + N; ++b)
                            if (bar_or_n(1, THREADS, (bad_local >> b) & 1)) bad |= 1 << b;
                        flags = 0;
                        if (bad >> 1) {
#pragma unroll
                            for (int c = 0; c < NPV; ++c) {
                                const int jb = c < N ? c : ms.e_basis[c - N];
                                if ((bad >> (1 + jb)) & 1) {
#pragma unroll
                                    for (int r = 0; r < RPT; ++r) colm[(size_t)c * MP + tid + r * THREADS] = 0.0;
                                }
                            }
                            redo = true;
                        }
                        bad &= 1;
                    }
                } while (redo);
            }

            if (a.dbg && tid == 0) t_sw0 += global_timer_ns() - tb;
            // 3. Householder steps. Iteration J: warp 0 finishes the pending reduction (dots of column J over the rows
            //    >= J) and forms the reflector; everybody applies it in one sweep that also accumulates what comes next.
            int dropped = 0;
            static_for<N>([&](auto Jc) {
                constexpr int J = decltype(Jc)::value;
                constexpr int K = NPV - J + 1; // sigma_J, dots with the NPV - J - 1 columns to the right, dot with y
                if (warp == 0) {
                    const double tot = block_sum_total<K, NW>(red, lane);
                    const double sigma = __shfl_sync(0xffffffffu, tot, 0);
                    const double ajj = rowpre[J];
                    const double nrm = sqrt(sigma);
                    const bool keep = isfinite(nrm) && nrm > 0.0; // near-dependence: rank policy on R1 in the LM phase
                    const double al = (ajj >= 0.0) ? -nrm : nrm;
                    const double vnorm2 = 2.0 * (sigma - ajj * al);
                    const double bt = (keep && vnorm2 > 0.0) ? 2.0 / vnorm2 : 0.0;
                    if (!keep) dropped |= 1 << J;
                    if (lane >= 1 && lane < K) bc[lane] = bt * (tot - al * rowpre[J + lane]); // tau_k; lane K - 1: tau_y
                    if (lane == 0) { bc[0] = ajj - al; tl.rdiag[J] = keep ? al : 0.0; }
                }
                bar_sync_n(1, THREADS);
                const double vjj = bc[0];
                double tau[K];
#pragma unroll
                for (int k = 1; k < K; ++k) tau[k] = bc[k];
                const double tau_y = tau[K - 1];
#pragma unroll
                for (int k = 0; k < KMAX; ++k) s[k] = 0.0;
#pragma unroll
                for (int r = 0; r < RPT; ++r) {
                    double *rowp = colm + tid + r * THREADS;
                    double vi = rowp[(size_t)J * MP]; // zero in the finished rows < J
                    if (r == 0) vi = (tid == J) ? vjj : vi;
                    double row[NPV]; // the updated entries of the columns > J of this row
#pragma unroll
                    for (int k = J + 1; k < NPV; ++k) row[k] = fma(-tau[k - J], vi, rowp[(size_t)k * MP]);
                    double yv = fma(-tau_y, vi, yt[r]);
                    if (r == 0) {
                        if (tid == J) { // the final row J: R[J][k > J], (Q^T D)[J][:], (Q^T y)[J]; then out of the system
#pragma unroll
                            for (int k = J + 1; k < NPV; ++k) { tl.top[J][k] = row[k]; row[k] = 0.0; }
                            tl.top[J][NPV] = yv;
                            yv = 0.0;
                        }
                    }
                    yt[r] = yv;
                    if constexpr (J + 1 < N) {
#pragma unroll
                        for (int k = J + 1; k < NPV; ++k) rowp[(size_t)k * MP] = row[k];
                        const double aj = row[J + 1]; // dots of the updated column J + 1 for the next reflector
#pragma unroll
                        for (int k = 0; k < NPV - J - 1; ++k) s[k] = fma(aj, row[J + 1 + k], s[k]);
                        s[NPV - J - 1] = fma(aj, yv, s[NPV - J - 1]);
                        if (r == 0) {
                            if (tid == J + 1) { // row J + 1 before ITS reflector
#pragma unroll
                                for (int k = J + 1; k < NPV; ++k) rowpre[k] = row[k];
                                rowpre[NPV] = yv;
                            }
                        }
                    } else { // tail sums over the rows >= n (the untruncated projector); the finished rows are zero
                        s[0] = fma(yv, yv, s[0]);
#pragma unroll
                        for (int e = 0; e < P; ++e) s[1 + e] = fma(row[N + e], yv, s[1 + e]);
                        int t = 1 + P;
#pragma unroll
                        for (int e = 0; e < P; ++e)
#pragma unroll
                            for (int f2 = e; f2 < P; ++f2) { s[t] = fma(row[N + e], row[N + f2], s[t]); ++t; }
                    }
                }
                if constexpr (J + 1 < N) block_sum_begin<K - 1, THREADS>(s, red, 0);
                else block_sum_begin<NTAIL, THREADS>(s, red, 0);
            });
            if (warp == 0) {
                const double tot = block_sum_total<NTAIL, NW>(red, lane);
                if (lane < NTAIL) tl.tv[lane] = tot;
                if (lane == 0) { tl.dropped = dropped; tl.bad = bad; }
            }
            if (a.dbg && tid == 0) { t_eval += global_timer_ns() - te; ++n_evals; }
            } // live slot
            if ((g == GH - 1 || g == G - 1) && warp == 0) { // warp 0 completes every tail: hand the group to the LM warp
                __syncwarp();
                if (lane == 0) st_flag(&ev_seq[grp], round);
            }
        }
        if (nlive == 0) break;
    }
    if (a.dbg && tid == 0) {
        unsigned long long *d = a.dbg + (size_t)blockIdx.x * 8;
        d[0] = t_eval; d[2] = t_wait; d[3] = n_evals; d[5] = t_basis; d[6] = t_sw0;
    }
}

} // namespace vp

