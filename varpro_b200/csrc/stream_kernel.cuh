// stream_kernel.cuh -- K2: the Y-streaming reduce (the HBM-bound kernel).
//
// One pass over the weighted observations Y_w (m x S, column-major) replaces,
// for the current alpha, everything the reference does that is O(m*S):
//   C = Phi_w^+ Y_w, R = Y_w - Phi_w C     src/solvers/levmar/mod.rs:52-59
//   residuals() -> vec(R)                  src/solvers/levmar/mod.rs:91-95
//   jacobian()  -> Kaufman J               src/solvers/levmar/mod.rs:101-201
//   pivoted QR of J, Q^T r                 levenberg-marquardt crate (lmder)
// Per column y_s (read ONCE from HBM, staged in shared memory by a 1-D bulk
// async copy / TMA, completion on an mbarrier):
//   b_s = Q^T y_s, u_s = E^T y_s           phase 1: (n+p) dot products
//   c_s = R1^-1 b_s                        -> coefficient matrix C (n x S)
//   r_s = y_s - Q b_s, ||r_s||^2           phase 2 (explicit residual: no
//                                          ||y||^2-||b||^2 cancellation)
// accumulated per CTA into  sum ||r_s||^2,  G = sum c_s c_s^T,
// V_e = sum c_{s,j(e)} u_{s,e}; the last CTA to finish folds the per-CTA
// partials in a fixed order (deterministic) into
//   g_k = -(sum_{e in k} V_e)          = (J^T r)_k
//   H_kl = sum_{e in k, f in l} M_ef G_{j(e) j(f)} = (J^T J)_kl.
//
// Mapping: the rows of a tile are spread over all threads of the CTA (each
// thread owns CHUNKS vectors of VEC rows), so the thread's slice of the panel
// [Q|E] lives in REGISTERS for the whole kernel and shared-memory traffic is
// just the two reads of each y element. A tile is CT whole columns, contiguous
// in HBM, fetched with one bulk copy.
#pragma once

#include "device_common.cuh"
#include "lm_step.cuh"
#include "panel_kernel.cuh"

namespace vp {

struct EvalOut {
    double rnorm2;
    double g[VP_MAX_Q];
    double H[VP_MAX_Q * VP_MAX_Q]; // column-major, ld = q
    int finite;
    int pad;
};

// Device-resident state of one fit (graph path): the lmder state machine of lm_step.cuh
// advanced by the last CTA of the streaming kernel, so that a whole fit runs without the host.
struct FitDevice {
    LmState st;
    LmConfig cfg;
    LmEval accepted; // evaluation belonging to st.x
    int cur;         // coefficient buffer holding C(st.x): the kernel writes the trial into cur^1
    int evals;       // evaluations made by the graph (diagnostics)
    double trace[4 * 48]; // per evaluation: fnorm_trial, par, delta, accepted (diagnostics; VP_TRACE)
};
constexpr int FIT_WORDS = (int)((sizeof(FitDevice) + 7) / 8);

template <typename T>
struct StreamArgs {
    const T *Y;      // m x S weighted observations, ld rows per column
    int ld;          // padded rows (multiple of 16/sizeof(T))
    int S;
    const T *Pq;     // panel [Q | E | 0]: (n+p+1) columns of ldp rows; the last column is all zero
    const T *Pe;     // = Pq + n*ldp
    int ldp;         // panel column stride (>= ld and >= the rows a kernel touches; rows >= m are zero)
    unsigned long long *dbg;   // optional timeline buffer (gridDim.x * VP_DBG_SLOTS), or nullptr
    int tiles_base, tiles_rem; // CTA b processes tiles_base + (b < tiles_rem) tiles: b, b+grid, b+2*grid, ...
    const PanelSmall *small;
    T *C0, *C1;      // n x S coefficient buffers
    int cdst;        // host-driven path: buffer to write (0/1); graph path: ignored (fit->cur ^ 1)
    FitDevice *fit;  // graph path: device-resident LM state, or nullptr
    unsigned long long cond; // graph path: cudaGraphConditionalHandle of the while node
    double *partials; // gridDim.x * red_stride
    int red_stride;
    unsigned int *ticket;
    EvalOut *out;
    int nstages;
    int q;
    int e_basis[VP_MAX_P];
    int e_param[VP_MAX_P];
};

constexpr int STREAM_MAX_STAGES = 16;
constexpr int DMMA_CT = 8; // columns per tile of stream_kernel_dmma (stream_kernel_dmma.cuh)

// coefficient buffer this launch writes: the one NOT holding the accepted coefficients
template <typename T>
__device__ __forceinline__ T *stream_cout(const StreamArgs<T> &a)
{
    const int dst = a.fit ? (__ldcg(&a.fit->cur) ^ 1) : a.cdst;
    return dst ? a.C1 : a.C0;
}

template <typename T> struct VecOf;
template <> struct VecOf<double> { using type = double2; static constexpr int N = 2; };
template <> struct VecOf<float> { using type = float4; static constexpr int N = 4; };

template <typename T> __device__ __forceinline__ void vec_unpack(const double2 &v, T (&o)[2]) { o[0] = v.x; o[1] = v.y; }
template <typename T> __device__ __forceinline__ void vec_unpack(const float4 &v, T (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }

// Halving fold of NACC = CT*NPV per-lane partial sums across a warp: each of
// the log2(CT) exchange steps halves the number of values a lane carries, so
// a lane ends up owning the NPV totals of ONE column; a plain butterfly over
// the remaining lane bits finishes. ~2.2*NACC shuffles instead of 10*NACC.
// All indices are compile-time constants (the array must stay in registers).
template <int NACC, int CNT, int OFF, int NPV>
__device__ __forceinline__ void warp_fold(double (&red)[NACC], const int lane)
{
    if constexpr (CNT > NPV) {
        constexpr int half = CNT / 2;
        const bool up = (lane & OFF) != 0;
#pragma unroll
        for (int t = 0; t < half; ++t) {
            const double send = up ? red[t] : red[t + half];
            const double keep = up ? red[t + half] : red[t];
            red[t] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
        }
        warp_fold<NACC, half, OFF / 2, NPV>(red, lane);
    } else {
#pragma unroll
        for (int off = OFF; off > 0; off >>= 1)
#pragma unroll
            for (int k = 0; k < NPV; ++k) red[k] += __shfl_xor_sync(0xffffffffu, red[k], off);
    }
}

// Final fold executed by the last CTA: partials -> (rnorm2, g, H).
// 16 threads per value: thread (k = tid%16, c0 = tid/16) sums value k over the
// CTAs c0, c0 + blockDim/16, ... (independent loads in flight), the 16-column
// table is then folded in a fixed order => bitwise reproducible for a given
// grid, no shuffles. sh: >= 64 doubles; scratch: >= 16 * blockDim/16 + p*p doubles.
constexpr int FIN_SCRATCH = 16 * 32 + VP_MAX_P * VP_MAX_P + 64 + 4 * 48; // also holds a FitDevice copy (static_assert below)
static_assert(FIT_WORDS <= FIN_SCRATCH, "FitDevice must fit in the finalize scratch");
// SMEM_SMALL: a.small points to shared memory (fit_kernel_dmma keeps the panel's small outputs in
// the CTA) -- plain loads instead of ld.global.cg. GRAPH_TAIL: run the CUDA-graph LM tail.
template <typename T, bool SMEM_SMALL = false, bool GRAPH_TAIL = true>
__device__ void stream_finalize(const StreamArgs<T> &a, int n, int p, int nparts, double *sh, double *scratch)
{
    const int nv = red_count(n, p);
    const int tid = threadIdx.x, nt = blockDim.x;
    const int k = tid & 15, c0 = tid >> 4, nc0 = nt >> 4;
    double *Msh = scratch + 16 * 32;
    if (tid < p * p) { // Msh[f*p + e]
        const double *msrc = &a.small->M[(tid / p) * VP_MAX_P + (tid % p)];
        Msh[tid] = SMEM_SMALL ? *msrc : __ldcg(msrc);
    }
    int nonfinite = 0;
    if (tid == 0) nonfinite = SMEM_SMALL ? a.small->nonfinite : __ldcg(&a.small->nonfinite);
    dbg_mark(a.dbg, 11);
#pragma unroll 1
    for (int base = 0; base < nv; base += 16) {
        double s = 0.0;
        if (base + k < nv) {
            // four independent accumulators: the loads of one trip are all in flight together
            const double *src = a.partials + base + k;
            const size_t rs = (size_t)a.red_stride;
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            int c = c0;
            for (; c + 3 * nc0 < nparts; c += 4 * nc0) {
                const double l0 = __ldcg(src + (size_t)c * rs), l1 = __ldcg(src + (size_t)(c + nc0) * rs);
                const double l2 = __ldcg(src + (size_t)(c + 2 * nc0) * rs), l3 = __ldcg(src + (size_t)(c + 3 * nc0) * rs);
                s0 += l0; s1 += l1; s2 += l2; s3 += l3;
            }
            for (; c < nparts; c += nc0) s0 += __ldcg(src + (size_t)c * rs);
            s = (s0 + s1) + (s2 + s3);
        }
        scratch[c0 * 16 + k] = s;
        __syncthreads();
        dbg_mark(a.dbg, 12);
        if (tid < 16 && base + tid < nv) {
            double t = 0.0;
            for (int c = 0; c < nc0; ++c) t += scratch[c * 16 + tid];
            sh[base + tid] = t;
        }
        __syncthreads();
    }
    if (tid == 0) {
        EvalOut *o = a.out;
        const int q = a.q;
        double rn2 = sh[0];
        const int resid_ok = isfinite(rn2) && !nonfinite;
        int finite = 1; // derivatives
        for (int kk = 0; kk < q; ++kk) {
            double gk = 0.0;
            for (int e = 0; e < p; ++e)
                if (a.e_param[e] == kk) gk -= sh[1 + n * (n + 1) / 2 + e];
            o->g[kk] = gk;
            finite = finite && isfinite(gk);
        }
        for (int kk = 0; kk < q; ++kk)
            for (int l = 0; l <= kk; ++l) {
                double h = 0.0;
                for (int e = 0; e < p; ++e) {
                    if (a.e_param[e] != kk) continue;
                    for (int f = 0; f < p; ++f) {
                        if (a.e_param[f] != l) continue;
                        int i = a.e_basis[e], j = a.e_basis[f];
                        if (i > j) { int t = i; i = j; j = t; }
                        h += Msh[f * p + e] * sh[g_index(n, i, j)];
                    }
                }
                o->H[l * q + kk] = h;
                o->H[kk * q + l] = h;
                finite = finite && isfinite(h);
            }
        o->rnorm2 = rn2;
        o->finite = (resid_ok ? VP_EVAL_RESIDUAL_OK : 0) | (finite ? VP_EVAL_DERIVS_OK : 0);
        *a.ticket = 0; // re-arm for the next launch
        dbg_mark(a.dbg, 13);
    }
    if (GRAPH_TAIL && a.fit) {
        // graph path: advance the lmder state machine on a shared-memory copy of the state
        // (cooperative load/store: one 8-byte word per thread) and steer the while node
        __syncthreads();
        unsigned long long *fw = reinterpret_cast<unsigned long long *>(a.fit);
        unsigned long long *lw = reinterpret_cast<unsigned long long *>(scratch); // FIN_SCRATCH >= FIT_WORDS
        for (int i = tid; i < FIT_WORDS; i += nt) lw[i] = __ldcg(fw + i);
        __syncthreads();
        if (tid == 0) {
            FitDevice *f = reinterpret_cast<FitDevice *>(lw);
            LmEval ev;
            ev.rnorm2 = a.out->rnorm2;
            ev.finite = a.out->finite;
            for (int kk = 0; kk < VP_LM_MAXQ; ++kk) ev.g[kk] = kk < a.q ? a.out->g[kk] : 0.0;
            for (int kk = 0; kk < VP_LM_MAXQ * VP_LM_MAXQ; ++kk) ev.H[kk] = kk < a.q * a.q ? a.out->H[kk] : 0.0;
            const bool more = lm_advance(f->st, f->cfg, ev);
            if (f->st.last_accepted) {
                f->cur ^= 1;
                f->accepted = ev;
            }
            if (f->evals < 48) {
                double *tr = f->trace + 4 * f->evals;
                tr[0] = sqrt(ev.rnorm2); tr[1] = f->st.par; tr[2] = f->st.delta; tr[3] = f->st.last_accepted;
            }
            f->evals += 1;
            cudaGraphSetConditional(a.cond, more ? 1u : 0u);
        }
        __syncthreads();
        for (int i = tid; i < FIT_WORDS; i += nt) fw[i] = lw[i];
        dbg_mark(a.dbg, 15);
    }
}

// CTA partial (already reduced, in shared memory) -> global; the last CTA to arrive folds all partials.
template <typename T>
__device__ __forceinline__ void stream_epilogue(const StreamArgs<T> &a, int n, int p, const double *cta_vals /* smem, nv */,
                                                double *sh, double *scratch, int *is_last)
{
    const int nv = red_count(n, p);
    const int tid = threadIdx.x;
    if (tid < nv) a.partials[(size_t)blockIdx.x * a.red_stride + tid] = cta_vals[tid];
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int prev = atomicAdd(a.ticket, 1u);
        *is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (*is_last) {
        __threadfence();
        stream_finalize<T>(a, n, p, gridDim.x, sh, scratch);
    }
}

// Fold the per-thread accumulators of one CTA (rn2 in every thread; G, V in
// threads < CT) into the CTA's partial vector, publish it, and let the last CTA
// finish. Written for few instructions: it runs once per CTA and these kernels
// live for a few microseconds.
template <typename T, int N, int P, int CT, int NW>
__device__ __forceinline__ void cta_publish_partial(const StreamArgs<T> &a, double rn2, const double *Gacc, const double *Vacc,
                                                    double *wsum /* NW */, double *gv /* CT*(NG+P) */, int *is_last,
                                                    const int row = blockIdx.x, const int nrows = gridDim.x)
{
    constexpr int NG = N * (N + 1) / 2, NGP = NG + P, NVR = 1 + NGP;
    static_assert(NVR <= 32, "the partial row must be written by one warp");
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    dbg_mark(a.dbg, 4);
    const double t = warp_sum(rn2);
    if (lane == 0) wsum[warp] = t;
    if (tid < CT) {
#pragma unroll
        for (int u = 0; u < NG; ++u) gv[tid * NGP + u] = Gacc[u];
#pragma unroll
        for (int e = 0; e < P; ++e) gv[tid * NGP + NG + e] = Vacc[e];
    }
    __syncthreads();
    if (tid < NVR) {
        double s = 0.0;
        if (tid == 0) {
#pragma unroll
            for (int w = 0; w < NW; ++w) s += wsum[w];
        } else {
#pragma unroll
            for (int c = 0; c < CT; ++c) s += gv[c * NGP + tid - 1];
        }
        a.partials[(size_t)row * a.red_stride + tid] = s;
    }
    // publish: the partial row is written by warp 0; its lane 0 then takes a ticket with
    // release/acquire semantics at gpu scope (orders the row before the ticket and, in
    // the last CTA, the ticket before the reads of everybody's rows)
    dbg_mark(a.dbg, 14);
    if (warp == 0) {
        __syncwarp();
        if (lane == 0) {
            unsigned int prev;
            asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(prev) : "l"(a.ticket) : "memory");
            *is_last = (prev == (unsigned int)nrows - 1u);
        }
    }
    __syncthreads();
    dbg_mark(a.dbg, 5);
}

template <typename T, int N, int P, int CT, int NW>
__device__ __forceinline__ void cta_publish(const StreamArgs<T> &a, double rn2, const double *Gacc, const double *Vacc,
                                            double *wsum /* NW */, double *gv /* CT*(NG+P) */, double *sh,
                                            double *scratch, int *is_last)
{
    cta_publish_partial<T, N, P, CT, NW>(a, rn2, Gacc, Vacc, wsum, gv, is_last);
    if (*is_last) {
        stream_finalize<T>(a, N, P, gridDim.x, sh, scratch);
        dbg_mark(a.dbg, 6);
    }
}

template <typename T, int N, int P, int CHUNKS, int CT, int THREADS>
__global__ void __launch_bounds__(THREADS)
stream_kernel(const StreamArgs<T> a)
{
    using VecT = typename VecOf<T>::type;
    constexpr int VEC = VecOf<T>::N;
    constexpr int NPV = N + P;
    constexpr int NW = THREADS / 32;
    constexpr int NACC = CT * NPV;
    constexpr int LOG2CT = (CT == 1) ? 0 : (CT == 2) ? 1 : (CT == 4) ? 2 : (CT == 8) ? 3 : 4;
    static_assert((1 << LOG2CT) == CT && CT <= 16, "CT must be a power of two <= 16");
    constexpr int NVR = 1 + N * (N + 1) / 2 + P;
    static_assert(NACC <= THREADS, "tile too wide for the cross-warp fold");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[STREAM_MAX_STAGES];
    __shared__ double part[NW * NACC];
    __shared__ double bu[NACC];
    __shared__ double rinv_s[N * N];
    __shared__ double fin_scratch[FIN_SCRATCH];
    __shared__ double wsum_s[NW];
    __shared__ double gv_s[CT * (N * (N + 1) / 2 + P)];
    __shared__ double fin_sh[64];
    __shared__ int is_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld, S = a.S, nst = a.nstages;
    const size_t stage_elems = (size_t)CT * ld;
    T *tiles = reinterpret_cast<T *>(smem_raw);
    T *Cout = stream_cout(a);

    const int my = a.tiles_base + ((int)blockIdx.x < a.tiles_rem ? 1 : 0);

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    // producer state (thread 0): next tile to fetch and the stage it goes to
    int next_i = 0, next_st = 0;
    auto issue = [&]() {
        const int tile = blockIdx.x + next_i * gridDim.x;
        const int col0 = tile * CT;
        const int nc = min(CT, S - col0);
        const uint32_t bytes = (uint32_t)((size_t)nc * ld * sizeof(T));
        mbar_arrive_expect_tx(&full_bar[next_st], bytes);
        bulk_copy_g2s(tiles + (size_t)next_st * stage_elems, a.Y + (size_t)col0 * ld, bytes, &full_bar[next_st]);
        ++next_i;
        if (++next_st == nst) next_st = 0;
    };
    // the observations do not depend on the panel: start fetching immediately
    if (tid == 0)
        for (int i = 0; i < nst && i < my; ++i) issue();

    // this thread's slice of the panel [Q | E], kept in registers
    T pan[CHUNKS][VEC][NPV];
    bool valid[CHUNKS];
#pragma unroll
    for (int ch = 0; ch < CHUNKS; ++ch) {
        const int r0 = VEC * (tid + ch * THREADS);
        valid[ch] = r0 < ld;
#pragma unroll
        for (int k = 0; k < NPV; ++k) {
            T tmp[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) tmp[v] = (T)0;
            if (valid[ch]) {
                const T *src = a.Pq + (size_t)k * a.ldp + r0;
                vec_unpack<T>(*reinterpret_cast<const VecT *>(src), tmp);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) pan[ch][v][k] = tmp[v];
        }
    }
    if (tid < N * N) rinv_s[tid] = a.small->Rinv[(tid / N) * VP_MAX_N + (tid % N)];

    double rn2 = 0.0;            // every thread: sum of r^2 over its rows
    double Gacc[N * (N + 1) / 2]; // threads < CT: sum c c^T over their columns
    double Vacc[P > 0 ? P : 1];
#pragma unroll
    for (int i = 0; i < N * (N + 1) / 2; ++i) Gacc[i] = 0.0;
#pragma unroll
    for (int e = 0; e < (P > 0 ? P : 1); ++e) Vacc[e] = 0.0;

    int st = 0;
    uint32_t parity = 0;
    for (int i = 0; i < my; ++i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int col0 = tile * CT;
        const int nc = min(CT, S - col0);
        const T *tp = tiles + (size_t)st * stage_elems;
        mbar_wait(&full_bar[st], parity);
        if (++st == nst) { st = 0; parity ^= 1u; }

        // ---- phase 1: partial dot products of the panel with CT columns ----
        T acc[CT][NPV];
#pragma unroll
        for (int c = 0; c < CT; ++c)
#pragma unroll
            for (int k = 0; k < NPV; ++k) acc[c][k] = (T)0;
#pragma unroll
        for (int ch = 0; ch < CHUNKS; ++ch) {
            if (!valid[ch]) continue;
            const int r0 = VEC * (tid + ch * THREADS);
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                T y[VEC];
                vec_unpack<T>(*reinterpret_cast<const VecT *>(tp + (size_t)c * ld + r0), y);
#pragma unroll
                for (int v = 0; v < VEC; ++v)
#pragma unroll
                    for (int k = 0; k < NPV; ++k) acc[c][k] = fma(pan[ch][v][k], y[v], acc[c][k]);
            }
        }
        // warp fold: LOG2CT halving exchanges (each lane ends up owning one
        // column), then a plain butterfly over the remaining lanes.
        double red[NACC];
#pragma unroll
        for (int c = 0; c < CT; ++c)
#pragma unroll
            for (int k = 0; k < NPV; ++k) red[c * NPV + k] = (double)acc[c][k];
        warp_fold<NACC, NACC, 16, NPV>(red, lane);
        if ((lane & ((32 >> LOG2CT) - 1)) == 0) {
            const int c = lane >> (5 - LOG2CT);
#pragma unroll
            for (int k = 0; k < NPV; ++k) part[warp * NACC + c * NPV + k] = red[k];
        }
        __syncthreads(); // (A) every thread is past phase 2 of the previous tile
        if (tid == 0 && i >= 1 && next_i < my) issue(); // refill the stage tile i-1 used
        if (tid < NACC) {
            double s = 0.0;
#pragma unroll
            for (int w2 = 0; w2 < NW; ++w2) s += part[w2 * NACC + tid];
            bu[tid] = s;
        }
        __syncthreads(); // (B) b_s, u_s of the CT columns are complete

        // ---- solve: c_s = R1^-1 b_s ; accumulate G and V (one thread per column)
        if (tid < nc) {
            double coef[N];
#pragma unroll
            for (int r = 0; r < N; ++r) {
                double s = 0.0;
#pragma unroll
                for (int c2 = 0; c2 < N; ++c2) s += rinv_s[c2 * N + r] * bu[tid * NPV + c2]; // full product: the rank policy may make Rinv non-triangular
                coef[r] = s;
                Cout[(size_t)(col0 + tid) * N + r] = (T)s;
            }
            int gi = 0;
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int c2 = r; c2 < N; ++c2) Gacc[gi++] += coef[r] * coef[c2];
#pragma unroll
            for (int e = 0; e < P; ++e) {
                double cj = 0.0;
#pragma unroll
                for (int r = 0; r < N; ++r) cj = (a.e_basis[e] == r) ? coef[r] : cj;
                Vacc[e] += cj * bu[tid * NPV + N + e];
            }
        }

        // ---- phase 2: explicit residual r = y - Q b and its squared norm ----
        {
            T rsum = (T)0;
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                if (c >= nc) break;
                T b[N];
#pragma unroll
                for (int k = 0; k < N; ++k) b[k] = (T)bu[c * NPV + k];
#pragma unroll
                for (int ch = 0; ch < CHUNKS; ++ch) {
                    if (!valid[ch]) continue;
                    const int r0 = VEC * (tid + ch * THREADS);
                    T y[VEC];
                    vec_unpack<T>(*reinterpret_cast<const VecT *>(tp + (size_t)c * ld + r0), y);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        T r = y[v];
#pragma unroll
                        for (int k = 0; k < N; ++k) r = fma(-pan[ch][v][k], b[k], r);
                        rsum = fma(r, r, rsum);
                    }
                }
            }
            rn2 += (double)rsum;
        }
    }

    // ---- CTA partial -> global, last CTA folds -------------------------------
    __syncthreads();
    cta_publish<T, N, P, CT, NW>(a, rn2, Gacc, Vacc, wsum_s, gv_s, fin_sh, fin_scratch, &is_last);
}

// -----------------------------------------------------------------------------
// Generic fallback (any n <= VP_MAX_N, p <= VP_MAX_P, any m): the streaming pass for model shapes without a
// specialised kernel. One persistent CTA per SM keeps the panel [Q | E] in shared memory (staged once per launch;
// read from global memory through L1/L2 if it does not fit); every warp takes TWO columns at a time, so each
// panel entry read from shared memory serves two right-hand sides; y is read twice (the second time from L2).
// NB = compile-time bound of n + p (8 / 12 / 20: the accumulators are registers with static indices).
// Per-warp accumulators in shared memory, folded over the warps in order: no floating-point atomics, the result
// is reproducible. Same outputs and partial layout as the specialised kernels.
// -----------------------------------------------------------------------------
template <typename T, int THREADS, int NB>
__global__ void __launch_bounds__(THREADS)
stream_kernel_generic(const StreamArgs<T> a, int n, int p, int m, int mp /* row stride of the staged panel, 0 = not staged */)
{
    constexpr int NW = THREADS / 32;
    constexpr int NVMAX = 1 + VP_MAX_N * (VP_MAX_N + 1) / 2 + VP_MAX_P;
    extern __shared__ __align__(16) double pan_s[]; // (n + p) columns of mp doubles
    __shared__ double acc_s[NVMAX];
    __shared__ double wacc_s[NW][NVMAX];
    __shared__ double rinv_s[VP_MAX_N * VP_MAX_N];
    __shared__ double fin_scratch[FIN_SCRATCH];
    __shared__ double fin_sh[64];
    __shared__ int is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld, npv = n + p;
    const int nv = red_count(n, p);
    T *Cout = stream_cout(a);
    for (int t = tid; t < NW * NVMAX; t += THREADS) (&wacc_s[0][0])[t] = 0.0;
    for (int t = tid; t < VP_MAX_N * VP_MAX_N; t += THREADS) rinv_s[t] = a.small->Rinv[t];
    const bool staged = mp > 0;
    if (staged)
        for (int idx = tid; idx < npv * mp; idx += THREADS) {
            const int k = idx / mp, i = idx - k * mp;
            pan_s[idx] = i < m ? (double)a.Pq[(size_t)k * a.ldp + i] : 0.0; // (Pe = Pq + n * ldp: the columns are contiguous)
        }
    __syncthreads();
    auto pan = [&](const int k, const int i) -> double {
        return staged ? pan_s[k * mp + i] : (double)a.Pq[(size_t)k * a.ldp + i];
    };

    double *wacc = wacc_s[warp];
    for (long long s0 = 2ll * ((long long)blockIdx.x * NW + warp); s0 < a.S; s0 += 2ll * gridDim.x * NW) {
        const bool two = s0 + 1 < a.S;
        const T *y0 = a.Y + (size_t)s0 * ld, *y1 = a.Y + (size_t)(two ? s0 + 1 : s0) * ld;
        double d0[NB], d1[NB];
#pragma unroll
        for (int k = 0; k < NB; ++k) d0[k] = d1[k] = 0.0;
#pragma unroll 4 // (eight observation loads in flight per lane: the pass is latency-bound otherwise)
        for (int i = lane; i < m; i += 32) {
            const double v0 = (double)y0[i], v1 = (double)y1[i];
#pragma unroll
            for (int k = 0; k < NB; ++k)
                if (k < npv) {
                    const double pk = pan(k, i);
                    d0[k] = fma(pk, v0, d0[k]);
                    d1[k] = fma(pk, v1, d1[k]);
                }
        }
#pragma unroll
        for (int k = 0; k < NB; ++k)
            if (k < npv) { d0[k] = warp_sum(d0[k]); d1[k] = warp_sum(d1[k]); }
        double rs0 = 0.0, rs1 = 0.0;
#pragma unroll 4
        for (int i = lane; i < m; i += 32) {
            double r0 = (double)y0[i], r1 = (double)y1[i];
#pragma unroll
            for (int k = 0; k < NB; ++k)
                if (k < n) {
                    const double pk = pan(k, i);
                    r0 = fma(-pk, d0[k], r0);
                    r1 = fma(-pk, d1[k], r1);
                }
            rs0 = fma(r0, r0, rs0);
            rs1 = fma(r1, r1, rs1);
        }
        rs0 = warp_sum(rs0);
        rs1 = warp_sum(rs1);
        if (lane == 0) {
#pragma unroll 1
            for (int c = 0; c < (two ? 2 : 1); ++c) {
                double coef[VP_MAX_N], dd[NB];
#pragma unroll
                for (int k = 0; k < NB; ++k) dd[k] = c ? d1[k] : d0[k];
                for (int r = 0; r < n; ++r) {
                    double sacc = 0.0;
#pragma unroll
                    for (int c2 = 0; c2 < VP_MAX_N; ++c2)
                        if (c2 < n && c2 < NB) sacc += rinv_s[c2 * VP_MAX_N + r] * dd[c2];
                    coef[r] = sacc;
                    Cout[(size_t)(s0 + c) * n + r] = (T)sacc;
                }
                wacc[0] += c ? rs1 : rs0;
                for (int r = 0; r < n; ++r)
                    for (int c2 = r; c2 < n; ++c2) wacc[g_index(n, r, c2)] += coef[r] * coef[c2];
#pragma unroll
                for (int e = 0; e < VP_MAX_P; ++e)
                    if (e < p && n + e < NB) {
                        double de = 0.0;
#pragma unroll
                        for (int k = 0; k < NB; ++k) de = (k == n + e) ? dd[k] : de;
                        wacc[1 + n * (n + 1) / 2 + e] += coef[a.e_basis[e]] * de;
                    }
            }
        }
    }
    __syncthreads();
    if (tid < nv) { // fold the warps' accumulators in order
        double t = 0.0;
        for (int w = 0; w < NW; ++w) t += wacc_s[w][tid];
        acc_s[tid] = t;
    }
    __syncthreads();
    stream_epilogue<T>(a, n, p, acc_s, fin_sh, fin_scratch, &is_last);
}

} // namespace vp
