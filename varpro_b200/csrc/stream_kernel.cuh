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
#include "panel_kernel.cuh"

namespace vp {

struct EvalOut {
    double rnorm2;
    double g[VP_MAX_Q];
    double H[VP_MAX_Q * VP_MAX_Q]; // column-major, ld = q
    int finite;
    int pad;
};

template <typename T>
struct StreamArgs {
    const T *Y;      // m x S weighted observations, ld rows per column
    int ld;          // padded rows (multiple of 16/sizeof(T))
    int S;
    const T *Pq;     // n x ld
    const T *Pe;     // p x ld
    const PanelSmall *small;
    T *Cout;         // n x S coefficients (trial buffer)
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

template <typename T> struct VecOf;
template <> struct VecOf<double> { using type = double2; static constexpr int N = 2; };
template <> struct VecOf<float> { using type = float4; static constexpr int N = 4; };

template <typename T> __device__ __forceinline__ void vec_unpack(const double2 &v, T (&o)[2]) { o[0] = v.x; o[1] = v.y; }
template <typename T> __device__ __forceinline__ void vec_unpack(const float4 &v, T (&o)[4]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w; }

// Final fold executed by the last CTA: partials -> (rnorm2, g, H).
template <typename T>
__device__ void stream_finalize(const StreamArgs<T> &a, int n, int p, int nparts, double *sh /* >= 64 doubles */)
{
    const int nv = red_count(n, p);
    const int tid = threadIdx.x;
    if (tid < nv) {
        // fixed summation order over CTAs => bitwise reproducible for a given grid
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        int c = 0;
        for (; c + 4 <= nparts; c += 4) {
            s0 += __ldcg(a.partials + (size_t)(c + 0) * a.red_stride + tid);
            s1 += __ldcg(a.partials + (size_t)(c + 1) * a.red_stride + tid);
            s2 += __ldcg(a.partials + (size_t)(c + 2) * a.red_stride + tid);
            s3 += __ldcg(a.partials + (size_t)(c + 3) * a.red_stride + tid);
        }
        for (; c < nparts; ++c) s0 += __ldcg(a.partials + (size_t)c * a.red_stride + tid);
        sh[tid] = (s0 + s1) + (s2 + s3);
    }
    __syncthreads();
    if (tid == 0) {
        const PanelSmall *sm = a.small;
        EvalOut *o = a.out;
        const int q = a.q;
        double rn2 = sh[0];
        int finite = isfinite(rn2) && !sm->nonfinite;
        for (int k = 0; k < q; ++k) {
            double gk = 0.0;
            for (int e = 0; e < p; ++e)
                if (a.e_param[e] == k) gk -= sh[1 + n * (n + 1) / 2 + e];
            o->g[k] = gk;
            finite = finite && isfinite(gk);
        }
        for (int k = 0; k < q; ++k)
            for (int l = 0; l <= k; ++l) {
                double h = 0.0;
                for (int e = 0; e < p; ++e) {
                    if (a.e_param[e] != k) continue;
                    for (int f = 0; f < p; ++f) {
                        if (a.e_param[f] != l) continue;
                        int i = a.e_basis[e], j = a.e_basis[f];
                        if (i > j) { int t = i; i = j; j = t; }
                        h += sm->M[f * VP_MAX_P + e] * sh[g_index(n, i, j)];
                    }
                }
                o->H[l * q + k] = h;
                o->H[k * q + l] = h;
                finite = finite && isfinite(h);
            }
        o->rnorm2 = rn2;
        o->finite = finite;
        *a.ticket = 0; // re-arm for the next launch
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
    __shared__ double fin_scratch[(NW + 1) * NVR + 64];
    __shared__ int is_last;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld, S = a.S, nst = a.nstages;
    const size_t stage_elems = (size_t)CT * ld;
    T *tiles = reinterpret_cast<T *>(smem_raw);

    const int ntiles = (S + CT - 1) / CT;
    const int my = ((int)blockIdx.x < ntiles) ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

    if (tid == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(&full_bar[s], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue = [&](int i) {
        const int tile = blockIdx.x + i * gridDim.x;
        const int col0 = tile * CT;
        const int nc = min(CT, S - col0);
        const uint32_t bytes = (uint32_t)((size_t)nc * ld * sizeof(T));
        const int st = i % nst;
        mbar_arrive_expect_tx(&full_bar[st], bytes);
        bulk_copy_g2s(tiles + (size_t)st * stage_elems, a.Y + (size_t)col0 * ld, bytes, &full_bar[st]);
    };
    // the observations do not depend on the panel: start fetching immediately
    if (tid == 0)
        for (int i = 0; i < nst && i < my; ++i) issue(i);

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
                const T *src = (k < N) ? (a.Pq + (size_t)k * ld + r0) : (a.Pe + (size_t)(k - N) * ld + r0);
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

    for (int i = 0; i < my; ++i) {
        const int st = i % nst;
        const uint32_t parity = (uint32_t)((i / nst) & 1);
        const int tile = blockIdx.x + i * gridDim.x;
        const int col0 = tile * CT;
        const int nc = min(CT, S - col0);
        const T *tp = tiles + (size_t)st * stage_elems;
        mbar_wait(&full_bar[st], parity);

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
        {
            int cnt = NACC;
#pragma unroll
            for (int h = 0; h < LOG2CT; ++h) {
                const int off = 16 >> h;
                const int half = cnt / 2;
                const bool up = (lane & off) != 0;
#pragma unroll
                for (int t = 0; t < half; ++t) {
                    const double send = up ? red[t] : red[t + half];
                    const double keep = up ? red[t + half] : red[t];
                    red[t] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
                cnt = half;
            }
#pragma unroll
            for (int off = (16 >> LOG2CT); off > 0; off >>= 1)
#pragma unroll
                for (int k = 0; k < NPV; ++k) red[k] += __shfl_xor_sync(0xffffffffu, red[k], off);
        }
        if ((lane & ((32 >> LOG2CT) - 1)) == 0) {
            const int c = lane >> (5 - LOG2CT);
#pragma unroll
            for (int k = 0; k < NPV; ++k) part[warp * NACC + c * NPV + k] = red[k];
        }
        __syncthreads(); // (A) every thread is past phase 2 of the previous tile
        if (tid == 0 && i >= 1 && (i - 1 + nst) < my) issue(i - 1 + nst); // refill the stage tile i-1 used
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
                for (int c2 = r; c2 < N; ++c2) s += rinv_s[c2 * N + r] * bu[tid * NPV + c2];
                coef[r] = s;
                a.Cout[(size_t)(col0 + tid) * N + r] = (T)s;
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
    {
        double fin[NVR];
        fin[0] = rn2;
#pragma unroll
        for (int t = 0; t < N * (N + 1) / 2; ++t) fin[1 + t] = (tid < CT) ? Gacc[t] : 0.0;
#pragma unroll
        for (int e = 0; e < P; ++e) fin[1 + N * (N + 1) / 2 + e] = (tid < CT) ? Vacc[e] : 0.0;
        block_sum<NVR>(fin, fin_scratch);
        if (tid == 0) {
#pragma unroll
            for (int t = 0; t < NVR; ++t) a.partials[(size_t)blockIdx.x * a.red_stride + t] = fin[t];
            __threadfence();
            const unsigned int prev = atomicAdd(a.ticket, 1u);
            is_last = (prev == gridDim.x - 1);
        }
        __syncthreads();
        if (is_last) {
            __threadfence();
            stream_finalize<T>(a, N, P, gridDim.x, fin_scratch);
        }
    }
}

// -----------------------------------------------------------------------------
// Generic fallback (any n <= VP_MAX_N, p <= VP_MAX_P, any m): one warp per
// column, panel read through L1/L2, y read twice (second read hits L1). Same
// outputs and partial layout as the fast kernel. Correctness path for shapes
// the register-panel kernel is not instantiated for; not tuned.
// -----------------------------------------------------------------------------
template <typename T, int THREADS>
__global__ void __launch_bounds__(THREADS)
stream_kernel_generic(const StreamArgs<T> a, int n, int p, int m)
{
    constexpr int NW = THREADS / 32;
    constexpr int NVMAX = 1 + VP_MAX_N * (VP_MAX_N + 1) / 2 + VP_MAX_P;
    __shared__ double acc_s[NVMAX];
    __shared__ double rinv_s[VP_MAX_N * VP_MAX_N];
    __shared__ double fin_scratch[64 + NVMAX];
    __shared__ int is_last;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = a.ld;
    const int nv = red_count(n, p);
    for (int t = tid; t < NVMAX; t += THREADS) acc_s[t] = 0.0;
    for (int t = tid; t < VP_MAX_N * VP_MAX_N; t += THREADS) rinv_s[t] = a.small->Rinv[t];
    __syncthreads();

    for (int s = blockIdx.x * NW + warp; s < a.S; s += gridDim.x * NW) {
        const T *y = a.Y + (size_t)s * ld;
        double d[VP_MAX_N + VP_MAX_P];
#pragma unroll
        for (int k = 0; k < VP_MAX_N + VP_MAX_P; ++k) d[k] = 0.0;
        for (int i = lane; i < m; i += 32) {
            const double yi = (double)y[i];
#pragma unroll
            for (int k = 0; k < VP_MAX_N; ++k)
                if (k < n) d[k] += (double)a.Pq[(size_t)k * ld + i] * yi;
#pragma unroll
            for (int e = 0; e < VP_MAX_P; ++e)
                if (e < p) d[VP_MAX_N + e] += (double)a.Pe[(size_t)e * ld + i] * yi;
        }
#pragma unroll
        for (int k = 0; k < VP_MAX_N + VP_MAX_P; ++k) d[k] = warp_sum(d[k]);
        double rs = 0.0;
        for (int i = lane; i < m; i += 32) {
            double r = (double)y[i];
#pragma unroll
            for (int k = 0; k < VP_MAX_N; ++k)
                if (k < n) r -= (double)a.Pq[(size_t)k * ld + i] * d[k];
            rs += r * r;
        }
        rs = warp_sum(rs);
        if (lane == 0) {
            double coef[VP_MAX_N];
            for (int r = 0; r < n; ++r) {
                double sacc = 0.0;
                for (int c2 = r; c2 < n; ++c2) sacc += rinv_s[c2 * VP_MAX_N + r] * d[c2];
                coef[r] = sacc;
                a.Cout[(size_t)s * n + r] = (T)sacc;
            }
            atomicAdd(&acc_s[0], rs);
            for (int r = 0; r < n; ++r)
                for (int c2 = r; c2 < n; ++c2) atomicAdd(&acc_s[g_index(n, r, c2)], coef[r] * coef[c2]);
            for (int e = 0; e < p; ++e)
                atomicAdd(&acc_s[1 + n * (n + 1) / 2 + e], coef[a.e_basis[e]] * d[VP_MAX_N + e]);
        }
    }
    __syncthreads();
    if (tid == 0) {
        for (int t = 0; t < nv; ++t) a.partials[(size_t)blockIdx.x * a.red_stride + t] = acc_s[t];
        __threadfence();
        const unsigned int prev = atomicAdd(a.ticket, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
        stream_finalize<T>(a, n, p, gridDim.x, fin_scratch);
    }
}

} // namespace vp
