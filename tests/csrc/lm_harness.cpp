// Host-side harness around the product's lm_step.cuh (which is __host__ __device__), so that the
// CPU test-suite can drive the exact LM state machine the GPU runs, with evaluations supplied by
// the test (e.g. H = J^T J, g = J^T r computed from the oracle's explicit J and r).
// Test infrastructure only.
#include <cstring>
#include "../../varpro_b200/csrc/lm_step.cuh"
#include "../../varpro_b200/csrc/rank_policy.cuh"

using namespace vp;

struct Harness {
    LmState st;
    LmConfig cfg;
    int generic = 0; // 1: force the run-time-q instantiation (lm_advance_generic) instead of the q-specialised one
};

extern "C" {
Harness *lmh_new(int q, const double *x0, double ftol, double xtol, double gtol, double stepbound, int maxfev,
                 int scale_diag, double epsmch)
{
    Harness *h = new Harness();
    h->cfg.ftol = ftol; h->cfg.xtol = xtol; h->cfg.gtol = gtol; h->cfg.stepbound = stepbound;
    h->cfg.maxfev = maxfev; h->cfg.scale_diag = scale_diag; h->cfg.epsmch = epsmch;
    lm_init(h->st, q, x0);
    return h;
}
void lmh_free(Harness *h) { delete h; }
// feed an evaluation made at lmh_trial(); returns 1 if another evaluation is needed
int lmh_advance(Harness *h, double rnorm2, const double *g, const double *H /* q x q col-major */, int finite)
{
    LmEval ev;
    std::memset(&ev, 0, sizeof(ev));
    const int q = h->st.q;
    ev.rnorm2 = rnorm2; ev.finite = finite ? VP_EVAL_ALL_OK : 0; // finite = 0: the evaluation failed (residuals() is None)
    for (int k = 0; k < q; ++k) ev.g[k] = g[k];
    for (int i = 0; i < q * q; ++i) ev.H[i] = H[i];
    return (h->generic ? lm_advance_generic(h->st, h->cfg, ev) : lm_advance(h->st, h->cfg, ev)) ? 1 : 0;
}
void lmh_set_generic(Harness *h, int on) { h->generic = on; }
// raw state words (bitwise comparison of the specialised and the generic instantiation)
int lmh_state_bytes(void) { return (int)sizeof(LmState); }
void lmh_state(const Harness *h, void *out) { std::memcpy(out, &h->st, sizeof(LmState)); }
void lmh_trial(const Harness *h, double *x) { for (int k = 0; k < h->st.q; ++k) x[k] = h->st.x_trial[k]; }
void lmh_accepted(const Harness *h, double *x) { for (int k = 0; k < h->st.q; ++k) x[k] = h->st.x[k]; }
int lmh_termination(const Harness *h) { return h->st.termination; }
int lmh_nfev(const Harness *h) { return h->st.nfev; }
int lmh_last_accepted(const Harness *h) { return h->st.last_accepted; }
double lmh_fnorm(const Harness *h) { return h->st.fnorm; }
double lmh_par(const Harness *h) { return h->st.par; }
double lmh_delta(const Harness *h) { return h->st.delta; }

// rank policy (rank_policy.cuh): R is n x n upper triangular, column-major with leading dimension n.
// Returns 1 if the cheap bound proves full rank (then the outputs are untouched), else runs the SVD path:
// truncated flag, Urot = Ur diag(keep), RinvEff = V diag(keep / sigma), both n x n column-major.
int rph_policy(int n, const double *R, const double *Rinv, double tol, int *truncated, double *Urot, double *RinvEff)
{
    if (rank_surely_full(n, R, n, Rinv, n, tol)) return 1;
    SmallSvd sv;
    rank_policy_svd(n, R, n, tol, &sv);
    *truncated = sv.truncated;
    for (int i = 0; i < n * n; ++i) { Urot[i] = sv.Urot[i]; RinvEff[i] = sv.RinvEff[i]; }
    return 0;
}
}
