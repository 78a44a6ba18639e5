// kernel_tables.h -- registry of the templated sm_100a kernel instantiations.
//
// The instantiations live in separate translation units (inst.cu compiled once per
// (dtype, n, p, part) so that nvcc can build them in parallel); each unit exports one
// vp_kernel_group_<tag>() returning its entries. vp_abi.cu concatenates the groups and
// launches through cudaLaunchKernel on the host stub addresses.
#pragma once

struct StreamKernelEntry {
    int dtype, n, p, threads, chunks, ct;
    const void *fn;
};
struct DmmaKernelEntry {
    int n, p, ksteps, nwarps, exact;
    const void *fn;
};
struct PanelHHEntry {
    int dtype, n, p, rpt, threads;
    const void *fn;
};
struct FitKernelEntry { // fit_kernel_dmma: fused panel + streaming reduce (+ whole LM loop)
    int dtype, n, p, ksteps, nwarps, exact;
    const void *fn;
};
struct QueueKernelEntry { // fit_queue_kernel: many fits on one persistent grid with a device-side work queue
    int dtype, n, p, ksteps, nwarps, exact;
    const void *fn;
};
struct BatchKernelEntry { // batch_fit_kernel: `threads` compute threads + `lm_warps` LM warps; `slots` problems in flight per CTA
    int n, p, rpt, threads, slots, lm_warps;
    const void *fn;
};
struct KernelGroup {
    const StreamKernelEntry *simt; int nsimt;
    const DmmaKernelEntry *dmma;   int ndmma;
    const PanelHHEntry *panel;     int npanel;
    const FitKernelEntry *fit;     int nfit;
    const BatchKernelEntry *batch; int nbatch;
    const QueueKernelEntry *queue; int nqueue;
};
typedef const KernelGroup *(*KernelGroupFn)();

// (tag, C type, vp_dtype, n, p, part): part 0 = SIMT streaming, 1 = DMMA streaming, 2 = Householder panel,
// 3 = fused evaluation / persistent fit kernel, 4 = independent-batch fit kernel,
// 5 = multi-fit work-queue kernel.
// The model shapes with a compiled fast path; everything else runs the generic kernels. The benchmark
// shape (3 basis functions, 2 derivative columns) has every variant; the other shapes keep the row
// tilings their reference workloads use (clean-build time is dominated by these instantiations).
#define VP_KERNEL_GROUPS(X)                 \
    X(f64_3_2_simt, double, VP_F64, 3, 2, 0) /* double exponential + offset (benches, C1/C2/C5) */ \
    X(f64_3_2_dmma, double, VP_F64, 3, 2, 1) \
    X(f64_3_2_panel, double, VP_F64, 3, 2, 2) \
    X(f64_3_2_fit0, double, VP_F64, 3, 2, 3) \
    X(f64_3_2_fit1, double, VP_F64, 3, 2, 3) \
    X(f64_3_2_fit2, double, VP_F64, 3, 2, 3) \
    X(f64_3_2_queue0, double, VP_F64, 3, 2, 5) \
    X(f64_3_2_queue1, double, VP_F64, 3, 2, 5) \
    X(f64_3_2_queue2, double, VP_F64, 3, 2, 5) \
    X(f32_3_2_simt, float, VP_F32, 3, 2, 0)  /* the same in fp32 (C4) */ \
    X(f32_3_2_panel, float, VP_F32, 3, 2, 2) \
    X(f32_3_2_fit0, float, VP_F32, 3, 2, 3) \
    X(f32_3_2_fit1, float, VP_F32, 3, 2, 3) \
    X(f32_3_2_fit2, float, VP_F32, 3, 2, 3) \
    X(f32_3_2_queue0, float, VP_F32, 3, 2, 5) \
    X(f32_3_2_queue1, float, VP_F32, 3, 2, 5) \
    X(f32_3_2_queue2, float, VP_F32, 3, 2, 5) \
    X(f64_3_3_simt, double, VP_F64, 3, 3, 0) /* triple exponential (C3 shape) */ \
    X(f64_3_3_dmma, double, VP_F64, 3, 3, 1) \
    X(f64_3_3_panel, double, VP_F64, 3, 3, 2) \
    X(f64_3_3_fit0, double, VP_F64, 3, 3, 3) \
    X(f64_3_3_fit1, double, VP_F64, 3, 3, 3) \
    X(f64_3_3_queue0, double, VP_F64, 3, 3, 5) /* the small row tilings, like fit0: vp_fit_many stays bitwise vp_fit at every m */ \
    X(f64_3_3_queue1, double, VP_F64, 3, 3, 5) \
    X(f64_2_4_dmma, double, VP_F64, 2, 4, 1) \
    X(f64_2_4_panel, double, VP_F64, 2, 4, 2) \
    X(f64_2_4_fit0, double, VP_F64, 2, 4, 3) \
    X(f64_2_4_fit1, double, VP_F64, 2, 4, 3) \
    X(f64_2_4_queue0, double, VP_F64, 2, 4, 5) /* the small row tilings, like fit0: vp_fit_many stays bitwise vp_fit at every m */ \
    X(f64_2_4_queue1, double, VP_F64, 2, 4, 5) \
    X(f64_3_3_batch, double, VP_F64, 3, 3, 4) \
    X(f64_3_2_batch, double, VP_F64, 3, 2, 4) \
    X(f64_2_4_batch, double, VP_F64, 2, 4, 4) \
    X(f64_2_1_dmma, double, VP_F64, 2, 1, 1) /* one exponential + offset */ \
    X(f64_2_1_panel, double, VP_F64, 2, 1, 2) \
    X(f64_2_1_fit0, double, VP_F64, 2, 1, 3) \
    X(f64_2_1_fit1, double, VP_F64, 2, 1, 3) \
    X(f64_2_1_queue0, double, VP_F64, 2, 1, 5) /* the small row tilings, like fit0: vp_fit_many stays bitwise vp_fit at every m */ \
    X(f64_2_1_queue1, double, VP_F64, 2, 1, 5) \
    X(f64_2_1_batch, double, VP_F64, 2, 1, 4) \
    X(f64_2_2_dmma, double, VP_F64, 2, 2, 1) /* double exponential without offset */ \
    X(f64_2_2_panel, double, VP_F64, 2, 2, 2) \
    X(f64_2_2_fit0, double, VP_F64, 2, 2, 3) \
    X(f64_2_2_fit1, double, VP_F64, 2, 2, 3) \
    X(f64_2_2_queue0, double, VP_F64, 2, 2, 5) /* the small row tilings, like fit0: vp_fit_many stays bitwise vp_fit at every m */ \
    X(f64_2_2_queue1, double, VP_F64, 2, 2, 5) \
    X(f64_2_2_batch, double, VP_F64, 2, 2, 4) \
    X(f64_4_3_dmma, double, VP_F64, 4, 3, 1) /* triple exponential + offset */ \
    X(f64_4_3_panel, double, VP_F64, 4, 3, 2) \
    X(f64_4_3_fit0, double, VP_F64, 4, 3, 3) \
    X(f64_4_3_fit1, double, VP_F64, 4, 3, 3) \
    X(f64_4_3_queue0, double, VP_F64, 4, 3, 5) /* the small row tilings, like fit0: vp_fit_many stays bitwise vp_fit at every m */ \
    X(f64_4_3_queue1, double, VP_F64, 4, 3, 5) \
    X(f64_4_3_batch, double, VP_F64, 4, 3, 4)

#define VP_DECLARE_GROUP(tag, T, DT, N, P, PART) const KernelGroup *vp_kernel_group_##tag();
VP_KERNEL_GROUPS(VP_DECLARE_GROUP)
#undef VP_DECLARE_GROUP
