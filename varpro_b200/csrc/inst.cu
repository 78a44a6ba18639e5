// inst.cu -- one group of kernel instantiations; compiled once per entry of VP_KERNEL_GROUPS with
//   -DVP_INST_TAG=<tag> -DVP_INST_T=<double|float> -DVP_INST_DT=<VP_F64|VP_F32> -DVP_INST_N=n -DVP_INST_P=p -DVP_INST_PART=<0|1|2>
// Each part includes only the kernel header it instantiates (build.py follows these conditionals when it decides
// which objects a header edit invalidates).
#include "kernel_tables.h"
#if VP_INST_PART == 0
#include "stream_kernel.cuh"
#elif VP_INST_PART == 1
#include "stream_kernel_dmma.cuh"
#elif VP_INST_PART == 3
#include "fit_kernel_dmma.cuh"
#elif VP_INST_PART == 4
#include "batch_fit_kernel.cuh"
#elif VP_INST_PART == 5
#include "fit_queue_kernel.cuh"
#else
#include "panel_kernel_hh.cuh"
#endif

using namespace vp;

#define VP_CAT2(a, b) a##b
#define VP_CAT(a, b) VP_CAT2(a, b)

constexpr int N_ = VP_INST_N, P_ = VP_INST_P;
typedef VP_INST_T T_;

#if VP_INST_PART == 0
#define VP_SK(THREADS, CHUNKS, CT) \
    {VP_INST_DT, N_, P_, THREADS, CHUNKS, CT, (const void *)&stream_kernel<T_, N_, P_, CHUNKS, CT, THREADS>}
static const StreamKernelEntry simt_tab[] = {VP_SK(128, 1, 4), VP_SK(128, 4, 4), VP_SK(128, 4, 8), VP_SK(256, 4, 4),
                                             VP_SK(256, 8, 4)};
static const KernelGroup group = {simt_tab, (int)(sizeof(simt_tab) / sizeof(simt_tab[0])), nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0};
#elif VP_INST_PART == 1
// EXACT (unpredicated fragment loads) only for the 1024-row tiling the benchmark shapes use
#define VP_DK(KS, NW, EX) {N_, P_, KS, NW, EX, (const void *)&stream_kernel_dmma<N_, P_, KS, NW, (EX) != 0>}
static const DmmaKernelEntry dmma_tab[] = {VP_DK(8, 4, 0), VP_DK(16, 8, 0), VP_DK(32, 8, 0), VP_DK(32, 8, 1), VP_DK(32, 16, 0)};
static const KernelGroup group = {nullptr, 0, dmma_tab, (int)(sizeof(dmma_tab) / sizeof(dmma_tab[0])), nullptr, 0, nullptr, 0};
#elif VP_INST_PART == 3
#define VP_FK(KS, NW, EX) {VP_INST_DT, N_, P_, KS, NW, EX, (const void *)&fit_kernel_dmma<T_, N_, P_, KS, NW, (EX) != 0>}
#if VP_INST_VARIANT == 0
static const FitKernelEntry fit_tab[] = {VP_FK(8, 4, 0), VP_FK(16, 8, 0)};
#elif VP_INST_VARIANT == 1
static const FitKernelEntry fit_tab[] = {VP_FK(32, 8, 0), VP_FK(32, 8, 1)};
#else
static const FitKernelEntry fit_tab[] = {VP_FK(32, 16, 0)};
#endif
static const KernelGroup group = {nullptr, 0, nullptr, 0, nullptr, 0, fit_tab, (int)(sizeof(fit_tab) / sizeof(fit_tab[0]))};
#elif VP_INST_PART == 5
#define VP_QK(KS, NW, EX) {VP_INST_DT, N_, P_, KS, NW, EX, (const void *)&fit_queue_kernel<T_, N_, P_, KS, NW, (EX) != 0>}
#if VP_INST_VARIANT == 0
static const QueueKernelEntry queue_tab[] = {VP_QK(8, 4, 0), VP_QK(16, 8, 0)};
#elif VP_INST_VARIANT == 1
static const QueueKernelEntry queue_tab[] = {VP_QK(32, 8, 0), VP_QK(32, 8, 1)};
#else
static const QueueKernelEntry queue_tab[] = {VP_QK(32, 16, 0)};
#endif
static const KernelGroup group = {nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, queue_tab, (int)(sizeof(queue_tab) / sizeof(queue_tab[0]))};
#elif VP_INST_PART == 4
constexpr int BATCH_THREADS = 512;
#define VP_BK(RPT, G) {N_, P_, RPT, BATCH_THREADS, G, 1, (const void *)&batch_fit_kernel<N_, P_, RPT, BATCH_THREADS, G>}
// slots: two groups of G / 2; one group's LM steps (lockstep in the LM warp, ~40 us) must fit into the other group's evaluations
constexpr int G_SMALL = (N_ + P_ <= 6) ? 32 : 16; // (static shared memory <= 48 KB)
static const BatchKernelEntry batch_tab[] = {VP_BK(1, G_SMALL), VP_BK(2, G_SMALL), VP_BK(4, G_SMALL), VP_BK(8, 16)};
static const KernelGroup group = {nullptr, 0, nullptr, 0, nullptr, 0, nullptr, 0, batch_tab, (int)(sizeof(batch_tab) / sizeof(batch_tab[0]))};
#else
constexpr int PANEL_HH_THREADS = 512;
#define VP_PK(RPT) {VP_INST_DT, N_, P_, RPT, PANEL_HH_THREADS, (const void *)&panel_kernel_hh<T_, N_, P_, RPT, PANEL_HH_THREADS>}
static const PanelHHEntry panel_tab[] = {VP_PK(1), VP_PK(2), VP_PK(4), VP_PK(8)};
static const KernelGroup group = {nullptr, 0, nullptr, 0, panel_tab, (int)(sizeof(panel_tab) / sizeof(panel_tab[0])), nullptr, 0};
#endif

const KernelGroup *VP_CAT(vp_kernel_group_, VP_INST_TAG)() { return &group; }
