// Shared device helpers for the HydraNet sm_100a kernels: error plumbing, PTX wrappers for
// mbarrier / TMA / tcgen05 / TMEM, small numeric helpers.  No reference code is restated here.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/hydranet_b200.h"

// ------------------------------------------------------------------------------------------------
// error plumbing (C-ABI: every entry point returns int, message kept per thread)
// ------------------------------------------------------------------------------------------------
void hn_set_error(const char* fmt, ...);

#define HN_CHECK_CUDA(expr)                                                                        \
    do {                                                                                           \
        cudaError_t _e = (expr);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            hn_set_error("%s:%d CUDA error %d (%s) in %s", __FILE__, __LINE__, (int)_e,            \
                         cudaGetErrorString(_e), #expr);                                           \
            return HN_ERR_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

#define HN_REQUIRE(cond, ...)                                                                      \
    do {                                                                                           \
        if (!(cond)) {                                                                             \
            hn_set_error(__VA_ARGS__);                                                             \
            return HN_ERR_ARG;                                                                     \
        }                                                                                          \
    } while (0)

static inline int hn_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// cuTensorMapEncodeTiled is a DRIVER entry point and needs a current context on the calling thread.  Threads that have
// only ever been handed device pointers (PyTorch's autograd worker running a backward) may not have one yet: bind the
// runtime's primary context once per thread.
static inline void hn_ensure_context() {
    static thread_local bool done = false;
    if (!done) {
        cudaFree(nullptr);
        done = true;
    }
}

// ------------------------------------------------------------------------------------------------
// device-side PTX wrappers
// ------------------------------------------------------------------------------------------------
#ifdef __CUDACC__

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ uint32_t hn_smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ bool hn_elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .b32 %%rx;\n\t"
        ".reg .pred %%px;\n\t"
        "elect.sync %%rx|%%px, %1;\n\t"
        "@%%px mov.s32 %0, 1;\n\t"
        "}\n"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}

// ---- mbarrier ----
__device__ __forceinline__ void hn_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(hn_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void hn_mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void hn_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(hn_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void hn_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(hn_smem_u32(bar)) : "memory");
}
// The suspend-time hint lets the hardware park a waiting warp until the phase completes instead of returning
// early: a polling producer / MMA lane otherwise eats a fifth of its scheduler's issue slots, which the epilogue
// warps of the same SM partition need.
static constexpr uint32_t kMbarSuspendNs = 20000;
__device__ __forceinline__ void hn_mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = hn_smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(kMbarSuspendNs)
            : "memory");
    } while (!done);
}

// ---- programmatic dependent launch ----
// With hn_set_pdl(1) every kernel of the forward plan is launched with cudaLaunchAttributeProgrammaticStreamSerialization: its CTAs may
// be scheduled while the previous kernel's last CTAs are still running, run their prologue (barrier init, TMEM
// allocation, descriptor prefetch, index math) and then block in hn_pdl_wait() until the previous grid has
// completed and its writes are visible.  NOTHING produced by an earlier kernel may be read, and nothing an earlier
// kernel may still read may be written, before hn_pdl_wait().
__device__ __forceinline__ void hn_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void hn_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- TMA ----
__device__ __forceinline__ void hn_tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)tmap) : "memory");
}
__device__ __forceinline__ void hn_tma_load_2d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            hn_smem_u32(dst)),
        "l"((uint64_t)tmap), "r"(hn_smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void hn_tma_load_4d(void* dst, const void* tmap, uint64_t* bar, int c0, int c1, int c2,
                                               int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(hn_smem_u32(dst)),
        "l"((uint64_t)tmap), "r"(hn_smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- CTA pair (cta_group::2): both CTAs load their halves, the barrier of the leader CTA (cluster rank 0)
// collects the transaction bytes of both ----
__device__ __forceinline__ uint32_t hn_mapa(uint32_t smem_addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(cta_rank));
    return r;
}
__device__ __forceinline__ void hn_tma_load_2d_pair(void* dst, const void* tmap, uint32_t leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            hn_smem_u32(dst)),
        "l"((uint64_t)tmap), "r"(leader_bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void hn_tma_load_4d_pair(void* dst, const void* tmap, uint32_t leader_bar, int c0, int c1, int c2,
                                                    int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
        "[%2];" ::"r"(hn_smem_u32(dst)),
        "l"((uint64_t)tmap), "r"(leader_bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void hn_mbar_arrive_remote(uint32_t cluster_bar_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar_addr) : "memory");
}
__device__ __forceinline__ uint32_t hn_cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void hn_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__device__ __forceinline__ void hn_tma_store_4d(const void* tmap, const void* src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"((uint64_t)tmap),
                 "r"(hn_smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void hn_tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void hn_tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void hn_tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void hn_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void hn_tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hn_smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void hn_tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void hn_tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void hn_tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void hn_tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem], bf16 x bf16 -> fp32, single CTA
__device__ __forceinline__ void hn_umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive once all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void hn_umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(hn_smem_u32(bar))
                 : "memory");
}
// ---- cta_group::2 flavours: one thread of the leader CTA drives the tensor cores of both SMs (M = 256) ----
__device__ __forceinline__ void hn_tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(hn_smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void hn_tmem_relinquish_pair() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void hn_tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void hn_umma_bf16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// arrives on the barrier at this offset in both CTAs of the pair once the issued MMAs have retired
__device__ __forceinline__ void hn_umma_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     hn_smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void hn_tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void hn_tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),
          "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]),
          "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void hn_tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, 128-byte-swizzled operand tile: rows of 128 B, 8-row groups 1024 B apart
// (cute/arch/mma_sm100_desc.hpp SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
//  version=1 [46,48), layout_type SWIZZLE_128B=2 [61,64)).
__device__ __forceinline__ uint64_t hn_umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;            // LBO (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;  // SBO
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M x N
__device__ __forceinline__ uint32_t hn_umma_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- numerics ----
__device__ __forceinline__ float hn_ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// sigmoid for the swish activations: 0.5 * tanh(x/2) + 0.5 with the hardware tanh (one MUFU op, 3 instructions; abs error
// < 2.5e-4, an order of magnitude below the bf16 rounding of the value it feeds).  The heads' final sigmoid (scores
// that are thresholded) uses the accurate expf form in hn_act.
__device__ __forceinline__ float hn_sigmoid(float x) {
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
}
__device__ __forceinline__ float hn_act(float x, int act) {
    switch (act) {
        case HN_ACT_RELU: return fmaxf(x, 0.0f);
        case HN_ACT_SWISH: return x * hn_sigmoid(x);
        case HN_ACT_ELU: return x > 0.0f ? x : expm1f(x);
        case HN_ACT_SIGMOID: return 1.0f / (1.0f + expf(-x));
        default: return x;
    }
}
__device__ __forceinline__ uint32_t hn_pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 hn_unpack_bf16x2(uint32_t u) {
    __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
    return __bfloat1622float2(v);
}

#endif  // __CUDACC__

// host: launch with the programmatic-serialization attribute (hn_set_pdl(0) turns it into a plain launch)
extern int g_hn_pdl;
template <typename... KArgs, typename... Args>
static inline cudaError_t hn_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_hn_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
