// Convolution weight gradient on the 5th-generation tensor cores (reference: the dW half of autograd's conv backward behind
// model/train.py:262).  dW[co][tap][ci] = sum over pixels of dY[pix][co] * X[pix + tap][ci]: a GEMM whose K dimension is
// the PIXEL axis.  NHWC tensors are [pixel][channel] matrices, i.e. both operands are "MN-major" for this product, which
// tcgen05 reads natively: a TMA box {64 channels, 128 pixels} lands as 128 rows of 128 bytes, 128B-swizzled -- exactly the
// canonical MN-major SWIZZLE_128B layout (8-pixel x 64-channel atoms of 1024 bytes; SBO = 1024 B between pixel atoms,
// LBO = one box between 64-channel atoms).  No transposes, no im2col.
//
// Work item (one CTA) = (128 output channels) x (up to 4 taps = 64-input-channel blocks) x (a slice of the pixel tiles);
// the taps of an item share the dY tile of every stage.  Accumulators: 128 TMEM lanes (co) x 64 columns per tap.  Split-K
// partial results are added into dW with fp32 reductions (red.global.add) after a per-warp shared-memory transpose that
// makes a warp's 32 lanes cover 32 consecutive input channels of one output channel.
// CTA = 6 warps: TMA producer, MMA issuer (+ TMEM allocator), 4 epilogue warps.
#include <mutex>

#include "hn_ops.h"

static constexpr int kWgThreads = 192;
static constexpr int kBoxBytes = 128 * 128;  // 128 pixels x 64 channels bf16
static constexpr int kMaxTapsPerItem = 4;

struct alignas(64) WgradParams {
    CUtensorMap tmDy;
    CUtensorMap tmX[HN_MAX_SRC];
    int flat, TH, TW, n_img, H, W, tiles_x, tiles_y;
    int k_tiles, k_per_split, splits;
    int num_taps, nb, tap_groups, m_tiles;
    unsigned char grp_first[HN_MAX_TAPS], grp_len[HN_MAX_TAPS], grp_merge[HN_MAX_TAPS];
    int cout, grouped, stages, tmem_cols;
    long long s_co, s_ci;
    float* dw;
    hn_tap taps[HN_MAX_TAPS];
    long long tap_off[HN_MAX_TAPS];
    int tap_cin[HN_MAX_TAPS];
};

// MN-major SWIZZLE_128B operand: start address, LBO (between 64-element atoms along M / N), SBO (between 8-row atoms along K)
__device__ __forceinline__ uint64_t hn_umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// kind::f16, D = f32, A = B = bf16, both MN-major
__device__ __forceinline__ uint32_t hn_umma_idesc_bf16_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void hn_red_add_f32(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__global__ void __launch_bounds__(kWgThreads, 1) hn_conv_wgrad_kernel(const __grid_constant__ WgradParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // work item
    const int split = blockIdx.x % p.splits;
    const int rest = blockIdx.x / p.splits;
    const int tg = rest % p.tap_groups, mt = rest / p.tap_groups;
    const int kt0 = split * p.k_per_split, kt1 = min(kt0 + p.k_per_split, p.k_tiles);
    if (kt0 >= kt1) return;  // whole CTA: nothing to add
    const int t0 = p.grp_first[tg], nb = p.grp_len[tg];
    const bool merged = p.grp_merge[tg] != 0;
    const int m0 = mt * 128;

    const int stage_bytes = (2 + p.nb) * kBoxBytes;
    uint8_t* sStage = smem;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + p.stages * stage_bytes);
    uint64_t* bar_empty = bar_full + p.stages;
    uint64_t* bar_acc = bar_empty + p.stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 1);
    float* tr_base = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~uintptr_t(15));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0 && lane == 0) {
        hn_tma_prefetch_desc(&p.tmDy);
        for (int s = 0; s < p.stages; ++s) {
            hn_mbar_init(&bar_full[s], 1);
            hn_mbar_init(&bar_empty[s], 1);
        }
        hn_mbar_init(bar_acc, 1);
        hn_mbar_fence_init();
    }
    if (warp == 1) {
        hn_tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols);
        hn_tmem_relinquish();
    }
    hn_tc_fence_before();
    __syncthreads();
    hn_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int kt = kt0; kt < kt1; ++kt) {
                int img = 0, y0 = 0, x0;
                if (p.flat) {
                    x0 = kt * 128;
                } else {
                    const int per_img = p.tiles_x * p.tiles_y;
                    img = kt / per_img;
                    const int r = kt - img * per_img, ry = r / p.tiles_x;
                    y0 = ry * p.TH;
                    x0 = (r - ry * p.tiles_x) * p.TW;
                }
                hn_mbar_wait(&bar_empty[s], ph ^ 1);
                uint8_t* st = sStage + s * stage_bytes;
                // the second 64-channel block of dY only if it holds real output channels: the rows of a block that is not
                // loaded multiply whatever the slot held before -- each accumulator row depends on its own A row only, and
                // the epilogue never reads rows >= cout
                const bool second = m0 + 64 < p.cout;
                hn_mbar_expect_tx(&bar_full[s], (uint32_t)((1 + (second ? 1 : 0) + nb) * kBoxBytes));
                hn_tma_load_4d(st, &p.tmDy, &bar_full[s], m0, x0, y0, img);
                if (second) hn_tma_load_4d(st + kBoxBytes, &p.tmDy, &bar_full[s], m0 + 64, x0, y0, img);
                for (int j = 0; j < nb; ++j) {
                    const hn_tap tp = p.taps[t0 + j];
                    const int c0 = (int)tp.c0 + (p.grouped ? m0 : 0);
                    hn_tma_load_4d(st + (2 + j) * kBoxBytes, &p.tmX[tp.src], &bar_full[s], c0, x0 + (int)tp.dx, y0 + (int)tp.dy, img);
                }
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = hn_umma_idesc_bf16_mn(128, merged ? 64 * nb : 64);
            int s = 0;
            uint32_t ph = 0;
            for (int kt = kt0; kt < kt1; ++kt) {
                hn_mbar_wait(&bar_full[s], ph);
                hn_tc_fence_after();
                const uint32_t a_addr = hn_smem_u32(sStage + s * stage_bytes);
                for (int j = 0; j < (merged ? 1 : nb); ++j) {
                    // merged: the group's boxes are consecutive 64-channel blocks of one shifted window -> one N = 64 * nb operand
                    // (64-channel atoms one box apart: LBO)
                    const uint32_t b_addr = a_addr + (uint32_t)((2 + j) * kBoxBytes);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {  // 16 pixels (K) per instruction = 16 rows of 128 bytes
                        const uint64_t da = hn_umma_desc_mn_sw128(a_addr + kk * 2048, kBoxBytes);
                        const uint64_t db = hn_umma_desc_mn_sw128(b_addr + kk * 2048, kBoxBytes);
                        hn_umma_bf16(tmem_base + (uint32_t)(j * 64), da, db, idesc, (uint32_t)((kt != kt0) | (kk != 0)));
                    }
                }
                hn_umma_commit(&bar_empty[s]);
                if (++s == p.stages) { s = 0; ph ^= 1; }
            }
            hn_umma_commit(bar_acc);
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        const int q = warp & 3;  // TMEM lane quarter
        float* tr = tr_base + (warp - 2) * (32 * 33);
        hn_mbar_wait(bar_acc, 0);
        hn_tc_fence_after();
        const int co_base = m0 + q * 32;
        for (int j = 0; j < nb; ++j) {
            const int t = t0 + j;
            const int cin = p.tap_cin[t];
            const long long toff = p.tap_off[t];
            const int c_abs0 = p.grouped ? (int)p.taps[t].c0 + m0 : 0;  // grouped: absolute input channel of column 0
            for (int c = 0; c < 64; c += 32) {
                if (c >= cin) break;  // warp-uniform
                uint32_t v[32];
                hn_tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(j * 64 + c), v);
                hn_tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 32; ++k) tr[lane * 33 + k] = __uint_as_float(v[k]);
                __syncwarp();
                const int ci = c + lane;  // this lane's column within the block
                for (int rr = 0; rr < 32; ++rr) {
                    const int co = co_base + rr;
                    if (co >= p.cout) break;  // warp-uniform
                    if (ci < cin) {
                        const float val = tr[rr * 33 + lane];
                        if (!p.grouped) {
                            hn_red_add_f32(p.dw + (long long)co * p.s_co + (long long)ci * p.s_ci + toff, val);
                        } else {
                            const int ca = c_abs0 + ci;
                            if ((ca >> 3) == (co >> 3)) hn_red_add_f32(p.dw + (long long)co * p.s_co + (long long)(ca & 7) * p.s_ci + toff, val);
                        }
                    }
                }
                __syncwarp();
            }
        }
        hn_tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        hn_tc_fence_after();
        hn_tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled wg_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(f);
    });
    return fn;
}
static int wg_encode_view(CUtensorMap* tm, const hn_view& v, int box_w, int box_h) {
    PFN_encodeTiled enc = wg_encode_fn();
    HN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    hn_ensure_context();
    HN_REQUIRE(v.ptr != nullptr && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0, "wgrad: view base must be 16-byte aligned");
    HN_REQUIRE(v.C % 8 == 0 && v.stride_x % 8 == 0 && v.stride_y % 8 == 0 && v.stride_n % 8 == 0,
               "wgrad: view channels / strides must be multiples of 8 (C=%d)", v.C);
    cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
    cuuint64_t strides[3] = {(cuuint64_t)v.stride_x * 2, (cuuint64_t)v.stride_y * 2, (cuuint64_t)v.stride_n * 2};
    for (int i = 0; i < 3; ++i)
        if (strides[i] == 0) strides[i] = (i == 0 ? (cuuint64_t)v.C * 2 : strides[i - 1] * dims[i]);
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HN_REQUIRE(r == CUDA_SUCCESS, "wgrad: cuTensorMapEncodeTiled failed: %d (C=%d W=%d H=%d N=%d)", (int)r, v.C, v.W, v.H, v.N);
    return HN_OK;
}

extern "C" int hn_conv_wgrad(const hn_wgrad_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d != nullptr && d->dw != nullptr, "wgrad: null pointer");
    HN_REQUIRE(d->n_src >= 1 && d->n_src <= HN_MAX_SRC && d->num_taps >= 1 && d->num_taps <= HN_MAX_TAPS, "wgrad: n_src / num_taps out of range");
    HN_REQUIRE(d->cout >= 1 && d->cout <= d->dy.C, "wgrad: cout=%d exceeds the gradient's %d channels", d->cout, d->dy.C);
    WgradParams p;
    memset(&p, 0, sizeof(p));
    const int TH = d->flat ? 1 : d->tile_h, TW = d->flat ? 128 : d->tile_w;
    HN_REQUIRE(TH * TW == 128 && TW <= 256 && TH <= 256, "wgrad: tile %dx%d must cover 128 pixels", TH, TW);
    if (int rc = wg_encode_view(&p.tmDy, d->dy, TW, TH)) return rc;
    for (int i = 0; i < d->n_src; ++i) {
        if (int rc = wg_encode_view(&p.tmX[i], d->src[i], TW, TH)) return rc;
        if (!d->flat)
            HN_REQUIRE(d->src[i].N == d->dy.N, "wgrad: source %d batch mismatch", i);
    }
    p.flat = d->flat;
    p.TH = TH;
    p.TW = TW;
    if (d->flat) {
        HN_REQUIRE(d->dy.N == 1 && d->dy.H == 1, "wgrad: flat mode needs a [1,1,rows,C] gradient view");
        p.k_tiles = hn_cdiv(d->dy.W, 128);
    } else {
        p.n_img = d->dy.N;
        p.H = d->dy.H;
        p.W = d->dy.W;
        p.tiles_x = hn_cdiv(p.W, TW);
        p.tiles_y = hn_cdiv(p.H, TH);
        p.k_tiles = p.tiles_x * p.tiles_y * p.n_img;
    }
    HN_REQUIRE(p.k_tiles >= 1, "wgrad: empty gradient");
    p.num_taps = d->num_taps;
    for (int t = 0; t < d->num_taps; ++t) {
        HN_REQUIRE(d->taps[t].src >= 0 && d->taps[t].src < d->n_src, "wgrad: tap %d: bad source", t);
        HN_REQUIRE(d->tap_cin[t] >= 1 && d->tap_cin[t] <= 64, "wgrad: tap %d: tap_cin=%d", t, d->tap_cin[t]);
        p.taps[t] = d->taps[t];
        p.tap_off[t] = d->tap_off[t];
        p.tap_cin[t] = d->tap_cin[t];
    }
    p.cout = d->cout;
    p.grouped = d->grouped;
    p.s_co = d->s_co;
    p.s_ci = d->s_ci;
    p.dw = d->dw;
    p.m_tiles = hn_cdiv(d->cout, 128);
    {   // tap groups: consecutive taps, at most kMaxTapsPerItem, cut where (source, shift) changes so that a group can be one operand
        int g = 0, longest = 1;
        for (int t = 0; t < d->num_taps;) {
            int len = 1;
            while (len < kMaxTapsPerItem && t + len < d->num_taps && !d->grouped) {  // run of channel-consecutive blocks of one window
                const hn_tap &a = d->taps[t + len - 1], &b = d->taps[t + len];
                if (a.src != b.src || a.dy != b.dy || a.dx != b.dx || b.c0 != a.c0 + 64) break;
                ++len;
            }
            const bool merge = len > 1;
            if (!merge)  // single blocks: bundle up to four arbitrary taps, they still share the dY tile of every stage
                while (len < kMaxTapsPerItem && t + len < d->num_taps) ++len;
            p.grp_first[g] = (unsigned char)t;
            p.grp_len[g] = (unsigned char)len;
            p.grp_merge[g] = merge ? 1 : 0;
            if (len > longest) longest = len;
            t += len;
            ++g;
        }
        p.tap_groups = g;
        p.nb = longest;
    }
    p.tmem_cols = p.nb * 64 <= 64 ? 64 : (p.nb * 64 <= 128 ? 128 : 256);
    const int stage_bytes = (2 + p.nb) * kBoxBytes;
    int stages = (int)((200 * 1024) / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 2) stages = 2;
    p.stages = stages;
    // Split K: CTAs = pairs x splits run in waves of `sms` (one CTA per SM: ~200 KB of shared memory each); every CTA walks
    // ceil(k_tiles / splits) pixel tiles and then pays a fixed prologue + reduction epilogue (~kFixed tiles' worth).  Pick the split
    // that minimises waves x (tiles per CTA + fixed): it lands on whole waves instead of e.g. 324 CTAs = 2.2 waves.
    int sms = hn_device_sm_count();
    if (sms <= 0) sms = 148;
    const long long pairs = (long long)p.m_tiles * p.tap_groups;
    {
        const int kFixed = 6;
        long long best_cost = -1;
        int best = 1;
        const int max_s = p.k_tiles < 256 ? p.k_tiles : 256;
        for (int sp = 1; sp <= max_s; ++sp) {
            const long long per = (p.k_tiles + sp - 1) / sp;
            const long long ctas = pairs * ((p.k_tiles + per - 1) / per);
            const long long waves = (ctas + sms - 1) / sms;
            const long long cost = waves * (per + kFixed);
            if (best_cost < 0 || cost < best_cost) { best_cost = cost; best = sp; }
        }
        p.k_per_split = hn_cdiv(p.k_tiles, best);
        p.splits = hn_cdiv(p.k_tiles, p.k_per_split);
    }
    const size_t smem = 1024 + (size_t)stages * stage_bytes + (2 * stages + 1) * 8 + 16 + 16 + 4 * 32 * 33 * 4 + 64;
    HN_REQUIRE(smem <= 227 * 1024, "wgrad: %zu bytes of shared memory", smem);
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] { attr_err = cudaFuncSetAttribute(hn_conv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024); });
    HN_CHECK_CUDA(attr_err);
    const long long grid = pairs * p.splits;
    HN_REQUIRE(grid < 0x7fffffffLL, "wgrad: grid too large");
    hn_conv_wgrad_kernel<<<(unsigned)grid, kWgThreads, smem, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}
