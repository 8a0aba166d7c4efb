// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged by
// TMA.  One kernel serves every dense / grouped / strided / up-sampled / concatenated convolution of
// the HydraNet forward: the host describes the K loop as a list of "taps" (source view, spatial
// shift, 64-channel slice); the A tile of a tap is one 4-D TMA box {64 ch, tile_w, tile_h, 1 image}
// of an NHWC bf16 view (out-of-bounds -> zero fill == zero padding), the B tile is a {64, BN} box of
// the pre-packed K-major weight matrix.  Both land 128-byte swizzled, exactly the canonical K-major
// UMMA layout, so four tcgen05.mma (K=16 each) consume a stage.
//
// CTA = 192 threads: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 epilogue (TMEM -> registers -> bias/activation/residual -> global, plus halo mirrors).
// Accumulator: 128 lanes x BN fp32 columns in TMEM.
#include <mutex>

#include "hn_ops.h"

static constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 bf16

__device__ __forceinline__ void store_row_chunk_bf16(bf16* base, const long long* offs, int ndst, int col,
                                                     const uint32_t (&pk)[8]) {
    for (int d = 0; d < ndst; ++d) {
        uint4* dst = reinterpret_cast<uint4*>(base + offs[d] + col);
        dst[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        dst[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
}

__global__ void __launch_bounds__(192) hn_conv_gemm_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int BN = p.bn;
    const int stages = p.stages;
    const int b_tile_bytes = BN * 128;
    uint8_t* sA = smem;
    uint8_t* sB = smem + stages * kATileBytes;
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sB + stages * b_tile_bytes);
    uint64_t* bar_empty = bar_full + stages;
    uint64_t* bar_acc = bar_empty + stages;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc + 1);
    float* s_bias = reinterpret_cast<float*>(tmem_slot + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int n0 = blockIdx.y * BN;

    // tile origin
    int img = 0, y0 = 0, x0 = 0;
    if (p.flat) {
        x0 = blockIdx.x * 128;
    } else {
        int per_img = p.tiles_x * p.tiles_y;
        img = blockIdx.x / per_img;
        int r = blockIdx.x - img * per_img;
        y0 = (r / p.tiles_x) * p.TH;
        x0 = (r % p.tiles_x) * p.TW;
    }

    if (warp == 0 && lane == 0) {
        hn_tma_prefetch_desc(&p.tmB);
        hn_tma_prefetch_desc(&p.tmA[0]);
        for (int s = 0; s < stages; ++s) {
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
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < BN; i += 128) s_bias[i] = p.bias ? p.bias[n0 + i] : 0.0f;
    }
    hn_tc_fence_before();
    __syncthreads();
    hn_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            const uint32_t stage_bytes = (uint32_t)(kATileBytes + b_tile_bytes);
            const int c_shift = p.grouped ? n0 : 0;
            int s = 0;
            uint32_t ph = 0;
            for (int k = 0; k < p.num_taps; ++k) {
                hn_mbar_wait(&bar_empty[s], ph ^ 1);
                hn_mbar_expect_tx(&bar_full[s], stage_bytes);
                const hn_tap t = p.taps[k];
                hn_tma_load_4d(sA + s * kATileBytes, &p.tmA[t.src], &bar_full[s], (int)t.c0 + c_shift, x0 + (int)t.dx,
                               y0 + (int)t.dy, img);
                hn_tma_load_2d(sB + s * b_tile_bytes, &p.tmB, &bar_full[s], k * 64, n0);
                if (++s == stages) { s = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            const uint32_t idesc = hn_umma_idesc_bf16(128, BN);
            int s = 0;
            uint32_t ph = 0;
            for (int k = 0; k < p.num_taps; ++k) {
                hn_mbar_wait(&bar_full[s], ph);
                hn_tc_fence_after();
                const uint32_t a_addr = hn_smem_u32(sA + s * kATileBytes);
                const uint32_t b_addr = hn_smem_u32(sB + s * b_tile_bytes);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    uint64_t da = hn_umma_desc_sw128(a_addr + kk * 32);
                    uint64_t db = hn_umma_desc_sw128(b_addr + kk * 32);
                    hn_umma_bf16(tmem_base, da, db, idesc, (uint32_t)((k | kk) != 0));
                }
                hn_umma_commit(&bar_empty[s]);  // frees the smem slot once these MMAs retire
                if (++s == stages) { s = 0; ph ^= 1; }
            }
            hn_umma_commit(bar_acc);  // accumulator complete
        }
    } else {
        // ------------------------------ epilogue ------------------------------
        const int q = warp & 3;  // TMEM lane quarter this warp may read
        const int row = q * 32 + lane;
        bool valid;
        int n_i, Y = 0, X = 0;
        long long off0, roff;
        if (p.flat) {
            long long m = (long long)x0 + row;
            valid = m < p.flat_m;
            n_i = (int)(m / p.flat_hw);
            long long pix = m - (long long)n_i * p.flat_hw;
            off0 = (long long)n_i * p.osn + pix * p.osx;
            roff = (long long)n_i * p.rsn + pix * p.rsx;
        } else {
            int ty = row / p.TW, tx = row - ty * p.TW;
            int y = y0 + ty, x = x0 + tx;
            valid = (y < p.H) && (x < p.W);
            n_i = img;
            Y = y * p.oscale + p.ooy;
            X = x * p.oscale + p.oox;
            off0 = (long long)n_i * p.osn + (long long)Y * p.osy + (long long)X * p.osx;
            roff = (long long)n_i * p.rsn + (long long)Y * p.rsy + (long long)X * p.rsx;
        }
        // destinations: the pixel itself plus mirrored halo copies
        long long offs[9];
        int ndst = 1;
        offs[0] = off0;
        if (!p.flat && p.halo != HN_HALO_NONE && valid) {
            const int OH = p.H * p.oscale, OW = p.W * p.oscale;
            int ys[3], xs[3], ny = 1, nx = 1;
            ys[0] = Y;
            xs[0] = X;
            if (p.halo == HN_HALO_REFLECT) {
                if (Y == 1) ys[ny++] = -1;
                if (Y == OH - 2) ys[ny++] = OH;
                if (X == 1) xs[nx++] = -1;
                if (X == OW - 2) xs[nx++] = OW;
            } else {
                if (Y == 0) ys[ny++] = -1;
                if (Y == OH - 1) ys[ny++] = OH;
                if (X == 0) xs[nx++] = -1;
                if (X == OW - 1) xs[nx++] = OW;
            }
            ndst = 0;
            for (int a = 0; a < ny; ++a)
                for (int b = 0; b < nx; ++b)
                    offs[ndst++] = (long long)n_i * p.osn + (long long)ys[a] * p.osy + (long long)xs[b] * p.osx;
        }

        hn_mbar_wait(bar_acc, 0);
        hn_tc_fence_after();
        const uint32_t t_row = tmem_base + ((uint32_t)(q * 32) << 16);

        if (p.epi == HN_EPI_SEGOUT) {
            // columns = 4 sub-pixel parities x 8 (n_cls valid): fp32 NCHW logits + fused argmax
            const int OH = p.H * 2, OW = p.W * 2;
            float* outf = reinterpret_cast<float*>(p.out);
            for (int c = 0; c < BN; c += 16) {
                uint32_t v[16];
                hn_tmem_ld16(t_row + c, v);
                hn_tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        int par = (c >> 3) + h;
                        int yy = Y * 2 + (par >> 1), xx = X * 2 + (par & 1);
                        float best = 0.f;
                        int bi = 0;
                        for (int k = 0; k < p.n_cls; ++k) {
                            float f = __uint_as_float(v[h * 8 + k]) + s_bias[c + h * 8 + k];
                            outf[(((long long)n_i * p.n_cls + k) * OH + yy) * OW + xx] = f;
                            if (k == 0 || f > best) { best = f; bi = k; }
                        }
                        if (p.out2) p.out2[((long long)n_i * OH + yy) * OW + xx] = (uint8_t)bi;
                    }
                }
            }
        } else if (p.out_fp32) {
            float* outf = reinterpret_cast<float*>(p.out);
            for (int c = 0; c < BN; c += 16) {
                uint32_t v[16];
                hn_tmem_ld16(t_row + c, v);
                hn_tmem_ld_wait();
                if (valid) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        int n = n0 + c + j;
                        if (n < p.cout) outf[off0 + n] = hn_act(__uint_as_float(v[j]) + s_bias[c + j], p.act);
                    }
                }
            }
        } else {
            bf16* outb = reinterpret_cast<bf16*>(p.out);
            for (int c = 0; c < BN; c += 16) {
                uint32_t v[16];
                hn_tmem_ld16(t_row + c, v);
                hn_tmem_ld_wait();
                const int n = n0 + c;
                if (valid && n < p.cout) {
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) f[j] = hn_act(__uint_as_float(v[j]) + s_bias[c + j], p.act);
                    const bool second = (n + 8) < p.cout;
                    if (p.res) {
                        const uint4* r = reinterpret_cast<const uint4*>(p.res + roff + n);
                        uint4 r0 = r[0];
                        uint4 r1 = second ? r[1] : make_uint4(0, 0, 0, 0);
                        uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            float2 t = hn_unpack_bf16x2(rr[j]);
                            f[2 * j] += t.x;
                            f[2 * j + 1] += t.y;
                        }
                        if (p.res_relu) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
                        }
                    }
                    uint32_t pk[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) pk[j] = hn_pack_bf16x2(f[2 * j], f[2 * j + 1]);
                    if (second) {
                        store_row_chunk_bf16(outb, offs, ndst, n, pk);
                    } else {
                        for (int d = 0; d < ndst; ++d)
                            *reinterpret_cast<uint4*>(outb + offs[d] + n) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
            }
        }
    }

    hn_tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        hn_tc_fence_after();
        hn_tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------
// host side: tensor-map encoding and launch
// ------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(f);
    });
    return fn;
}

static int encode_view_map(CUtensorMap* tm, const hn_view& v, int box_w, int box_h) {
    PFN_encodeTiled enc = get_encode_fn();
    HN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    HN_REQUIRE(v.ptr != nullptr && (reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0, "view base must be 16-byte aligned");
    HN_REQUIRE(v.C % 8 == 0 && v.stride_x % 8 == 0 && v.stride_y % 8 == 0 && v.stride_n % 8 == 0,
               "view channels/strides must be multiples of 8 elements (16 bytes): C=%d sx=%lld sy=%lld sn=%lld", v.C,
               (long long)v.stride_x, (long long)v.stride_y, (long long)v.stride_n);
    cuuint64_t dims[4] = {(cuuint64_t)v.C, (cuuint64_t)v.W, (cuuint64_t)v.H, (cuuint64_t)v.N};
    cuuint64_t strides[3] = {(cuuint64_t)v.stride_x * 2, (cuuint64_t)v.stride_y * 2, (cuuint64_t)v.stride_n * 2};
    // degenerate dims still need a legal (non-zero, 16B-multiple) stride: use the contiguous one
    for (int i = 0; i < 3; ++i)
        if (strides[i] == 0) strides[i] = (i == 0 ? (cuuint64_t)v.C * 2 : strides[i - 1] * dims[i]);
    cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(v.ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(A) failed: %d (C=%d W=%d H=%d N=%d)", (int)r, v.C, v.W, v.H,
               v.N);
    return HN_OK;
}

static int encode_weight_map(CUtensorMap* tm, const void* w, int rows, int kcols, int bn) {
    PFN_encodeTiled enc = get_encode_fn();
    HN_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled entry point unavailable");
    HN_REQUIRE(w != nullptr && (reinterpret_cast<uintptr_t>(w) & 15) == 0, "weights must be 16-byte aligned");
    cuuint64_t dims[2] = {(cuuint64_t)kcols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)kcols * 2};
    cuuint32_t box[2] = {64, (cuuint32_t)bn};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(w), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    HN_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(B) failed: %d (rows=%d k=%d bn=%d)", (int)r, rows, kcols, bn);
    return HN_OK;
}

static int round_pow2_cols(int bn) {
    int c = 32;
    while (c < bn) c <<= 1;
    return c;
}

int hn_conv_prepare(const hn_conv_desc* d, ConvLaunch* L) {
    HN_REQUIRE(d != nullptr, "null conv desc");
    HN_REQUIRE(d->n_src >= 1 && d->n_src <= HN_MAX_SRC, "n_src=%d out of range", d->n_src);
    HN_REQUIRE(d->num_taps >= 1 && d->num_taps <= HN_MAX_TAPS, "num_taps=%d out of range", d->num_taps);
    HN_REQUIRE(d->bn >= 16 && d->bn <= 256 && d->bn % 16 == 0, "bn=%d must be a multiple of 16 in [16,256]", d->bn);
    HN_REQUIRE(d->stages >= 2 && d->stages <= 8, "stages=%d out of range", d->stages);
    HN_REQUIRE(d->out != nullptr && d->cout >= 1, "missing output");
    ConvParams& p = L->prm;
    memset(&p, 0, sizeof(p));
    int TH = d->flat ? 1 : d->tile_h, TW = d->flat ? 128 : d->tile_w;
    HN_REQUIRE(TH * TW == 128 && TW <= 256 && TH <= 256, "tile %dx%d must cover 128 pixels", TH, TW);
    for (int i = 0; i < d->n_src; ++i) {
        int rc = encode_view_map(&p.tmA[i], d->src[i], TW, TH);
        if (rc) return rc;
    }
    int rc = encode_weight_map(&p.tmB, d->weight, d->w_rows, d->num_taps * 64, d->bn);
    if (rc) return rc;
    for (int k = 0; k < d->num_taps; ++k) {
        HN_REQUIRE(d->taps[k].src >= 0 && d->taps[k].src < d->n_src, "tap %d: bad source %d", k, d->taps[k].src);
        p.taps[k] = d->taps[k];
    }
    p.flat = d->flat;
    p.TH = TH;
    p.TW = TW;
    p.num_taps = d->num_taps;
    p.cout = d->cout;
    p.bn = d->bn;
    p.stages = d->stages;
    p.tmem_cols = round_pow2_cols(d->bn);
    p.bias = d->bias;
    p.act = d->act;
    p.epi = d->epi;
    p.out = d->out;
    p.out_fp32 = d->out_fp32;
    p.osn = d->out_stride_n;
    p.osy = d->out_stride_y;
    p.osx = d->out_stride_x;
    p.oscale = d->out_scale > 0 ? d->out_scale : 1;
    p.ooy = d->out_oy;
    p.oox = d->out_ox;
    p.halo = d->halo;
    p.res = reinterpret_cast<const bf16*>(d->res);
    p.rsn = d->res_stride_n;
    p.rsy = d->res_stride_y;
    p.rsx = d->res_stride_x;
    p.res_relu = d->res_relu;
    p.grouped = d->grouped;
    HN_REQUIRE(!d->grouped || d->bn == 64, "grouped conv needs bn == 64");
    p.out2 = reinterpret_cast<uint8_t*>(d->out2);
    p.n_cls = d->n_cls;
    int m_tiles;
    if (d->flat) {
        HN_REQUIRE(d->flat_hw > 0, "flat conv needs flat_hw");
        p.flat_m = d->src[0].W;
        p.flat_hw = d->flat_hw;
        m_tiles = hn_cdiv(p.flat_m, 128);
    } else {
        HN_REQUIRE(d->n_img > 0 && d->out_h > 0 && d->out_w > 0, "spatial conv needs n_img/out_h/out_w");
        p.n_img = d->n_img;
        p.H = d->out_h;
        p.W = d->out_w;
        p.tiles_x = hn_cdiv(p.W, TW);
        p.tiles_y = hn_cdiv(p.H, TH);
        m_tiles = p.tiles_x * p.tiles_y * p.n_img;
    }
    if (d->epi == HN_EPI_SEGOUT) {
        HN_REQUIRE(!d->flat && d->bn == 32 && d->n_cls >= 1 && d->n_cls <= 8, "segout epilogue needs bn=32, n_cls<=8");
    } else if (!d->out_fp32) {
        HN_REQUIRE(d->cout % 8 == 0 && d->out_stride_x % 8 == 0 && d->out_stride_y % 8 == 0 && d->out_stride_n % 8 == 0 &&
                       (reinterpret_cast<uintptr_t>(d->out) & 15) == 0,
                   "bf16 output needs cout and strides in multiples of 8 and a 16-byte aligned base");
    }
    int n_tiles = hn_cdiv(d->cout, d->bn);
    if (d->epi == HN_EPI_SEGOUT) n_tiles = 1;
    L->grid = dim3((unsigned)m_tiles, (unsigned)n_tiles, 1);
    L->smem = 1024 + (size_t)d->stages * (kATileBytes + d->bn * 128) + (2 * d->stages + 1) * 8 + 16 + d->bn * 4 + 64;
    HN_REQUIRE(L->smem <= 227 * 1024, "conv needs %zu bytes of shared memory (> 227 KB): lower stages/bn", L->smem);
    return HN_OK;
}

int hn_conv_launch(const ConvLaunch* L, cudaStream_t stream) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(hn_conv_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    HN_CHECK_CUDA(attr_err);
    hn_conv_gemm_kernel<<<L->grid, 192, L->smem, stream>>>(L->prm);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

extern "C" int hn_conv_fwd(const hn_conv_desc* d, void* stream) {
    ConvLaunch L;
    int rc = hn_conv_prepare(d, &L);
    if (rc) return rc;
    return hn_conv_launch(&L, reinterpret_cast<cudaStream_t>(stream));
}
