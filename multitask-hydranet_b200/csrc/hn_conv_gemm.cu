// Implicit-GEMM convolution on the 5th-generation tensor cores (tcgen05 + TMEM), operands staged by
// TMA.  One kernel serves every dense / grouped / strided / up-sampled / concatenated convolution of
// the HydraNet forward: the host describes the K loop as a list of "taps" (source view, spatial
// shift, 64-channel slice); the A tile of a tap is one 4-D TMA box {64 ch, tile_w, tile_h, 1 image}
// of an NHWC bf16 view (out-of-bounds -> zero fill == zero padding), the B tile is a {64, BN} box of
// the pre-packed K-major weight matrix.  Both land 128-byte swizzled, exactly the canonical K-major
// UMMA layout, so four tcgen05.mma (K=16 each) consume a stage.
//
// CTA = 320 threads: warp 0 TMA producer, warp 1 TMEM allocator + MMA issuer (one elected lane),
// warps 2..9 epilogue (TMEM -> registers -> bias/activation/residual -> global, plus halo mirrors): two warps
// per TMEM lane quarter, each taking every other 32-column chunk -- the epilogue is latency-bound per warp, and
// for narrow / short-K tiles it, not the MMA, sets the tile rate.
// Accumulator: 128 lanes x BN fp32 columns in TMEM.
#include <algorithm>
#include <mutex>

#include "hn_ops.h"

static constexpr int kATileBytes = 128 * 128;  // 128 rows x 64 bf16
static constexpr int kEpiThreads = 256;
static constexpr int kCtaThreads = 64 + kEpiThreads;

__device__ __forceinline__ long long hn_globaltimer() {
    long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void hn_named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
template <int NC>
__device__ __forceinline__ void hn_tmem_ldN(uint32_t taddr, uint32_t (&v)[NC]) {
    if constexpr (NC == 32) hn_tmem_ld32(taddr, v);
    else hn_tmem_ld16(taddr, v);
}

// n / d for a launch-constant divisor: magic = floor(2^64 / d) + 1 (0 stands for d == 1); exact for every 32-bit n
// (the error term n * (magic*d - 2^64) stays below 2^64).  The per-tile index math runs in every epilogue thread of
// every tile; hardware integer division costs ~20 (32-bit) to ~100 (64-bit) instructions.
__device__ __forceinline__ unsigned hn_fastdiv(unsigned n, unsigned long long magic) {
    return magic ? (unsigned)__umul64hi((unsigned long long)n, magic) : n;
}

// tile index -> origin
struct TileOrigin {
    int n0, img, y0, x0;
};
// work item t = (n tile, group of `cluster` consecutive m tiles); CTA `rank` of the cluster takes m tile group*cluster+rank
// (an m tile index past the end yields out-of-bounds coordinates: zero-filled loads, clipped / masked stores)
__device__ __forceinline__ TileOrigin tile_origin(const ConvParams& p, int t, int rank) {
    TileOrigin o;
    const int nt = p.n_tiles == 1 ? 0 : (int)hn_fastdiv((unsigned)t, p.div_m_groups), mt = (t - nt * p.m_groups) * p.cluster + rank;
    o.n0 = nt * p.bn;
    if (p.flat) {
        o.img = 0; o.y0 = 0; o.x0 = mt * 128;
    } else {
        const int per_img = p.tiles_x * p.tiles_y;
        o.img = (int)hn_fastdiv((unsigned)mt, p.div_per_img);
        const int r = mt - o.img * per_img;
        const int ry = (int)hn_fastdiv((unsigned)r, p.div_tiles_x);
        o.y0 = ry * p.TH;
        o.x0 = (r - ry * p.tiles_x) * p.TW;
    }
    return o;
}

template <int NC>
__device__ __forceinline__ void apply_act(float (&f)[NC], int act) {
    switch (act) {
        case HN_ACT_RELU:
#pragma unroll
            for (int j = 0; j < NC; ++j) f[j] = fmaxf(f[j], 0.0f);
            break;
        case HN_ACT_SWISH:
#pragma unroll
            for (int j = 0; j < NC; ++j) f[j] = f[j] * hn_sigmoid(f[j]);
            break;
        case HN_ACT_ELU:
#pragma unroll
            for (int j = 0; j < NC; ++j) f[j] = f[j] > 0.0f ? f[j] : hn_ex2(f[j] * 1.4426950408889634f) - 1.0f;
            break;
        case HN_ACT_SIGMOID:
#pragma unroll
            // ex2.approx + fast reciprocal: ~2 ulp of the fp32 sigmoid (the head scores; far inside the 1e-2 parity budget)
            for (int j = 0; j < NC; ++j) f[j] = __fdividef(1.0f, 1.0f + __expf(-f[j]));
            break;
        default: break;
    }
}

// per-row epilogue state
struct EpiRow {
    bool valid;
    int n_i, Y, X;
    long long off0, roff;
    int ym1, ym2, xm1, xm2;  // mirrored halo coordinates (INT_MIN when absent)
};
static constexpr int kNoCoord = -2147483647;

// `stage_row`: this thread's 128-byte row of the shared-memory staging tile (nullptr = direct global stores);
// 16-byte chunk j of the row lives at chunk (j ^ (row & 7)) -- the 128B swizzle the output tensor map expects.
template <int NC>
__device__ __forceinline__ void epi_chunk_std(const ConvParams& p, const EpiRow& e, const float* s_bias, const float* s_scale,
                                              int c, int n0, uint32_t (&v)[NC], uint8_t* stage_row, int row, float* tr = nullptr) {
    const int n = n0 + c;
    if (n >= p.cout) return;                               // warp-uniform
    if (!e.valid && !(p.out_fp32 && tr)) return;           // the transposing fp32 path needs the whole warp
    float f[NC];
    if (s_scale) {
#pragma unroll
        for (int j = 0; j < NC; ++j) f[j] = fmaf(__uint_as_float(v[j]), s_scale[c + j], s_bias[c + j]);
    } else {
#pragma unroll
        for (int j = 0; j < NC; ++j) f[j] = __uint_as_float(v[j]) + s_bias[c + j];
    }
    apply_act<NC>(f, p.act);
    if (p.out_fp32) {
        if (tr) {
            // fp32 head outputs ([.., HW*9, 4 or 9]: a row is 36 or 81 floats): a thread owns a ROW, so direct stores hit
            // 32 different sectors per instruction.  Transpose the 32x32 block through the warp's scratch so that a store
            // instruction writes 32 consecutive floats of one row.
            const int lane = threadIdx.x & 31;
#pragma unroll
            for (int j = 0; j < NC; ++j) tr[lane * 33 + j] = f[j];
            __syncwarp();
            const int ncol = min(NC, p.cout - n);
            float* outf = reinterpret_cast<float*>(p.out);
#pragma unroll 4
            for (int rr = 0; rr < 32; ++rr) {
                const long long off = __shfl_sync(0xffffffffu, e.off0, rr);
                const int ok = __shfl_sync(0xffffffffu, (int)e.valid, rr);
                if (ok && lane < ncol) outf[off + n + lane] = tr[rr * 33 + lane];
            }
            __syncwarp();
            return;
        }
        float* outf = reinterpret_cast<float*>(p.out) + e.off0;
#pragma unroll
        for (int j = 0; j < NC; ++j)
            if (n + j < p.cout) outf[n + j] = f[j];
        return;
    }
    const int nvec = min(NC / 8, (p.cout - n) / 8);  // 8-channel vectors to store (cout % 8 == 0)
    if (p.res) {
        const uint4* r = reinterpret_cast<const uint4*>(p.res + e.roff + n);
#pragma unroll
        for (int q = 0; q < NC / 8; ++q) {
            if (q < nvec) {
                uint4 rv = r[q];
                float2 t0 = hn_unpack_bf16x2(rv.x), t1 = hn_unpack_bf16x2(rv.y), t2 = hn_unpack_bf16x2(rv.z), t3 = hn_unpack_bf16x2(rv.w);
                f[q * 8 + 0] += t0.x; f[q * 8 + 1] += t0.y; f[q * 8 + 2] += t1.x; f[q * 8 + 3] += t1.y;
                f[q * 8 + 4] += t2.x; f[q * 8 + 5] += t2.y; f[q * 8 + 6] += t3.x; f[q * 8 + 7] += t3.y;
            }
        }
        if (p.res_relu) {
#pragma unroll
            for (int j = 0; j < NC; ++j) f[j] = fmaxf(f[j], 0.0f);
        }
    }
    uint4 pk[NC / 8];
#pragma unroll
    for (int q = 0; q < NC / 8; ++q)
        pk[q] = make_uint4(hn_pack_bf16x2(f[q * 8 + 0], f[q * 8 + 1]), hn_pack_bf16x2(f[q * 8 + 2], f[q * 8 + 3]),
                           hn_pack_bf16x2(f[q * 8 + 4], f[q * 8 + 5]), hn_pack_bf16x2(f[q * 8 + 6], f[q * 8 + 7]));
    bf16* outb = reinterpret_cast<bf16*>(p.out);
    if (stage_row) {
        const int j0 = (c & 63) >> 3;
#pragma unroll
        for (int q = 0; q < NC / 8; ++q)
            *reinterpret_cast<uint4*>(stage_row + (((j0 + q) ^ (row & 7)) << 4)) = pk[q];
    } else {
        uint4* dst = reinterpret_cast<uint4*>(outb + e.off0 + n);
#pragma unroll
        for (int q = 0; q < NC / 8; ++q)
            if (q < nvec) dst[q] = pk[q];
    }
    if (p.halo != HN_HALO_NONE && (e.ym1 != kNoCoord || e.ym2 != kNoCoord || e.xm1 != kNoCoord || e.xm2 != kNoCoord)) {  // border rows only
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            const int yy = a == 0 ? e.Y : (a == 1 ? e.ym1 : e.ym2);
            if (yy == kNoCoord) continue;
#pragma unroll
            for (int b = 0; b < 3; ++b) {
                const int xx = b == 0 ? e.X : (b == 1 ? e.xm1 : e.xm2);
                if (xx == kNoCoord || (a == 0 && b == 0)) continue;
                uint4* dst = reinterpret_cast<uint4*>(outb + (long long)e.n_i * p.osn + (long long)yy * p.osy + (long long)xx * p.osx + n);
#pragma unroll
                for (int q = 0; q < NC / 8; ++q)
                    if (q < nvec) dst[q] = pk[q];
            }
        }
    }
}

// columns = 4 sub-pixel parities x 8 (n_cls valid): fp32 NCHW logits + fused arg-max.  The two x-parities of
// a class are adjacent in memory: one float2 (uchar2 for the class map) per (class, y-parity).
__device__ __forceinline__ void epi_segout(const ConvParams& p, const EpiRow& e, const float* s_bias, uint32_t (&v)[16], int py) {
    if (!e.valid) return;
    const int OH = p.H * 2, OW = p.W * 2;
    float* outf = reinterpret_cast<float*>(p.out);
    {
        const int yy = e.Y * 2 + py, xx = e.X * 2;
        float b0 = 0.f, b1 = 0.f;
        int i0 = 0, i1 = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            if (k < p.n_cls) {
                float f0 = __uint_as_float(v[k]) + s_bias[(py * 2 + 0) * 8 + k];
                float f1 = __uint_as_float(v[8 + k]) + s_bias[(py * 2 + 1) * 8 + k];
                *reinterpret_cast<float2*>(outf + (((long long)e.n_i * p.n_cls + k) * OH + yy) * OW + xx) = make_float2(f0, f1);
                if (k == 0 || f0 > b0) { b0 = f0; i0 = k; }
                if (k == 0 || f1 > b1) { b1 = f1; i1 = k; }
            }
        }
        if (p.out2) *reinterpret_cast<uchar2*>(p.out2 + ((long long)e.n_i * OH + yy) * OW + xx) = make_uchar2((uint8_t)i0, (uint8_t)i1);
    }
}

// Persistent kernel: every CTA walks tiles t = blockIdx.x, blockIdx.x + gridDim.x, ... (m fastest, so
// concurrently running CTAs share the weight tile in L2).  Shared-memory ring (full/empty mbarriers)
// between the TMA producer and the MMA issuer; two TMEM accumulators (acc_full/acc_empty) so the
// epilogue of tile i overlaps the main loop of tile i+1.
// kPair = true: the cta_group::2 build (must be launched as 2-CTA clusters); false: no pair instructions at all
// kMinBlocks: 2 caps the registers (96) so that two CTAs share an SM (narrow tiles); 1 lets the epilogue keep its
// values in registers (168) when the CTA owns the SM anyway
template <bool kPair, int kMinBlocks = kPair ? 1 : 2>
__global__ void __launch_bounds__(kCtaThreads, kMinBlocks) hn_conv_gemm_kernel(const __grid_constant__ ConvParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int BN = p.bn;
    const int stages = p.stages;
    const int b_tile_bytes = (kPair ? BN / 2 : BN) * 128;  // a CTA of a pair holds half of the weight rows
    // a stage holds one tap RUN: the A box of (TH + run_max - 1) tile rows -- consecutive dy taps of one (source,
    // dx, channel slice) are row-shifted windows of it -- and run_max weight tiles
    const int a_stage = p.a_stage_bytes, b_stage = p.run_max * b_tile_bytes;
    const int a_row_bytes = p.TW * 128;  // one tile row of the A box (a multiple of the 1024-byte swizzle atom)
    uint8_t* sA = smem;
    uint8_t* sB = smem + stages * a_stage;
    uint8_t* sO = sB + stages * b_stage;  // n_staging x 16 KB output staging (1024-byte aligned)
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(sO + p.n_staging * kATileBytes);
    uint64_t* bar_empty = bar_full + stages;
    uint64_t* bar_acc_full = bar_empty + stages;   // [2]
    uint64_t* bar_acc_empty = bar_acc_full + 2;    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_acc_empty + 2);
    // per accumulator: [BN] bias, or with row groups [n_groups][BN] shift followed by [n_groups][BN] scale
    float* s_bias = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(tmem_slot + 2) + 15) & ~uintptr_t(15));  // float4 reads
    const int bias_slots = p.n_groups > 0 ? 2 * p.n_groups : 1;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    hn_pdl_launch_dependents();
    // persistent loop over work items, one item per CTA (or per CTA pair) per iteration
    constexpr int CS = kPair ? 2 : 1;  // 2: CTA pair driving one 256-row cta_group::2 MMA
    constexpr bool PAIR = kPair;
    int rank = 0;
    if constexpr (PAIR) rank = (int)hn_cluster_ctarank();
    const bool leader = rank == 0;
    const int total_tiles = p.m_groups * p.n_tiles;
    const int t_first = blockIdx.x / CS, t_step = gridDim.x / CS;
    long long* dbg = p.dbg ? p.dbg + (long long)blockIdx.x * 16 : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = hn_globaltimer();

    if (warp == 0 && lane == 0) {
        hn_tma_prefetch_desc(&p.tmB);
        hn_tma_prefetch_desc(&p.tmA[0]);
        if (p.n_staging) hn_tma_prefetch_desc(&p.tmO);
        for (int s = 0; s < stages; ++s) {
            hn_mbar_init(&bar_full[s], 1);
            hn_mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            hn_mbar_init(&bar_acc_full[a], 1);
            hn_mbar_init(&bar_acc_empty[a], CS * (kEpiThreads / 32));  // pair: the peer's epilogue warps arrive here too
        }
        hn_mbar_fence_init();
    }
    if (warp == 1) {
        if constexpr (PAIR) { hn_tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols); hn_tmem_relinquish_pair(); }
        else { hn_tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); hn_tmem_relinquish(); }
    }
    hn_tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) hn_cluster_sync();  // the peer's barriers are initialised before anything remote touches them
    hn_tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (dbg && threadIdx.x == 0) dbg[1] = hn_globaltimer();
    // prologue done (it overlapped the previous kernel's tail): from here on we read what that kernel wrote
    hn_pdl_wait();

    if (warp == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int t = t_first; t < total_tiles; t += t_step) {
                const TileOrigin o = tile_origin(p, t, rank);
                const int c_shift = p.grouped ? o.n0 : 0;
                for (int k = 0; k < p.num_taps;) {
                    hn_mbar_wait(&bar_empty[s], ph ^ 1);
                    const hn_tap tp = p.taps[k];
                    const int run = tp.rsv0;  // taps k .. k+run-1 share this A box (run == 1: the plain 128-row tile)
                    const CUtensorMap* tma = run > 1 ? &p.tmArun[tp.src] : &p.tmA[tp.src];
                    const uint32_t stage_bytes = (uint32_t)((run > 1 ? a_stage : kATileBytes) + run * b_tile_bytes);
                    if constexpr (!PAIR) {
                        hn_mbar_expect_tx(&bar_full[s], stage_bytes);
                        hn_tma_load_4d(sA + s * a_stage, tma, &bar_full[s], (int)tp.c0 + c_shift, o.x0 + (int)tp.dx,
                                       o.y0 + (int)tp.dy, o.img);
                        for (int j = 0; j < run; ++j)
                            hn_tma_load_2d(sB + s * b_stage + j * b_tile_bytes, &p.tmB, &bar_full[s], (int)p.taps[k + j].rsv1 * 64, o.n0);
                    } else {
                        // each CTA loads its own 128 A rows and its half of the weight rows; all bytes are counted on the
                        // leader's barrier, which is the one the (single) MMA issuer waits on
                        if (leader) hn_mbar_expect_tx(&bar_full[s], 2u * stage_bytes);
                        const uint32_t lbar = hn_mapa(hn_smem_u32(&bar_full[s]), 0);
                        hn_tma_load_4d_pair(sA + s * a_stage, tma, lbar, (int)tp.c0 + c_shift, o.x0 + (int)tp.dx,
                                            o.y0 + (int)tp.dy, o.img);
                        for (int j = 0; j < run; ++j)
                            hn_tma_load_2d_pair(sB + s * b_stage + j * b_tile_bytes, &p.tmBpart, lbar, (int)p.taps[k + j].rsv1 * 64,
                                                o.n0 + rank * (BN / 2));
                    }
                    k += run;
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                if (dbg && t == t_first) dbg[2] = hn_globaltimer();
            }
        }
    } else if (warp == 1) {
        // ------------------------------ MMA issuer (pair: leader CTA only) ------------------------------
        if (lane == 0 && leader) {
            const uint32_t idesc = hn_umma_idesc_bf16(PAIR ? 256 : 128, BN);
            int s = 0, a = 0;
            uint32_t ph = 0, aph = 0;
            for (int t = t_first; t < total_tiles; t += t_step) {
                hn_mbar_wait(&bar_acc_empty[a], aph ^ 1);  // epilogue has drained this accumulator
                hn_tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(a * p.acc_stride);
                for (int k = 0; k < p.num_taps;) {
                    const int run = p.taps[k].rsv0;
                    hn_mbar_wait(&bar_full[s], ph);
                    hn_tc_fence_after();
                    if (dbg && t == t_first && k == 0) dbg[3] = hn_globaltimer();
                    for (int j = 0; j < run; ++j) {
                        // tap k+j reads the A box from tile row j on
                        const uint32_t a_addr = hn_smem_u32(sA + s * a_stage + j * a_row_bytes);
                        const uint32_t b_addr = hn_smem_u32(sB + s * b_stage + j * b_tile_bytes);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk) {
                            uint64_t da = hn_umma_desc_sw128(a_addr + kk * 32);
                            uint64_t db = hn_umma_desc_sw128(b_addr + kk * 32);
                            if constexpr (PAIR) hn_umma_bf16_pair(d_tmem, da, db, idesc, (uint32_t)((k | j | kk) != 0));
                            else hn_umma_bf16(d_tmem, da, db, idesc, (uint32_t)((k | j | kk) != 0));
                        }
                    }
                    // frees the smem slot once these MMAs retire (pair: in both CTAs)
                    if constexpr (PAIR) hn_umma_commit_pair(&bar_empty[s]); else hn_umma_commit(&bar_empty[s]);
                    k += run;
                    if (++s == stages) { s = 0; ph ^= 1; }
                }
                // accumulator complete (pair: each CTA's epilogue drains its own 128 rows)
                if constexpr (PAIR) hn_umma_commit_pair(&bar_acc_full[a]); else hn_umma_commit(&bar_acc_full[a]);
                if (dbg && t == t_first) dbg[4] = hn_globaltimer();
                if (++a == 2) { a = 0; aph ^= 1; }
            }
        }
    } else {
        // ------------------------------ epilogue (warps 2..9) ------------------------------
        const int q = warp & 3;          // TMEM lane quarter this warp may read
        const int half = (warp - 2) >> 2;  // which of the quarter's two warps: it takes the odd or the even 32-column chunks
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;
        int a = 0;
        uint32_t aph = 0;
        int ntile = 0, st_count = 0;
        // bias (or per-row-group shift / scale) of output channels n0 .. n0+BN into one accumulator's slot
        auto stage_bias = [&](float* dst, int n0) {
            if (p.n_groups > 0 && p.group_shift) {
                for (int i = et; i < p.n_groups * BN; i += kEpiThreads) {
                    const int gi = i / BN, ci = i - gi * BN;
                    dst[i] = p.group_shift[(long long)gi * p.gstride + n0 + ci];
                    dst[p.n_groups * BN + i] = p.group_scale ? p.group_scale[(long long)gi * p.gstride + n0 + ci] : 1.0f;
                }
            } else {
                for (int i = et; i < BN; i += kEpiThreads) dst[i] = (p.bias && n0 + i < p.cout) ? p.bias[n0 + i] : 0.0f;  // callers may pass an unpadded [cout] vector
            }
        };
        // a single N tile: every work item has the same channels -- stage once (a global-load round trip and a barrier
        // less per tile, which is what bounds short-K tiles)
        const bool bias_once = p.n_tiles == 1;
        if (bias_once) {
            stage_bias(s_bias, 0);
            stage_bias(s_bias + bias_slots * BN, 0);
            hn_named_bar_sync(1, kEpiThreads);
        }
        for (int t = t_first; t < total_tiles; t += t_step, ++ntile) {
            const TileOrigin o = tile_origin(p, t, rank);
            float* bias_s = s_bias + a * bias_slots * BN;
            const float* scale_s = nullptr;
            if (!bias_once) stage_bias(bias_s, o.n0);
            EpiRow e;
            e.ym1 = e.ym2 = e.xm1 = e.xm2 = kNoCoord;
            e.Y = e.X = 0;
            if (p.flat) {
                long long m = (long long)o.x0 + row;
                e.valid = m < p.flat_m;
                int g = 0;
                if (p.n_groups > 0) {
                    while (g < p.n_groups - 1 && m >= p.group_end[g]) ++g;
                    if (p.group_shift) {  // this row's shift / scale vectors
                        bias_s += g * BN;
                        scale_s = bias_s + p.n_groups * BN;
                    }
                }
                if (p.n_groups > 0 && p.group_addr) {
                    const long long ml = m - (g > 0 ? p.group_end[g - 1] : 0);
                    e.n_i = (int)(ml / p.group_hw[g]);
                    const long long pix = ml - (long long)e.n_i * p.group_hw[g];
                    e.off0 = (long long)e.n_i * p.osn + p.group_out_base[g] + pix * p.osx;
                    e.roff = 0;
                } else {
                    e.off0 = e.roff = 0;
                    e.n_i = 0;
                    if (!p.n_staging || p.res) {
                        e.n_i = (int)hn_fastdiv((unsigned)m, p.div_flat_hw);  // m < flat_m < 2^31
                        long long pix = m - (long long)e.n_i * p.flat_hw;
                        e.off0 = (long long)e.n_i * p.osn + pix * p.osx;
                        e.roff = (long long)e.n_i * p.rsn + pix * p.rsx;
                    }
                }
            } else {
                int ty = row >> p.tw_shift, tx = row - (ty << p.tw_shift);  // TW is a power of two (TH*TW == 128)
                int y = o.y0 + ty, x = o.x0 + tx;
                e.valid = (y < p.H) && (x < p.W) && (o.img < p.n_img);
                e.n_i = o.img;
                e.Y = y * p.oscale + p.ooy;
                e.X = x * p.oscale + p.oox;
                // element offsets only where something addresses global memory directly (the staged path stores
                // through the tensor map): this setup runs in every epilogue thread for every tile
                e.off0 = e.roff = 0;
                if (!p.n_staging) e.off0 = (long long)e.n_i * p.osn + (long long)e.Y * p.osy + (long long)e.X * p.osx;
                if (p.res) e.roff = (long long)e.n_i * p.rsn + (long long)e.Y * p.rsy + (long long)e.X * p.rsx;
                if (p.halo != HN_HALO_NONE && p.epi == HN_EPI_STD) {
                    const int OH = p.H * p.oscale, OW = p.W * p.oscale;
                    if (p.halo == HN_HALO_REFLECT) {
                        if (e.Y == 1) e.ym1 = -1;
                        if (e.Y == OH - 2) e.ym2 = OH;
                        if (e.X == 1) e.xm1 = -1;
                        if (e.X == OW - 2) e.xm2 = OW;
                    } else {
                        if (e.Y == 0) e.ym1 = -1;
                        if (e.Y == OH - 1) e.ym2 = OH;
                        if (e.X == 0) e.xm1 = -1;
                        if (e.X == OW - 1) e.xm2 = OW;
                    }
                }
            }
            if (!bias_once) hn_named_bar_sync(1, kEpiThreads);  // bias staged
            hn_mbar_wait(&bar_acc_full[a], aph);
            hn_tc_fence_after();
            if (dbg && ntile == 0 && et == 0) dbg[5] = hn_globaltimer();
            const uint32_t t_row = tmem_base + (uint32_t)(a * p.acc_stride) + ((uint32_t)(q * 32) << 16);
            if (p.epi == HN_EPI_SEGOUT) {
                uint32_t v[16];  // this warp's output-row parity: columns [half][x parity][8]
                hn_tmem_ld16(t_row + half * 16, v);
                hn_tmem_ld_wait();
                epi_segout(p, e, bias_s, v, half);
            } else if (p.n_staging) {
                // 64-channel slabs: registers -> swizzled shared-memory tile -> TMA store.  Each TMEM lane quarter (its
                // two warps, 32 tile rows = a 4 KB sub-slab) stores its own box and synchronises on its own 64-thread
                // named barriers: no quarter waits for another.
                const bool qlead = half == 0 && lane == 0;
                const int q_dx = (q * 32) & (p.TW - 1), q_dy = (q * 32) >> p.tw_shift;  // the quarter's first tile row
                for (int c = 0; c < BN; c += 64) {
                    uint8_t* slab = sO + (st_count & (p.n_staging - 1)) * kATileBytes;  // n_staging is 1 or 2
                    if (qlead) {
                        if (p.n_staging == 2) hn_tma_store_wait_read<1>(); else hn_tma_store_wait_read<0>();
                    }
                    hn_named_bar_sync(2 + q, 64);  // the quarter's sub-slab is free again
                    uint8_t* stage_row = slab + row * 128;
                    {
                        const int cc = c + half * 32;
                        if (cc + 32 <= BN) {
                            uint32_t v[32];
                            hn_tmem_ld32(t_row + cc, v);
                            hn_tmem_ld_wait();
                            if (dbg && ntile == 0 && et == 0 && c == 0) dbg[9] = hn_globaltimer();
                            epi_chunk_std<32>(p, e, bias_s, scale_s, cc, o.n0, v, stage_row, row);
                            if (dbg && ntile == 0 && et == 0 && c == 0) dbg[10] = hn_globaltimer();
                        } else if (cc + 16 <= BN) {
                            uint32_t v[16];
                            hn_tmem_ld16(t_row + cc, v);
                            hn_tmem_ld_wait();
                            epi_chunk_std<16>(p, e, bias_s, scale_s, cc, o.n0, v, stage_row, row);
                        }
                    }
                    hn_fence_proxy_async();
                    hn_named_bar_sync(6 + q, 64);  // sub-slab fully written
                    if (dbg && ntile == 0 && et == 0 && c == 0) dbg[11] = hn_globaltimer();
                    if (qlead) {
                        hn_tma_store_4d(&p.tmO, slab + q * 4096, o.n0 + c, o.x0 + q_dx, o.y0 + q_dy, o.img);
                        hn_tma_store_commit();
                    }
                    if (dbg && ntile == 0 && et == 0 && c == 0) dbg[12] = hn_globaltimer();
                    if (dbg && ntile == 0 && et == 0 && c == 64) dbg[13] = hn_globaltimer();
                    ++st_count;
                }
            } else {
                float* tr = p.tr_off ? reinterpret_cast<float*>(smem + p.tr_off) + (warp - 2) * (32 * 33) : nullptr;
                int c = half * 32;
                for (; c + 32 <= BN; c += 64) {
                    uint32_t v[32];
                    hn_tmem_ld32(t_row + c, v);
                    hn_tmem_ld_wait();
                    epi_chunk_std<32>(p, e, bias_s, scale_s, c, o.n0, v, nullptr, row, tr);
                }
                if (c < BN) {  // a 16-column tail chunk (BN % 32 == 16)
                    uint32_t v[16];
                    hn_tmem_ld16(t_row + c, v);
                    hn_tmem_ld_wait();
                    epi_chunk_std<16>(p, e, bias_s, scale_s, c, o.n0, v, nullptr, row, tr);
                }
            }
            // all TMEM reads of this accumulator are complete: hand it back to the MMA issuer
            hn_tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (PAIR) hn_mbar_arrive_remote(hn_mapa(hn_smem_u32(&bar_acc_empty[a]), 0));  // the issuer lives in the leader CTA
                else hn_mbar_arrive(&bar_acc_empty[a]);
            }
            if (dbg && ntile == 0 && et == 0) dbg[6] = hn_globaltimer();
            if (++a == 2) { a = 0; aph ^= 1; }
        }
        if (p.n_staging && half == 0 && lane == 0) hn_tma_store_wait_all();  // shared memory must outlive the bulk stores
        if (dbg && et == 0) dbg[8] = ntile;
    }

    hn_tc_fence_before();
    __syncthreads();
    if constexpr (PAIR) hn_cluster_sync();  // neither CTA leaves while the pair's MMAs / remote arrivals may still touch it
    if (warp == 1) {
        hn_tc_fence_after();
        if constexpr (PAIR) hn_tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
        else hn_tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
    }
    if (dbg && threadIdx.x == 0) dbg[7] = hn_globaltimer();
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
    hn_ensure_context();
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
    hn_ensure_context();
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

static void* g_conv_dbg = nullptr;
extern "C" void hn_conv_set_debug_buffer(void* p) { g_conv_dbg = p; }
static int g_conv_cluster = 0;  // 0 = default policy
extern "C" void hn_conv_set_cluster(int cs) { g_conv_cluster = cs; }
static int g_conv_pair_min_bn = 192;  // narrowest N tile that runs on CTA pairs
extern "C" void hn_conv_set_pair_min_bn(int bn) { g_conv_pair_min_bn = bn; }
static int g_conv_runs = 1;  // 0: never share A boxes between taps; 1: default policy (narrow N tiles); 2: wherever they fit
extern "C" void hn_conv_set_tap_runs(int mode) { g_conv_runs = mode; }
static constexpr int kMaxRun = 3;

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
        p.taps[k].rsv0 = 1;            // run length (set on the first tap of a run)
        p.taps[k].rsv1 = (int16_t)k;  // K block of the weight matrix (the caller's tap index)
    }
    p.flat = d->flat;
    p.TH = TH;
    p.TW = TW;
    p.num_taps = d->num_taps;
    p.cout = d->cout;
    p.bn = d->bn;
    p.stages = d->stages;
    p.acc_stride = round_pow2_cols(d->bn);
    p.tmem_cols = 2 * p.acc_stride;  // two accumulators: epilogue of tile i overlaps main loop of tile i+1
    p.dbg = reinterpret_cast<long long*>(g_conv_dbg);
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
    HN_REQUIRE(d->n_groups >= 0 && d->n_groups <= HN_MAX_GROUPS, "n_groups=%d out of range", d->n_groups);
    HN_REQUIRE(d->n_groups == 0 || d->flat, "row groups need flat mode");
    p.n_groups = d->n_groups;
    p.group_addr = d->group_addr;
    p.group_scale = d->group_scale;
    p.group_shift = d->group_shift;
    for (int g = 0; g < d->n_groups; ++g) {
        p.group_end[g] = d->group_end[g];
        p.group_hw[g] = d->group_hw[g] > 0 ? d->group_hw[g] : 1;
        p.group_out_base[g] = d->group_out_base[g];
    }
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
    p.m_tiles = m_tiles;
    p.n_tiles = n_tiles;
    // CTA pairs (cta_group::2, M = 256): each SM stages and reads only half of the weight tile.  The 1-CTA main
    // loop of a wide tile is bound by shared-memory bandwidth (TMA fill + operand reads ~ 96 KB per 64-deep K step);
    // the pair cuts that to 64 KB.  Used for N tiles >= 192 when there are at least two M tiles (narrower tiles do
    // better as two independent CTAs per SM).
    int cs = (g_conv_cluster > 0) ? g_conv_cluster : 2;
    if (cs != 2 || d->bn < g_conv_pair_min_bn || d->bn % 32 != 0 || m_tiles < 2 || d->epi != HN_EPI_STD) cs = 1;
    p.cluster = cs;
    p.m_groups = hn_cdiv(m_tiles, cs);
    {
        auto magic = [](long long dv) -> unsigned long long { return dv <= 1 ? 0ull : ~0ull / (unsigned long long)dv + 1ull; };
        p.div_m_groups = magic(p.m_groups);
        p.div_per_img = magic(d->flat ? 1 : (long long)p.tiles_x * p.tiles_y);
        p.div_tiles_x = magic(d->flat ? 1 : p.tiles_x);
        p.div_flat_hw = magic(d->flat ? p.flat_hw : 1);
        p.tw_shift = 0;
        while ((1 << p.tw_shift) < TW) ++p.tw_shift;
        HN_REQUIRE((1 << p.tw_shift) == TW, "tile width %d must be a power of two", TW);
    }
    int stages = d->stages;
    if (cs == 2) {
        int rc3 = encode_weight_map(&p.tmBpart, d->weight, d->w_rows, d->num_taps * 64, d->bn / 2);
        if (rc3) return rc3;
        // half-size B stages: room for a deeper ring
        int fit = (int)((227 * 1024 - 16 * 1024 - 6 * 1024) / (kATileBytes + d->bn * 64));
        stages = fit;  // the ring runs across tiles (persistent CTAs): its depth is not tied to the taps of one tile
        if (stages > 8) stages = 8;
        if (stages < 2) stages = 2;
        p.stages = stages;
    }
    const int b_rows_cta = cs == 2 ? d->bn / 2 : d->bn;
    p.gstride = n_tiles * d->bn;
    const int bias_slots = d->n_groups > 0 ? 2 * d->n_groups : 1;
    // bf16 outputs leave through shared memory + TMA store (coalesced, clipped by the tensor map) when the
    // 64-channel slabs of an N tile never spill into the next tile's channels
    p.n_staging = 0;
    bool tma_out = d->epi == HN_EPI_STD && !d->out_fp32 && !d->group_addr && (n_tiles == 1 || d->bn % 64 == 0) &&
                   (!d->flat || d->out_stride_n == (int64_t)d->flat_hw * d->out_stride_x);
    // a dense flat output of at most 32 channels: a warp's direct 16-byte stores already cover one contiguous span
    // (32 rows x cout*2 bytes), so the staging slab, its two barriers and the TMA store would be pure overhead
    if (d->flat && d->cout <= 32 && d->out_stride_x == d->cout) tma_out = false;
    const size_t fixed_smem = 1024 + (2 * 8 + 4) * 8 + 16 + 2 * (size_t)bias_slots * d->bn * 4 + 64;
    // Tap runs: consecutive taps that differ only by dy = +1 read row-shifted windows of ONE (TH+2)-row A box, so a
    // 3x3 window costs 3 box loads (3.75 tiles' worth of bytes) instead of 9 tile loads.  A window starts dy*TW rows
    // into the box; with TW % 8 == 0 that is a whole number of 1024-byte swizzle atoms, so the shifted window is
    // still a canonical K-major SWIZZLE_128B operand.  Narrow N tiles are bound by exactly this L2 -> shared-memory
    // traffic.  The ring then holds fewer, fatter stages (one run each); runs are used when two such stages plus
    // one output staging slab fit next to a second CTA on the SM (narrow tiles) or in the SM (wide tiles).
    p.run_max = 1;
    p.a_stage_bytes = kATileBytes;
    size_t smem_cap = 227 * 1024;
    if (g_conv_runs && !d->flat && TW % 8 == 0) {
        // candidate order: runs adjacent (dy innermost), and consecutive runs one pixel apart (dx next)
        int order[HN_MAX_TAPS];
        for (int k = 0; k < d->num_taps; ++k) order[k] = k;
        auto key_less = [&](int a, int b) {
            const hn_tap &x = d->taps[a], &y = d->taps[b];
            if (x.src != y.src) return x.src < y.src;
            if (x.c0 != y.c0) return x.c0 < y.c0;
            if (x.dx != y.dx) return x.dx < y.dx;
            return x.dy < y.dy;
        };
        std::stable_sort(order, order + d->num_taps, key_less);
        int n_runs = 0, longest = 1, shortest = kMaxRun;
        uint8_t run_len[HN_MAX_TAPS];
        for (int k = 0; k < d->num_taps;) {
            int r = 1;
            const hn_tap& t0 = d->taps[order[k]];
            while (r < kMaxRun && k + r < d->num_taps) {
                const hn_tap& t1 = d->taps[order[k + r]];
                if (t1.src != t0.src || t1.dx != t0.dx || t1.c0 != t0.c0 || t1.dy != t0.dy + r) break;
                ++r;
            }
            run_len[k] = (uint8_t)r;
            if (r > longest) longest = r;
            if (r < shortest) shortest = r;
            k += r;
            ++n_runs;
        }
        const size_t run_stage = (size_t)(TH + kMaxRun - 1) * TW * 128 + (size_t)kMaxRun * b_rows_cta * 128;
        // default policy: narrow N tiles (bound by L2 -> shared-memory traffic), and the big plain 3x3 layers on CTA
        // pairs; wide tiles with 2-tap runs (the parity-collapsed up-convolutions) are MMA-bound and only lose
        // pipeline granularity
        const bool want = longest > 1 && (g_conv_runs > 1 || d->bn <= 64 || (cs == 2 && shortest == kMaxRun && d->num_taps >= 36));
        const bool two_per_sm = d->bn <= 128 && 2 * p.tmem_cols <= 512;
        const size_t cap = two_per_sm ? (227 * 1024) / 2 - 1024 : 227 * 1024;
        const size_t avail = cap - fixed_smem - (tma_out ? kATileBytes : 0);
        if (want && 2 * run_stage <= avail) {
            p.run_max = kMaxRun;
            p.a_stage_bytes = (TH + kMaxRun - 1) * TW * 128;
            for (int k = 0; k < d->num_taps; ++k) {
                p.taps[k] = d->taps[order[k]];
                p.taps[k].rsv0 = 1;
                p.taps[k].rsv1 = (int16_t)order[k];
            }
            for (int k = 0; k < d->num_taps; k += run_len[k]) p.taps[k].rsv0 = (int8_t)run_len[k];
            for (int i = 0; i < d->n_src; ++i) {
                int rc4 = encode_view_map(&p.tmArun[i], d->src[i], TW, TH + kMaxRun - 1);
                if (rc4) return rc4;
            }
            stages = (int)(avail / run_stage);
            if (stages > 8) stages = 8;
            p.stages = stages;
            smem_cap = cap;
        }
    }
    const size_t base_smem = 1024 + (size_t)stages * (p.a_stage_bytes + (size_t)p.run_max * b_rows_cta * 128) +
                             (2 * stages + 4) * 8 + 16 + 2 * (size_t)bias_slots * d->bn * 4 + 64;
    // two staging slabs if that still leaves room for a second CTA on the SM (narrow tiles), else one, else whatever fits
    const auto two_per_sm_ok = [&](size_t bytes) { return 2 * (bytes + 1024) <= 227 * 1024 && 2 * p.tmem_cols <= 512; };
    if (tma_out && base_smem + kATileBytes <= smem_cap) {
        if (two_per_sm_ok(base_smem + 2 * kATileBytes)) p.n_staging = 2;
        else if (two_per_sm_ok(base_smem + kATileBytes)) p.n_staging = 1;
        else p.n_staging = (base_smem + 2 * kATileBytes <= smem_cap) ? 2 : 1;
        hn_view ov;
        if (d->flat) {
            ov.ptr = d->out; ov.N = 1; ov.H = 1; ov.W = p.flat_m; ov.C = d->cout;
            ov.stride_n = 0; ov.stride_y = 0; ov.stride_x = d->out_stride_x;
        } else {
            ov.ptr = reinterpret_cast<const bf16*>(d->out) + (int64_t)p.ooy * p.osy + (int64_t)p.oox * p.osx;
            ov.N = p.n_img; ov.H = p.H; ov.W = p.W; ov.C = d->cout;
            ov.stride_n = p.osn; ov.stride_y = p.osy * p.oscale; ov.stride_x = p.osx * p.oscale;
        }
        // the output box is one TMEM lane quarter of the tile: 32 rows = 32 pixels of a tile row, or 32/TW whole tile rows
        int rc2 = TW >= 32 ? encode_view_map(&p.tmO, ov, 32, 1) : encode_view_map(&p.tmO, ov, TW, 32 / TW);
        if (rc2) return rc2;
    }
    L->smem = base_smem + (size_t)p.n_staging * kATileBytes;
    p.tr_off = 0;
    if (d->out_fp32 && d->epi == HN_EPI_STD) {  // per-warp 32x33 transpose scratch for coalesced fp32 stores
        const size_t tr_bytes = (size_t)(kEpiThreads / 32) * 32 * 33 * 4;
        // offsets count from the 1024-byte aligned base; L->smem carries 1024 bytes of slack for that alignment
        const size_t off = (L->smem - 1024 + 15) & ~size_t(15);
        if (off + tr_bytes + 1024 <= 227 * 1024) {
            p.tr_off = (int)off;
            L->smem = off + tr_bytes + 1024;
        }
    }
    HN_REQUIRE(L->smem <= 227 * 1024, "conv needs %zu bytes of shared memory (> 227 KB): lower stages/bn", L->smem);
    // persistent grid: one CTA per SM, or two when shared memory and TMEM (512 columns) allow it
    int sms = hn_device_sm_count();
    if (sms <= 0) sms = 148;
    int per_sm = (2 * (L->smem + 1024) <= 227 * 1024 && 2 * p.tmem_cols <= 512) ? 2 : 1;
    long long total = (long long)p.m_groups * n_tiles;      // work items (one per cluster per iteration)
    long long g = (long long)sms * per_sm / cs;               // resident clusters
    L->grid = dim3((unsigned)((total < g ? total : g) * cs), 1, 1);
    L->cluster = cs;
    // the 168-register build whenever no SM gets a second CTA of this launch anyway
    L->per_sm = (per_sm == 1 || (long long)L->grid.x <= sms) ? 1 : 2;
    return HN_OK;
}

int hn_conv_launch(const ConvLaunch* L, cudaStream_t stream) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(hn_conv_gemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(hn_conv_gemm_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (attr_err == cudaSuccess)
            attr_err = cudaFuncSetAttribute(hn_conv_gemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    });
    HN_CHECK_CUDA(attr_err);
    if (L->cluster <= 1) {
        if (L->per_sm == 1) HN_CHECK_CUDA(hn_launch(hn_conv_gemm_kernel<false, 1>, L->grid, dim3(kCtaThreads), L->smem, stream, L->prm));
        else HN_CHECK_CUDA(hn_launch(hn_conv_gemm_kernel<false>, L->grid, dim3(kCtaThreads), L->smem, stream, L->prm));
        HN_CHECK_CUDA(cudaGetLastError());
        return HN_OK;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = L->grid;
    cfg.blockDim = dim3(kCtaThreads);
    cfg.dynamicSmemBytes = L->smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)L->cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_hn_pdl ? 2 : 1;
    HN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, hn_conv_gemm_kernel<true>, L->prm));
    return HN_OK;
}

extern "C" int hn_conv_fwd(const hn_conv_desc* d, void* stream) {
    ConvLaunch L;
    int rc = hn_conv_prepare(d, &L);
    if (rc) return rc;
    return hn_conv_launch(&L, reinterpret_cast<cudaStream_t>(stream));
}
