// HBM-bound operators of the training step (reference: the autograd graph behind
// model/train.py:248-264): train-mode BatchNorm forward / backward, activation derivatives, the BiFPN fusion sum,
// re-sampling adjoints, the segmentation decoder's pad / up-sample / concat assembly and its adjoint, head-gradient
// layout conversion, depthwise and stem weight gradients, squeeze-excite FC layers, weight packing and Adam.
// Activations / activation gradients are bf16 (NHWC views or [pixels][channels] row matrices); every thread owns
// 8-channel (16-byte) vectors; all reductions are deterministic (two-level: per-chunk partials summed in chunk order).
#include "hn_ops.h"

// ------------------------------------------------------------------------------------------------
// row-matrix helpers
// ------------------------------------------------------------------------------------------------
struct Mat {
    bf16* ptr;
    long long rows, ld;
    int cols;
};
static inline Mat to_mat(const hn_mat& m) {
    Mat r;
    r.ptr = reinterpret_cast<bf16*>(m.ptr);
    r.rows = m.rows;
    r.ld = m.ld;
    r.cols = m.cols;
    return r;
}
static inline int check_mat(const hn_mat& m, const char* what) {
    HN_REQUIRE(m.ptr != nullptr && m.rows >= 0 && m.cols > 0, "%s: null / empty matrix", what);
    HN_REQUIRE((reinterpret_cast<uintptr_t>(m.ptr) & 15) == 0 && m.cols % 8 == 0 && m.ld % 8 == 0 && m.ld >= m.cols,
               "%s: matrix must be 16-byte aligned with cols / ld multiples of 8 (cols=%d ld=%lld)", what, m.cols, (long long)m.ld);
    HN_REQUIRE(m.cols <= 2048, "%s: at most 2048 columns (got %d)", what, m.cols);
    return HN_OK;
}
static inline bool same_shape(const hn_mat& a, const hn_mat& b) { return a.rows == b.rows && a.cols == b.cols; }

// ex2.approx + rcp.approx (~2 ulp): the result feeds a value that is rounded to bf16; the IEEE division cost more than the rest of
// the swish gradient together
__device__ __forceinline__ float hn_sigmoid_acc(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
    const float2 a = hn_unpack_bf16x2(u.x), b = hn_unpack_bf16x2(u.y), c = hn_unpack_bf16x2(u.z), d = hn_unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
// derivative of the activation; `ref` = output for ReLU / ELU / sigmoid, pre-activation for swish
__device__ __forceinline__ float act_grad(float ref, int act) {
    switch (act) {
        case HN_ACT_RELU: return ref > 0.0f ? 1.0f : 0.0f;
        case HN_ACT_ELU: return ref > 0.0f ? 1.0f : ref + 1.0f;
        case HN_ACT_SIGMOID: return ref * (1.0f - ref);
        case HN_ACT_SWISH: {
            const float s = hn_sigmoid_acc(ref);
            return s * (1.0f + ref * (1.0f - s));
        }
        default: return 1.0f;
    }
}

// ------------------------------------------------------------------------------------------------
// two-level column reduction framework
// ------------------------------------------------------------------------------------------------
// Rows are cut into chunks that never straddle a segment; stage 1 (one CTA per chunk) writes per-chunk sums
// partial[chunk][k][C] (k = 0,1), stage 2 adds a segment's chunks in order.
struct RedGeom {
    long long rows;
    int C, CV;
    int rows_per_chunk, n_chunks, n_seg;
    // table mode (n_seg <= HN_MAX_SEG): explicit segment ends; uniform mode: segments of `uniform` rows each
    long long uniform;
    int chunks_per_seg;  // uniform mode
    long long seg_end[HN_MAX_SEG];
    int chunk_start[HN_MAX_SEG + 1];
};

__device__ __forceinline__ void red_locate(const RedGeom& g, int chunk, int& seg, long long& r0, long long& r1) {
    if (g.uniform > 0) {
        seg = chunk / g.chunks_per_seg;
        const long long b = (long long)seg * g.uniform;
        r0 = b + (long long)(chunk - seg * g.chunks_per_seg) * g.rows_per_chunk;
        r1 = min(r0 + g.rows_per_chunk, b + g.uniform);
    } else {
        seg = 0;
        while (seg < g.n_seg - 1 && chunk >= g.chunk_start[seg + 1]) ++seg;
        const long long b = seg > 0 ? g.seg_end[seg - 1] : 0;
        r0 = b + (long long)(chunk - g.chunk_start[seg]) * g.rows_per_chunk;
        r1 = min(r0 + g.rows_per_chunk, g.seg_end[seg]);
    }
}
__device__ __forceinline__ int seg_of_row(const RedGeom& g, long long row) {
    if (g.uniform > 0) return (int)(row / g.uniform);
    int s = 0;
    while (s < g.n_seg - 1 && row >= g.seg_end[s]) ++s;
    return s;
}

static int make_geom(RedGeom* g, long long rows, int C, int n_seg, const int64_t* seg_end, long long uniform, int target_chunks, int min_rpc = 32) {
    memset(g, 0, sizeof(*g));
    g->rows = rows;
    g->C = C;
    g->CV = C / 8;
    long long rpc = (rows + target_chunks - 1) / target_chunks;
    if (rpc < min_rpc) rpc = min_rpc;
    g->rows_per_chunk = (int)rpc;
    if (uniform > 0) {
        HN_REQUIRE(rows % uniform == 0, "col reduce: rows %lld not a multiple of the segment size %lld", rows, uniform);
        g->uniform = uniform;
        g->n_seg = (int)(rows / uniform);
        // keep at least ~target_chunks chunks in total, at least one per segment
        long long per = (target_chunks + g->n_seg - 1) / g->n_seg;
        rpc = (uniform + per - 1) / per;
        if (rpc < min_rpc) rpc = min_rpc;
        g->rows_per_chunk = (int)rpc;
        g->chunks_per_seg = (int)((uniform + rpc - 1) / rpc);
        g->n_chunks = g->chunks_per_seg * g->n_seg;
    } else {
        HN_REQUIRE(n_seg >= 1 && n_seg <= HN_MAX_SEG, "n_seg=%d out of range", n_seg);
        g->n_seg = n_seg;
        long long prev = 0;
        int c = 0;
        for (int s = 0; s < n_seg; ++s) {
            const long long e = n_seg == 1 && !seg_end ? rows : seg_end[s];
            HN_REQUIRE(e > prev && e <= rows, "segment %d end %lld out of order (rows %lld)", s, e, rows);
            g->seg_end[s] = e;
            g->chunk_start[s] = c;
            c += (int)((e - prev + rpc - 1) / rpc);
            prev = e;
        }
        HN_REQUIRE(prev == rows, "segments cover %lld of %lld rows", prev, rows);
        g->chunk_start[n_seg] = c;
        g->n_chunks = c;
    }
    return HN_OK;
}
static inline size_t partial_bytes(const RedGeom& g) { return (size_t)g.n_chunks * 2 * g.C * sizeof(float); }

template <class F>
__global__ void __launch_bounds__(256) hn_red1_kernel(const RedGeom g, const F f, float* __restrict__ partial) {
    __shared__ float sm[256 * 16];
    const int CV = g.CV;
    const int rpi = 256 / CV;  // rows in flight per iteration (CV <= 256)
    const int cv = threadIdx.x % CV, rl = threadIdx.x / CV;
    const bool active = rl < rpi;
    int seg;
    long long r0, r1;
    red_locate(g, blockIdx.x, seg, r0, r1);
    float a0[8], a1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a0[j] = a1[j] = 0.0f;
    if (active) {
        typename F::Ctx ctx;  // per-thread constants of the functor (the channel vector's statistics), loaded once
        f.prepare(seg, cv, ctx);
#pragma unroll 4
        for (long long r = r0 + rl; r < r1; r += rpi) f(ctx, seg, r, cv, a0, a1);
    }
    float* mine = sm + (size_t)threadIdx.x * 16;
#pragma unroll
    for (int j = 0; j < 8; ++j) { mine[j] = a0[j]; mine[8 + j] = a1[j]; }
    __syncthreads();
    if (rl == 0) {
        for (int q = 1; q < rpi; ++q) {
            const float* o = sm + (size_t)(q * CV + cv) * 16;
#pragma unroll
            for (int j = 0; j < 8; ++j) { a0[j] += o[j]; a1[j] += o[8 + j]; }
        }
        float* p0 = partial + ((size_t)blockIdx.x * 2) * g.C + cv * 8;
        float* p1 = p0 + g.C;
#pragma unroll
        for (int j = 0; j < 8; ++j) { p0[j] = a0[j]; p1[j] = a1[j]; }
    }
}

// Stage 2, cooperative: a CTA of (32 channels) x (8 chunk slices); slice y adds the chunks cb+y, cb+y+8, ... with four loads
// in flight, the eight slice sums are combined in slice order through shared memory: deterministic, and the walk over a
// segment's chunks is 8 x 4 wide instead of one dependent chain per channel.  Every thread of the CTA must call it.
static constexpr int kFinX = 32, kFinY = 8;
__device__ __forceinline__ void red_sum_chunks(const RedGeom& g, const float* __restrict__ partial, int seg, int c, bool valid, double& s0, double& s1) {
    __shared__ double sh[2][kFinY][kFinX];
    int cb, ce;
    if (g.uniform > 0) { cb = seg * g.chunks_per_seg; ce = cb + g.chunks_per_seg; }
    else { cb = g.chunk_start[seg]; ce = g.chunk_start[seg + 1]; }
    double a0 = 0.0, a1 = 0.0;
    if (valid) {
        int k = cb + threadIdx.y;
        for (; k + 3 * kFinY < ce; k += 4 * kFinY) {
            float v0[4], v1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                v0[j] = partial[((size_t)(k + j * kFinY) * 2) * g.C + c];
                v1[j] = partial[((size_t)(k + j * kFinY) * 2 + 1) * g.C + c];
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) { a0 += (double)v0[j]; a1 += (double)v1[j]; }
        }
        for (; k < ce; k += kFinY) {
            a0 += (double)partial[((size_t)k * 2) * g.C + c];
            a1 += (double)partial[((size_t)k * 2 + 1) * g.C + c];
        }
    }
    sh[0][threadIdx.y][threadIdx.x] = a0;
    sh[1][threadIdx.y][threadIdx.x] = a1;
    __syncthreads();
    s0 = s1 = 0.0;
#pragma unroll
    for (int y = 0; y < kFinY; ++y) { s0 += sh[0][y][threadIdx.x]; s1 += sh[1][y][threadIdx.x]; }
}
// finalize launch geometry: grid (ceil(C / 32), n_seg), block (32, 8); thread (x, 0) writes channel blockIdx.x * 32 + x
static inline dim3 fin_grid(const RedGeom& g) { return dim3((unsigned)((g.C + kFinX - 1) / kFinX), (unsigned)g.n_seg); }
static inline dim3 fin_block() { return dim3(kFinX, kFinY); }
__device__ __forceinline__ long long seg_rows(const RedGeom& g, int seg) {
    if (g.uniform > 0) return g.uniform;
    return g.seg_end[seg] - (seg > 0 ? g.seg_end[seg - 1] : 0);
}

// ------------------------------------------------------------------------------------------------
// BatchNorm (training)
// ------------------------------------------------------------------------------------------------
struct BnPtrs {
    const float* gamma[HN_MAX_SEG];
    const float* beta[HN_MAX_SEG];
    float* rmean[HN_MAX_SEG];
    float* rvar[HN_MAX_SEG];
    float* dgamma[HN_MAX_SEG];
    float* dbeta[HN_MAX_SEG];
};

struct NoCtx {};
struct StatsF {
    Mat z;
    typedef NoCtx Ctx;
    __device__ __forceinline__ void prepare(int, int, Ctx&) const {}
    __device__ __forceinline__ void operator()(const Ctx&, int, long long r, int cv, float (&a0)[8], float (&a1)[8]) const {
        float v[8];
        load8(z.ptr + r * z.ld + cv * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += v[j]; a1[j] = fmaf(v[j], v[j], a1[j]); }
    }
};

// stats layout: [seg][4][C] = mean, invstd, scale (gamma * invstd), shift (beta - mean * scale)
__global__ void hn_bn_finalize_kernel(const RedGeom g, const float* __restrict__ partial, BnPtrs bp, float eps, float momentum,
                                      float* __restrict__ stats) {
    const int seg = blockIdx.y, c = blockIdx.x * kFinX + threadIdx.x;
    double s0, s1;
    red_sum_chunks(g, partial, seg, c, c < g.C, s0, s1);
    if (c >= g.C || threadIdx.y != 0) return;
    const double n = (double)seg_rows(g, seg);
    const double mean = s0 / n;
    double var = s1 / n - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float ga = bp.gamma[seg] ? bp.gamma[seg][c] : 1.0f, be = bp.beta[seg] ? bp.beta[seg][c] : 0.0f;
    float* st = stats + (size_t)seg * 4 * g.C;
    st[c] = (float)mean;
    st[g.C + c] = invstd;
    st[2 * g.C + c] = ga * invstd;
    st[3 * g.C + c] = be - (float)mean * ga * invstd;
    if (bp.rmean[seg]) {
        const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
        bp.rmean[seg][c] = (1.0f - momentum) * bp.rmean[seg][c] + momentum * (float)mean;
        bp.rvar[seg][c] = (1.0f - momentum) * bp.rvar[seg][c] + momentum * (float)unbiased;
    }
}

// Element-wise passes of BatchNorm.  Mapping as in stage 1: a CTA owns a chunk of rows of ONE segment, a thread owns one 8-channel
// vector (its scale / shift live in registers) and walks the chunk's rows, kApplyU independent rows in flight.  The earlier
// flat mapping (item = row * CV + vector) spent most of its issue slots on a 64-bit division and 16-32 statistic loads per
// 16 bytes moved: the mid-size swish layers (BiFPN, detection towers) ran at 1 TB/s.
static constexpr int kEwU = 1;  // flat element-wise kernels below (activation backward, fusion): one item per thread and iteration
static constexpr int kApplyU = 4;
static constexpr int kApplyChunks = 148 * 8;
__device__ __forceinline__ void load_f8(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__global__ void __launch_bounds__(256) hn_bn_apply_kernel(const RedGeom g, Mat z, const float* __restrict__ stats, int act, Mat res, Mat y) {
    const int CV = g.CV, rpi = 256 / CV;
    const int cv = threadIdx.x % CV, rl = threadIdx.x / CV;
    if (rl >= rpi) return;
    int seg;
    long long r0, r1;
    red_locate(g, blockIdx.x, seg, r0, r1);
    float sc[8], sh[8];
    load_f8(stats + ((size_t)seg * 4 + 2) * g.C + cv * 8, sc);
    load_f8(stats + ((size_t)seg * 4 + 3) * g.C + cv * 8, sh);
    const int c = cv * 8;
    for (long long rb = r0 + rl; rb < r1; rb += (long long)rpi * kApplyU) {
        uint4 zv[kApplyU], rv[kApplyU];
#pragma unroll
        for (int u = 0; u < kApplyU; ++u) {
            const long long r = rb + (long long)u * rpi;
            if (r < r1) {
                zv[u] = *reinterpret_cast<const uint4*>(z.ptr + r * z.ld + c);
                if (res.ptr) rv[u] = *reinterpret_cast<const uint4*>(res.ptr + r * res.ld + c);
            }
        }
#pragma unroll
        for (int u = 0; u < kApplyU; ++u) {
            const long long r = rb + (long long)u * rpi;
            if (r >= r1) break;
            float v[8];
            unpack8(zv[u], v);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(v[j], sc[j], sh[j]);
            if (res.ptr) {
                float q[8];
                unpack8(rv[u], q);
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += q[j];
            }
            if (act == HN_ACT_RELU) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.0f);
            } else if (act == HN_ACT_SWISH) {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = v[j] * hn_sigmoid_acc(v[j]);
            }
            store8(y.ptr + r * y.ld + c, v);
        }
    }
}
// chunk geometry of the element-wise passes: enough chunks to fill the machine, at least one unrolled iteration per thread
static int make_apply_geom(RedGeom* g, long long rows, int C, int n_seg, const int64_t* seg_end) {
    const int rpi = 256 / (C / 8);
    return make_geom(g, rows, C, n_seg, seg_end, 0, kApplyChunks, rpi * kApplyU);
}

static int fill_bn_ptrs(const hn_bn_desc* d, BnPtrs* bp) {
    for (int s = 0; s < HN_MAX_SEG; ++s) {
        const bool on = s < d->n_seg;
        bp->gamma[s] = on ? d->gamma[s] : nullptr;
        bp->beta[s] = on ? d->beta[s] : nullptr;
        bp->rmean[s] = on ? d->running_mean[s] : nullptr;
        bp->rvar[s] = on ? d->running_var[s] : nullptr;
        bp->dgamma[s] = on ? d->dgamma[s] : nullptr;
        bp->dbeta[s] = on ? d->dbeta[s] : nullptr;
        if (on) HN_REQUIRE((bp->rmean[s] == nullptr) == (bp->rvar[s] == nullptr), "bn: running_mean / running_var must come together");
    }
    return HN_OK;
}
static inline int ew_grid(long long total_vec) {
    long long b = (total_vec + 256 * kEwU - 1) / (256 * kEwU);
    long long cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b < cap ? b : cap));
}
static constexpr int kTargetChunks = 296;  // 148 SMs x 2: stage 1 is HBM-bound, stage 2 walks the partials per channel
// chunks of the BatchNorm / column reductions (HN_RED_CHUNKS overrides, for A/B runs)
static int red_chunks() {
    static const int v = [] {
        const char* e = getenv("HN_RED_CHUNKS");
        const int n = e ? atoi(e) : 0;
        return n >= 32 && n <= 4096 ? n : kTargetChunks;
    }();
    return v;
}

extern "C" int hn_bn_train_fwd(const hn_bn_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d != nullptr && d->stats != nullptr && d->scratch != nullptr, "bn fwd: null pointer");
    if (int rc = check_mat(d->z, "bn.z")) return rc;
    if (int rc = check_mat(d->y, "bn.y")) return rc;
    HN_REQUIRE(same_shape(d->z, d->y), "bn: z / y shape mismatch");
    if (d->res.ptr) {
        if (int rc = check_mat(d->res, "bn.res")) return rc;
        HN_REQUIRE(same_shape(d->z, d->res), "bn: residual shape mismatch");
    }
    HN_REQUIRE(d->act == HN_ACT_NONE || d->act == HN_ACT_RELU || d->act == HN_ACT_SWISH, "bn: unsupported activation %d", d->act);
    if (d->z.rows == 0) return HN_OK;
    RedGeom g;
    if (int rc = make_geom(&g, d->z.rows, d->z.cols, d->n_seg, d->seg_end, 0, red_chunks())) return rc;
    HN_REQUIRE((int64_t)partial_bytes(g) <= d->scratch_bytes, "bn: scratch too small (%zu > %lld)", partial_bytes(g), (long long)d->scratch_bytes);
    BnPtrs bp;
    if (int rc = fill_bn_ptrs(d, &bp)) return rc;
    StatsF f{to_mat(d->z)};
    hn_red1_kernel<StatsF><<<g.n_chunks, 256, 0, stream>>>(g, f, d->scratch);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_bn_finalize_kernel<<<fin_grid(g), fin_block(), 0, stream>>>(g, d->scratch, bp, d->eps, d->momentum, d->stats);
    HN_CHECK_CUDA(cudaGetLastError());
    Mat res{nullptr, 0, 0, 0};
    if (d->res.ptr) res = to_mat(d->res);
    RedGeom ga;
    if (int rc = make_apply_geom(&ga, d->z.rows, d->z.cols, d->n_seg, d->seg_end)) return rc;
    hn_bn_apply_kernel<<<ga.n_chunks, 256, 0, stream>>>(ga, to_mat(d->z), d->stats, d->act, res, to_mat(d->y));
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// backward stage 1: dz_act = dy * act'(.), sums of dz_act and dz_act * xhat
struct BnBwdF {
    Mat dy, z, y;
    const float* stats;
    int C, act;
    struct Raw { uint4 g, z, y; };
    __device__ __forceinline__ void load(long long r, int cv, Raw& w) const {
        w.g = *reinterpret_cast<const uint4*>(dy.ptr + r * dy.ld + cv * 8);
        w.z = *reinterpret_cast<const uint4*>(z.ptr + r * z.ld + cv * 8);
        if (act == HN_ACT_RELU) w.y = *reinterpret_cast<const uint4*>(y.ptr + r * y.ld + cv * 8);
    }
    // the channel vector's statistics live in registers: loaded once per thread, not once per row
    struct St { float mean[8], invstd[8], scale[8], shift[8]; };
    __device__ __forceinline__ void load_stats(int seg, int cv, St& s) const {
        const float* st = stats + (size_t)seg * 4 * C + cv * 8;
        load_f8(st, s.mean);
        load_f8(st + C, s.invstd);
        load_f8(st + 2 * C, s.scale);
        if (act == HN_ACT_SWISH) load_f8(st + 3 * C, s.shift);
    }
    __device__ __forceinline__ void compute(const St& s, const Raw& w, float (&dzv)[8], float (&xh)[8]) const {
        float g[8], zz[8];
        unpack8(w.g, g);
        unpack8(w.z, zz);
#pragma unroll
        for (int j = 0; j < 8; ++j) xh[j] = (zz[j] - s.mean[j]) * s.invstd[j];
        if (act == HN_ACT_RELU) {
            float o[8];
            unpack8(w.y, o);
#pragma unroll
            for (int j = 0; j < 8; ++j) dzv[j] = o[j] > 0.0f ? g[j] : 0.0f;
        } else if (act == HN_ACT_SWISH) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dzv[j] = g[j] * act_grad(fmaf(zz[j], s.scale[j], s.shift[j]), HN_ACT_SWISH);
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) dzv[j] = g[j];
        }
    }
    typedef St Ctx;
    __device__ __forceinline__ void prepare(int seg, int cv, Ctx& s) const { load_stats(seg, cv, s); }
    __device__ __forceinline__ void operator()(const Ctx& s, int, long long r, int cv, float (&a0)[8], float (&a1)[8]) const {
        Raw w;
        load(r, cv, w);
        float dzv[8], xh[8];
        compute(s, w, dzv, xh);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += dzv[j]; a1[j] = fmaf(dzv[j], xh[j], a1[j]); }
    }
};

// sums[seg][2][C] = (mean of dz_act, mean of dz_act * xhat); parameter gradients written
__global__ void hn_bn_bwd_finalize_kernel(const RedGeom g, const float* __restrict__ partial, BnPtrs bp, float* __restrict__ sums) {
    const int seg = blockIdx.y, c = blockIdx.x * kFinX + threadIdx.x;
    double s0, s1;
    red_sum_chunks(g, partial, seg, c, c < g.C, s0, s1);
    if (c >= g.C || threadIdx.y != 0) return;
    const double n = (double)seg_rows(g, seg);
    if (bp.dbeta[seg]) bp.dbeta[seg][c] = (float)s0;
    if (bp.dgamma[seg]) bp.dgamma[seg][c] = (float)s1;
    sums[((size_t)seg * 2) * g.C + c] = (float)(s0 / n);
    sums[((size_t)seg * 2 + 1) * g.C + c] = (float)(s1 / n);
}

static constexpr int kBwdApplyU = 2;  // three loads per row
__global__ void __launch_bounds__(256) hn_bn_bwd_apply_kernel(const RedGeom g, const BnBwdF f, const float* __restrict__ sums, Mat dz, Mat dres) {
    const int CV = g.CV, rpi = 256 / CV;
    const int cv = threadIdx.x % CV, rl = threadIdx.x / CV;
    if (rl >= rpi) return;
    int seg;
    long long r0, r1;
    red_locate(g, blockIdx.x, seg, r0, r1);
    BnBwdF::St st;
    f.load_stats(seg, cv, st);
    float m0[8], m1[8];
    load_f8(sums + ((size_t)seg * 2) * g.C + cv * 8, m0);
    load_f8(sums + ((size_t)seg * 2 + 1) * g.C + cv * 8, m1);
    const int c = cv * 8;
    for (long long rb = r0 + rl; rb < r1; rb += (long long)rpi * kBwdApplyU) {
        BnBwdF::Raw w[kBwdApplyU];
#pragma unroll
        for (int u = 0; u < kBwdApplyU; ++u) {
            const long long r = rb + (long long)u * rpi;
            if (r < r1) f.load(r, cv, w[u]);
        }
#pragma unroll
        for (int u = 0; u < kBwdApplyU; ++u) {
            const long long r = rb + (long long)u * rpi;
            if (r >= r1) break;
            float dzv[8], xh[8];
            f.compute(st, w[u], dzv, xh);
            if (dres.ptr) store8(dres.ptr + r * dres.ld + c, dzv);
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = st.scale[j] * (dzv[j] - m0[j] - xh[j] * m1[j]);
            store8(dz.ptr + r * dz.ld + c, o);
        }
    }
}

extern "C" int hn_bn_train_bwd(const hn_bn_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d != nullptr && d->stats != nullptr && d->scratch != nullptr, "bn bwd: null pointer");
    if (int rc = check_mat(d->z, "bn.z")) return rc;
    if (int rc = check_mat(d->dy, "bn.dy")) return rc;
    if (int rc = check_mat(d->dz, "bn.dz")) return rc;
    HN_REQUIRE(same_shape(d->z, d->dy) && same_shape(d->z, d->dz), "bn bwd: shape mismatch");
    if (d->act == HN_ACT_RELU) {
        if (int rc = check_mat(d->y, "bn.y")) return rc;
        HN_REQUIRE(same_shape(d->z, d->y), "bn bwd: y shape mismatch");
    }
    if (d->dres.ptr) {
        if (int rc = check_mat(d->dres, "bn.dres")) return rc;
        HN_REQUIRE(same_shape(d->z, d->dres), "bn bwd: dres shape mismatch");
    }
    if (d->z.rows == 0) return HN_OK;
    RedGeom g;
    if (int rc = make_geom(&g, d->z.rows, d->z.cols, d->n_seg, d->seg_end, 0, red_chunks())) return rc;
    const size_t need = partial_bytes(g) + (size_t)g.n_seg * 2 * g.C * sizeof(float);
    HN_REQUIRE((int64_t)need <= d->scratch_bytes, "bn bwd: scratch too small (%zu > %lld)", need, (long long)d->scratch_bytes);
    BnPtrs bp;
    if (int rc = fill_bn_ptrs(d, &bp)) return rc;
    Mat ym{nullptr, 0, 0, 0};
    if (d->act == HN_ACT_RELU) ym = to_mat(d->y);
    BnBwdF f{to_mat(d->dy), to_mat(d->z), ym, d->stats, g.C, d->act};
    float* sums = d->scratch + (size_t)g.n_chunks * 2 * g.C;
    hn_red1_kernel<BnBwdF><<<g.n_chunks, 256, 0, stream>>>(g, f, d->scratch);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_bn_bwd_finalize_kernel<<<fin_grid(g), fin_block(), 0, stream>>>(g, d->scratch, bp, sums);
    HN_CHECK_CUDA(cudaGetLastError());
    Mat dres{nullptr, 0, 0, 0};
    if (d->dres.ptr) dres = to_mat(d->dres);
    RedGeom ga;
    {
        const int rpi = 256 / g.CV;
        if (int rc = make_geom(&ga, d->z.rows, d->z.cols, d->n_seg, d->seg_end, 0, kApplyChunks, rpi * kBwdApplyU)) return rc;
    }
    hn_bn_bwd_apply_kernel<<<ga.n_chunks, 256, 0, stream>>>(ga, f, sums, to_mat(d->dz), dres);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// generic column reductions
// ------------------------------------------------------------------------------------------------
struct SumF {
    Mat a;
    typedef NoCtx Ctx;
    __device__ __forceinline__ void prepare(int, int, Ctx&) const {}
    __device__ __forceinline__ void operator()(const Ctx&, int, long long r, int cv, float (&a0)[8], float (&a1)[8]) const {
        float v[8];
        load8(a.ptr + r * a.ld + cv * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) { a0[j] += v[j]; a1[j] = fmaf(v[j], v[j], a1[j]); }
    }
};
struct DotF {
    Mat a, b;
    typedef NoCtx Ctx;
    __device__ __forceinline__ void prepare(int, int, Ctx&) const {}
    __device__ __forceinline__ void operator()(const Ctx&, int, long long r, int cv, float (&a0)[8], float (&)[8]) const {
        float v[8], w[8];
        load8(a.ptr + r * a.ld + cv * 8, v);
        load8(b.ptr + r * b.ld + cv * 8, w);
#pragma unroll
        for (int j = 0; j < 8; ++j) a0[j] = fmaf(v[j], w[j], a0[j]);
    }
};
__global__ void hn_red2_kernel(const RedGeom g, const float* __restrict__ partial, float scale, float* __restrict__ out0, float* __restrict__ out1) {
    const int seg = blockIdx.y, c = blockIdx.x * kFinX + threadIdx.x;
    double s0, s1;
    red_sum_chunks(g, partial, seg, c, c < g.C, s0, s1);
    if (c >= g.C || threadIdx.y != 0) return;
    const int i = seg * g.C + c;
    out0[i] = (float)(s0 * (double)scale);
    if (out1) out1[i] = (float)(s1 * (double)scale);
}

extern "C" int hn_col_reduce(const hn_mat* a, const hn_mat* b, int32_t mode, int64_t rows_per_seg, float* out0, float* out1, float scale,
                             float* scratch, int64_t scratch_bytes, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(a != nullptr && out0 != nullptr && scratch != nullptr, "col reduce: null pointer");
    if (int rc = check_mat(*a, "reduce.a")) return rc;
    HN_REQUIRE(a->rows > 0, "col reduce: empty matrix");
    RedGeom g;
    if (int rc = make_geom(&g, a->rows, a->cols, 1, nullptr, rows_per_seg > 0 ? rows_per_seg : a->rows, red_chunks())) return rc;
    HN_REQUIRE((int64_t)partial_bytes(g) <= scratch_bytes, "col reduce: scratch too small (%zu > %lld)", partial_bytes(g), (long long)scratch_bytes);
    if (mode == 0) {
        SumF f{to_mat(*a)};
        hn_red1_kernel<SumF><<<g.n_chunks, 256, 0, stream>>>(g, f, scratch);
    } else if (mode == 1) {
        HN_REQUIRE(b != nullptr, "col reduce: dot needs b");
        if (int rc = check_mat(*b, "reduce.b")) return rc;
        HN_REQUIRE(same_shape(*a, *b), "col reduce: a / b shape mismatch");
        DotF f{to_mat(*a), to_mat(*b)};
        hn_red1_kernel<DotF><<<g.n_chunks, 256, 0, stream>>>(g, f, scratch);
        out1 = nullptr;
    } else {
        HN_REQUIRE(false, "col reduce: unknown mode %d", mode);
    }
    HN_CHECK_CUDA(cudaGetLastError());
    hn_red2_kernel<<<fin_grid(g), fin_block(), 0, stream>>>(g, scratch, scale, out0, out1);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// element-wise: activation backward (+ scaled copies), fusion sum + swish, squeeze-excite apply
// ------------------------------------------------------------------------------------------------
struct ActBwdParams {
    Mat dy, ref, dz;
    int act, n_scaled;
    Mat scaled[3];
    const float* w;
};
__global__ void __launch_bounds__(256) hn_act_bwd_kernel(const ActBwdParams p) {
    const int CV = p.dy.cols / 8;
    const long long total = p.dy.rows * CV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / CV;
        const int c = (int)(i - r * CV) * 8;
        float g[8];
        load8(p.dy.ptr + r * p.dy.ld + c, g);
        if (p.act != HN_ACT_NONE) {
            float q[8];
            load8(p.ref.ptr + r * p.ref.ld + c, q);
#pragma unroll
            for (int j = 0; j < 8; ++j) g[j] *= act_grad(q[j], p.act);
        }
        if (p.dz.ptr) store8(p.dz.ptr + r * p.dz.ld + c, g);
        for (int k = 0; k < p.n_scaled; ++k) {
            float o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = g[j] * p.w[k];
            store8(p.scaled[k].ptr + r * p.scaled[k].ld + c, o);
        }
    }
}
extern "C" int hn_act_bwd(const hn_actbwd_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d != nullptr, "act bwd: null desc");
    if (int rc = check_mat(d->dy, "actbwd.dy")) return rc;
    ActBwdParams p;
    memset(&p, 0, sizeof(p));
    p.dy = to_mat(d->dy);
    p.act = d->act;
    if (d->act != HN_ACT_NONE) {
        if (int rc = check_mat(d->ref, "actbwd.ref")) return rc;
        HN_REQUIRE(same_shape(d->dy, d->ref), "act bwd: ref shape mismatch");
        p.ref = to_mat(d->ref);
    }
    if (d->dz.ptr) {
        if (int rc = check_mat(d->dz, "actbwd.dz")) return rc;
        HN_REQUIRE(same_shape(d->dy, d->dz), "act bwd: dz shape mismatch");
        p.dz = to_mat(d->dz);
    }
    HN_REQUIRE(d->n_scaled >= 0 && d->n_scaled <= 3, "act bwd: n_scaled=%d", d->n_scaled);
    p.n_scaled = d->n_scaled;
    for (int k = 0; k < d->n_scaled; ++k) {
        if (int rc = check_mat(d->scaled[k], "actbwd.scaled")) return rc;
        HN_REQUIRE(same_shape(d->dy, d->scaled[k]), "act bwd: scaled output shape mismatch");
        p.scaled[k] = to_mat(d->scaled[k]);
    }
    HN_REQUIRE(d->n_scaled == 0 || d->w != nullptr, "act bwd: scaled outputs need the weight vector");
    p.w = d->w;
    if (d->dy.rows == 0) return HN_OK;
    hn_act_bwd_kernel<<<ew_grid(p.dy.rows * (p.dy.cols / 8)), 256, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

struct WsumParams {
    int n_in;
    Mat in[3];
    const float* w;
    Mat s, a;
};
__global__ void __launch_bounds__(256) hn_wsum_kernel(const WsumParams p) {
    const int CV = p.s.cols / 8;
    const long long total = p.s.rows * CV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / CV;
        const int c = (int)(i - r * CV) * 8;
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
        for (int k = 0; k < p.n_in; ++k) {
            float v[8];
            load8(p.in[k].ptr + r * p.in[k].ld + c, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaf(p.w[k], v[j], acc[j]);
        }
        store8(p.s.ptr + r * p.s.ld + c, acc);
        // the swish reads the ROUNDED sum: the backward differentiates exactly what was stored
        float sr[8];
        load8(p.s.ptr + r * p.s.ld + c, sr);
#pragma unroll
        for (int j = 0; j < 8; ++j) sr[j] = sr[j] * hn_sigmoid_acc(sr[j]);
        store8(p.a.ptr + r * p.a.ld + c, sr);
    }
}
extern "C" int hn_wsum_swish_fwd(const hn_wsum_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d != nullptr && d->n_in >= 1 && d->n_in <= 3, "wsum: bad desc");
    if (int rc = check_mat(d->s, "wsum.s")) return rc;
    if (int rc = check_mat(d->a, "wsum.a")) return rc;
    HN_REQUIRE(same_shape(d->s, d->a), "wsum: s / a shape mismatch");
    WsumParams p;
    memset(&p, 0, sizeof(p));
    p.n_in = d->n_in;
    for (int k = 0; k < d->n_in; ++k) {
        if (int rc = check_mat(d->in[k], "wsum.in")) return rc;
        HN_REQUIRE(same_shape(d->s, d->in[k]), "wsum: input %d shape mismatch", k);
        p.in[k] = to_mat(d->in[k]);
    }
    HN_REQUIRE(d->w != nullptr, "wsum: null weight vector");
    p.w = d->w;
    p.s = to_mat(d->s);
    p.a = to_mat(d->a);
    if (d->s.rows == 0) return HN_OK;
    hn_wsum_kernel<<<ew_grid(p.s.rows * (p.s.cols / 8)), 256, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

__global__ void __launch_bounds__(256) hn_se_apply_kernel(Mat x, const float* __restrict__ gate, const float* __restrict__ add, long long rpi, Mat y) {
    const int CV = x.cols / 8;
    const long long total = x.rows * CV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / CV;
        const int c = (int)(i - r * CV) * 8;
        const long long n = r / rpi;
        float v[8];
        load8(x.ptr + r * x.ld + c, v);
        const float* gp = gate + n * x.cols + c;
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] *= gp[j];
        if (add) {
            const float* ap = add + n * x.cols + c;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += ap[j];
        }
        store8(y.ptr + r * y.ld + c, v);
    }
}
extern "C" int hn_se_apply(const hn_mat* x, const float* gate, const float* add, int64_t rows_per_img, const hn_mat* y, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(x && y && gate && rows_per_img > 0, "se apply: bad arguments");
    if (int rc = check_mat(*x, "se.x")) return rc;
    if (int rc = check_mat(*y, "se.y")) return rc;
    HN_REQUIRE(same_shape(*x, *y) && x->rows % rows_per_img == 0, "se apply: shape mismatch");
    if (x->rows == 0) return HN_OK;
    hn_se_apply_kernel<<<ew_grid(x->rows * (x->cols / 8)), 256, 0, stream>>>(to_mat(*x), gate, add, rows_per_img, to_mat(*y));
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// re-sampling: nearest x2 up-sampling, the two 3x3 stride-2 max-pools, and their adjoints
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void decode4(unsigned idx, const View& v, int& n, int& y, int& x, int& cv) {
    const unsigned CV = (unsigned)(v.C >> 3);
    cv = (int)(idx % CV);
    unsigned t = idx / CV;
    x = (int)(t % (unsigned)v.W);
    t /= (unsigned)v.W;
    y = (int)(t % (unsigned)v.H);
    n = (int)(t / (unsigned)v.H);
}
__global__ void hn_up2_kernel(View in, View out) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)out.N * out.H * out.W * (out.C >> 3)) return;
    int n, y, x, cv;
    decode4(idx, out, n, y, x, cv);
    *reinterpret_cast<uint4*>(const_cast<bf16*>(vptr(out, n, y, x, cv * 8))) =
        *reinterpret_cast<const uint4*>(vptr(in, n, min(y >> 1, in.H - 1), min(x >> 1, in.W - 1), cv * 8));
}
__global__ void hn_up2_bwd_kernel(View dy, View dx) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)dx.N * dx.H * dx.W * (dx.C >> 3)) return;
    int n, y, x, cv;
    decode4(idx, dx, n, y, x, cv);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            const int Y = 2 * y + a, X = 2 * x + b;
            if (Y >= dy.H || X >= dy.W) continue;
            float g[8];
            load8(vptr(dy, n, Y, X, cv * 8), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += g[j];
        }
    store8(const_cast<bf16*>(vptr(dx, n, y, x, cv * 8)), acc);
}
// Pool backward, gather form: input pixel (Y, X) collects the gradient of every output window that contains it and whose
// FIRST maximum in (ky, kx) scan order is (Y, X).  pad = 0: window rows 2y .. 2y+2, zeros beyond the bottom / right edge
// take part (and swallow the gradient when they win); pad = 1: window rows 2y-1 .. 2y+1, outside = -inf.
// A thread owns a 2x2 block of dx whose first row / column are the ones SHARED by two windows (rows 2*by - pad and 2*by - pad + 1):
// the block is touched by exactly the windows (by-1 .. by) x (bx-1 .. bx), so the arg-max of a window is evaluated once per block
// it reaches (9 loads per input pixel instead of the 36 of a per-pixel gather).  Deterministic: no atomics, fixed order.
__global__ void hn_pool_bwd_kernel(View x, View dy, View dx, int pad, int BH, int BW) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned CV = (unsigned)(dx.C >> 3);
    if (idx >= (unsigned)dx.N * BH * BW * CV) return;
    const int cv = (int)(idx % CV);
    unsigned t = idx / CV;
    const int bx = (int)(t % (unsigned)BW);
    t /= (unsigned)BW;
    const int by = (int)(t % (unsigned)BH), n = (int)(t / (unsigned)BH);
    const int c = cv * 8;
    const int Ya = 2 * by - pad, Xa = 2 * bx - pad;  // block rows Ya, Ya + 1; columns Xa, Xa + 1
    float acc[2][2][8];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[a][b2][j] = 0.0f;
#pragma unroll
    for (int wy = 0; wy < 2; ++wy) {
#pragma unroll
        for (int wx = 0; wx < 2; ++wx) {
            const int oy = by - 1 + wy, ox = bx - 1 + wx;
            if (oy < 0 || oy >= dy.H || ox < 0 || ox >= dy.W) continue;
            float best[8];
            int arg[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = -1; }
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int iy = 2 * oy - pad + ky, ix = 2 * ox - pad + kx;
                    float v[8];
                    if (iy >= 0 && iy < x.H && ix >= 0 && ix < x.W) {
                        load8(vptr(x, n, iy, ix, c), v);
                    } else if (pad == 0) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = 0.0f;
                    } else {
                        continue;
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (v[j] > best[j] || arg[j] < 0) { best[j] = v[j]; arg[j] = ky * 3 + kx; }
                }
            float g[8];
            load8(vptr(dy, n, oy, ox, c), g);
            // window cell (ky, kx) is block cell (ky - 2 + 2*wy... ): window row 2*oy - pad + ky = Ya + (ky - 2 + 2 * wy)
#pragma unroll
            for (int a = 0; a < 2; ++a)
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    const int ky = a + 2 - 2 * wy, kx = b2 + 2 - 2 * wx;  // the window's cell at block position (a, b2)
                    if (ky > 2 || kx > 2) continue;
                    const int cell = ky * 3 + kx;
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        if (arg[j] == cell) acc[a][b2][j] += g[j];
                }
        }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
            const int Y = Ya + a, X = Xa + b2;
            if (Y >= 0 && Y < dx.H && X >= 0 && X < dx.W) store8(const_cast<bf16*>(vptr(dx, n, Y, X, c)), acc[a][b2]);
        }
}

extern "C" int hn_pool_fwd(const hn_pool_desc* d, void* stream);
extern "C" int hn_resample_fwd(const hn_resample_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d != nullptr, "resample: null desc");
    if (d->mode == HN_RS_UP2) {
        if (int rc = check_view(d->x, "up2.x")) return rc;
        if (int rc = check_view(d->y, "up2.y")) return rc;
        HN_REQUIRE(d->y.H == 2 * d->x.H && d->y.W == 2 * d->x.W && d->y.C == d->x.C && d->y.N == d->x.N, "up2: shape mismatch");
        const long long total = (long long)d->y.N * d->y.H * d->y.W * (d->y.C / 8);
        HN_REQUIRE(total < 0x7fffffffLL, "up2: too many work items");
        if (total == 0) return HN_OK;
        hn_up2_kernel<<<hn_cdiv(total, 256), 256, 0, stream>>>(to_view(d->x), to_view(d->y));
        HN_CHECK_CUDA(cudaGetLastError());
        return HN_OK;
    }
    hn_pool_desc pd;
    pd.in = d->x;
    pd.out = d->y;
    pd.mode = d->mode == HN_RS_POOL_ZERO ? HN_POOL_ZERO_RB : HN_POOL_NEGINF;
    return hn_pool_fwd(&pd, stream_);
}
extern "C" int hn_resample_bwd(const hn_resample_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d != nullptr, "resample bwd: null desc");
    if (int rc = check_view(d->dy, "resample.dy")) return rc;
    if (int rc = check_view(d->dx, "resample.dx")) return rc;
    HN_REQUIRE(d->dy.C == d->dx.C && d->dy.N == d->dx.N, "resample bwd: channel / batch mismatch");
    const long long total = (long long)d->dx.N * d->dx.H * d->dx.W * (d->dx.C / 8);
    HN_REQUIRE(total < 0x7fffffffLL, "resample bwd: too many work items");
    if (total == 0) return HN_OK;
    if (d->mode == HN_RS_UP2) {
        HN_REQUIRE(d->dy.H == 2 * d->dx.H && d->dy.W == 2 * d->dx.W, "up2 bwd: shape mismatch");
        hn_up2_bwd_kernel<<<hn_cdiv(total, 256), 256, 0, stream>>>(to_view(d->dy), to_view(d->dx));
    } else {
        if (int rc = check_view(d->x, "pool.x")) return rc;
        HN_REQUIRE(d->x.H == d->dx.H && d->x.W == d->dx.W && d->x.C == d->dx.C && d->x.N == d->dx.N, "pool bwd: x / dx mismatch");
        const int pad = d->mode == HN_RS_POOL_ZERO ? 0 : 1;
        if (pad == 0) HN_REQUIRE((d->x.H - 2) / 2 + 1 == d->dy.H && (d->x.W - 2) / 2 + 1 == d->dy.W, "pool bwd: size mismatch");
        else HN_REQUIRE((d->x.H - 1) / 2 + 1 == d->dy.H && (d->x.W - 1) / 2 + 1 == d->dy.W, "pool bwd: size mismatch");
        const int BH = (d->dx.H - 1 + pad) / 2 + 1, BW = (d->dx.W - 1 + pad) / 2 + 1;  // 2x2 blocks of dx (rows 2*by - pad, + 1)
        const long long nb = (long long)d->dx.N * BH * BW * (d->dx.C / 8);
        hn_pool_bwd_kernel<<<hn_cdiv(nb, 128), 128, 0, stream>>>(to_view(d->x), to_view(d->dy), to_view(d->dx), pad, BH, BW);
    }
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// segmentation decoder input: ReflectionPad2d(1)(cat(up2(low), skip)) and its adjoint
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__global__ void hn_seggather_kernel(View low, View skip, View out, int Cl) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)out.N * out.H * out.W * (out.C >> 3)) return;
    int n, Y, X, cv;
    decode4(idx, out, n, Y, X, cv);
    const int H = out.H - 2, W = out.W - 2;
    const int y = reflect1(Y - 1, H), x = reflect1(X - 1, W), c = cv * 8;
    const bf16* src = c < Cl ? vptr(low, n, y >> 1, x >> 1, c) : vptr(skip, n, y, x, c - Cl);
    *reinterpret_cast<uint4*>(const_cast<bf16*>(vptr(out, n, Y, X, c))) = *reinterpret_cast<const uint4*>(src);
}
// folded gradient of interior pixel (y, x): the padded positions that mirror onto it
__device__ __forceinline__ void seg_fold(const View& d, int n, int y, int x, int c, int H, int W, float (&acc)[8]) {
    int ys[3], xs[3], ny = 0, nx = 0;
    ys[ny++] = y + 1;
    if (y == 1) ys[ny++] = 0;
    if (y == H - 2) ys[ny++] = H + 1;
    xs[nx++] = x + 1;
    if (x == 1) xs[nx++] = 0;
    if (x == W - 2) xs[nx++] = W + 1;
    for (int a = 0; a < ny; ++a)
        for (int b = 0; b < nx; ++b) {
            float g[8];
            load8(vptr(d, n, ys[a], xs[b], c), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += g[j];
        }
}
__global__ void hn_seggather_bwd_skip_kernel(View dpad, View dskip, int Cl) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)dskip.N * dskip.H * dskip.W * (dskip.C >> 3)) return;
    int n, y, x, cv;
    decode4(idx, dskip, n, y, x, cv);
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    seg_fold(dpad, n, y, x, Cl + cv * 8, dskip.H, dskip.W, acc);
    store8(const_cast<bf16*>(vptr(dskip, n, y, x, cv * 8)), acc);
}
__global__ void hn_seggather_bwd_low_kernel(View dpad, View dlow) {
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (unsigned)dlow.N * dlow.H * dlow.W * (dlow.C >> 3)) return;
    int n, y, x, cv;
    decode4(idx, dlow, n, y, x, cv);
    const int H = dpad.H - 2, W = dpad.W - 2;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) seg_fold(dpad, n, 2 * y + a, 2 * x + b, cv * 8, H, W, acc);
    store8(const_cast<bf16*>(vptr(dlow, n, y, x, cv * 8)), acc);
}
static int seggather_check(const hn_seggather_desc* d, const hn_view& lo, const hn_view& sk, int* Cl) {
    if (int rc = check_view(d->out, "seggather.out")) return rc;
    const int H = d->out.H - 2, W = d->out.W - 2;
    HN_REQUIRE(H >= 3 && W >= 3, "seggather: map too small for reflection padding");
    *Cl = 0;
    if (lo.ptr) {
        if (int rc = check_view(lo, "seggather.low")) return rc;
        HN_REQUIRE(lo.H * 2 == H && lo.W * 2 == W && lo.N == d->out.N, "seggather: low-resolution input shape mismatch");
        *Cl = lo.C;
    }
    int Cs = 0;
    if (sk.ptr) {
        if (int rc = check_view(sk, "seggather.skip")) return rc;
        HN_REQUIRE(sk.H == H && sk.W == W && sk.N == d->out.N, "seggather: skip input shape mismatch");
        Cs = sk.C;
    }
    HN_REQUIRE(*Cl + Cs == d->out.C && d->out.C > 0, "seggather: channels %d + %d != %d", *Cl, Cs, d->out.C);
    HN_REQUIRE((long long)d->out.N * d->out.H * d->out.W * (d->out.C / 8) < 0x7fffffffLL, "seggather: too many work items");
    return HN_OK;
}
extern "C" int hn_seggather_fwd(const hn_seggather_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d != nullptr, "seggather: null desc");
    int Cl;
    if (int rc = seggather_check(d, d->low, d->skip, &Cl)) return rc;
    View lo, sk;
    memset(&lo, 0, sizeof(lo));
    memset(&sk, 0, sizeof(sk));
    if (d->low.ptr) lo = to_view(d->low);
    if (d->skip.ptr) sk = to_view(d->skip);
    const long long total = (long long)d->out.N * d->out.H * d->out.W * (d->out.C / 8);
    if (total == 0) return HN_OK;
    hn_seggather_kernel<<<hn_cdiv(total, 256), 256, 0, stream>>>(lo, sk, to_view(d->out), Cl);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}
extern "C" int hn_seggather_bwd(const hn_seggather_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d != nullptr, "seggather bwd: null desc");
    int Cl;
    if (int rc = seggather_check(d, d->dlow, d->dskip, &Cl)) return rc;
    if (d->dskip.ptr) {
        const long long total = (long long)d->dskip.N * d->dskip.H * d->dskip.W * (d->dskip.C / 8);
        if (total) hn_seggather_bwd_skip_kernel<<<hn_cdiv(total, 256), 256, 0, stream>>>(to_view(d->out), to_view(d->dskip), Cl);
        HN_CHECK_CUDA(cudaGetLastError());
    }
    if (d->dlow.ptr) {
        const long long total = (long long)d->dlow.N * d->dlow.H * d->dlow.W * (d->dlow.C / 8);
        if (total) hn_seggather_bwd_low_kernel<<<hn_cdiv(total, 256), 256, 0, stream>>>(to_view(d->out), to_view(d->dlow));
        HN_CHECK_CUDA(cudaGetLastError());
    }
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// head gradients: fp32 tensors in the reference's output layouts -> bf16 gradient rows
// ------------------------------------------------------------------------------------------------
struct HeadGradParams {
    const float* dout;
    const float* out;
    int act, cols_valid, n_groups;
    long long sn, spix, sc, rows_per_img;
    long long group_end[HN_MAX_GROUPS], group_hw[HN_MAX_GROUPS], group_out_base[HN_MAX_GROUPS];
    Mat dz;
};
__global__ void __launch_bounds__(256) hn_head_grad_kernel(const HeadGradParams p) {
    const int CV = p.dz.cols / 8;
    const long long total = p.dz.rows * CV;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        // rows fastest within a vector column: neighbouring threads read neighbouring pixels (coalesced for NCHW sources)
        const int cv = (int)(i / p.dz.rows);
        const long long r = i - (long long)cv * p.dz.rows;
        long long base;
        if (p.n_groups > 0) {
            int g = 0;
            while (g < p.n_groups - 1 && r >= p.group_end[g]) ++g;
            const long long ml = r - (g > 0 ? p.group_end[g - 1] : 0);
            const long long n = ml / p.group_hw[g], pix = ml - n * p.group_hw[g];
            base = n * p.sn + p.group_out_base[g] + pix * p.spix;
        } else {
            const long long n = r / p.rows_per_img, pix = r - n * p.rows_per_img;
            base = n * p.sn + pix * p.spix;
        }
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j;
            float g = 0.0f;
            if (c < p.cols_valid) {
                g = p.dout[base + c * p.sc];
                if (p.act == HN_ACT_SIGMOID) {
                    const float o = p.out[base + c * p.sc];
                    g *= o * (1.0f - o);
                }
            }
            v[j] = g;
        }
        store8(p.dz.ptr + r * p.dz.ld + cv * 8, v);
    }
}
extern "C" int hn_head_grad(const hn_headgrad_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d != nullptr && d->dout != nullptr, "head grad: null pointer");
    if (int rc = check_mat(d->dz, "headgrad.dz")) return rc;
    HN_REQUIRE(d->cols_valid >= 1 && d->cols_valid <= d->dz.cols, "head grad: cols_valid=%d of %d", d->cols_valid, d->dz.cols);
    HN_REQUIRE(d->act == HN_ACT_NONE || (d->act == HN_ACT_SIGMOID && d->out != nullptr), "head grad: unsupported activation");
    HN_REQUIRE(d->n_groups >= 0 && d->n_groups <= HN_MAX_GROUPS, "head grad: n_groups=%d", d->n_groups);
    HN_REQUIRE(d->n_groups > 0 || d->rows_per_img > 0, "head grad: rows_per_img missing");
    HeadGradParams p;
    memset(&p, 0, sizeof(p));
    p.dout = d->dout;
    p.out = d->out;
    p.act = d->act;
    p.cols_valid = d->cols_valid;
    p.n_groups = d->n_groups;
    p.sn = d->stride_n;
    p.spix = d->stride_pix;
    p.sc = d->stride_c;
    p.rows_per_img = d->rows_per_img;
    for (int g = 0; g < d->n_groups; ++g) {
        p.group_end[g] = d->group_end[g];
        p.group_hw[g] = d->group_hw[g] > 0 ? d->group_hw[g] : 1;
        p.group_out_base[g] = d->group_out_base[g];
    }
    p.dz = to_mat(d->dz);
    if (d->dz.rows == 0) return HN_OK;
    hn_head_grad_kernel<<<ew_grid(p.dz.rows * (p.dz.cols / 8)), 256, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// depthwise 3x3 weight gradient
// ------------------------------------------------------------------------------------------------
// CTA = chunk of pixels; thread = (8-channel vector, pixel lane) with 9 x 8 accumulators; per-chunk partial sums
// [chunk][9][C], summed in chunk order by the finalize kernel.
__global__ void __launch_bounds__(256) hn_dw_wgrad_kernel(View x, View dy, int pix_per_chunk, float* __restrict__ partial) {
    __shared__ float sm[256 * 8];
    const int CV = x.C >> 3;
    const int ppi = 256 / CV;
    const int cv = threadIdx.x % CV, pl = threadIdx.x / CV;
    const bool active = pl < ppi;
    const long long total = (long long)x.N * x.H * x.W;
    const long long p0 = (long long)blockIdx.x * pix_per_chunk, p1 = min(p0 + pix_per_chunk, total);
    float acc[9][8];
#pragma unroll
    for (int t = 0; t < 9; ++t)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[t][j] = 0.0f;
    if (active)
        for (long long p = p0 + pl; p < p1; p += ppi) {
            const int xx = (int)(p % x.W);
            const long long q = p / x.W;
            const int yy = (int)(q % x.H), n = (int)(q / x.H);
            float g[8];
            load8(vptr(dy, n, yy, xx, cv * 8), g);
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const int iy = yy + ky - 1;
                if (iy < 0 || iy >= x.H) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int ix = xx + kx - 1;
                    if (ix < 0 || ix >= x.W) continue;
                    float v[8];
                    load8(vptr(x, n, iy, ix, cv * 8), v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[ky * 3 + kx][j] = fmaf(g[j], v[j], acc[ky * 3 + kx][j]);
                }
            }
        }
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[threadIdx.x * 8 + j] = acc[t][j];
        __syncthreads();
        if (pl == 0) {
            float s[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) s[j] = acc[t][j];
            for (int q = 1; q < ppi; ++q)
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] += sm[(q * CV + cv) * 8 + j];
            float* o = partial + ((size_t)blockIdx.x * 9 + t) * x.C + cv * 8;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = s[j];
        }
    }
}
__global__ void hn_sum_partials_kernel(const float* __restrict__ partial, int n_chunks, int n, float* __restrict__ out, int accumulate) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s = accumulate ? (double)out[i] : 0.0;
    for (int k = 0; k < n_chunks; ++k) s += (double)partial[(size_t)k * n + i];
    out[i] = (float)s;
}
extern "C" int hn_dw_wgrad(const hn_view* x, const hn_view* dy, float* dw, int32_t accumulate, float* scratch, int64_t scratch_bytes, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(x && dy && dw && scratch, "dw wgrad: null pointer");
    if (int rc = check_view(*x, "dwwgrad.x")) return rc;
    if (int rc = check_view(*dy, "dwwgrad.dy")) return rc;
    HN_REQUIRE(x->N == dy->N && x->H == dy->H && x->W == dy->W && x->C == dy->C && x->C <= 2048, "dw wgrad: shape mismatch");
    const long long total = (long long)x->N * x->H * x->W;
    if (total == 0) return HN_OK;
    long long ppc = (total + kTargetChunks - 1) / kTargetChunks;
    if (ppc < 64) ppc = 64;
    const int chunks = (int)((total + ppc - 1) / ppc);
    HN_REQUIRE((int64_t)((size_t)chunks * 9 * x->C * 4) <= scratch_bytes, "dw wgrad: scratch too small");
    hn_dw_wgrad_kernel<<<chunks, 256, 0, stream>>>(to_view(*x), to_view(*dy), (int)ppc, scratch);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_sum_partials_kernel<<<hn_cdiv(9 * x->C, 128), 128, 0, stream>>>(scratch, chunks, 9 * x->C, dw, accumulate);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// stem weight gradient: dW[co][k] (k = ci*9 + ky*3 + kx), warp = strip of output pixels, lane = output channel
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hn_stem_wgrad_kernel(const float* __restrict__ x, int N, int H, int W, View dz, int pix_per_chunk,
                                                            float* __restrict__ partial) {
    __shared__ float sm[8][27][32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int OH = dz.H, OW = dz.W;
    const long long total = (long long)N * OH * OW;
    const long long p0 = (long long)blockIdx.x * pix_per_chunk, p1 = min(p0 + pix_per_chunk, total);
    float acc[27];
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[k] = 0.0f;
    const int ci = lane / 9, kk = lane - ci * 9, ky = kk / 3, kx = kk - ky * 3;  // lanes 0..26 fetch one input sample each
    for (long long p = p0 + warp; p < p1; p += 8) {
        const int ox = (int)(p % OW);
        const long long q = p / OW;
        const int oy = (int)(q % OH), n = (int)(q / OH);
        const float g = __bfloat162float(*vptr(dz, n, oy, ox, lane));
        float v = 0.0f;
        if (lane < 27) {
            const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = __ldg(x + (((long long)n * 3 + ci) * H + iy) * W + ix);
        }
#pragma unroll
        for (int k = 0; k < 27; ++k) acc[k] = fmaf(g, __shfl_sync(0xffffffffu, v, k), acc[k]);
    }
#pragma unroll
    for (int k = 0; k < 27; ++k) sm[warp][k][lane] = acc[k];
    __syncthreads();
    for (int i = threadIdx.x; i < 27 * 32; i += 256) {
        const int k = i >> 5, co = i & 31;
        float s = 0.0f;
#pragma unroll
        for (int w = 0; w < 8; ++w) s += sm[w][k][co];
        partial[(size_t)blockIdx.x * 864 + co * 27 + k] = s;  // OIHW order: [co][ci][ky][kx]
    }
}
extern "C" int hn_stem_wgrad(const float* x, int32_t N, int32_t H, int32_t W, const hn_view* dz, float* dw, float* scratch, int64_t scratch_bytes,
                             void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(x && dz && dw && scratch, "stem wgrad: null pointer");
    if (int rc = check_view(*dz, "stemwgrad.dz")) return rc;
    HN_REQUIRE(dz->C == 32 && dz->N == N && dz->H == (H + 1) / 2 && dz->W == (W + 1) / 2, "stem wgrad: shape mismatch");
    const long long total = (long long)N * dz->H * dz->W;
    if (total == 0) return HN_OK;
    long long ppc = (total + 2 * kTargetChunks - 1) / (2 * kTargetChunks);
    if (ppc < 64) ppc = 64;
    const int chunks = (int)((total + ppc - 1) / ppc);
    HN_REQUIRE((int64_t)((size_t)chunks * 864 * 4) <= scratch_bytes, "stem wgrad: scratch too small");
    hn_stem_wgrad_kernel<<<chunks, 256, 0, stream>>>(x, N, H, W, to_view(*dz), (int)ppc, scratch);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_sum_partials_kernel<<<hn_cdiv(864, 128), 128, 0, stream>>>(scratch, chunks, 864, dw, 0);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// squeeze-excite FC layers (fp32, one CTA per image)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Both FC layers serve EVERY image per CTA (the weights are read once per CTA, not once per image): FC1 = one CTA per
// group of 4 hidden units, FC2 = one CTA per group of 32 channels; warps take images, lanes split the reduction.
static constexpr int kSeU = 4, kSeCh = 32;
__global__ void __launch_bounds__(256) hn_se_fc1_train_kernel(const hn_sefc_desc d) {
    extern __shared__ float sm[];  // [kSeU][C]
    const int s0 = blockIdx.x * kSeU;
    for (int i = threadIdx.x; i < kSeU * d.C; i += 256) {
        const int u = i / d.C, c = i - u * d.C;
        sm[i] = s0 + u < d.S ? d.w1[(size_t)(s0 + u) * d.C + c] : 0.0f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = warp; n < d.N; n += 8) {
        float acc[kSeU];
#pragma unroll
        for (int u = 0; u < kSeU; ++u) acc[u] = 0.0f;
        for (int c = lane; c < d.C; c += 32) {
            const float m = d.mean[(size_t)n * d.C + c];
#pragma unroll
            for (int u = 0; u < kSeU; ++u) acc[u] = fmaf(sm[u * d.C + c], m, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < kSeU; ++u) {
            const float a = warp_sum(acc[u]);
            if (lane == 0 && s0 + u < d.S) d.h[(size_t)n * d.S + s0 + u] = fmaxf(a + d.b1[s0 + u], 0.0f);
        }
    }
}
__global__ void __launch_bounds__(256) hn_se_fc2_train_kernel(const hn_sefc_desc d) {
    extern __shared__ float sm[];  // [kSeCh][S + 1]
    const int c0 = blockIdx.x * kSeCh, ld = d.S + 1;
    for (int i = threadIdx.x; i < kSeCh * d.S; i += 256) {
        const int u = i / d.S, s = i - u * d.S;
        sm[u * ld + s] = c0 + u < d.C ? d.w2[(size_t)(c0 + u) * d.S + s] : 0.0f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c = c0 + lane;
    for (int n = warp; n < d.N; n += 8) {
        const float* h = d.h + (size_t)n * d.S;
        float a = 0.0f;
        for (int s = 0; s < d.S; ++s) a = fmaf(sm[lane * ld + s], h[s], a);
        if (c < d.C) d.gate[(size_t)n * d.C + c] = 1.0f / (1.0f + expf(-(a + d.b2[c])));
    }
}
// backward: ds2 = dgate * gate * (1 - gate) (tmp[n][0:C]); dh = (h > 0) * W2^T ds2 (tmp[n][C:C+S]); dmean = W1^T dh
__global__ void __launch_bounds__(256) hn_se_ds2_kernel(const hn_sefc_desc d) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= d.N * d.C) return;
    const int n = i / d.C, c = i - n * d.C;
    const float g = d.gate[i];
    d.tmp[(size_t)n * (d.C + d.S) + c] = d.dgate[i] * g * (1.0f - g);
}
__global__ void __launch_bounds__(256) hn_se_dh_kernel(const hn_sefc_desc d) {
    extern __shared__ float sm[];  // [kSeU][C]: columns s0 .. s0+3 of W2
    const int s0 = blockIdx.x * kSeU, T = d.C + d.S;
    for (int i = threadIdx.x; i < kSeU * d.C; i += 256) {
        const int c = i / kSeU, u = i - c * kSeU;
        sm[u * d.C + c] = s0 + u < d.S ? d.w2[(size_t)c * d.S + s0 + u] : 0.0f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int n = warp; n < d.N; n += 8) {
        float acc[kSeU];
#pragma unroll
        for (int u = 0; u < kSeU; ++u) acc[u] = 0.0f;
        for (int c = lane; c < d.C; c += 32) {
            const float v = d.tmp[(size_t)n * T + c];
#pragma unroll
            for (int u = 0; u < kSeU; ++u) acc[u] = fmaf(sm[u * d.C + c], v, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < kSeU; ++u) {
            const float a = warp_sum(acc[u]);
            if (lane == 0 && s0 + u < d.S) d.tmp[(size_t)n * T + d.C + s0 + u] = d.h[(size_t)n * d.S + s0 + u] > 0.0f ? a : 0.0f;
        }
    }
}
__global__ void __launch_bounds__(256) hn_se_dmean_kernel(const hn_sefc_desc d) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i >= d.N * d.C) return;
    const int n = i / d.C, c = i - n * d.C;
    const float* dh = d.tmp + (size_t)n * (d.C + d.S) + d.C;
    float a = 0.0f;
    for (int s = 0; s < d.S; ++s) a = fmaf(d.w1[(size_t)s * d.C + c], dh[s], a);
    d.dmean[i] = a;
}
// parameter gradients: sums over the batch, one thread per element
__global__ void hn_se_fc_wgrad_kernel(const hn_sefc_desc d) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long n2 = (long long)d.C * d.S;
    const int T = d.C + d.S;
    if (i < n2) {  // dW2[c][s] = sum_n ds2[n][c] * h[n][s]
        const int c = (int)(i / d.S), s = (int)(i - (long long)c * d.S);
        float a = 0.0f;
        for (int n = 0; n < d.N; ++n) a = fmaf(d.tmp[(size_t)n * T + c], d.h[(size_t)n * d.S + s], a);
        d.dw2[i] = a;
    } else if (i < 2 * n2) {  // dW1[s][c] = sum_n dh[n][s] * mean[n][c]
        const long long k = i - n2;
        const int s = (int)(k / d.C), c = (int)(k - (long long)s * d.C);
        float a = 0.0f;
        for (int n = 0; n < d.N; ++n) a = fmaf(d.tmp[(size_t)n * T + d.C + s], d.mean[(size_t)n * d.C + c], a);
        d.dw1[k] = a;
    } else if (i < 2 * n2 + d.C) {
        const int c = (int)(i - 2 * n2);
        float a = 0.0f;
        for (int n = 0; n < d.N; ++n) a += d.tmp[(size_t)n * T + c];
        d.db2[c] = a;
    } else if (i < 2 * n2 + d.C + d.S) {
        const int s = (int)(i - 2 * n2 - d.C);
        float a = 0.0f;
        for (int n = 0; n < d.N; ++n) a += d.tmp[(size_t)n * T + d.C + s];
        d.db1[s] = a;
    }
}
extern "C" int hn_se_fc_fwd(const hn_sefc_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d && d->mean && d->w1 && d->b1 && d->w2 && d->b2 && d->h && d->gate, "se fc: null pointer");
    HN_REQUIRE(d->N >= 1 && d->C >= 1 && d->S >= 1 && (size_t)(d->C + d->S) * 4 <= 48 * 1024, "se fc: bad sizes N=%d C=%d S=%d", d->N, d->C, d->S);
    HN_REQUIRE((size_t)kSeU * d->C * 4 <= 48 * 1024 && (size_t)kSeCh * (d->S + 1) * 4 <= 48 * 1024, "se fc: layer too wide");
    hn_se_fc1_train_kernel<<<hn_cdiv(d->S, kSeU), 256, (size_t)kSeU * d->C * 4, stream>>>(*d);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_se_fc2_train_kernel<<<hn_cdiv(d->C, kSeCh), 256, (size_t)kSeCh * (d->S + 1) * 4, stream>>>(*d);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}
extern "C" int hn_se_fc_bwd(const hn_sefc_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(d && d->mean && d->w1 && d->w2 && d->h && d->gate && d->dgate && d->dmean && d->dw1 && d->db1 && d->dw2 && d->db2 && d->tmp,
               "se fc bwd: null pointer");
    HN_REQUIRE(d->N >= 1 && d->C >= 1 && d->S >= 1 && (size_t)(d->C + d->S) * 4 <= 48 * 1024, "se fc bwd: bad sizes");
    HN_REQUIRE((size_t)kSeU * d->C * 4 <= 48 * 1024, "se fc bwd: layer too wide");
    hn_se_ds2_kernel<<<hn_cdiv((long)d->N * d->C, 256), 256, 0, stream>>>(*d);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_se_dh_kernel<<<hn_cdiv(d->S, kSeU), 256, (size_t)kSeU * d->C * 4, stream>>>(*d);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_se_dmean_kernel<<<hn_cdiv((long)d->N * d->C, 256), 256, 0, stream>>>(*d);
    HN_CHECK_CUDA(cudaGetLastError());
    const long long total = 2LL * d->C * d->S + d->C + d->S;
    hn_se_fc_wgrad_kernel<<<hn_cdiv(total, 256), 256, 0, stream>>>(*d);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// weight packing: fp32 parameters -> bf16 K-major 64-column blocks, every layer in one launch
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hn_pack_kernel(const hn_pack_entry* __restrict__ entries) {
    const hn_pack_entry e = entries[blockIdx.x];
    const int r = blockIdx.y * 32 + (threadIdx.x >> 3);
    if (r >= e.rows_pad) return;
    const int j0 = (threadIdx.x & 7) * 8;
    float v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int j = j0 + q;
        float x = 0.0f;
        if (r < e.rows) {
            if (e.grouped == 0) {
                if (j < e.cols) x = e.src[(long long)r * e.s_r + (long long)j * e.s_c];
            } else if (e.grouped == 1) {  // forward: row = co, column = input channel within the 64-block; group = 8 channels
                if ((j >> 3) == ((r & 63) >> 3)) x = e.src[(long long)r * e.s_r + (long long)(j & 7) * e.s_c];
            } else {  // dgrad: row = ci, column = co within the 64-block
                const int co = (r & ~63) + j;
                if (co < e.rows && (j >> 3) == ((r & 63) >> 3)) x = e.src[(long long)co * e.s_r + (long long)(r & 7) * e.s_c];
            }
        }
        v[q] = x;
    }
    store8(reinterpret_cast<bf16*>(e.dst) + (long long)r * e.dst_ld + j0, v);
}
extern "C" int hn_pack_weights(const hn_pack_entry* entries_device, int32_t n, int32_t max_rows_pad, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(entries_device != nullptr && n >= 1 && max_rows_pad >= 1, "pack: bad arguments");
    HN_REQUIRE(hn_cdiv(max_rows_pad, 32) <= 65535, "pack: too many rows");
    hn_pack_kernel<<<dim3((unsigned)n, (unsigned)hn_cdiv(max_rows_pad, 32)), 256, 0, stream>>>(entries_device);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// Adam (torch.optim.Adam: L2 weight decay folded into the gradient, bias-corrected moments)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) hn_adam_kernel(const hn_adam_tensor* __restrict__ tensors, const int32_t* __restrict__ chunk_tensor,
                                                      const int32_t* __restrict__ chunk_index, int chunk_elems, float lr, float beta1, float beta2,
                                                      float eps, float wd, float bc1, float bc2_sqrt, float grad_scale, const float* __restrict__ dyn) {
    if (dyn) {  // hyper-parameters on the device: {lr, steps taken so far}
        lr = dyn[0];
        const float t = dyn[1] + 1.0f;
        bc1 = 1.0f - powf(beta1, t);
        bc2_sqrt = sqrtf(1.0f - powf(beta2, t));
    }
    const hn_adam_tensor t = tensors[chunk_tensor[blockIdx.x]];
    const long long b = (long long)chunk_index[blockIdx.x] * chunk_elems;
    const long long e = min(b + chunk_elems, (long long)t.n);
    const float step_size = lr / bc1;
    for (long long i = b + threadIdx.x; i < e; i += 256) {
        const float p = t.p[i];
        float g = t.g[i] * grad_scale;
        g = fmaf(wd, p, g);
        const float m = beta1 * t.m[i] + (1.0f - beta1) * g;
        const float v = beta2 * t.v[i] + (1.0f - beta2) * g * g;
        t.m[i] = m;
        t.v[i] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        t.p[i] = p - step_size * (m / denom);
    }
}
__global__ void hn_adam_advance_kernel(float* dyn) { dyn[1] += 1.0f; }
extern "C" int hn_adam_step(const hn_adam_tensor* tensors_device, const int32_t* chunk_tensor_device, const int32_t* chunk_index_device,
                            int32_t n_chunks, int32_t chunk_elems, float lr, float beta1, float beta2, float eps, float weight_decay, int32_t step,
                            float grad_scale, float* dyn_device, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(tensors_device && chunk_tensor_device && chunk_index_device && n_chunks >= 1 && chunk_elems >= 256 && (step >= 1 || dyn_device),
               "adam: bad arguments");
    const double bc1 = 1.0 - pow((double)beta1, (double)(step > 0 ? step : 1)), bc2 = 1.0 - pow((double)beta2, (double)(step > 0 ? step : 1));
    hn_adam_kernel<<<n_chunks, 256, 0, stream>>>(tensors_device, chunk_tensor_device, chunk_index_device, chunk_elems, lr, beta1, beta2, eps, weight_decay,
                                                 (float)bc1, (float)sqrt(bc2), grad_scale, dyn_device);
    HN_CHECK_CUDA(cudaGetLastError());
    if (dyn_device) {
        hn_adam_advance_kernel<<<1, 1, 0, stream>>>(dyn_device);
        HN_CHECK_CUDA(cudaGetLastError());
    }
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// Detection loss (reference head_detect/detection_loss.py:111-267, FocalLoss.forward):
// IoU anchor assignment, focal classification loss and smooth-L1 box loss, forward AND gradient in two launches.
//   kernel 1 (assign): per (image, anchor) IoU against the image's valid boxes -> best box (first maximum, as torch.max),
//                      state {ignore, negative, positive + box}, positives counted per image;
//   kernel 2 (loss):   per (image, anchor) the focal terms of its K classes and, for positives, the smooth-L1 of its 4 offsets;
//                      per-image sums in double (order-insensitive to ~1e-16) and the gradients w.r.t. classification / regression
//                      of  mean_b(cls_loss_b)  and  mean_b(reg_loss_b)  written directly.
// Annotations [B][M][5] = (x1, y1, x2, y2, class), class -1 = padding; anchors [A][4] = (y1, x1, y2, x2).
// ------------------------------------------------------------------------------------------------
struct DetLossParams {
    const float *cls, *reg, *anchors, *ann;
    int B, A, K, M;
    float alpha, gamma;
    int* assign;     // [B][A]: -2 ignore, -1 negative, >= 0 index of the assigned box
    int* num_pos;    // [B]
    int* has_box;    // [B]
    double* sums;    // [B][2]: classification, regression (un-normalised)
    float *dcls, *dreg;
    float *cls_loss, *reg_loss;  // [B] per-image losses (normalised as the reference does)
};
__global__ void __launch_bounds__(256) hn_det_assign_kernel(const DetLossParams p) {
    const int b = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float s_ann[64 * 5];
    __shared__ int s_pos;
    if (threadIdx.x == 0) s_pos = 0;
    for (int i = threadIdx.x; i < p.M * 5; i += blockDim.x) s_ann[i] = p.ann[(size_t)b * p.M * 5 + i];
    __syncthreads();
    int any = 0, positive = 0;
    if (a < p.A) {
        const float ay1 = p.anchors[a * 4], ax1 = p.anchors[a * 4 + 1], ay2 = p.anchors[a * 4 + 2], ax2 = p.anchors[a * 4 + 3];
        const float aarea = __fmul_rn(__fsub_rn(ay2, ay1), __fsub_rn(ax2, ax1));
        float best = -1.0f;
        int arg = 0;
        for (int m = 0; m < p.M; ++m) {
            const float* g = s_ann + m * 5;
            if (g[4] == -1.0f) continue;
            any = 1;
            const float area = __fmul_rn(__fsub_rn(g[2], g[0]), __fsub_rn(g[3], g[1]));
            const float iw = fmaxf(__fsub_rn(fminf(ax2, g[2]), fmaxf(ax1, g[0])), 0.0f);
            const float ih = fmaxf(__fsub_rn(fminf(ay2, g[3]), fmaxf(ay1, g[1])), 0.0f);
            const float inter = __fmul_rn(iw, ih);
            const float ua = fmaxf(__fsub_rn(__fadd_rn(aarea, area), inter), 1e-8f);
            const float iou = __fdiv_rn(inter, ua);
            if (iou > best) { best = iou; arg = m; }
        }
        int st = -1;  // no valid box: every anchor is a negative
        if (any) st = best >= 0.5f ? arg : (best < 0.4f ? -1 : -2);
        positive = st >= 0;
        p.assign[(size_t)b * p.A + a] = st;
    }
    if (positive) atomicAdd(&s_pos, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_pos) atomicAdd(p.num_pos + b, s_pos);
        if (blockIdx.x == 0) {
            int hb = 0;
            for (int m = 0; m < p.M; ++m) hb |= s_ann[m * 5 + 4] != -1.0f;
            p.has_box[b] = hb;
        }
    }
}
__global__ void __launch_bounds__(256) hn_det_loss_kernel(const DetLossParams p) {
    const int b = blockIdx.y;
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ double s_sum[2][8];
    double lc = 0.0, lr = 0.0;
    if (a < p.A) {
        const int st = p.assign[(size_t)b * p.A + a];
        const int npos = p.num_pos[b];
        const bool hb = p.has_box[b] != 0;
        const float cnorm = hb ? 1.0f / fmaxf((float)npos, 1.0f) : 1.0f;  // images without boxes are NOT normalised (detection_loss.py:138-160)
        const float invB = 1.0f / (float)p.B;
        const float* c = p.cls + ((size_t)b * p.A + a) * p.K;
        float* dc = p.dcls + ((size_t)b * p.A + a) * p.K;
        const float* g = p.ann + ((size_t)b * p.M + (st >= 0 ? st : 0)) * 5;
        const int tcls = st >= 0 ? (int)g[4] : -1;
        for (int k = 0; k < p.K; ++k) {
            float grad = 0.0f;
            if (st != -2) {
                const float x = c[k];
                const float v = fminf(fmaxf(x, 1e-4f), 1.0f - 1e-4f);
                const bool pass = x >= 1e-4f && x <= 1.0f - 1e-4f;  // clamp passes the gradient inside [min, max]
                float l, dl;
                if (k == tcls) {  // target 1: -alpha (1 - v)^gamma log v
                    const float w = powf(1.0f - v, p.gamma), lg = logf(v);
                    l = -p.alpha * w * lg;
                    dl = p.alpha * (p.gamma * powf(1.0f - v, p.gamma - 1.0f) * lg - w / v);
                } else {          // target 0: -(1 - alpha) v^gamma log(1 - v)
                    const float w = powf(v, p.gamma), lg = logf(1.0f - v);
                    l = -(1.0f - p.alpha) * w * lg;
                    dl = (1.0f - p.alpha) * (-p.gamma * powf(v, p.gamma - 1.0f) * lg + w / (1.0f - v));
                }
                lc += (double)l;
                grad = pass ? dl * cnorm * invB : 0.0f;
            }
            dc[k] = grad;
        }
        float* dr = p.dreg + ((size_t)b * p.A + a) * 4;
        if (st >= 0) {
            const float ay1 = p.anchors[a * 4], ax1 = p.anchors[a * 4 + 1], ay2 = p.anchors[a * 4 + 2], ax2 = p.anchors[a * 4 + 3];
            const float aw = ax2 - ax1, ah = ay2 - ay1, acx = ax1 + 0.5f * aw, acy = ay1 + 0.5f * ah;
            float gw = g[2] - g[0], gh = g[3] - g[1];
            const float gcx = g[0] + 0.5f * gw, gcy = g[1] + 0.5f * gh;
            gw = fmaxf(gw, 1.0f);
            gh = fmaxf(gh, 1.0f);
            const float t[4] = {(gcy - acy) / ah, (gcx - acx) / aw, logf(gh / ah), logf(gw / aw)};  // (dy, dx, dh, dw)
            const float* r = p.reg + ((size_t)b * p.A + a) * 4;
            const float rnorm = invB / (4.0f * fmaxf((float)npos, 1.0f));
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = t[j] - r[j], ad = fabsf(d);
                float l, dl;  // d/d(diff)
                if (ad <= 1.0f / 9.0f) { l = 0.5f * 9.0f * ad * ad; dl = 9.0f * ad; }
                else { l = ad - 0.5f / 9.0f; dl = 1.0f; }
                lr += (double)l;
                const float sgn = d > 0.0f ? 1.0f : (d < 0.0f ? -1.0f : 0.0f);
                dr[j] = -sgn * dl * rnorm;  // d|t - r|/dr = -sign(t - r)
            }
        } else {
            dr[0] = dr[1] = dr[2] = dr[3] = 0.0f;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        lc += __shfl_xor_sync(0xffffffffu, lc, o);
        lr += __shfl_xor_sync(0xffffffffu, lr, o);
    }
    if ((threadIdx.x & 31) == 0) { s_sum[0][threadIdx.x >> 5] = lc; s_sum[1][threadIdx.x >> 5] = lr; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a0 = 0.0, a1 = 0.0;
        for (int w = 0; w < 8; ++w) { a0 += s_sum[0][w]; a1 += s_sum[1][w]; }
        atomicAdd(p.sums + b * 2, a0);
        atomicAdd(p.sums + b * 2 + 1, a1);
    }
}
__global__ void hn_det_loss_finish_kernel(const DetLossParams p) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.B) return;
    const int npos = p.num_pos[b];
    const bool hb = p.has_box[b] != 0;
    p.cls_loss[b] = hb ? (float)(p.sums[b * 2] / fmax((double)npos, 1.0)) : (float)p.sums[b * 2];
    p.reg_loss[b] = (hb && npos > 0) ? (float)(p.sums[b * 2 + 1] / (4.0 * (double)npos)) : 0.0f;
}
extern "C" int hn_det_loss(const float* classification, const float* regression, const float* anchors, const float* annotations, int32_t B, int32_t A,
                           int32_t K, int32_t M, float alpha, float gamma, void* workspace, int64_t workspace_bytes, float* cls_loss, float* reg_loss,
                           float* dcls, float* dreg, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(classification && regression && anchors && annotations && workspace && cls_loss && reg_loss && dcls && dreg, "det loss: null pointer");
    HN_REQUIRE(B >= 1 && A >= 1 && K >= 1 && M >= 1 && M <= 64, "det loss: bad sizes (B=%d A=%d K=%d M=%d; at most 64 boxes per image)", B, A, K, M);
    const size_t need = (size_t)B * A * 4 + (size_t)B * 8 + (size_t)B * 16 + 64;
    HN_REQUIRE((int64_t)need <= workspace_bytes, "det loss: workspace too small (%zu > %lld)", need, (long long)workspace_bytes);
    DetLossParams p;
    p.cls = classification; p.reg = regression; p.anchors = anchors; p.ann = annotations;
    p.B = B; p.A = A; p.K = K; p.M = M; p.alpha = alpha; p.gamma = gamma;
    uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
    p.sums = reinterpret_cast<double*>(w);          w += (size_t)B * 16;
    p.num_pos = reinterpret_cast<int*>(w);          w += (size_t)B * 4;
    p.has_box = reinterpret_cast<int*>(w);          w += (size_t)B * 4;
    p.assign = reinterpret_cast<int*>(w);
    p.dcls = dcls; p.dreg = dreg; p.cls_loss = cls_loss; p.reg_loss = reg_loss;
    HN_CHECK_CUDA(cudaMemsetAsync(workspace, 0, (size_t)B * 24, stream));
    dim3 grid((unsigned)hn_cdiv(A, 256), (unsigned)B);
    hn_det_assign_kernel<<<grid, 256, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_det_loss_kernel<<<grid, 256, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_det_loss_finish_kernel<<<hn_cdiv(B, 128), 128, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// segmentation loss (head_segment/segmentation_loss.py:48-65): class-weighted cross-entropy per
// pixel, the k largest per image, mean over the N*k kept values -- and its gradient
// ------------------------------------------------------------------------------------------------
// Top-k without a sort: the k-th largest loss of an image is found by a 3-pass radix select on the fp32 bit pattern (losses are
// >= 0, so the unsigned pattern orders like the value): per pass a shared-memory histogram of 11 / 11 / 10 key bits of the
// elements that match the prefix found so far, then one CTA per image picks the bin that holds the k-th element.  After the last
// pass tau[b] is exact, `need` = how many elements EQUAL to tau belong to the top k and n_eq = how many there are; the sum of the
// kept values is sum(l > tau) + need * tau.  Ties at tau share their weight (need / n_eq each) in the gradient: torch.topk keeps
// an unspecified subset of them, any choice is a valid subgradient, and the loss value is the same.
static constexpr int kSegPx = 4096;   // pixels per CTA in the per-image passes
static constexpr int kSelBins = 2048;
struct SegSel {  // per image
    unsigned prefix, mask;
    long long k_rem;  // elements still to take from the matching set
    float tau, frac;
    long long n_eq;
};
struct SegLossParams {
    const float* logits;
    const long long* target;
    const float* weight;
    int N, C;
    long long HW, k;
    int ignore;
    float* loss_px;
    unsigned* hist;    // [N][kSelBins]
    SegSel* sel;       // [N]
    double* partial;   // [N][chunks]
    int chunks;
    float* loss;       // [1]
    float* state;      // [N][2] = tau, frac (kept for the backward)
};
__global__ void __launch_bounds__(256) hn_segce_kernel(const SegLossParams p) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= p.HW) return;
    if (i == 0) {
        SegSel s;
        s.prefix = 0u; s.mask = 0u; s.k_rem = p.k; s.tau = 0.0f; s.frac = 0.0f; s.n_eq = 0;
        p.sel[b] = s;
    }
    const long long t = p.target[(long long)b * p.HW + i];
    float l = 0.0f;
    if (t != p.ignore && t >= 0 && t < p.C) {
        const float* x = p.logits + (long long)b * p.C * p.HW + i;
        float m = x[0];
        for (int c = 1; c < p.C; ++c) m = fmaxf(m, x[(long long)c * p.HW]);
        float se = 0.0f;
        for (int c = 0; c < p.C; ++c) se += expf(x[(long long)c * p.HW] - m);
        l = p.weight[t] * (m + logf(se) - x[t * p.HW]);
        l = fmaxf(l, 0.0f);  // -0.0 / rounding below zero would break the unsigned key order
    }
    p.loss_px[(long long)b * p.HW + i] = l;
}
__global__ void __launch_bounds__(256) hn_segsel_hist_kernel(const SegLossParams p, int shift, int bits) {
    __shared__ unsigned sh[kSelBins];
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const SegSel s = p.sel[b];
    const long long i0 = (long long)blockIdx.x * kSegPx, i1 = min(i0 + kSegPx, p.HW);
    const float* l = p.loss_px + (long long)b * p.HW;
    const unsigned bmask = (1u << bits) - 1u;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const unsigned key = __float_as_uint(l[i]);
        if ((key & s.mask) == s.prefix) atomicAdd(&sh[(key >> shift) & bmask], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSelBins; i += blockDim.x)
        if (sh[i]) atomicAdd(p.hist + (long long)b * kSelBins + i, sh[i]);
}
// one CTA of 256 threads per image: the bin (from the top) in which the running count reaches k_rem
__global__ void __launch_bounds__(256) hn_segsel_pick_kernel(const SegLossParams p, int shift, int bits, int last) {
    __shared__ unsigned long long tot[256];
    const int b = blockIdx.x, t = threadIdx.x;
    unsigned* h = p.hist + (long long)b * kSelBins;
    unsigned v[8];
    unsigned long long mine = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) { v[j] = h[t * 8 + j]; mine += v[j]; h[t * 8 + j] = 0u; }  // zeroed for the next pass
    tot[t] = mine;
    __syncthreads();
    unsigned long long above = 0;
    for (int u = t + 1; u < 256; ++u) above += tot[u];
    SegSel s = p.sel[b];
    const unsigned long long k = (unsigned long long)s.k_rem;
    __syncthreads();
    if (above < k && k <= above + mine) {  // exactly one thread (k >= 1, k <= matching elements)
        unsigned long long run = above;
        int bin = t * 8 + 7;
        for (int j = 7; j >= 0; --j) {
            if (run + v[j] >= k) { bin = t * 8 + j; break; }
            run += v[j];
        }
        s.prefix |= (unsigned)bin << shift;
        s.mask |= ((1u << bits) - 1u) << shift;
        s.k_rem = (long long)(k - run);
        if (last) {
            s.tau = __uint_as_float(s.prefix);
            s.n_eq = (long long)v[bin - t * 8];
            s.frac = (float)((double)s.k_rem / (double)s.n_eq);
            p.state[b * 2] = s.tau;
            p.state[b * 2 + 1] = s.frac;
        }
        p.sel[b] = s;
    }
}
__global__ void __launch_bounds__(256) hn_segsel_sum_kernel(const SegLossParams p) {
    __shared__ double sh[256];
    const int b = blockIdx.y;
    const float tau = p.sel[b].tau;
    const long long i0 = (long long)blockIdx.x * kSegPx, i1 = min(i0 + kSegPx, p.HW);
    const float* l = p.loss_px + (long long)b * p.HW;
    double a = 0.0;
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const float v = l[i];
        if (v > tau) a += (double)v;
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {  // fixed tree: deterministic
        if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) p.partial[(long long)b * p.chunks + blockIdx.x] = sh[0];
}
__global__ void hn_segsel_finish_kernel(const SegLossParams p) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    double total = 0.0;
    for (int b = 0; b < p.N; ++b) {
        double a = 0.0;
        for (int c = 0; c < p.chunks; ++c) a += p.partial[(long long)b * p.chunks + c];
        total += a + (double)p.sel[b].k_rem * (double)p.sel[b].tau;
    }
    p.loss[0] = (float)(total / ((double)p.N * (double)p.k));
}
__global__ void __launch_bounds__(256) hn_segce_bwd_kernel(const SegLossParams p, const float* __restrict__ gout, float* __restrict__ dlogits) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (i >= p.HW) return;
    const float l = p.loss_px[(long long)b * p.HW + i];
    const float tau = p.state[b * 2], frac = p.state[b * 2 + 1];
    const long long t = p.target[(long long)b * p.HW + i];
    float sel = l > tau ? 1.0f : (l == tau ? frac : 0.0f);
    if (t == p.ignore || t < 0 || t >= p.C) sel = 0.0f;
    float* d = dlogits + (long long)b * p.C * p.HW + i;
    if (sel == 0.0f) {
        for (int c = 0; c < p.C; ++c) d[(long long)c * p.HW] = 0.0f;
        return;
    }
    const float* x = p.logits + (long long)b * p.C * p.HW + i;
    float m = x[0];
    for (int c = 1; c < p.C; ++c) m = fmaxf(m, x[(long long)c * p.HW]);
    float se = 0.0f;
    for (int c = 0; c < p.C; ++c) se += expf(x[(long long)c * p.HW] - m);
    const float scale = sel * p.weight[t] * gout[0] / ((float)p.N * (float)p.k), inv = 1.0f / se;
    for (int c = 0; c < p.C; ++c) {
        const float pr = expf(x[(long long)c * p.HW] - m) * inv;
        d[(long long)c * p.HW] = scale * (pr - (c == t ? 1.0f : 0.0f));
    }
}
static size_t seg_loss_layout(int N, long long HW, SegLossParams* p, void* base) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        void* q = base ? static_cast<char*>(base) + off : nullptr;
        off += (bytes + 255) & ~size_t(255);
        return q;
    };
    const int chunks = (int)((HW + kSegPx - 1) / kSegPx);
    unsigned* hist = static_cast<unsigned*>(take((size_t)N * kSelBins * sizeof(unsigned)));
    SegSel* sel = static_cast<SegSel*>(take((size_t)N * sizeof(SegSel)));
    double* partial = static_cast<double*>(take((size_t)N * chunks * sizeof(double)));
    float* loss_px = static_cast<float*>(take((size_t)N * HW * sizeof(float)));
    float* state = static_cast<float*>(take((size_t)N * 2 * sizeof(float)));
    if (p) { p->hist = hist; p->sel = sel; p->partial = partial; p->loss_px = loss_px; p->state = state; p->chunks = chunks; }
    return off;
}
extern "C" int64_t hn_seg_loss_workspace_bytes(int32_t N, int64_t HW) { return (int64_t)seg_loss_layout(N, HW, nullptr, nullptr); }
static int seg_loss_params(const hn_segloss_desc* d, SegLossParams* p) {
    HN_REQUIRE(d != nullptr && d->logits && d->target && d->weight && d->workspace, "seg loss: null pointer");
    HN_REQUIRE(d->N >= 1 && d->C >= 1 && d->C <= 64 && d->HW >= 1 && d->k >= 1 && d->k <= d->HW, "seg loss: bad sizes (N=%d C=%d HW=%lld k=%lld)", d->N,
               d->C, (long long)d->HW, (long long)d->k);
    HN_REQUIRE((int64_t)seg_loss_layout(d->N, d->HW, nullptr, nullptr) <= d->workspace_bytes, "seg loss: workspace too small");
    memset(p, 0, sizeof(*p));
    p->logits = d->logits; p->target = reinterpret_cast<const long long*>(d->target); p->weight = d->weight;
    p->N = d->N; p->C = d->C; p->HW = d->HW; p->k = d->k; p->ignore = d->ignore_index;
    seg_loss_layout(d->N, d->HW, p, d->workspace);
    return HN_OK;
}
extern "C" int hn_seg_loss_fwd(const hn_segloss_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    SegLossParams p;
    if (int rc = seg_loss_params(d, &p)) return rc;
    HN_REQUIRE(d->loss != nullptr, "seg loss: null output");
    p.loss = d->loss;
    const dim3 gpx((unsigned)hn_cdiv(p.HW, 256), (unsigned)p.N), gch((unsigned)p.chunks, (unsigned)p.N);
    HN_CHECK_CUDA(cudaMemsetAsync(p.hist, 0, (size_t)p.N * kSelBins * sizeof(unsigned), stream));
    hn_segce_kernel<<<gpx, 256, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    const int shifts[3] = {21, 10, 0}, bits[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        hn_segsel_hist_kernel<<<gch, 256, 0, stream>>>(p, shifts[pass], bits[pass]);
        HN_CHECK_CUDA(cudaGetLastError());
        hn_segsel_pick_kernel<<<p.N, 256, 0, stream>>>(p, shifts[pass], bits[pass], pass == 2);
        HN_CHECK_CUDA(cudaGetLastError());
    }
    hn_segsel_sum_kernel<<<gch, 256, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_segsel_finish_kernel<<<1, 32, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}
extern "C" int hn_seg_loss_bwd(const hn_segloss_desc* d, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    SegLossParams p;
    if (int rc = seg_loss_params(d, &p)) return rc;
    HN_REQUIRE(d->gout != nullptr && d->dlogits != nullptr, "seg loss bwd: null pointer");
    const dim3 gpx((unsigned)hn_cdiv(p.HW, 256), (unsigned)p.N);
    hn_segce_bwd_kernel<<<gpx, 256, 0, stream>>>(p, d->gout, d->dlogits);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// lane losses (head_lane/lanedetect_loss.py:18-78): two-class log-softmax with online hard-negative
// mining (the negative_num negatives with the smallest background log-probability, ties included) and the masked Huber
// regression loss over the positive anchors -- values and gradients
// ------------------------------------------------------------------------------------------------
struct LaneLossParams {
    const float* cls_t;   // [T][2]  (column 1 > 0: positive)
    const float* cls_p;   // [T][2]
    const float* loc_t;   // [T][L]
    const float* loc_p;   // [T][L]
    int T, L, wpos;       // entries wpos, wpos + 1 of a row carry the weight `alpha`
    float neg_ratio, alpha;
    float* out;           // [4] = total_pos, total_neg, loc, positive_num
    float* dcls;          // [2][T][2]: d total_pos / d cls_p, d total_neg / d cls_p
    float* dloc;          // [T][L]
    float* per_anchor;    // [T] scratch
    unsigned char* pmask; // [T] scratch
};
__device__ __forceinline__ unsigned asc_key(float f) {
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ void lane_logp(const float* x, float& l0, float& l1) {
    const float m = fmaxf(x[0], x[1]);
    const float lse = m + logf(expf(x[0] - m) + expf(x[1] - m));
    l0 = x[0] - lse;
    l1 = x[1] - lse;
}
// deterministic block sums (fixed tree) of up to two doubles / two ints
__device__ __forceinline__ double lane_block_sum(double v, double* sh) {
    __syncthreads();
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    return sh[0];
}
__global__ void __launch_bounds__(1024) hn_lane_cls_loss_kernel(const LaneLossParams p) {
    __shared__ double shd[1024];
    __shared__ unsigned hist[256];
    __shared__ unsigned s_prefix, s_mask;
    __shared__ long long s_krem;
    const int T = p.T;
    double np = 0.0;
    float* bg = p.per_anchor;  // the background log-probabilities, computed ONCE: the select and the final test must see the same bits
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const bool pos = p.cls_t[i * 2 + 1] > 0.0f;
        p.pmask[i] = pos ? 1 : 0;
        np += pos ? 1.0 : 0.0;
        float l0, l1;
        lane_logp(p.cls_p + i * 2, l0, l1);
        bg[i] = l0;
    }
    const double n_pos = lane_block_sum(np, shd), n_neg = (double)T - n_pos;
    const double pos_num = fmax(n_pos, 1.0);
    long long neg_num = (long long)fmin(fmax(n_pos * (double)p.neg_ratio, 1.0), n_neg);
    // k-th smallest background log-probability among the negatives: radix select on an order-preserving key, 4 x 8 bits
    if (threadIdx.x == 0) { s_prefix = 0u; s_mask = 0u; s_krem = neg_num; }
    float kth = -INFINITY;
    if (neg_num >= 1) {
        for (int pass = 0; pass < 4; ++pass) {
            const int shift = 24 - 8 * pass;
            if (threadIdx.x < 256) hist[threadIdx.x] = 0u;
            __syncthreads();
            const unsigned prefix = s_prefix, mask = s_mask;
            for (int i = threadIdx.x; i < T; i += blockDim.x) {
                if (p.cls_t[i * 2 + 1] > 0.0f) continue;
                const unsigned key = asc_key(bg[i]);
                if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                unsigned long long run = 0;
                const unsigned long long k = (unsigned long long)s_krem;
                int bin = 255;
                for (int b = 0; b < 256; ++b) {
                    if (run + hist[b] >= k) { bin = b; break; }
                    run += hist[b];
                }
                s_prefix = prefix | ((unsigned)bin << shift);
                s_mask = mask | (255u << shift);
                s_krem = (long long)(k - run);
            }
            __syncthreads();
        }
        const unsigned key = s_prefix;  // invert asc_key
        kth = __uint_as_float((key & 0x80000000u) ? (key & 0x7FFFFFFFu) : ~key);
    }
    double sp = 0.0, sn = 0.0;
    const float cp = -p.alpha / (float)pos_num;
    for (int i = threadIdx.x; i < T; i += blockDim.x) {
        float l0, l1;
        lane_logp(p.cls_p + i * 2, l0, l1);
        const bool pos = p.cls_t[i * 2 + 1] > 0.0f;
        const bool hard = !pos && bg[i] <= kth;
        const float p0 = expf(l0), p1 = expf(l1);
        // d logp1 / dx = (-p0, 1 - p1); d logp0 / dx = (1 - p0, -p1)
        p.dcls[i * 2] = pos ? cp * -p0 : 0.0f;
        p.dcls[i * 2 + 1] = pos ? cp * (1.0f - p1) : 0.0f;
        p.dcls[(T + i) * 2] = hard ? cp * (1.0f - p0) : 0.0f;
        p.dcls[(T + i) * 2 + 1] = hard ? cp * -p1 : 0.0f;
        if (pos) sp += (double)(p.alpha * l1);
        if (hard) sn += (double)(p.alpha * l0);
    }
    const double tp = lane_block_sum(sp, shd);
    const double tn = lane_block_sum(sn, shd);
    if (threadIdx.x == 0) {
        p.out[0] = (float)(-tp / pos_num);
        p.out[1] = (float)(-tn / pos_num);
        p.out[3] = (float)pos_num;
    }
}
// one warp per anchor
__global__ void __launch_bounds__(256) hn_lane_loc_loss_kernel(const LaneLossParams p) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= p.T) return;
    const float* t = p.loc_t + (long long)i * p.L;
    const float* q = p.loc_p + (long long)i * p.L;
    float* d = p.dloc + (long long)i * p.L;
    if (!p.pmask[i]) {
        for (int j = lane; j < p.L; j += 32) d[j] = 0.0f;
        if (lane == 0) p.per_anchor[i] = 0.0f;
        return;
    }
    float s = 0.0f, nv = 0.0f;
    for (int j = lane; j < p.L; j += 32) {
        const float tv = t[j];
        if (tv != 0.0f) {
            const float w = (j == p.wpos || j == p.wpos + 1) ? p.alpha : 1.0f;
            const float e = q[j] - tv, ae = fabsf(e);
            s += w * (ae < 1.0f ? e * e * 0.5f : ae - 0.5f);
            nv += 1.0f;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        nv += __shfl_xor_sync(0xffffffffu, nv, o);
    }
    const float inv = 1.0f / fmaxf(nv, 1.0f), pn = p.out[3];
    if (lane == 0) p.per_anchor[i] = s * inv;
    for (int j = lane; j < p.L; j += 32) {
        const float tv = t[j];
        float g = 0.0f;
        if (tv != 0.0f) {
            const float w = (j == p.wpos || j == p.wpos + 1) ? p.alpha : 1.0f;
            const float e = q[j] - tv;
            g = w * fminf(fmaxf(e, -1.0f), 1.0f) * inv / pn;
        }
        d[j] = g;
    }
}
__global__ void __launch_bounds__(1024) hn_lane_loc_finish_kernel(const LaneLossParams p) {
    __shared__ double shd[1024];
    double a = 0.0;
    for (int i = threadIdx.x; i < p.T; i += blockDim.x) a += (double)p.per_anchor[i];
    const double tot = lane_block_sum(a, shd);
    if (threadIdx.x == 0) p.out[2] = (float)(tot / (double)p.out[3]);
}
extern "C" int64_t hn_lane_loss_workspace_bytes(int32_t T) { return (int64_t)T * 4 + (((int64_t)T + 255) & ~255LL) + 256; }
extern "C" int hn_lane_loss(const float* cls_targets, const float* cls_preds, const float* loc_targets, const float* loc_preds, int32_t T, int32_t L,
                            int32_t weighted_index, float negative_ratio, float alpha, void* workspace, int64_t workspace_bytes, float* out4, float* dcls,
                            float* dloc, void* stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    HN_REQUIRE(cls_targets && cls_preds && loc_targets && loc_preds && workspace && out4 && dcls && dloc, "lane loss: null pointer");
    HN_REQUIRE(T >= 1 && T <= (1 << 24) && L >= 1 && weighted_index >= 0 && weighted_index + 1 < L, "lane loss: bad sizes (T=%d L=%d)", T, L);
    HN_REQUIRE(workspace_bytes >= hn_lane_loss_workspace_bytes(T), "lane loss: workspace too small");
    LaneLossParams p;
    memset(&p, 0, sizeof(p));
    p.cls_t = cls_targets; p.cls_p = cls_preds; p.loc_t = loc_targets; p.loc_p = loc_preds;
    p.T = T; p.L = L; p.wpos = weighted_index; p.neg_ratio = negative_ratio; p.alpha = alpha;
    p.out = out4; p.dcls = dcls; p.dloc = dloc;
    p.per_anchor = static_cast<float*>(workspace);
    p.pmask = static_cast<unsigned char*>(workspace) + (size_t)T * 4;
    hn_lane_cls_loss_kernel<<<1, 1024, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_lane_loc_loss_kernel<<<hn_cdiv(T, 8), 256, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_lane_loc_finish_kernel<<<1, 1024, 0, stream>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}
