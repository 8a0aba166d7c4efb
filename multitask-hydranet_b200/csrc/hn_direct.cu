// HBM-bound operators of the HydraNet forward: stem, BiFPN fusion node (+depthwise), max-pools,
// lane multi-scale fuse, squeeze-excite.  All activations are NHWC bf16 views; threads own 8-channel
// (16-byte) vectors so every global access is a coalesced 128-bit transaction.
#include <mutex>

#include "hn_ops.h"

// ------------------------------------------------------------------------------------------------
// stem: 3x3 s2 p1, 3 -> 32, fp32 NCHW -> bf16 NHWC, BN folded, ReLU
// ------------------------------------------------------------------------------------------------
// One thread = two horizontally adjacent output pixels x 32 channels: the kernel is bound by the shared-memory reads
// of the weights (8 LDS.128 per tap), and two pixels share every weight vector (and 9 of their 54 input samples).
// The accumulation order per output (ci, ky, kx) is the same as a one-pixel loop.
__global__ void __launch_bounds__(128) hn_stem_kernel(const float* __restrict__ x, int N, int H, int W,
                                                      const float* __restrict__ w, const float* __restrict__ b,
                                                      View out, int no_relu) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    __shared__ float sw[27 * 32];
    __shared__ float sb[32];
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 32) sb[threadIdx.x] = b[threadIdx.x];
    __syncthreads();
    const int OH = out.H, OW = out.W, OW2 = (OW + 1) >> 1;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned total = (unsigned)N * OH * OW2;
    if (idx >= total) return;
    const int ox = (int)(idx % (unsigned)OW2) * 2;
    const int oy = (int)((idx / (unsigned)OW2) % (unsigned)OH);
    const int n = (int)(idx / ((unsigned)OW2 * OH));
    float acc[2][32];
#pragma unroll
    for (int c = 0; c < 32; ++c) acc[0][c] = acc[1][c] = sb[c];
    const float* xin = x + (long long)n * 3 * H * W;
#pragma unroll
    for (int ci = 0; ci < 3; ++ci) {
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * 2 - 1 + ky;
            const bool row_ok = iy >= 0 && iy < H;
            const float* rowp = xin + ((long long)ci * H + (row_ok ? iy : 0)) * W;
            float v[5];  // input columns 2*ox-1 .. 2*ox+3
#pragma unroll
            for (int q = 0; q < 5; ++q) {
                const int ix = ox * 2 - 1 + q;
                v[q] = (row_ok && ix >= 0 && ix < W) ? __ldg(rowp + ix) : 0.0f;
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const float4* wr = reinterpret_cast<const float4*>(sw + (ci * 9 + ky * 3 + kx) * 32);
                const float v0 = v[kx], v1 = v[kx + 2];
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4) {
                    const float4 ww = wr[c4];
                    acc[0][c4 * 4 + 0] = fmaf(v0, ww.x, acc[0][c4 * 4 + 0]);
                    acc[0][c4 * 4 + 1] = fmaf(v0, ww.y, acc[0][c4 * 4 + 1]);
                    acc[0][c4 * 4 + 2] = fmaf(v0, ww.z, acc[0][c4 * 4 + 2]);
                    acc[0][c4 * 4 + 3] = fmaf(v0, ww.w, acc[0][c4 * 4 + 3]);
                    acc[1][c4 * 4 + 0] = fmaf(v1, ww.x, acc[1][c4 * 4 + 0]);
                    acc[1][c4 * 4 + 1] = fmaf(v1, ww.y, acc[1][c4 * 4 + 1]);
                    acc[1][c4 * 4 + 2] = fmaf(v1, ww.z, acc[1][c4 * 4 + 2]);
                    acc[1][c4 * 4 + 3] = fmaf(v1, ww.w, acc[1][c4 * 4 + 3]);
                }
            }
        }
    }
#pragma unroll
    for (int px = 0; px < 2; ++px) {
        if (ox + px >= OW) break;
        bf16* o = const_cast<bf16*>(vptr(out, n, oy, ox + px, 0));
#pragma unroll
        for (int c8 = 0; c8 < 4; ++c8) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = no_relu ? acc[px][c8 * 8 + j] : fmaxf(acc[px][c8 * 8 + j], 0.0f);
            store8(o + c8 * 8, f);
        }
    }
}

// The same convolution on the tensor cores.  The fp32 kernel above is bound by its FMAs (864 per output pixel: 221 us at batch 32,
// a quarter of the HBM rate).  Here a warp owns 16 consecutive output pixels of a row: the 27 taps are the K dimension (padded to
// 32, k = ci*9 + ky*3 + kx) of a 16 x 32 x 32 product in mma.sync m16n8k16.  Inputs and weights are split into bf16 (hi, lo)
// pairs and three products are accumulated (hi*hi + lo*hi + hi*lo, fp32 accumulators): the result carries ~16 mantissa bits, i.e.
// fp32-conv accuracy at the bf16 output -- the first layer adds no rounding of its own.  A thread gathers exactly the 16 input
// samples its A fragments need (its 8 k indices x 2 pixel rows); the output tile goes through a padded shared-memory tile so that
// the global stores are 16-byte vectors of one pixel's channels.
__device__ __forceinline__ void stem_mma(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    hi = hn_pack_bf16x2(v0, v1);
    const float2 h = hn_unpack_bf16x2(hi);
    lo = hn_pack_bf16x2(v0 - h.x, v1 - h.y);
}
static constexpr int kStemRowWords = 20;  // 16 words of a pixel's 32 bf16 channels + 4 of padding: conflict-free staging
__global__ void __launch_bounds__(128, 3) hn_stem_mma_kernel(const float* __restrict__ x, int N, int H, int W, const float* __restrict__ w,
                                                          const float* __restrict__ b, View out, int no_relu, int tiles, int TX) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    __shared__ __align__(16) uint32_t s_out[4][16 * kStemRowWords];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tq = lane & 3;
    // this thread's k indices: slot (s, h, e) -> k = 16 s + 8 h + 2 tq + e
    int koff_c[8], k_ky[8], k_kx[8];
    uint32_t bhi[2][4][2], blo[2][4][2];
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k0 = 16 * s + 8 * h + 2 * tq;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int k = k0 + e, kk = k < 27 ? k : 0, ci = kk / 9, r = kk - ci * 9;
                koff_c[(s * 2 + h) * 2 + e] = k < 27 ? ci : -1;
                k_ky[(s * 2 + h) * 2 + e] = r / 3;
                k_kx[(s * 2 + h) * 2 + e] = r - (r / 3) * 3;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = 8 * j + gid;
                const float w0 = k0 < 27 ? __ldg(w + k0 * 32 + n) : 0.0f, w1 = k0 + 1 < 27 ? __ldg(w + (k0 + 1) * 32 + n) : 0.0f;
                split_pair(w0, w1, bhi[s][j][h], blo[s][j][h]);
            }
        }
    float bias[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { bias[j][0] = __ldg(b + 8 * j + 2 * tq); bias[j][1] = __ldg(b + 8 * j + 2 * tq + 1); }
    const int OH = out.H, OW = out.W;
    uint32_t* so = s_out[warp];
    // A warp walks whole output rows: the eight row pointers of a thread's taps are set up once per row, and a tile away from the
    // left / right border needs no bounds checks -- the kernel is bound by instruction issue, not by its 24 MMAs per tile.
    int ixb[8];  // input column of row gid of tile 0 for tap slot i: 2 * gid - 1 + kx
#pragma unroll
    for (int i = 0; i < 8; ++i) ixb[i] = 2 * gid - 1 + k_kx[i];
    const int rows = N * OH;
    for (int row = blockIdx.x * 4 + warp; row < rows; row += gridDim.x * 4) {
        const int n = row / OH, oy = row - n * OH;
        const float* xin = x + (long long)n * 3 * H * W;
        const float* rowp[8];
        bool rok[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int iy = 2 * oy - 1 + k_ky[i];
            rok[i] = koff_c[i] >= 0 && iy >= 0 && iy < H;
            rowp[i] = xin + ((long long)(rok[i] ? koff_c[i] : 0) * H + (rok[i] ? iy : 0)) * W + ixb[i];
        }
        bf16* orow = const_cast<bf16*>(vptr(out, n, oy, 0, 0));
        // the 16 samples of tile xb + 1 are requested before tile xb is computed (a warp would otherwise idle for a full memory
        // round trip per tile)
        auto gather = [&](int xb, float (&v)[8][2]) {
            const int ox0 = xb * 16;
            if (xb > 0 && 2 * (ox0 + 15) + 1 < W && ox0 + 15 < OW) {  // interior tile: every tap of every pixel is inside the image
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v[i][0] = rok[i] ? __ldg(rowp[i] + 2 * ox0) : 0.0f;
                    v[i][1] = rok[i] ? __ldg(rowp[i] + 2 * ox0 + 16) : 0.0f;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int ox = ox0 + gid + 8 * rr, ix = 2 * ox - 1 + k_kx[i];
                        v[i][rr] = (rok[i] && ix >= 0 && ix < W && ox < OW) ? __ldg(rowp[i] + 2 * ox0 + 16 * rr) : 0.0f;
                    }
            }
        };
        float vn[8][2];
        gather(0, vn);
        for (int xb = 0; xb < TX; ++xb) {
            const int ox0 = xb * 16;
            float v[8][2];
#pragma unroll
            for (int i = 0; i < 8; ++i) { v[i][0] = vn[i][0]; v[i][1] = vn[i][1]; }
            if (xb + 1 < TX) gather(xb + 1, vn);
            uint32_t ahi[2][4], alo[2][4];
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int rr = 0; rr < 2; ++rr) {
                        const int i = (s2 * 2 + h) * 2;
                        split_pair(v[i][rr], v[i + 1][rr], ahi[s2][h * 2 + rr], alo[s2][h * 2 + rr]);  // a0/a1: k pair h = 0, rows gid / gid+8; a2/a3: h = 1
                    }
            float acc[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { acc[j][0] = acc[j][2] = bias[j][0]; acc[j][1] = acc[j][3] = bias[j][1]; }
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    stem_mma(acc[j], ahi[s2], bhi[s2][j][0], bhi[s2][j][1]);
                    stem_mma(acc[j], alo[s2], bhi[s2][j][0], bhi[s2][j][1]);
                    stem_mma(acc[j], ahi[s2], blo[s2][j][0], blo[s2][j][1]);
                }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!no_relu) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) acc[j][q] = fmaxf(acc[j][q], 0.0f);
                }
                so[gid * kStemRowWords + 4 * j + tq] = hn_pack_bf16x2(acc[j][0], acc[j][1]);
                so[(gid + 8) * kStemRowWords + 4 * j + tq] = hn_pack_bf16x2(acc[j][2], acc[j][3]);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int idx = lane + 32 * i, px = idx >> 2, q = idx & 3;
                if (ox0 + px < OW)
                    *reinterpret_cast<uint4*>(orow + (ox0 + px) * out.sx + q * 8) = *reinterpret_cast<const uint4*>(so + px * kStemRowWords + q * 4);
            }
            __syncwarp();
        }
    }
}
static int g_stem_mma = 1;
extern "C" void hn_stem_set_mma(int on) { g_stem_mma = on ? 1 : 0; }  // 0: the fp32 CUDA-core kernel (A/B runs)

extern "C" int hn_stem_fwd(const hn_stem_desc* d, void* stream) {
    HN_REQUIRE(d && d->x && d->w && d->b, "stem: null pointer");
    if (int rc = check_view(d->out, "stem.out")) return rc;
    HN_REQUIRE(d->out.C == 32 && d->out.N == d->N && d->out.H == (d->H + 1) / 2 && d->out.W == (d->W + 1) / 2,
               "stem: output view %dx%dx%dx%d does not match input %dx3x%dx%d", d->out.N, d->out.H, d->out.W, d->out.C, d->N,
               d->H, d->W);
    long long total = (long long)d->N * d->out.H * ((d->out.W + 1) / 2);  // two output pixels per thread
    HN_REQUIRE(total < 0x7fffffffLL, "stem: too many output pixels for one launch");
    if (g_stem_mma) {
        const int TX = hn_cdiv(d->out.W, 16);
        const long long tiles = (long long)d->N * d->out.H * TX;
        HN_REQUIRE(tiles < 0x7fffffffLL, "stem: too many tiles for one launch");
        int sms = hn_device_sm_count();
        if (sms <= 0) sms = 148;
        long long blocks = hn_cdiv((long long)d->N * d->out.H, 4);       // a warp per output row, grid-stride
        if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;  // the weight fragments are built once per thread
        HN_CHECK_CUDA(hn_launch(hn_stem_mma_kernel, dim3((unsigned)blocks), dim3(128), (size_t)0, reinterpret_cast<cudaStream_t>(stream), d->x, d->N, d->H, d->W,
                                d->w, d->b, to_view(d->out), (int)d->no_relu, (int)tiles, TX));
        HN_CHECK_CUDA(cudaGetLastError());
        return HN_OK;
    }
    HN_CHECK_CUDA(hn_launch(hn_stem_kernel, dim3(hn_cdiv(total, 128)), dim3(128), (size_t)(0), reinterpret_cast<cudaStream_t>(stream), d->x, d->N, d->H, d->W, d->w,
                                                                                          d->b, to_view(d->out), (int)d->no_relu));
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// BiFPN node front half: weighted sum of <=3 inputs (same / nearest-up2 / zero-padded max-pool)
// -> swish -> depthwise 3x3 (zero pad 1).  The fused value of the (T+2)^2 halo tile is computed once
// into shared memory (fp32), the depthwise taps then read it from there.
// ------------------------------------------------------------------------------------------------
struct NodeParams {
    int n_in;
    View in[3];
    int mode[3];
    float w[3];
    int swish;
    const float* dw;
    View out;
};


__device__ __forceinline__ void node_fetch(const View& v, int mode, int n, int y, int x, int c, float (&f)[8]) {
    if (mode == HN_IN_SAME) {
        load8(vptr(v, n, y, x, c), f);
    } else if (mode == HN_IN_UP2) {
        load8(vptr(v, n, y >> 1, x >> 1, c), f);
    } else {
        // 3x3 stride-2 max over the input padded with one zero row/col at the bottom/right
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = -INFINITY;
        for (int dy = 0; dy < 3; ++dy) {
            int iy = 2 * y + dy;
            for (int dx = 0; dx < 3; ++dx) {
                int ix = 2 * x + dx;
                float g[8];
                if (iy < v.H && ix < v.W) {
                    load8(vptr(v, n, iy, ix, c), g);
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) g[j] = 0.0f;
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], g[j]);
            }
        }
    }
}

static constexpr int kNodeTH = 8, kNodeTW = 16;         // output tile
static constexpr int kNodeHH = kNodeTH + 2, kNodeHW = kNodeTW + 2;  // halo tile
static constexpr int kNodeCols = 1;                       // output columns per thread in phase 2
static constexpr int kNodeRows = kNodeTW / kNodeCols;     // threadIdx.y extent: one per group of kNodeCols output columns

// packed bf16x2 max (max commutes with the monotonic bf16 -> fp32 widening, so this equals max on the widened values)
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
    __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&a), y = *reinterpret_cast<__nv_bfloat162*>(&b);
    __nv_bfloat162 m = __hmax2(x, y);
    return *reinterpret_cast<uint32_t*>(&m);
}

// CTA = L x 8 threads over an 8x16 output tile of a channel slice (L four-channel vectors, blockIdx.y picks the
// slice; depthwise work never crosses channels).  threadIdx.x is the channel vector: a pixel's channels are
// contiguous, so global accesses coalesce and neighbouring lanes hit neighbouring shared-memory banks.
// Phase 1 computes the fused value (weighted sum, swish) of the (8+2)x(16+2) halo tile once into shared memory
// (fp32), kNodeBatch pixels per thread at a time.  The phase is bound by global-load latency, so its loads are
// branch-free (clamped coordinates, the up-sampling as a shift, the input count and the pooled input as template
// parameters): the compiler can then put a whole batch in flight before the first use.
// Phase 2: thread (lane, j) owns output columns 2j, 2j+1 and walks down the ten halo rows with the 9x4 weights in
// registers; each halo row is read once (4 vectors) and feeds the three output rows it touches, whose
// accumulators roll.
static constexpr int kNodeBatch = 6;
template <int kNin, bool kPoolLast>
__global__ void __launch_bounds__(512) hn_node_kernel(const __grid_constant__ NodeParams p) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    extern __shared__ float s_tile[];  // [kNodeHH * kNodeHW][CB] fused values
    constexpr int kDirect = kPoolLast ? kNin - 1 : kNin;
    const int C = p.out.C;
    const int CB = blockDim.x * 4;
    const int lc = threadIdx.x * 4;
    const int c = blockIdx.y * CB + lc;
    const int tiles_x = (p.out.W + kNodeTW - 1) / kNodeTW, tiles_y = (p.out.H + kNodeTH - 1) / kNodeTH;
    const int per_img = tiles_x * tiles_y;
    const int n = blockIdx.x / per_img;
    const int r = blockIdx.x - n * per_img;
    const int y0 = (r / tiles_x) * kNodeTH, x0 = (r % tiles_x) * kNodeTW;
    constexpr int kHalo = kNodeHH * kNodeHW;
    const bf16* base[kNin];
    int sh[kNin];
#pragma unroll
    for (int i = 0; i < kNin; ++i) {
        base[i] = p.in[i].ptr + n * p.in[i].sn + c;
        sh[i] = p.mode[i] == HN_IN_UP2 ? 1 : 0;
    }
    for (int hp0 = threadIdx.y; hp0 < kHalo; hp0 += kNodeRows * kNodeBatch) {
        uint2 raw[kNodeBatch][kNin];
        bool inside[kNodeBatch];
#pragma unroll
        for (int u = 0; u < kNodeBatch; ++u) {
            const int hp = min(hp0 + u * kNodeRows, kHalo - 1);
            const int hy = hp / kNodeHW, hx = hp - hy * kNodeHW;
            const int y = y0 + hy - 1, x = x0 + hx - 1;
            inside[u] = y >= 0 && y < p.out.H && x >= 0 && x < p.out.W;
            const int yc = min(max(y, 0), p.out.H - 1), xc = min(max(x, 0), p.out.W - 1);
#pragma unroll
            for (int i = 0; i < kDirect; ++i)
                raw[u][i] = *reinterpret_cast<const uint2*>(base[i] + (yc >> sh[i]) * p.in[i].sy + (xc >> sh[i]) * p.in[i].sx);
            if (kPoolLast) {
                // 3x3 stride-2 max over the input padded with one zero row/col at the bottom/right
                const View& v = p.in[kNin - 1];
                uint2 m = make_uint2(0xff80ff80u, 0xff80ff80u);  // -inf
#pragma unroll
                for (int dy = 0; dy < 3; ++dy) {
#pragma unroll
                    for (int dx = 0; dx < 3; ++dx) {
                        const int iy = 2 * yc + dy, ix = 2 * xc + dx;
                        const bool ok = iy < v.H && ix < v.W;
                        uint2 q = *reinterpret_cast<const uint2*>(base[kNin - 1] + min(iy, v.H - 1) * v.sy + min(ix, v.W - 1) * v.sx);
                        if (!ok) q = make_uint2(0u, 0u);
                        m.x = bf16x2_max(m.x, q.x);
                        m.y = bf16x2_max(m.y, q.y);
                    }
                }
                raw[u][kNin - 1] = m;
            }
        }
#pragma unroll
        for (int u = 0; u < kNodeBatch; ++u) {
            const int hp = hp0 + u * kNodeRows;
            if (hp >= kHalo) break;
            float acc[4];
            // the reference evaluates w0*a + w1*b (+ w2*c) left to right in fp32 (bifpn.py:170-231)
#pragma unroll
            for (int i = 0; i < kNin; ++i) {
                const float2 a = hn_unpack_bf16x2(raw[u][i].x), b = hn_unpack_bf16x2(raw[u][i].y);
                const float w = p.w[i];
                if (i == 0) { acc[0] = w * a.x; acc[1] = w * a.y; acc[2] = w * b.x; acc[3] = w * b.y; }
                else { acc[0] = acc[0] + w * a.x; acc[1] = acc[1] + w * a.y; acc[2] = acc[2] + w * b.x; acc[3] = acc[3] + w * b.y; }
            }
            if (p.swish) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = acc[j] * hn_sigmoid(acc[j]);
            }
            if (!inside[u]) acc[0] = acc[1] = acc[2] = acc[3] = 0.0f;  // the depthwise conv's zero padding
            *reinterpret_cast<float4*>(s_tile + hp * CB + lc) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        }
    }
    float4 wgt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wgt[k] = __ldg(reinterpret_cast<const float4*>(p.dw + k * C + c));
    __syncthreads();
    const int tx = threadIdx.y * kNodeCols;  // first of this thread's output columns
    float4 acc[kNodeTH][kNodeCols];
#pragma unroll
    for (int hy = 0; hy < kNodeHH; ++hy) {
        float4 a[kNodeCols + 2];
#pragma unroll
        for (int q = 0; q < kNodeCols + 2; ++q) a[q] = *reinterpret_cast<const float4*>(s_tile + (hy * kNodeHW + tx + q) * CB + lc);
#pragma unroll
        for (int ky = 2; ky >= 0; --ky) {  // output row o = hy - ky takes this halo row through kernel row ky
            const int o = hy - ky;
            if (o < 0 || o >= kNodeTH) continue;
#pragma unroll
            for (int col = 0; col < kNodeCols; ++col) {
                float4 v = ky == 0 ? make_float4(0.0f, 0.0f, 0.0f, 0.0f) : acc[o][col];
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4 w = wgt[ky * 3 + kx];
                    const float4 s = a[col + kx];
                    v.x = fmaf(s.x, w.x, v.x); v.y = fmaf(s.y, w.y, v.y); v.z = fmaf(s.z, w.z, v.z); v.w = fmaf(s.w, w.w, v.w);
                }
                acc[o][col] = v;
            }
        }
        const int o = hy - 2;  // complete after its third halo row
        if (o >= 0) {
            const int y = y0 + o;
#pragma unroll
            for (int col = 0; col < kNodeCols; ++col) {
                const int x = x0 + tx + col;
                if (y < p.out.H && x < p.out.W) {
                    const float4 v = acc[o][col];
                    *reinterpret_cast<uint2*>(const_cast<bf16*>(vptr(p.out, n, y, x, c))) =
                        make_uint2(hn_pack_bf16x2(v.x, v.y), hn_pack_bf16x2(v.z, v.w));
                }
            }
        }
    }
}

// Plain depthwise 3x3 (one input at the output resolution, no fusion).  One thread = one 8-channel vector of
// kDwPx consecutive pixels of a row: per kernel row it loads kDwPx+2 input vectors and the row's three weight
// vectors once and reuses them across the kDwPx outputs (4.5 instead of 9 data loads and 4.5 instead of 18 weight
// loads per output: the op is bound by L1 load issue, not by HBM).  No shared memory, no barrier.
static constexpr int kDwPx = 4;
__device__ __forceinline__ void dw_strip(const View& in, const View& out, const float* __restrict__ dw, int n, int y, int x0,
                                         int c) {
    const int C = out.C;
    float acc[kDwPx][8];
#pragma unroll
    for (int p = 0; p < kDwPx; ++p)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[p][j] = 0.0f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = y + ky - 1;
        if (iy < 0 || iy >= out.H) continue;
        uint4 raw[kDwPx + 2];
#pragma unroll
        for (int q = 0; q < kDwPx + 2; ++q) {
            const int ix = x0 + q - 1;
            raw[q] = (ix >= 0 && ix < out.W) ? *reinterpret_cast<const uint4*>(vptr(in, n, iy, ix, c)) : make_uint4(0, 0, 0, 0);
        }
        float w[3][8];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const float4* wv = reinterpret_cast<const float4*>(dw + (ky * 3 + kx) * C + c);
            const float4 w0 = __ldg(wv), w1 = __ldg(wv + 1);
            w[kx][0] = w0.x; w[kx][1] = w0.y; w[kx][2] = w0.z; w[kx][3] = w0.w;
            w[kx][4] = w1.x; w[kx][5] = w1.y; w[kx][6] = w1.z; w[kx][7] = w1.w;
        }
#pragma unroll
        for (int q = 0; q < kDwPx + 2; ++q) {
            float f[8];
            const float2 a = hn_unpack_bf16x2(raw[q].x), b = hn_unpack_bf16x2(raw[q].y), cc = hn_unpack_bf16x2(raw[q].z),
                         d = hn_unpack_bf16x2(raw[q].w);
            f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = cc.x; f[5] = cc.y; f[6] = d.x; f[7] = d.y;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int p = q - kx;  // input column q feeds output p through tap kx
                if (p >= 0 && p < kDwPx) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc[p][j] = fmaf(f[j], w[kx][j], acc[p][j]);
                }
            }
        }
    }
#pragma unroll
    for (int p = 0; p < kDwPx; ++p)
        if (x0 + p < out.W) store8(const_cast<bf16*>(vptr(out, n, y, x0 + p, c)), acc[p]);
}

__global__ void __launch_bounds__(256) hn_dw_kernel(View in, View out, const float* __restrict__ dw) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    const int CV = out.C >> 3, WB = (out.W + kDwPx - 1) / kDwPx;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned total = (unsigned)out.N * out.H * WB * CV;
    if (idx >= total) return;
    const int cv = (int)(idx % (unsigned)CV);
    unsigned t = idx / (unsigned)CV;
    const int xb = (int)(t % (unsigned)WB);
    t /= (unsigned)WB;
    const int y = (int)(t % (unsigned)out.H);
    const int n = (int)(t / (unsigned)out.H);
    dw_strip(in, out, dw, n, y, xb * kDwPx, cv * 8);
}

// the same over several (input, output) pairs sharing the weights: one launch for all pyramid levels
struct DwMultiParams {
    int n;
    View in[HN_MAX_GROUPS], out[HN_MAX_GROUPS];
    unsigned end[HN_MAX_GROUPS];  // exclusive prefix of work items
    const float* dw;
};
// thread = one 4-channel vector (threadIdx.x) with its 9x4 weights in registers, walking strips of kDwPx output pixels of
// a row (grid-stride over the strips of all levels).  Per kernel row a strip loads kDwPx+2 input vectors, branch-free
// (clamped coordinates + select), and reuses them across its outputs.  `end[g]` counts strips here.
__global__ void __launch_bounds__(256) hn_dw_multi_kernel(const __grid_constant__ DwMultiParams p) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    const int C = p.out[0].C;
    const int c = threadIdx.x * 4;
    float4 wgt[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) wgt[k] = __ldg(reinterpret_cast<const float4*>(p.dw + k * C + c));
    const unsigned total = p.end[p.n - 1];
    for (unsigned s = blockIdx.x * blockDim.y + threadIdx.y; s < total; s += gridDim.x * blockDim.y) {
        int g = 0;
        while (g < p.n - 1 && s >= p.end[g]) ++g;
        unsigned t = g > 0 ? s - p.end[g - 1] : s;
        const View& in = p.in[g];
        const View& out = p.out[g];
        const int H = out.H, W = out.W, WB = (W + kDwPx - 1) / kDwPx;
        const int xb = (int)(t % (unsigned)WB);
        t /= (unsigned)WB;
        const int y = (int)(t % (unsigned)H);
        const int n = (int)(t / (unsigned)H);
        const int x0 = xb * kDwPx;
        float4 acc[kDwPx];
#pragma unroll
        for (int q = 0; q < kDwPx; ++q) acc[q] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = y + ky - 1;
            const bool row_ok = iy >= 0 && iy < H;
            const bf16* rowp = in.ptr + n * in.sn + min(max(iy, 0), H - 1) * in.sy + c;
            uint2 raw[kDwPx + 2];
#pragma unroll
            for (int q = 0; q < kDwPx + 2; ++q) {
                const int ix = x0 + q - 1;
                raw[q] = *reinterpret_cast<const uint2*>(rowp + min(max(ix, 0), W - 1) * in.sx);
                if (!(row_ok && ix >= 0 && ix < W)) raw[q] = make_uint2(0u, 0u);
            }
#pragma unroll
            for (int q = 0; q < kDwPx + 2; ++q) {
                const float2 a = hn_unpack_bf16x2(raw[q].x), b = hn_unpack_bf16x2(raw[q].y);
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int o = q - kx;  // input column q feeds output o through tap kx
                    if (o >= 0 && o < kDwPx) {
                        const float4 w = wgt[ky * 3 + kx];
                        acc[o].x = fmaf(a.x, w.x, acc[o].x); acc[o].y = fmaf(a.y, w.y, acc[o].y);
                        acc[o].z = fmaf(b.x, w.z, acc[o].z); acc[o].w = fmaf(b.y, w.w, acc[o].w);
                    }
                }
            }
        }
#pragma unroll
        for (int q = 0; q < kDwPx; ++q)
            if (x0 + q < W)
                *reinterpret_cast<uint2*>(const_cast<bf16*>(vptr(out, n, y, x0 + q, c))) =
                    make_uint2(hn_pack_bf16x2(acc[q].x, acc[q].y), hn_pack_bf16x2(acc[q].z, acc[q].w));
    }
}

extern "C" int hn_dw_multi_fwd(const hn_dw_multi_desc* d, void* stream) {
    HN_REQUIRE(d && d->n >= 1 && d->n <= HN_MAX_GROUPS && d->dw, "dw_multi: bad descriptor");
    DwMultiParams p;
    memset(&p, 0, sizeof(p));
    p.n = d->n;
    p.dw = d->dw;
    long long total = 0;
    for (int i = 0; i < d->n; ++i) {
        if (int rc = check_view(d->in[i], "dw_multi.in")) return rc;
        if (int rc = check_view(d->out[i], "dw_multi.out")) return rc;
        HN_REQUIRE(d->in[i].N == d->out[i].N && d->in[i].H == d->out[i].H && d->in[i].W == d->out[i].W &&
                       d->in[i].C == d->out[i].C && d->in[i].C == d->in[0].C,
                   "dw_multi: pair %d shape mismatch", i);
        p.in[i] = to_view(d->in[i]);
        p.out[i] = to_view(d->out[i]);
        total += (long long)d->out[i].N * d->out[i].H * ((d->out[i].W + kDwPx - 1) / kDwPx);  // strips
        HN_REQUIRE(total < 0x7fffffffLL, "dw_multi: too many work items");
        p.end[i] = (unsigned)total;
    }
    const int lanes = d->out[0].C / 4;
    HN_REQUIRE(lanes >= 1 && lanes <= 256, "dw_multi: C=%d not supported", d->out[0].C);
    const int rows = 256 / lanes;
    int sms = hn_device_sm_count();
    if (sms <= 0) sms = 148;
    long long blocks = hn_cdiv(total, rows);
    if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;  // grid-stride: the weights are loaded once per thread
    HN_CHECK_CUDA(hn_launch(hn_dw_multi_kernel, dim3((unsigned)blocks), dim3(lanes, rows), (size_t)0, reinterpret_cast<cudaStream_t>(stream), p));
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

extern "C" int hn_node_fwd(const hn_node_desc* d, void* stream) {
    HN_REQUIRE(d && d->n_in >= 1 && d->n_in <= 3 && d->dw, "node: bad descriptor");
    if (int rc = check_view(d->out, "node.out")) return rc;
    NodeParams p;
    memset(&p, 0, sizeof(p));
    p.n_in = d->n_in;
    for (int i = 0; i < d->n_in; ++i) {
        if (int rc = check_view(d->in[i], "node.in")) return rc;
        const hn_view& v = d->in[i];
        HN_REQUIRE(v.C == d->out.C && v.N == d->out.N, "node: input %d channel/batch mismatch", i);
        if (d->mode[i] == HN_IN_SAME) HN_REQUIRE(v.H == d->out.H && v.W == d->out.W, "node: SAME input %d size mismatch", i);
        else if (d->mode[i] == HN_IN_UP2)
            HN_REQUIRE(v.H * 2 == d->out.H && v.W * 2 == d->out.W, "node: UP2 input %d is %dx%d for output %dx%d", i, v.H, v.W,
                       d->out.H, d->out.W);
        else if (d->mode[i] == HN_IN_POOL)
            HN_REQUIRE((v.H - 2) / 2 + 1 == d->out.H && (v.W - 2) / 2 + 1 == d->out.W,
                       "node: POOL input %d is %dx%d for output %dx%d", i, v.H, v.W, d->out.H, d->out.W);
        else HN_REQUIRE(false, "node: unknown input mode %d", d->mode[i]);
        p.in[i] = to_view(v);
        p.mode[i] = d->mode[i];
        p.w[i] = d->w[i];
    }
    p.swish = d->swish;
    p.dw = d->dw;
    p.out = to_view(d->out);
    if (d->n_in == 1 && d->mode[0] == HN_IN_SAME && !d->swish && d->w[0] == 1.0f) {
        long long total = (long long)d->out.N * d->out.H * ((d->out.W + kDwPx - 1) / kDwPx) * (d->out.C / 8);
        HN_REQUIRE(total < 0x7fffffffLL, "node: too many work items for one launch");
        HN_CHECK_CUDA(hn_launch(hn_dw_kernel, dim3(hn_cdiv(total, 256)), dim3(256), (size_t)(0), reinterpret_cast<cudaStream_t>(stream), p.in[0], p.out, d->dw));
        HN_CHECK_CUDA(cudaGetLastError());
        return HN_OK;
    }
    // Channel slices: a CTA takes L = (C/4)/split four-channel vectors.  Smaller slices mean a smaller shared tile
    // and more resident CTAs per SM, so the load phase of one overlaps the tap phase of another.
    const int L4 = d->out.C / 4;
    int split = 1;
    while (split < 4 && (L4 / split) % 2 == 0 && L4 / split > 8 &&
           (size_t)kNodeHH * kNodeHW * (L4 / split) * 4 * sizeof(float) > 44 * 1024)
        split *= 2;
    const int CB = L4 / split;
    HN_REQUIRE(CB >= 1 && CB * kNodeRows <= 512, "node: C=%d not supported", d->out.C);
    const int tiles = hn_cdiv(d->out.W, kNodeTW) * hn_cdiv(d->out.H, kNodeTH);
    size_t smem = (size_t)kNodeHH * kNodeHW * CB * 4 * sizeof(float);
    // pooled inputs: only as the last one (the BiFPN bottom-up nodes: same-level inputs, then the level below)
    bool pool_last = d->mode[d->n_in - 1] == HN_IN_POOL;
    for (int i = 0; i + 1 < d->n_in; ++i) HN_REQUIRE(d->mode[i] != HN_IN_POOL, "node: a pooled input must be the last input");
    void (*kern)(const NodeParams) = nullptr;
    switch (d->n_in * 2 + (pool_last ? 1 : 0)) {
        case 2: kern = hn_node_kernel<1, false>; break;
        case 3: kern = hn_node_kernel<1, true>; break;
        case 4: kern = hn_node_kernel<2, false>; break;
        case 5: kern = hn_node_kernel<2, true>; break;
        case 6: kern = hn_node_kernel<3, false>; break;
        default: kern = hn_node_kernel<3, true>; break;
    }
    if (smem > 48 * 1024) HN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    HN_CHECK_CUDA(hn_launch(kern, dim3(tiles * d->out.N, split), dim3(CB, kNodeRows), smem, reinterpret_cast<cudaStream_t>(stream), p));
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// standalone 3x3 stride-2 max-pool
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pool_neginf(const View& v, int n, int y, int x, int c, float (&f)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = -INFINITY;
    for (int dy = -1; dy <= 1; ++dy) {
        int iy = 2 * y + dy;
        if (iy < 0 || iy >= v.H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            int ix = 2 * x + dx;
            if (ix < 0 || ix >= v.W) continue;
            float g[8];
            load8(vptr(v, n, iy, ix, c), g);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], g[j]);
        }
    }
}

__global__ void hn_pool_kernel(View in, View out, int mode) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    const int CV = out.C >> 3;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned total = (unsigned)out.N * out.H * out.W * CV;
    if (idx >= total) return;
    int cv = (int)(idx % (unsigned)CV);
    unsigned t = idx / (unsigned)CV;
    int x = (int)(t % (unsigned)out.W);
    t /= (unsigned)out.W;
    int y = (int)(t % (unsigned)out.H);
    int n = (int)(t / (unsigned)out.H);
    float f[8];
    if (mode == HN_POOL_ZERO_RB) node_fetch(in, HN_IN_POOL, n, y, x, cv * 8, f);
    else pool_neginf(in, n, y, x, cv * 8, f);
    store8(const_cast<bf16*>(vptr(out, n, y, x, cv * 8)), f);
}

extern "C" int hn_pool_fwd(const hn_pool_desc* d, void* stream) {
    HN_REQUIRE(d != nullptr, "pool: null desc");
    if (int rc = check_view(d->in, "pool.in")) return rc;
    if (int rc = check_view(d->out, "pool.out")) return rc;
    HN_REQUIRE(d->in.C == d->out.C && d->in.N == d->out.N, "pool: channel/batch mismatch");
    if (d->mode == HN_POOL_ZERO_RB)
        HN_REQUIRE((d->in.H - 2) / 2 + 1 == d->out.H && (d->in.W - 2) / 2 + 1 == d->out.W, "pool: size mismatch");
    else
        HN_REQUIRE((d->in.H - 1) / 2 + 1 == d->out.H && (d->in.W - 1) / 2 + 1 == d->out.W, "pool: size mismatch");
    long long total = (long long)d->out.N * d->out.H * d->out.W * (d->out.C / 8);
    HN_REQUIRE(total < 0x7fffffffLL, "pool: too many work items for one launch");
    HN_CHECK_CUDA(hn_launch(hn_pool_kernel, dim3(hn_cdiv(total, 256)), dim3(256), (size_t)(0), reinterpret_cast<cudaStream_t>(stream), to_view(d->in), to_view(d->out),
                                                                                          d->mode));
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// lane fuse
// ------------------------------------------------------------------------------------------------
struct LaneFuseParams {
    View p3, p4, p5, p6, out;
    int stride;
};

__global__ void hn_lanefuse_kernel(const __grid_constant__ LaneFuseParams p) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    const int C = p.p3.C, CV = C >> 3;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned total = (unsigned)p.out.N * p.out.H * p.out.W * 4 * CV;
    if (idx >= total) return;
    int cv = (int)(idx % (unsigned)CV);
    unsigned t = idx / (unsigned)CV;
    int part = (int)(t % 4u);
    t /= 4u;
    int x = (int)(t % (unsigned)p.out.W);
    t /= (unsigned)p.out.W;
    int y = (int)(t % (unsigned)p.out.H);
    int n = (int)(t / (unsigned)p.out.H);
    const int c = cv * 8;
    float f[8];
    if (p.stride == 32) {
        if (part == 0) {
            // maxpool(maxpool(P3)): evaluate the inner pool at the (up to 9) valid outer taps
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = -INFINITY;
            const int H1 = (p.p3.H - 1) / 2 + 1, W1 = (p.p3.W - 1) / 2 + 1;
            for (int dy = -1; dy <= 1; ++dy) {
                int iy = 2 * y + dy;
                if (iy < 0 || iy >= H1) continue;
                for (int dx = -1; dx <= 1; ++dx) {
                    int ix = 2 * x + dx;
                    if (ix < 0 || ix >= W1) continue;
                    float g[8];
                    pool_neginf(p.p3, n, iy, ix, c, g);
#pragma unroll
                    for (int j = 0; j < 8; ++j) f[j] = fmaxf(f[j], g[j]);
                }
            }
        } else if (part == 1) {
            pool_neginf(p.p4, n, y, x, c, f);
        } else if (part == 2) {
            load8(vptr(p.p5, n, y, x, c), f);
        } else {
            load8(vptr(p.p6, n, y >> 1, x >> 1, c), f);
        }
    } else {
        if (part == 0) pool_neginf(p.p3, n, y, x, c, f);
        else if (part == 1) load8(vptr(p.p5, n, y >> 1, x >> 1, c), f);
        else if (part == 2) load8(vptr(p.p4, n, y, x, c), f);
        else load8(vptr(p.p6, n, y >> 2, x >> 2, c), f);
    }
    store8(const_cast<bf16*>(vptr(p.out, n, y, x, part * C + c)), f);
}

extern "C" int hn_lanefuse_fwd(const hn_lanefuse_desc* d, void* stream) {
    HN_REQUIRE(d != nullptr && (d->stride == 16 || d->stride == 32), "lanefuse: stride must be 16 or 32");
    const hn_view* vs[5] = {&d->p3, &d->p4, &d->p5, &d->p6, &d->out};
    for (int i = 0; i < 5; ++i)
        if (int rc = check_view(*vs[i], "lanefuse")) return rc;
    HN_REQUIRE(d->out.C == 4 * d->p3.C && d->p4.C == d->p3.C && d->p5.C == d->p3.C && d->p6.C == d->p3.C,
               "lanefuse: channel mismatch");
    if (d->stride == 32)
        HN_REQUIRE(d->out.H == d->p5.H && d->out.W == d->p5.W && d->p6.H * 2 == d->out.H && d->p6.W * 2 == d->out.W &&
                       (d->p4.H - 1) / 2 + 1 == d->out.H && ((d->p3.H - 1) / 2 + 1 - 1) / 2 + 1 == d->out.H,
                   "lanefuse: pyramid sizes inconsistent for stride 32");
    else
        HN_REQUIRE(d->out.H == d->p4.H && d->out.W == d->p4.W && d->p5.H * 2 == d->out.H && d->p6.H * 4 == d->out.H &&
                       (d->p3.H - 1) / 2 + 1 == d->out.H,
                   "lanefuse: pyramid sizes inconsistent for stride 16");
    LaneFuseParams p;
    p.p3 = to_view(d->p3); p.p4 = to_view(d->p4); p.p5 = to_view(d->p5); p.p6 = to_view(d->p6); p.out = to_view(d->out);
    p.stride = d->stride;
    long long total = (long long)d->out.N * d->out.H * d->out.W * 4 * (d->p3.C / 8);
    HN_REQUIRE(total < 0x7fffffffLL, "lanefuse: too many work items for one launch");
    HN_CHECK_CUDA(hn_launch(hn_lanefuse_kernel, dim3(hn_cdiv(total, 256)), dim3(256), (size_t)(0), reinterpret_cast<cudaStream_t>(stream), p));
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// squeeze-excite: global average pool (deterministic) + the two FC layers of the gate in one launch, and the
// in-place channel scaling
// ------------------------------------------------------------------------------------------------
static constexpr int kSeThreads = 512;

struct SeFc {
    int S;            // padded hidden width (multiple of 8), 0 = no FC stage
    const bf16* w1;   // [S][C]
    const float* b1;  // [S]
    const bf16* w2;   // [C][S]
    const float* b2;  // [C]
    bf16* gate;       // [N][C]
};

// dot product of `len` (a multiple of 8) bf16 weights with fp32 activations in shared memory, 16-byte weight loads
__device__ __forceinline__ float se_dot(const bf16* __restrict__ w, const float* act, int v0, int v1) {
    float acc = 0.0f;
    constexpr int U = 16;  // weight vectors in flight per thread: the FC stage is one CTA per image, bound by L2 latency
    for (int vb = v0; vb < v1; vb += U) {
        uint4 r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = vb + u < v1 ? __ldg(reinterpret_cast<const uint4*>(w) + vb + u) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (vb + u >= v1) break;
            const float2 a = hn_unpack_bf16x2(r[u].x), b = hn_unpack_bf16x2(r[u].y), c = hn_unpack_bf16x2(r[u].z), d = hn_unpack_bf16x2(r[u].w);
            const float* x = act + (vb + u) * 8;
            acc = fmaf(a.x, x[0], acc); acc = fmaf(a.y, x[1], acc); acc = fmaf(b.x, x[2], acc); acc = fmaf(b.y, x[3], acc);
            acc = fmaf(c.x, x[4], acc); acc = fmaf(c.y, x[5], acc); acc = fmaf(d.x, x[6], acc); acc = fmaf(d.y, x[7], acc);
        }
    }
    return acc;
}

__global__ void __launch_bounds__(kSeThreads) hn_se_pool_kernel(View x, float* __restrict__ partial, int* __restrict__ counter,
                                                                bf16* __restrict__ mean_out, float inv_hw, int kSePix, SeFc fc) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    extern __shared__ float sm[];  // [lanes][CG*8] partial sums of this CTA's channel group
    __shared__ int s_last;
    if (fc.S) {
        // The FC stage at the end is a chain of dependent loads run by ONE block per image; its weights were last
        // touched a whole step ago (evicted to DRAM).  All blocks pull them into the L2 now, while they pool.
        const long long nb = (long long)gridDim.x * gridDim.y * gridDim.z;
        const long long bid = ((long long)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        const long long wbytes = (long long)fc.S * x.C * 2;
        const long long lines = (wbytes + 127) >> 7;
        for (long long l = bid * blockDim.x + threadIdx.x; l < 2 * lines; l += nb * blockDim.x) {
            const char* p = l < lines ? reinterpret_cast<const char*>(fc.w1) + (l << 7) : reinterpret_cast<const char*>(fc.w2) + ((l - lines) << 7);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
        }
        if (bid == 0) {
            for (int l = threadIdx.x; l * 32 < fc.S; l += blockDim.x) asm volatile("prefetch.global.L2 [%0];" ::"l"(fc.b1 + l * 32));
            for (int l = threadIdx.x; l * 32 < x.C; l += blockDim.x) asm volatile("prefetch.global.L2 [%0];" ::"l"(fc.b2 + l * 32));
        }
    }
    const int C = x.C, CV = C >> 3;
    const int n = blockIdx.y;
    const int HW = x.H * x.W;
    const int p0 = blockIdx.x * kSePix, p1 = min(p0 + kSePix, HW);
    // blockIdx.z = channel group of CG 8-channel vectors: small maps with many channels (10x10x936) would otherwise
    // leave each thread a long serial chain of dependent-latency loads on a handful of CTAs
    const int CG = (CV + gridDim.z - 1) / gridDim.z, cv0 = blockIdx.z * CG;
    const int lanes = blockDim.x / CG;
    const int cvl = threadIdx.x % CG, pl = threadIdx.x / CG, cv = cv0 + cvl;
    if (pl < lanes && cv < CV) {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
        const bf16* base = x.ptr + n * x.sn + cv * 8;
        int px = p0 + pl;
        for (; px + 3 * lanes < p1; px += 4 * lanes) {  // four independent loads in flight
            uint4 r[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int q = px + u * lanes, y = q / x.W, xx = q - y * x.W;
                r[u] = *reinterpret_cast<const uint4*>(base + y * x.sy + xx * x.sx);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {  // same summation order as the one-at-a-time loop
                const float2 a = hn_unpack_bf16x2(r[u].x), b = hn_unpack_bf16x2(r[u].y), c2 = hn_unpack_bf16x2(r[u].z),
                             d2 = hn_unpack_bf16x2(r[u].w);
                acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
                acc[4] += c2.x; acc[5] += c2.y; acc[6] += d2.x; acc[7] += d2.y;
            }
        }
        for (; px < p1; px += lanes) {
            const int y = px / x.W, xx = px - y * x.W;
            float f[8];
            load8(base + y * x.sy + xx * x.sx, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += f[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) sm[(pl * CG + cvl) * 8 + j] = acc[j];
    }
    __syncthreads();
    const int nchunk = gridDim.x;
    for (int c = threadIdx.x; c < CG * 8; c += blockDim.x) {
        if (cv0 * 8 + c >= C) break;
        float a = 0.0f;
        for (int l = 0; l < lanes; ++l) a += sm[l * CG * 8 + c];
        partial[((long long)n * nchunk + blockIdx.x) * C + cv0 * 8 + c] = a;  // fixed summation order: deterministic
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(counter + n, 1) == nchunk * (int)gridDim.z - 1);
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float* pp = partial + (long long)n * nchunk * C + c;
        float a = 0.0f;
        int k = 0;
        for (; k + 8 <= nchunk; k += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = __ldcg(pp + (long long)(k + j) * C);
#pragma unroll
            for (int j = 0; j < 8; ++j) a += v[j];
        }
        for (; k < nchunk; ++k) a += __ldcg(pp + (long long)k * C);
        const bf16 mb = __float2bfloat16(a * inv_hw);
        mean_out[(long long)n * C + c] = mb;
        if (fc.S) sm[c] = __bfloat162float(mb);  // the FC stage reads the rounded mean, as a separate GEMM launch would
    }
    if (threadIdx.x == 0) counter[n] = 0;
    if (!fc.S) return;
    // ---- FC1 + ReLU, FC2 + sigmoid for this image (this block arrived last, so it is the only one left) ----
    float* s_mean = sm;
    float* s_hid = sm + ((C + 7) & ~7);
    __syncthreads();
    const int S = fc.S;
    {   // FC1: T adjacent lanes per hidden unit (T = largest power of two with S*T <= blockDim.x, at most 32)
        int T = 1;
        while (T < 32 && S * T * 2 <= (int)blockDim.x) T *= 2;
        const int nv = C >> 3, per = (nv + T - 1) / T;
        for (int s0 = 0; s0 < S; s0 += blockDim.x / T) {
            const int srow = s0 + threadIdx.x / T, part = threadIdx.x % T;
            float acc = 0.0f;
            if (srow < S) acc = se_dot(fc.w1 + (long long)srow * C, s_mean, min(part * per, nv), min(part * per + per, nv));
            for (int o = T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (srow < S && part == 0) s_hid[srow] = __bfloat162float(__float2bfloat16(fmaxf(acc + fc.b1[srow], 0.0f)));
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {  // FC2: one thread per channel, its weight row is contiguous
        const float acc = se_dot(fc.w2 + (long long)c * S, s_hid, 0, S >> 3) + fc.b2[c];
        fc.gate[(long long)n * C + c] = __float2bfloat16(1.0f / (1.0f + expf(-acc)));
    }
}


// ---- squeeze-excite FC layers as two batched launches (all images per CTA) ------------------------------------------
// The fused tail above runs the two FC layers of an image in ONE block (a chain of dependent L2 loads on 32 SMs while
// the other 116 idle).  Split form: pooling only, then FC1 with a CTA per group of hidden units and FC2 with a CTA per
// group of channels, each CTA serving every image -- the weights are read once per CTA instead of once per image, and both
// launches fill the machine.  Same arithmetic and roundings as the fused tail (bf16 mean, bf16 hidden, bf16 gate).
static constexpr int kFc1Units = 4;   // hidden units per CTA
static constexpr int kFc2Chans = 32;  // channels per CTA
__global__ void __launch_bounds__(256) hn_se_fc1_kernel(const bf16* __restrict__ mean, int N, int C, SeFc fc, bf16* __restrict__ hidden) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    extern __shared__ float sm[];  // [kFc1Units][C] weights as fp32
    const int s0 = blockIdx.x * kFc1Units;
    for (int i = threadIdx.x; i < kFc1Units * C; i += blockDim.x) {
        const int u = i / C, c = i - u * C;
        sm[i] = s0 + u < fc.S ? __bfloat162float(fc.w1[(long long)(s0 + u) * C + c]) : 0.0f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, CV = C >> 3;
    for (int n = warp; n < N; n += 8) {
        float acc[kFc1Units];
#pragma unroll
        for (int u = 0; u < kFc1Units; ++u) acc[u] = 0.0f;
        for (int v = lane; v < CV; v += 32) {
            float m[8];
            load8(mean + (long long)n * C + v * 8, m);
#pragma unroll
            for (int u = 0; u < kFc1Units; ++u) {
                const float* w = sm + u * C + v * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[u] = fmaf(w[j], m[j], acc[u]);
            }
        }
#pragma unroll
        for (int u = 0; u < kFc1Units; ++u) {
            float a = acc[u];
            for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
            if (lane == 0 && s0 + u < fc.S) hidden[(long long)n * fc.S + s0 + u] = __float2bfloat16(fmaxf(a + fc.b1[s0 + u], 0.0f));
        }
    }
}
__global__ void __launch_bounds__(256) hn_se_fc2_kernel(const bf16* __restrict__ hidden, int N, int C, SeFc fc) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    extern __shared__ float sm[];  // [kFc2Chans][S + 1] weights as fp32 (padded rows: no bank conflicts across channels)
    const int c0 = blockIdx.x * kFc2Chans, S = fc.S, ld = S + 1;
    for (int i = threadIdx.x; i < kFc2Chans * S; i += blockDim.x) {
        const int u = i / S, s = i - u * S;
        sm[u * ld + s] = c0 + u < C ? __bfloat162float(fc.w2[(long long)(c0 + u) * S + s]) : 0.0f;
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int c = c0 + lane;
    const float b = c < C ? fc.b2[c] : 0.0f;
    for (int n = warp; n < N; n += 8) {
        const bf16* h = hidden + (long long)n * S;
        float acc = 0.0f;
        for (int s = 0; s < S; s += 8) {
            float hv[8];
            load8(h + s, hv);  // the same 16 bytes for every lane: one broadcast transaction
#pragma unroll
            for (int j = 0; j < 8; ++j) acc = fmaf(sm[lane * ld + s + j], hv[j], acc);
        }
        if (c < C) fc.gate[(long long)n * C + c] = __float2bfloat16(1.0f / (1.0f + expf(-(acc + b))));
    }
}
static int g_se_split_fc = 0;  // measured on B200 (batch 32): the gate of a stage-4 block takes 73 us split vs 36 us fused -> off
extern "C" void hn_se_set_split_fc(int on) { g_se_split_fc = on ? 1 : 0; }
extern "C" int hn_se_pool_num_launches(const hn_se_pool_desc* d) { return (d && d->S > 0 && g_se_split_fc) ? 3 : 1; }

__global__ void hn_se_scale_kernel(View x, const bf16* __restrict__ scale) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    const int CV = x.C >> 3;
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned total = (unsigned)x.N * x.H * x.W * CV;
    if (idx >= total) return;
    int cv = (int)(idx % (unsigned)CV);
    unsigned t = idx / (unsigned)CV;
    int xx = (int)(t % (unsigned)x.W);
    t /= (unsigned)x.W;
    int y = (int)(t % (unsigned)x.H);
    int n = (int)(t / (unsigned)x.H);
    bf16* p = const_cast<bf16*>(vptr(x, n, y, xx, cv * 8));
    float f[8], sc[8];
    load8(p, f);
    load8(scale + (long long)n * x.C + cv * 8, sc);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] *= sc[j];
    store8(p, f);
}

extern "C" int hn_se_pool_fwd(const hn_se_pool_desc* d, void* stream) {
    HN_REQUIRE(d && d->partial && d->counter && d->mean, "se_pool: bad descriptor");
    if (int rc = check_view(d->x, "se_pool.x")) return rc;
    const int C = d->x.C, CV = C / 8, HW = d->x.H * d->x.W;
    HN_REQUIRE(CV <= kSeThreads, "se_pool: C=%d too wide", C);
    HN_REQUIRE(d->pix_per_block >= 128 && d->pix_per_block % 128 == 0, "se_pool: pix_per_block must be a multiple of 128");
    const int gz = hn_cdiv(CV, 32);  // channel groups of at most 32 vectors (>= 16 pixel lanes per CTA)
    const int CG = hn_cdiv(CV, gz);
    dim3 grid(hn_cdiv(HW, d->pix_per_block), d->x.N, gz);
    int lanes = kSeThreads / CG;
    size_t smem = (size_t)lanes * CG * 8 * sizeof(float);
    HN_REQUIRE(smem <= 48 * 1024, "se_pool: shared memory");
    SeFc fc;
    memset(&fc, 0, sizeof(fc));
    if (d->S > 0) {
        HN_REQUIRE(d->S % 8 == 0 && d->w1 && d->b1 && d->w2 && d->b2 && d->gate, "se_pool: FC stage needs S %% 8 == 0 and all five pointers");
        HN_REQUIRE(((reinterpret_cast<uintptr_t>(d->w1) | reinterpret_cast<uintptr_t>(d->w2)) & 15) == 0, "se_pool: FC weights must be 16-byte aligned");
        fc.S = d->S;
        fc.w1 = reinterpret_cast<const bf16*>(d->w1); fc.b1 = d->b1;
        fc.w2 = reinterpret_cast<const bf16*>(d->w2); fc.b2 = d->b2;
        fc.gate = reinterpret_cast<bf16*>(d->gate);
        const size_t need = (size_t)(((C + 7) & ~7) + d->S) * sizeof(float);  // mean + hidden reuse the pooling scratch
        if (need > smem) smem = need;
    }
    if (fc.S && g_se_split_fc && (size_t)kFc1Units * C * 4 <= 48 * 1024 && (size_t)kFc2Chans * (fc.S + 1) * 4 <= 48 * 1024) {
        // pooling only, then the two FC layers as batched launches; `partial` is free again once the pool kernel is done
        // and holds the hidden activations (bf16 [N][S], needs N*S*2 <= N*chunks*C*4 bytes: S <= C always)
        SeFc none;
        memset(&none, 0, sizeof(none));
        const size_t pool_smem = (size_t)lanes * CG * 8 * sizeof(float);
        HN_CHECK_CUDA(hn_launch(hn_se_pool_kernel, dim3(grid), dim3(kSeThreads), pool_smem, reinterpret_cast<cudaStream_t>(stream),
            to_view(d->x), d->partial, d->counter, reinterpret_cast<bf16*>(d->mean), 1.0f / (float)HW, d->pix_per_block, none));
        bf16* hidden = reinterpret_cast<bf16*>(d->partial);
        HN_CHECK_CUDA(hn_launch(hn_se_fc1_kernel, dim3(hn_cdiv(fc.S, kFc1Units)), dim3(256), (size_t)kFc1Units * C * 4, reinterpret_cast<cudaStream_t>(stream),
            reinterpret_cast<const bf16*>(d->mean), (int)d->x.N, C, fc, hidden));
        HN_CHECK_CUDA(hn_launch(hn_se_fc2_kernel, dim3(hn_cdiv(C, kFc2Chans)), dim3(256), (size_t)kFc2Chans * (fc.S + 1) * 4, reinterpret_cast<cudaStream_t>(stream),
            reinterpret_cast<const bf16*>(hidden), (int)d->x.N, C, fc));
        HN_CHECK_CUDA(cudaGetLastError());
        return HN_OK;
    }
    HN_CHECK_CUDA(hn_launch(hn_se_pool_kernel, dim3(grid), dim3(kSeThreads), (size_t)(smem), reinterpret_cast<cudaStream_t>(stream),
        to_view(d->x), d->partial, d->counter, reinterpret_cast<bf16*>(d->mean), 1.0f / (float)HW, d->pix_per_block, fc));
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

extern "C" int hn_se_scale_fwd(const hn_se_scale_desc* d, void* stream) {
    HN_REQUIRE(d && d->scale, "se_scale: bad descriptor");
    if (int rc = check_view(d->x, "se_scale.x")) return rc;
    long long total = (long long)d->x.N * d->x.H * d->x.W * (d->x.C / 8);
    HN_REQUIRE(total < 0x7fffffffLL, "se_scale: too many work items for one launch");
    HN_CHECK_CUDA(hn_launch(hn_se_scale_kernel, dim3(hn_cdiv(total, 256)), dim3(256), (size_t)(0), reinterpret_cast<cudaStream_t>(stream), 
        to_view(d->x), reinterpret_cast<const bf16*>(d->scale)));
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ---- squeeze-excite in ONE launch (small maps): [grouped 3x3 +] pool + FC1 + FC2 + channel scaling -------------------
// The two-launch form above ends in a serial tail: ONE block per image pulls both FC weight matrices (stage 4: 2 x 438 KB)
// through its SM's L2 port, 32 SMs busy and 116 idle, and the scaling is a separate launch behind it.  Here a cluster of
// kSeCl CTAs serves one image and every CTA owns a slice of the CHANNELS: it pools its channels over all pixels (keeping
// the slice in shared memory), computes 1/kSeCl of the hidden units, then the gate of its own channels, and scales its
// slice in place.  The mean and the hidden vector cross the cluster through distributed shared memory (st.shared::cluster into
// every CTA's copy, release / acquire cluster barriers); every sum runs in a fixed order.  Roundings as in the two-launch form: bf16 mean, bf16 hidden, bf16 gate.
//
// kConv: the block's grouped 3x3 convolution (group width 8, stride 1, folded BN + ReLU; anynet.py:60-66) runs in front, in
// the same launch: a group is exactly one 8-channel vector, so a channel slice is self-contained.  The CTA loads its slice
// of the 1x1 conv's output into shared memory and runs the 3x3 as warp-level mma.sync m16n8k16 (M = 16 pixels, N = the
// group's 8 output channels, K = 2 taps x 8 input channels; 5 k-steps, the 10th tap is zero) -- at 13-21 MFLOP per image
// this conv was a 22-27 us launch of the big GEMM kernel for ~1 us of math.  The result never leaves the SM before it is
// pooled, gated and scaled.
static constexpr int kSeCl = 4;
static constexpr int kGwTaps = 10;                    // 9 taps padded to an even count
static constexpr int kGwGroupElems = kGwTaps * 64;    // packed weights of one group: [tap][oc][ic] bf16
struct SeFusedParams {
    View x;         // the tensor that is pooled and scaled in place (kConv: the conv's output)
    View in;        // kConv: the conv's input
    const bf16* wg; // kConv: [C/8][10][8 oc][8 ic]
    const float* cbias;  // kConv: [C]
    bf16* mean;     // [N][C]
    SeFc fc;
    float inv_hw;
    int cvs;   // 8-channel vectors per CTA (the last CTA of a cluster may own fewer)
    int uss;   // hidden units per CTA
    int tile;  // the CTA's slice [HW][cvs] stays in shared memory between pooling and scaling
    long long* dbg;  // optional [CTA][8] globaltimer stamps at the phase boundaries (hn_se_fused_set_debug)
};

__device__ __forceinline__ void hn_mma_m16n8k16_bf16(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void se_stamp(const SeFusedParams& p, int k) {
    if (p.dbg && threadIdx.x == 0) {
        long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.dbg[(long long)blockIdx.x * 8 + k] = t;
    }
}

// one fp32 value into the same shared-memory slot of every CTA of the cluster (distributed shared memory)
__device__ __forceinline__ void se_bcast_f32(float* slot, float v) {
    const uint32_t a = hn_smem_u32(slot);
#pragma unroll
    for (int rk = 0; rk < kSeCl; ++rk) asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(hn_mapa(a, (uint32_t)rk)), "f"(v) : "memory");
}

template <bool kConv>
__global__ void __launch_bounds__(kSeThreads) hn_se_fused_kernel(const __grid_constant__ SeFusedParams p) {
    hn_pdl_launch_dependents();
    hn_pdl_wait();
    // every CTA of the cluster must be running before a peer stores into its shared memory: arrive now, wait before the first
    // remote store
    asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
    se_stamp(p, 0);
    extern __shared__ __align__(16) uint8_t se_smem[];
    const View& x = p.x;
    const int C = x.C, CV = C >> 3, HW = x.H * x.W, S = p.fc.S;
    const int r = (int)hn_cluster_ctarank(), n = blockIdx.x / kSeCl;
    const int cvs = p.cvs, v0 = min(CV, r * cvs), nvs = min(cvs, CV - v0);  // this CTA's vectors [v0, v0 + nvs)
    const int u0 = min(S, r * p.uss), nu = min(p.uss, S - u0);               // ... and hidden units [u0, u0 + nu)
    const int lanes = kSeThreads / cvs;
    const int pv = kConv ? (cvs | 1) : cvs;  // row pitch of the shared-memory tiles in 16-byte vectors (conv: odd, see below)
    float* s_mean = reinterpret_cast<float*>(se_smem);  // [C rounded up to 8]
    float* s_hid = s_mean + ((C + 7) & ~7);             // [S]
    float* s_gate = s_hid + S;                          // [cvs * 8]
    float* s_acc = s_gate + cvs * 8;                    // [lanes][cvs * 8]
    uint4* s_tile = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(s_acc + lanes * cvs * 8) + 15) & ~uintptr_t(15));  // [HW][pv]
    {   // the weight rows this CTA will read were last touched a whole step ago: pull them into the L2 while pooling
        const char* w1p = reinterpret_cast<const char*>(p.fc.w1 + (long long)u0 * C);
        const char* w2p = reinterpret_cast<const char*>(p.fc.w2 + (long long)v0 * 8 * S);
        const int l1 = (nu * C * 2 + 127) >> 7, l2 = (nvs * 8 * S * 2 + 127) >> 7;
        for (int l = threadIdx.x; l < l1 + l2; l += blockDim.x) {
            const char* q = l < l1 ? w1p + (l << 7) : w2p + ((l - l1) << 7);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        }
    }
    const int vl = threadIdx.x % cvs, pl = threadIdx.x / cvs;
    const bool active = vl < nvs && pl < lanes;
    bf16* base = const_cast<bf16*>(x.ptr) + n * x.sn + (v0 + vl) * 8;
    if constexpr (kConv) {
        // ---- grouped 3x3 of this CTA's groups: input slice and weights into shared memory, mma.sync, result into s_tile ----
        uint4* s_in = s_tile + HW * pv;                                     // [HW][pv]
        uint32_t* s_w = reinterpret_cast<uint32_t*>(s_in + HW * pv);        // [nvs][10][8 oc][4 ic pairs]
        uint4* s_zero = reinterpret_cast<uint4*>(s_w + cvs * (kGwGroupElems / 2));  // one zero vector: what a padding tap reads
        float* s_cb = reinterpret_cast<float*>(s_zero + 1);                 // [nvs * 8] conv bias of the slice
        if (threadIdx.x == 0) *s_zero = make_uint4(0u, 0u, 0u, 0u);
        for (int i = threadIdx.x; i < nvs * 8; i += blockDim.x) s_cb[i] = p.cbias[v0 * 8 + i];
        if (active) {
            const bf16* ib = p.in.ptr + n * p.in.sn + (v0 + vl) * 8;
            for (int px = pl; px < HW; px += 8 * lanes) {  // eight independent loads in flight (a load-store loop would serialise them)
                uint4 v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const int q = px + u * lanes, y = q / x.W, xx = q - y * x.W;
                    if (q < HW) v[u] = *reinterpret_cast<const uint4*>(ib + y * p.in.sy + xx * p.in.sx);
                }
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (px + u * lanes < HW) s_in[(px + u * lanes) * pv + vl] = v[u];
            }
        }
        {
            const uint4* wsrc = reinterpret_cast<const uint4*>(p.wg + (long long)v0 * kGwGroupElems);
            uint4* wdst = reinterpret_cast<uint4*>(s_w);
            for (int i = threadIdx.x; i < nvs * (kGwGroupElems / 8); i += blockDim.x) wdst[i] = __ldg(wsrc + i);
        }
        __syncthreads();
        se_stamp(p, 1);
        // Work item = (16-pixel m tile, group), m-tile-major; a warp takes a contiguous run.  The A fragment of a k-step (two taps x 8 channels) is ONE ldmatrix.x4: a
        // matrix row is a pixel's 16-byte group vector, lane l supplies the address of pixel (l & 15) of the tile at tap
        // 2j + (l >> 4); the odd row pitch keeps the eight rows of a matrix in different bank groups.
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gid = lane >> 2, tq = lane & 3;
        const int nwarps = kSeThreads / 32, MT = (HW + 15) >> 4;
        const int npair = MT * nvs, per = (npair + nwarps - 1) / nwarps;
        const int q0 = warp * per, q1 = min(npair, q0 + per);
        const int W = x.W, H = x.H;
        const int row_bytes = pv * 16;
        const uint32_t in_s = hn_smem_u32(s_in), zero_s = hn_smem_u32(s_zero);
        uint8_t* out_b = reinterpret_cast<uint8_t*>(s_tile) + tq * 4;
        const int lrow = lane & 15, ltap = lane >> 4;
        // This loop is bound by instruction issue (16 warps of integer address arithmetic around 5 MMAs), not by the tensor
        // pipe: items run m-tile-major so the five tap offsets of a lane are computed once per m tile and shared by its groups.
        int cur_mt = -1;
        int aoff[kGwTaps / 2];  // byte offset (within a row's first group) of this lane's ldmatrix row per k-step, -1 = padding
        int mt = q0 / nvs, g = q0 - mt * nvs;
        for (int q = q0; q < q1; ++q, ++g) {
            if (g == nvs) { g = 0; ++mt; }
            if (mt != cur_mt) {
                cur_mt = mt;
                const int ar = mt * 16 + lrow, ay = ar / W, ax = ar - ay * W;  // the pixel whose row address this lane supplies
#pragma unroll
                for (int j = 0; j < kGwTaps / 2; ++j) {
                    const int t = 2 * j + ltap, dy = t / 3 - 1, dx = t - (t / 3) * 3 - 1;
                    const bool ok = t < 9 && ar < HW && ay + dy >= 0 && ay + dy < H && ax + dx >= 0 && ax + dx < W;
                    aoff[j] = ok ? ((ay + dy) * W + ax + dx) * row_bytes : -1;
                }
            }
            const uint32_t in_g = in_s + (uint32_t)(g * 16);
            const uint32_t* wq = s_w + g * (kGwGroupElems / 2) + gid * 4 + tq;
            uint32_t a[kGwTaps / 2][4];
#pragma unroll
            for (int j = 0; j < kGwTaps / 2; ++j) {
                const uint32_t addr = aoff[j] >= 0 ? in_g + (uint32_t)aoff[j] : zero_s;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                             : "=r"(a[j][0]), "=r"(a[j][1]), "=r"(a[j][2]), "=r"(a[j][3]) : "r"(addr));
            }
            float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
            for (int j = 0; j < kGwTaps / 2; ++j)
                hn_mma_m16n8k16_bf16(acc, a[j][0], a[j][1], a[j][2], a[j][3], wq[(2 * j) * 32], wq[(2 * j + 1) * 32]);
            const int r0 = mt * 16 + gid, r1 = r0 + 8;
            const float2 bz = *reinterpret_cast<const float2*>(s_cb + g * 8 + tq * 2);
            uint8_t* o = out_b + r0 * row_bytes + g * 16;
            if (r0 < HW) *reinterpret_cast<uint32_t*>(o) = hn_pack_bf16x2(fmaxf(acc[0] + bz.x, 0.0f), fmaxf(acc[1] + bz.y, 0.0f));
            if (r1 < HW) *reinterpret_cast<uint32_t*>(o + 8 * row_bytes) = hn_pack_bf16x2(fmaxf(acc[2] + bz.x, 0.0f), fmaxf(acc[3] + bz.y, 0.0f));
        }
        __syncthreads();
    }
    se_stamp(p, 2);
    // ---- pooling of this CTA's channels ----
    auto fetch = [&](int q) -> uint4 {
        if constexpr (kConv) return s_tile[q * pv + vl];
        const int y = q / x.W, xx = q - y * x.W;
        const uint4 v = *reinterpret_cast<const uint4*>(base + y * x.sy + xx * x.sx);
        if (p.tile) s_tile[q * pv + vl] = v;
        return v;
    };
    {
        float acc[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = 0.0f;
        if (active) {
            int px = pl;
            for (; px + 3 * lanes < HW; px += 4 * lanes) {  // four independent loads in flight
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = fetch(px + u * lanes);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float2 a = hn_unpack_bf16x2(v[u].x), b = hn_unpack_bf16x2(v[u].y), c2 = hn_unpack_bf16x2(v[u].z), d2 = hn_unpack_bf16x2(v[u].w);
                    acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
                    acc[4] += c2.x; acc[5] += c2.y; acc[6] += d2.x; acc[7] += d2.y;
                }
            }
            for (; px < HW; px += lanes) {
                const uint4 v = fetch(px);
                const float2 a = hn_unpack_bf16x2(v.x), b = hn_unpack_bf16x2(v.y), c2 = hn_unpack_bf16x2(v.z), d2 = hn_unpack_bf16x2(v.w);
                acc[0] += a.x; acc[1] += a.y; acc[2] += b.x; acc[3] += b.y;
                acc[4] += c2.x; acc[5] += c2.y; acc[6] += d2.x; acc[7] += d2.y;
            }
        }
        if (pl < lanes) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s_acc[(pl * cvs + vl) * 8 + j] = acc[j];
        }
    }
    __syncthreads();
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");  // the peers are running (arrived at kernel start)
    for (int c = threadIdx.x; c < nvs * 8; c += blockDim.x) {
        float a = 0.0f;
        for (int l = 0; l < lanes; ++l) a += s_acc[l * cvs * 8 + c];  // fixed order: deterministic
        const bf16 mb = __float2bfloat16(a * p.inv_hw);
        p.mean[(long long)n * C + v0 * 8 + c] = mb;
        se_bcast_f32(s_mean + v0 * 8 + c, __bfloat162float(mb));  // the FC layers read the rounded mean, as separate launches would
    }
    se_stamp(p, 3);
    hn_cluster_sync();  // release / acquire: the four slices of the mean are in every CTA's s_mean
    se_stamp(p, 4);
    // ---- FC1 + ReLU: this CTA's hidden units ----
    {
        int T = 1;  // adjacent lanes per hidden unit
        while (T < 32 && p.uss * T * 2 <= (int)blockDim.x) T *= 2;
        const int nv = C >> 3, per = (nv + T - 1) / T;
        for (int s0 = 0; s0 < nu; s0 += blockDim.x / T) {
            const int srow = s0 + threadIdx.x / T, part = threadIdx.x % T;
            float acc = 0.0f;
            if (srow < nu) acc = se_dot(p.fc.w1 + (long long)(u0 + srow) * C, s_mean, min(part * per, nv), min(part * per + per, nv));
            for (int o = T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (srow < nu && part == 0)
                se_bcast_f32(s_hid + u0 + srow, __bfloat162float(__float2bfloat16(fmaxf(acc + p.fc.b1[u0 + srow], 0.0f))));
        }
    }
    hn_cluster_sync();
    se_stamp(p, 5);
    // ---- FC2 + sigmoid: the gate of this CTA's channels ----
    {
        const int nch = nvs * 8;
        int T = 1;
        while (T < 32 && cvs * 8 * T * 2 <= (int)blockDim.x) T *= 2;
        const int nv = S >> 3, per = (nv + T - 1) / T;
        for (int c0 = 0; c0 < nch; c0 += blockDim.x / T) {
            const int cl = c0 + threadIdx.x / T, part = threadIdx.x % T;
            float acc = 0.0f;
            if (cl < nch) acc = se_dot(p.fc.w2 + (long long)(v0 * 8 + cl) * S, s_hid, min(part * per, nv), min(part * per + per, nv));
            for (int o = T >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (cl < nch && part == 0) {
                const bf16 g = __float2bfloat16(1.0f / (1.0f + expf(-(acc + p.fc.b2[v0 * 8 + cl]))));
                p.fc.gate[(long long)n * C + v0 * 8 + cl] = g;
                s_gate[cl] = __bfloat162float(g);
            }
        }
    }
    __syncthreads();
    se_stamp(p, 6);
    // ---- x *= gate, in place (kConv: x is written here for the first time) ----
    if (active) {
        float g[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) g[j] = s_gate[vl * 8 + j];
        for (int px = pl; px < HW; px += lanes) {
            const int y = px / x.W, xx = px - y * x.W;
            bf16* dst = base + y * x.sy + xx * x.sx;
            const uint4 v = (kConv || p.tile) ? s_tile[px * pv + vl] : *reinterpret_cast<const uint4*>(dst);
            const float2 a = hn_unpack_bf16x2(v.x), b = hn_unpack_bf16x2(v.y), c2 = hn_unpack_bf16x2(v.z), d2 = hn_unpack_bf16x2(v.w);
            const float f[8] = {a.x * g[0], a.y * g[1], b.x * g[2], b.y * g[3], c2.x * g[4], c2.y * g[5], d2.x * g[6], d2.y * g[7]};
            store8(dst, f);
        }
    }
    se_stamp(p, 7);
}

static long long* g_se_dbg = nullptr;
extern "C" void hn_se_fused_set_debug(long long* buf) { g_se_dbg = buf; }  // device buffer of 8 stamps per CTA (N * 4 CTAs), or null

static constexpr size_t kSeFusedMaxSmem = 220 * 1024;
// shared-memory bytes of the fused launch; *tile = whether the channel slice fits next to the vectors (conv: it must, twice,
// plus the packed weights of the slice; 0 is returned when it does not)
static size_t se_fused_smem(int HW, int C, int S, bool conv, int* tile) {
    const int CV = C / 8, cvs = hn_cdiv(CV, kSeCl), lanes = kSeThreads / cvs;
    const size_t vec = (size_t)(((C + 7) & ~7) + S + cvs * 8 + lanes * cvs * 8) * sizeof(float) + 16;
    const size_t t = (size_t)HW * cvs * 16;
    if (conv) {
        const size_t need = vec + 2 * (size_t)HW * (cvs | 1) * 16 + (size_t)cvs * kGwGroupElems * 2 + 16 + (size_t)cvs * 8 * sizeof(float);
        *tile = 1;
        return need <= kSeFusedMaxSmem ? need : 0;
    }
    *tile = vec + t <= kSeFusedMaxSmem;
    return vec + (*tile ? t : 0);
}
static int se_fused_shape_ok(int H, int W, int C, int S) {
    if (H <= 0 || W <= 0 || C % 8 != 0 || S <= 0 || S % 8 != 0) return 0;
    const int CV = C / 8, cvs = hn_cdiv(CV, kSeCl);
    if (CV < 2 * kSeCl || cvs > 64) return 0;           // every CTA of the cluster needs channels of its own
    if ((long long)H * W > 4096) return 0;              // larger maps: many blocks per image (hn_se_pool_fwd)
    return 1;
}
extern "C" int hn_se_fused_supported(int H, int W, int C, int S) {
    int tile;
    return se_fused_shape_ok(H, W, C, S) && se_fused_smem(H * W, C, S, false, &tile) <= kSeFusedMaxSmem;
}
extern "C" int hn_gconv_se_supported(int H, int W, int C, int S) {
    int tile;
    return se_fused_shape_ok(H, W, C, S) && se_fused_smem(H * W, C, S, true, &tile) != 0;
}

static int se_fused_launch(const hn_se_pool_desc* d, const hn_gconv_se_desc* cv, void* stream) {
    const char* what = cv ? "gconv_se" : "se_fused";
    HN_REQUIRE(d->partial && d->mean && d->gate && d->w1 && d->b1 && d->w2 && d->b2, "%s: bad descriptor", what);
    if (int rc = check_view(d->x, "se_fused.x")) return rc;
    HN_REQUIRE(cv ? hn_gconv_se_supported(d->x.H, d->x.W, d->x.C, d->S) : hn_se_fused_supported(d->x.H, d->x.W, d->x.C, d->S),
               "%s: unsupported shape %dx%dx%d, S=%d", what, d->x.H, d->x.W, d->x.C, d->S);
    HN_REQUIRE(((reinterpret_cast<uintptr_t>(d->w1) | reinterpret_cast<uintptr_t>(d->w2)) & 15) == 0, "%s: FC weights must be 16-byte aligned", what);
    SeFusedParams p;
    memset(&p, 0, sizeof(p));
    p.x = to_view(d->x);
    if (cv) {
        if (int rc = check_view(cv->in, "gconv_se.in")) return rc;
        HN_REQUIRE(cv->in.N == d->x.N && cv->in.H == d->x.H && cv->in.W == d->x.W && cv->in.C == d->x.C, "gconv_se: input and output shapes differ");
        HN_REQUIRE(cv->weight && cv->bias && (reinterpret_cast<uintptr_t>(cv->weight) & 15) == 0, "gconv_se: weights");
        HN_REQUIRE(cv->in.ptr != d->x.ptr, "gconv_se: input and output must not alias (a CTA reads its neighbours' input pixels)");
        p.in = to_view(cv->in);
        p.wg = reinterpret_cast<const bf16*>(cv->weight);
        p.cbias = cv->bias;
    }
    p.mean = reinterpret_cast<bf16*>(d->mean);
    p.fc.S = d->S;
    p.fc.w1 = reinterpret_cast<const bf16*>(d->w1); p.fc.b1 = d->b1;
    p.fc.w2 = reinterpret_cast<const bf16*>(d->w2); p.fc.b2 = d->b2;
    p.fc.gate = reinterpret_cast<bf16*>(d->gate);
    p.inv_hw = 1.0f / (float)(d->x.H * d->x.W);
    p.cvs = hn_cdiv(d->x.C / 8, kSeCl);
    p.uss = hn_cdiv(d->S, kSeCl);
    p.dbg = g_se_dbg;
    const size_t smem = se_fused_smem(d->x.H * d->x.W, d->x.C, d->S, cv != nullptr, &p.tile);
    static std::once_flag once;
    static cudaError_t attr_rc = cudaSuccess;
    std::call_once(once, [] {
        attr_rc = cudaFuncSetAttribute(hn_se_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSeFusedMaxSmem);
        if (attr_rc == cudaSuccess)
            attr_rc = cudaFuncSetAttribute(hn_se_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSeFusedMaxSmem);
    });
    HN_CHECK_CUDA(attr_rc);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(d->x.N * kSeCl);
    cfg.blockDim = dim3(kSeThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = reinterpret_cast<cudaStream_t>(stream);
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kSeCl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_hn_pdl ? 2 : 1;
    if (cv) HN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, hn_se_fused_kernel<true>, p));
    else HN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, hn_se_fused_kernel<false>, p));
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}
extern "C" int hn_se_fused_fwd(const hn_se_pool_desc* d, void* stream) {
    HN_REQUIRE(d != nullptr, "se_fused: null descriptor");
    return se_fused_launch(d, nullptr, stream);
}
extern "C" int hn_gconv_se_fwd(const hn_gconv_se_desc* d, void* stream) {
    HN_REQUIRE(d != nullptr, "gconv_se: null descriptor");
    return se_fused_launch(&d->se, d, stream);
}

// ------------------------------------------------------------------------------------------------
// image pre-processing (demo.py:191-196): BGR u8 HWC -> resized, normalised fp32 planar RGB
// ------------------------------------------------------------------------------------------------
struct PreParams {
    const uint8_t* src;
    int N, sh, sw;
    long long pitch, istride;
    float* dst;
    int dh, dw;
    double scale_x, scale_y;  // 1 / (dst / src), as cv2 computes them
    int area2;                // exact 2x2 down-scale: cv2 switches INTER_LINEAR to INTER_AREA
};

// cv2's fixed-point coefficient pair for one destination coordinate (resize.cpp, INTER_RESIZE_COEF_BITS = 11)
__device__ __forceinline__ void pre_coeff(int d, double scale, int src, bool clamp_fraction, int& s, int& w0, int& w1) {
    float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);  // separate multiply and subtract, as the host code does
    s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    if (clamp_fraction) {
        if (s < 0) { f = 0.0f; s = 0; }
        if (s >= src - 1) { f = 0.0f; s = src - 1; }
    }
    w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));  // saturate_cast<short> = round half to even
    w1 = __float2int_rn(__fmul_rn(f, 2048.0f));
}

__global__ void __launch_bounds__(256) hn_preprocess_kernel(const PreParams p) {
    __shared__ float lut[3][256];  // value of every uint8 level per output channel (R, G, B), float64 arithmetic
    for (int i = threadIdx.x; i < 768; i += blockDim.x) {
        const int c = i >> 8, v = i & 255;
        const double mean = c == 0 ? 0.485 : (c == 1 ? 0.456 : 0.406), sd = c == 0 ? 0.229 : (c == 1 ? 0.224 : 0.225);
        lut[c][v] = (float)__ddiv_rn(__dsub_rn(__ddiv_rn((double)(float)v, 255.0), mean), sd);
    }
    __syncthreads();
    const unsigned idx = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned total = (unsigned)p.N * p.dh * p.dw;
    if (idx >= total) return;
    const int x = (int)(idx % (unsigned)p.dw);
    const int y = (int)((idx / (unsigned)p.dw) % (unsigned)p.dh);
    const int n = (int)(idx / ((unsigned)p.dw * p.dh));
    const uint8_t* img = p.src + (long long)n * p.istride;
    int out[3];
    if (p.area2) {
        const uint8_t* r0 = img + (long long)(2 * y) * p.pitch + (long long)(2 * x) * 3;
        const uint8_t* r1 = r0 + p.pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) out[c] = ((int)r0[c] + (int)r0[3 + c] + (int)r1[c] + (int)r1[3 + c] + 2) >> 2;
    } else {
        int sx, ax0, ax1, sy, by0, by1;
        pre_coeff(x, p.scale_x, p.sw, true, sx, ax0, ax1);
        pre_coeff(y, p.scale_y, p.sh, false, sy, by0, by1);
        const int sx1 = min(sx + 1, p.sw - 1);
        const int y0 = min(max(sy, 0), p.sh - 1), y1 = min(max(sy + 1, 0), p.sh - 1);
        const uint8_t* r0 = img + (long long)y0 * p.pitch;
        const uint8_t* r1 = img + (long long)y1 * p.pitch;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int h0 = (int)r0[sx * 3 + c] * ax0 + (int)r0[sx1 * 3 + c] * ax1;
            const int h1 = (int)r1[sx * 3 + c] * ax0 + (int)r1[sx1 * 3 + c] * ax1;
            out[c] = (((by0 * (h0 >> 4)) >> 16) + ((by1 * (h1 >> 4)) >> 16) + 2) >> 2;
        }
    }
    const long long plane = (long long)p.dh * p.dw;
    float* o = p.dst + (long long)n * 3 * plane + (long long)y * p.dw + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c * plane] = lut[c][out[2 - c] & 255];  // plane c = R,G,B <- source byte 2-c
}

extern "C" int hn_preprocess_fwd(const hn_preprocess_desc* d, void* stream) {
    HN_REQUIRE(d && d->src && d->dst, "preprocess: null pointer");
    HN_REQUIRE(d->N >= 0 && d->src_h >= 1 && d->src_w >= 1 && d->dst_h >= 1 && d->dst_w >= 1, "preprocess: bad sizes");
    HN_REQUIRE(d->src_pitch >= (int64_t)d->src_w * 3 && d->src_stride >= d->src_pitch * (d->src_h - 1) + (int64_t)d->src_w * 3,
               "preprocess: pitch/stride smaller than the image");
    if (d->N == 0) return HN_OK;
    const long long total = (long long)d->N * d->dst_h * d->dst_w;
    HN_REQUIRE(total < 0x7fffffffLL, "preprocess: too many output pixels for one launch");
    PreParams p;
    p.src = d->src; p.N = d->N; p.sh = d->src_h; p.sw = d->src_w; p.pitch = d->src_pitch; p.istride = d->src_stride;
    p.dst = d->dst; p.dh = d->dst_h; p.dw = d->dst_w;
    p.scale_x = 1.0 / ((double)d->dst_w / (double)d->src_w);
    p.scale_y = 1.0 / ((double)d->dst_h / (double)d->src_h);
    p.area2 = (d->src_w == 2 * d->dst_w && d->src_h == 2 * d->dst_h) ? 1 : 0;
    hn_preprocess_kernel<<<hn_cdiv(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

