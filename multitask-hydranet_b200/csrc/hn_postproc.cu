// Post-processing of the three heads on the GPU.
//   seg : arg-max over classes (torch.argmax semantics: first maximum, NaN counts as maximum)
//   det : box decode + clip + score threshold + class-aware greedy NMS with torchvision's exact
//         fp32 arithmetic and tie order
//   lane: per-anchor polyline decode + greedy lane NMS with the reference's fp32/py-float arithmetic
// Arithmetic that must be bit-exact uses __f*_rn intrinsics so nvcc cannot contract it into FMAs.
#include <cooperative_groups.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "hn_ops.h"

// ------------------------------------------------------------------------------------------------
// segmentation arg-max
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool greater_or_nan(float v, float best) {
    return (v > best) || (v != v && best == best);
}

__global__ void hn_seg_argmax_kernel(const float* __restrict__ logits, int C, long long HW, long long total_px,
                                     int64_t* __restrict__ out64, uint8_t* __restrict__ out8) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i >= total_px) return;
    long long n = i / HW, p = i - n * HW;
    const float* base = logits + n * C * HW + p;
    if (p + 3 < HW && (HW & 3) == 0) {
        float4 best = *reinterpret_cast<const float4*>(base);
        int b0 = 0, b1 = 0, b2 = 0, b3 = 0;
        for (int k = 1; k < C; ++k) {
            float4 v = *reinterpret_cast<const float4*>(base + k * HW);
            if (greater_or_nan(v.x, best.x)) { best.x = v.x; b0 = k; }
            if (greater_or_nan(v.y, best.y)) { best.y = v.y; b1 = k; }
            if (greater_or_nan(v.z, best.z)) { best.z = v.z; b2 = k; }
            if (greater_or_nan(v.w, best.w)) { best.w = v.w; b3 = k; }
        }
        if (out64) { out64[i] = b0; out64[i + 1] = b1; out64[i + 2] = b2; out64[i + 3] = b3; }
        if (out8) *reinterpret_cast<uchar4*>(out8 + i) = make_uchar4((uint8_t)b0, (uint8_t)b1, (uint8_t)b2, (uint8_t)b3);
    } else {
        for (int j = 0; j < 4 && i + j < total_px; ++j) {
            long long ii = i + j;
            long long nn = ii / HW, pp = ii - nn * HW;
            const float* b = logits + nn * C * HW + pp;
            float best = b[0];
            int bi = 0;
            for (int k = 1; k < C; ++k) {
                float v = b[k * HW];
                if (greater_or_nan(v, best)) { best = v; bi = k; }
            }
            if (out64) out64[ii] = bi;
            if (out8) out8[ii] = (uint8_t)bi;
        }
    }
}

extern "C" int hn_seg_argmax(const float* logits, int32_t N, int32_t C, int64_t HW, int64_t* out_i64, uint8_t* out_u8,
                             void* stream) {
    HN_REQUIRE(N >= 0 && C >= 1 && C <= 255 && HW >= 0, "seg_argmax: bad arguments");
    long long total = (long long)N * HW;
    if (total == 0) return HN_OK;
    HN_REQUIRE(logits && (out_i64 || out_u8), "seg_argmax: bad arguments");
    HN_REQUIRE((reinterpret_cast<uintptr_t>(logits) & 15) == 0, "seg_argmax: logits must be 16-byte aligned");
    hn_seg_argmax_kernel<<<hn_cdiv(hn_cdiv(total, 4), 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        logits, C, HW, total, out_i64, out_u8);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

__global__ void hn_u8_to_i64_kernel(const uint8_t* __restrict__ in, int64_t* __restrict__ out, long long n) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}
extern "C" int hn_u8_to_i64(const uint8_t* in, int64_t* out, int64_t n, void* stream) {
    HN_REQUIRE(in && out && n >= 0, "u8_to_i64: bad arguments");
    if (n == 0) return HN_OK;
    hn_u8_to_i64_kernel<<<hn_cdiv(n, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(in, out, n);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// detection
// ------------------------------------------------------------------------------------------------
// sort key: [img:8][cls:4][~score:32][anchor:20]; invalid (below threshold) = all ones
static constexpr int kAnchorBits = 20, kScoreShift = 20, kClsShift = 52, kImgShift = 56;
static constexpr int kMaxCls = 16;

__device__ __forceinline__ uint32_t float_desc_key(float f) {
    uint32_t u = __float_as_uint(f);
    u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;  // ascending-sortable
    return ~u;                                    // descending
}
__device__ __forceinline__ int float_to_ordered(float f) {
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7FFFFFFF;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7FFFFFFF); }

struct DetWs {
    float4* boxes;      // [N*A] decoded + clipped
    float* scores;      // [N*A]
    uint64_t* keys;     // [N*A]
    uint64_t* keys_alt; // [N*A]
    int* n_cand;        // [N]
    int* max_coord;     // [N] ordered-int of the max coordinate over candidates
    int* seg_start;     // [N*16]
    int* seg_end;       // [N*16]
    int* seg_kept;      // [N*16]
    float4* kept_boxes; // [N*A] compacted per segment (NMS-space coordinates)
    uint64_t* kept_keys;// [N*A] compacted per segment
    int* kept_next;     // [N*A] linked list through the kept boxes of one grid cell
    int* heads;         // [N*16][kGridHeads] list heads of the per-level spatial grids (-1 = empty)
    void* cub_tmp;
    size_t cub_bytes;
    long long* dbg;     // optional [N*16][8] cycle counters of the NMS kernel
    // ---- parallel (dependency-round) NMS ----
    float4* sbox;        // [N*A] NMS-space boxes in sorted order
    uint32_t* ckey;      // [N*A] (segment, grid cell) key per sorted position, and its sorted copy
    uint32_t* ckey_alt;
    int* cval;           // [N*A] sorted position, and the cell-ordered permutation
    int* cval_alt;
    float4* cbox;        // [N*A] boxes in cell order
    int* cell_cnt;       // [N*16*kCellStride + 1] members per (segment, cell) ...
    int* cell_begin;     // ... and its exclusive scan = first cell-order index of every cell (rows are contiguous)
    int* preds;          // [N*A][kMaxPreds] earlier boxes of the same class with IoU > thr
    int* npred;          // [N*A] predecessors found by the box's own scan: preds[0 .. npred)
    int* nfor;           // [N*A] predecessors recorded by other boxes' scans: preds[kMaxPreds-1 .. kMaxPreds-nfor] (downwards)
    unsigned char* status; // [N*A] 0 undecided, 1 kept, 2 suppressed
    int* kflag;          // [N*A] kept as int, and its exclusive scan
    int* kscan;
    int* overflow;       // [N] image has a box with more than kMaxPreds predecessors -> sequential kernel
    int* changed;        // [8] worklist sizes (ping-pong), round counter
    int* wl_a;           // [N*A] undecided boxes, ping
    int* wl_b;           // [N*A] pong
    size_t cub2_bytes;
    void* cub2_tmp;
};
static constexpr int kCellStride = 32768;  // >= kGridHeads, power of two
static constexpr int kMaxPreds = 64;

static size_t align_up(size_t x) { return (x + 255) & ~size_t(255); }

// Spatial index over the kept boxes of one (image, class) segment.  Boxes are binned by size level
// l = floor(log2(max(w,h) / c0)) (cell size c0*2^l, so a level-l box is smaller than two cells) and
// by the cell of their centre.  IoU > thr > 0 implies (a) the boxes intersect and (b) their widths and
// heights are each within a factor 1/thr, so a candidate only visits levels within +-delta of its own
// and the cells its extent (grown by one cell) touches: exact pruning, not an approximation.
static constexpr int kGridLevels = 8;
static constexpr int kGridHeads = 24576;
// A level's cells are kGridFX x kGridFY per (level size)^2: narrow in x, where a window row is ONE contiguous index
// range whatever the cell width, so finer cells cost nothing and trim the scan to the window; in y every extra
// row is an extra range.
static constexpr float kGridFX = 4.0f, kGridFY = 2.0f;
struct GridGeom {
    int nlev, delta, prune;
    float c0;
    float rfac;  // IoU > thr needs IoU_x > thr, i.e. |centre offset| < (w1+w2)/2 * (1-thr)/(1+thr) (same in y)
    int nx[kGridLevels], ny[kGridLevels], base[kGridLevels];
};

static GridGeom make_grid(int img_h, int img_w, float iou_thr) {
    GridGeom g;
    memset(&g, 0, sizeof(g));
    int mx = img_w > img_h ? img_w : img_h;
    float c0 = 16.0f;
    while (mx / c0 > 64.0f) c0 *= 2.0f;
    g.c0 = c0;
    g.nlev = kGridLevels;
    int off = 0;
    for (int l = 0; l < kGridLevels; ++l) {
        float c = c0 * (float)(1 << l);
        int nx = (int)(img_w * kGridFX / c) + 2, ny = (int)(img_h * kGridFY / c) + 2;
        if (l == kGridLevels - 1) nx = ny = 1;  // unbounded sizes: one bucket
        if (off + nx * ny > kGridHeads) { nx = ny = 1; }
        g.nx[l] = nx; g.ny[l] = ny; g.base[l] = off;
        off += nx * ny;
    }
    g.prune = iou_thr > 0.0078125f;
    int d = 1;
    if (g.prune) { float r = 1.0f / iou_thr; while ((float)(1 << (d - 1)) < r) ++d; }  // floor(log2(1/thr)) + 1 or more
    g.delta = g.prune ? d : kGridLevels;
    g.rfac = g.prune ? (1.0f - iou_thr) / (1.0f + iou_thr) * 1.0001f : 1.0f;
    return g;
}

static size_t det_cub_bytes(long long n) {
    size_t bytes = 0;
    cub::DoubleBuffer<uint64_t> db(nullptr, nullptr);
    cub::DeviceRadixSort::SortKeys(nullptr, bytes, db, (int)n, kScoreShift, 64, (cudaStream_t)0);
    return bytes;
}

static size_t det_layout(int N, int A, void* base, DetWs* ws) {
    size_t off = 0;
    long long NA = (long long)N * A;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += align_up(bytes);
        return base ? reinterpret_cast<uint8_t*>(base) + o : nullptr;
    };
    DetWs w;
    w.boxes = reinterpret_cast<float4*>(take(NA * 16));
    w.scores = reinterpret_cast<float*>(take(NA * 4));
    w.keys = reinterpret_cast<uint64_t*>(take(NA * 8));
    w.keys_alt = reinterpret_cast<uint64_t*>(take(NA * 8));
    w.n_cand = reinterpret_cast<int*>(take(N * 4));
    w.max_coord = reinterpret_cast<int*>(take(N * 4));
    w.seg_start = reinterpret_cast<int*>(take(N * kMaxCls * 4));
    w.seg_end = reinterpret_cast<int*>(take(N * kMaxCls * 4));
    w.seg_kept = reinterpret_cast<int*>(take(N * kMaxCls * 4));
    w.kept_boxes = reinterpret_cast<float4*>(take(NA * 16));
    w.kept_keys = reinterpret_cast<uint64_t*>(take(NA * 8));
    w.kept_next = reinterpret_cast<int*>(take(NA * 4));
    w.heads = reinterpret_cast<int*>(take((size_t)N * kMaxCls * kGridHeads * 4));
    w.dbg = nullptr;
    w.cub_bytes = det_cub_bytes(NA);
    w.cub_tmp = take(w.cub_bytes);
    w.sbox = reinterpret_cast<float4*>(take(NA * 16));
    w.ckey = reinterpret_cast<uint32_t*>(take(NA * 4));
    w.ckey_alt = reinterpret_cast<uint32_t*>(take(NA * 4));
    w.cval = reinterpret_cast<int*>(take(NA * 4));
    w.cval_alt = reinterpret_cast<int*>(take(NA * 4));
    w.cbox = reinterpret_cast<float4*>(take(NA * 16));
    w.cell_cnt = reinterpret_cast<int*>(take(((size_t)N * kMaxCls * kCellStride + 1) * 4));
    w.cell_begin = reinterpret_cast<int*>(take(((size_t)N * kMaxCls * kCellStride + 1) * 4));
    w.preds = reinterpret_cast<int*>(take(NA * kMaxPreds * 4));
    w.npred = reinterpret_cast<int*>(take(NA * 4));
    w.nfor = reinterpret_cast<int*>(take(NA * 4));
    w.status = reinterpret_cast<unsigned char*>(take(NA));
    w.kflag = reinterpret_cast<int*>(take(NA * 4));
    w.kscan = reinterpret_cast<int*>(take(NA * 4));
    w.overflow = reinterpret_cast<int*>(take(N * 4));
    w.changed = reinterpret_cast<int*>(take(64));
    w.wl_a = reinterpret_cast<int*>(take(NA * 4));
    w.wl_b = reinterpret_cast<int*>(take(NA * 4));
    {
        size_t b1 = 0, b2 = 0;
        cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr);
        cub::DoubleBuffer<int> dv(nullptr, nullptr);
        cub::DeviceRadixSort::SortPairs(nullptr, b1, dk, dv, (int)NA, 0, 32, (cudaStream_t)0);
        long long T = (long long)N * kMaxCls * kCellStride + 1;
        cub::DeviceScan::ExclusiveSum(nullptr, b2, (int*)nullptr, (int*)nullptr, (int)(T > NA ? T : NA), (cudaStream_t)0);
        w.cub2_bytes = b1 > b2 ? b1 : b2;
    }
    w.cub2_tmp = take(w.cub2_bytes);
    if (ws) *ws = w;
    return off;
}

extern "C" int64_t hn_det_workspace_bytes(int32_t N, int32_t A) {
    if (N <= 0 || A <= 0) return 256;
    return (int64_t)det_layout(N, A, nullptr, nullptr);
}

int hn_det_num_launches(const hn_det_desc*) { return 12; }  // own kernels (the two CUB radix sorts and two scans add 20 library launches)

// decode (BBoxTransform + ClipBoxes, detection_loss.py:7-52), score = max over classes, threshold
__global__ void hn_det_decode_kernel(const float* __restrict__ anchors, const float* __restrict__ reg,
                                     const float* __restrict__ cls, const float* __restrict__ pre_boxes, int N, int A,
                                     int ncls, float wmax, float hmax, float thr, DetWs ws) {
    unsigned uidx = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = uidx < (unsigned)N * (unsigned)A;
    if (!valid) uidx = (unsigned)N * (unsigned)A - 1u;
    const int n = (int)(uidx / (unsigned)A), a = (int)(uidx - (unsigned)n * (unsigned)A);
    const long long idx = uidx;
    const float* c = cls + idx * ncls;
    float best = c[0];
    int bi = 0;
    for (int k = 1; k < ncls; ++k) {
        float v = c[k];
        if (greater_or_nan(v, best)) { best = v; bi = k; }
    }
    float4 box;
    if (pre_boxes) {
        box = *reinterpret_cast<const float4*>(pre_boxes + idx * 4);
    } else {
        float4 an = *reinterpret_cast<const float4*>(anchors + (long long)a * 4);  // y1,x1,y2,x2
        float4 rg = *reinterpret_cast<const float4*>(reg + idx * 4);               // dy,dx,dh,dw
        float yca = __fdiv_rn(__fadd_rn(an.x, an.z), 2.0f);
        float xca = __fdiv_rn(__fadd_rn(an.y, an.w), 2.0f);
        float ha = __fsub_rn(an.z, an.x);
        float wa = __fsub_rn(an.w, an.y);
        // exp evaluated in double and rounded once: correctly rounded fp32 (oracle does the same)
        float w = __fmul_rn((float)exp((double)rg.w), wa);
        float h = __fmul_rn((float)exp((double)rg.z), ha);
        float yc = __fadd_rn(__fmul_rn(rg.x, ha), yca);
        float xc = __fadd_rn(__fmul_rn(rg.y, wa), xca);
        float hh = __fdiv_rn(h, 2.0f), hw = __fdiv_rn(w, 2.0f);
        box.x = fmaxf(__fsub_rn(xc, hw), 0.0f);  // xmin
        box.y = fmaxf(__fsub_rn(yc, hh), 0.0f);  // ymin
        box.z = fminf(__fadd_rn(xc, hw), wmax);  // xmax
        box.w = fminf(__fadd_rn(yc, hh), hmax);  // ymax
    }
    const bool cand = valid && best > thr;
    uint64_t key = ~0ull;
    int ord = float_to_ordered(-INFINITY);
    if (cand) {
        key = ((uint64_t)n << kImgShift) | ((uint64_t)bi << kClsShift) | ((uint64_t)float_desc_key(best) << kScoreShift) |
              (uint64_t)a;
        ord = float_to_ordered(fmaxf(fmaxf(box.x, box.y), fmaxf(box.z, box.w)));
    }
    if (valid) {
        ws.boxes[idx] = box;
        ws.scores[idx] = best;
        ws.keys[idx] = key;
    }
    // per-image candidate count and max coordinate.  These are same-address atomics (one address per image), which
    // the L2 serialises: reduce per warp, then per block through shared memory, and issue ONE pair per block when the
    // whole block sits in one image (all but the blocks straddling an image boundary).
    __shared__ int s_cnt[8], s_max[8], s_img[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int n0 = __shfl_sync(0xffffffffu, n, 0);
    const unsigned m = __ballot_sync(0xffffffffu, cand);
    const bool warp_one = __all_sync(0xffffffffu, n == n0);
    const int wmaxo = __reduce_max_sync(0xffffffffu, ord);
    if (lane == 0) {
        s_cnt[wid] = __popc(m);
        s_max[wid] = wmaxo;
        s_img[wid] = warp_one ? n0 : -1;
    }
    __syncthreads();
    bool block_one = true;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) block_one = block_one && s_img[w] == s_img[0] && s_img[w] >= 0;
    if (block_one) {
        if (threadIdx.x == 0) {
            int cnt = 0, mx = s_max[0];
            for (int w = 0; w < nw; ++w) { cnt += s_cnt[w]; mx = max(mx, s_max[w]); }
            if (cnt) {
                atomicAdd(ws.n_cand + s_img[0], cnt);
                atomicMax(ws.max_coord + s_img[0], mx);
            }
        }
    } else if (warp_one) {
        if (lane == 0 && m) {
            atomicAdd(ws.n_cand + n0, __popc(m));
            atomicMax(ws.max_coord + n0, wmaxo);
        }
    } else if (cand) {
        atomicAdd(ws.n_cand + n, 1);
        atomicMax(ws.max_coord + n, ord);
    }
}

__global__ void hn_det_segments_kernel(const uint64_t* __restrict__ keys, long long total, int* seg_start, int* seg_end) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint64_t k = keys[i];
    if (k == ~0ull) return;
    int seg = (int)(k >> kClsShift);  // img*16 + cls
    if (i == 0 || (int)(keys[i - 1] >> kClsShift) != seg) seg_start[seg] = (int)i;
    if (i + 1 == total || keys[i + 1] == ~0ull || (int)(keys[i + 1] >> kClsShift) != seg) seg_end[seg] = (int)(i + 1);
}

__device__ __forceinline__ bool iou_gt(const float4& a, float area_a, const float4& b, float area_b, float thr) {
    // torchvision nms: inter / (area_i + area_j - inter) > thr, all fp32
    float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
    float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
    float w = fmaxf(0.0f, __fsub_rn(xx2, xx1)), h = fmaxf(0.0f, __fsub_rn(yy2, yy1));
    if (thr >= 0.0f && (w <= 0.0f || h <= 0.0f)) return false;  // inter == 0 -> ovr is 0 or NaN: never > thr
    float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    // the decision is fl(inter / uni) > thr; away from the boundary a multiplication settles it (relative margins of
    // 1e-5 dwarf the rounding of one division and one product), the exact division runs only inside the band
    if (thr > 0.0f && uni > 0.0f) {
        const float tu = thr * uni;
        if (inter < tu * 0.99999f) return false;
        if (inter > tu * 1.00001f) return true;
    }
    float ovr = __fdiv_rn(inter, uni);
    return ovr > thr;
}
__device__ __forceinline__ float box_area(const float4& b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

static constexpr int kNmsChunk = 512;

__device__ __forceinline__ int grid_level(float m, const GridGeom& g) {
    int l = 0;
    float c = g.c0 * 2.0f;
    while (l < g.nlev - 1 && m >= c) { c *= 2.0f; ++l; }
    return l;
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// One CTA per (image, class) segment of the sorted candidates: greedy NMS in sorted order, processed
// in chunks of 512: (1) every candidate of the chunk is tested against the boxes kept so far that the
// spatial index says can overlap it, (2) survivors are resolved against each other with a 512x512 bit
// matrix walked by one warp, which also inserts the newly kept boxes into the index.
__global__ void __launch_bounds__(kNmsChunk) hn_det_nms_kernel(DetWs ws, int A, float iou_thr, int nms_mode, GridGeom g,
                                                               int only_overflow) {
    __shared__ float4 s_box[kNmsChunk];
    __shared__ float s_area[kNmsChunk];
    __shared__ unsigned long long s_mask[kNmsChunk][kNmsChunk / 64];
    __shared__ unsigned long long s_alive[kNmsChunk / 64];
    __shared__ unsigned short s_keptlist[kNmsChunk];
    __shared__ int s_nk, s_new;
    const int seg = blockIdx.x;
    const int n = seg / kMaxCls, cls = seg % kMaxCls;
    const int s0 = ws.seg_start[seg], s1 = ws.seg_end[seg];
    const int tid = threadIdx.x;
    if (only_overflow && ws.overflow[n] == 0) return;  // this image was handled by the parallel path
    if (tid == 0) s_nk = 0;
    if (s1 <= s0) {
        if (tid == 0) ws.seg_kept[seg] = 0;
        return;
    }
    long long* dbgc = ws.dbg ? ws.dbg + (long long)seg * 8 : nullptr;  // cycles: phase1, phase2, resolve, publish
    long long t_mark = clock64();
    // coordinate trick (torchvision boxes.py _batched_nms_coordinate_trick): offset = cls * (max + 1)
    const int ncand = ws.n_cand[n];
    bool trick = nms_mode == HN_NMS_TRICK ||
                 (nms_mode == HN_NMS_AUTO_CUDA && (long long)ncand * 4 <= 100000) ||
                 (nms_mode == HN_NMS_AUTO_CPU && (long long)ncand * 4 <= 4000);
    float offset = 0.0f;
    if (trick) offset = __fmul_rn((float)cls, __fadd_rn(ordered_to_float(ws.max_coord[n]), 1.0f));
    float4* kept_boxes = ws.kept_boxes + s0;
    uint64_t* kept_keys = ws.kept_keys + s0;
    int* kept_next = ws.kept_next + s0;
    int* heads = ws.heads + (size_t)seg * kGridHeads;
    __syncthreads();

    for (int c0 = s0; c0 < s1; c0 += kNmsChunk) {
        const int j = c0 + tid;
        const bool have = j < s1;
        float4 b = make_float4(0, 0, 0, 0);
        if (have) {
            uint64_t key = ws.keys[j];
            int a = (int)(key & ((1u << kAnchorBits) - 1));
            b = ws.boxes[(long long)n * A + a];
            if (trick) {
                b.x = __fadd_rn(b.x, offset); b.y = __fadd_rn(b.y, offset);
                b.z = __fadd_rn(b.z, offset); b.w = __fadd_rn(b.w, offset);
            }
        }
        const float area = box_area(b);
        // (1) against the kept boxes that can overlap this one
        const int nk = s_nk;
        bool alive = have;
        if (alive && nk > 0) {
            if (!g.prune) {
                for (int k = 0; k < nk && alive; ++k) {
                    float4 kb = kept_boxes[k];
                    if (iou_gt(kb, box_area(kb), b, area, iou_thr)) alive = false;
                }
            } else {
                const float wj = b.z - b.x, hj = b.w - b.y;
                const float cx = 0.5f * (b.x + b.z) - offset, cy = 0.5f * (b.y + b.w) - offset;
                const int t = grid_level(fmaxf(wj, hj), g);
                const int l_lo = max(0, t - g.delta), l_hi = min(g.nlev - 1, t + g.delta);
                for (int l = l_lo; l <= l_hi && alive; ++l) {
                    const float cl = g.c0 * (float)(1 << l);
                    const float inv = 1.0f / cl;
                    const int nx = g.nx[l], ny = g.ny[l];
                    int x_lo = 0, x_hi = 0, y_lo = 0, y_hi = 0;
                    if (nx * ny > 1) {
                        const float rx = g.rfac * (0.5f * fmaxf(wj, 0.0f) + cl) + 1.0f, ry = g.rfac * (0.5f * fmaxf(hj, 0.0f) + cl) + 1.0f;
                        x_lo = clampi((int)floorf((cx - rx) * inv * kGridFX), 0, nx - 1);
                        x_hi = clampi((int)floorf((cx + rx) * inv * kGridFX), 0, nx - 1);
                        y_lo = clampi((int)floorf((cy - ry) * inv * kGridFY), 0, ny - 1);
                        y_hi = clampi((int)floorf((cy + ry) * inv * kGridFY), 0, ny - 1);
                    }
                    for (int yy = y_lo; yy <= y_hi && alive; ++yy) {
                        for (int xx = x_lo; xx <= x_hi && alive; ++xx) {
                            int k = heads[g.base[l] + yy * nx + xx];
                            while (k >= 0) {
                                float4 kb = kept_boxes[k];
                                if (iou_gt(kb, box_area(kb), b, area, iou_thr)) { alive = false; break; }
                                k = kept_next[k];
                            }
                        }
                    }
                }
            }
        }
        s_box[tid] = b;
        s_area[tid] = area;
        if (tid < kNmsChunk / 64) s_alive[tid] = 0ull;
        __syncthreads();
        if (alive) atomicOr(&s_alive[tid >> 6], 1ull << (tid & 63));
        __syncthreads();
        if (dbgc && tid == 0) { long long tt = clock64(); dbgc[0] += tt - t_mark; t_mark = tt; }
        // (2) intra-chunk suppression rows (only rows of live boxes are ever read)
        if (alive) {
            for (int w = 0; w < kNmsChunk / 64; ++w) {
                unsigned long long bits = 0ull;
                if (w >= (tid >> 6)) {
                    unsigned long long cand = s_alive[w];  // only live, later boxes can be suppressed by this one
                    if (w == (tid >> 6)) cand &= ~((2ull << (tid & 63)) - 1ull);
                    while (cand) {
                        int t = __ffsll((long long)cand) - 1;
                        cand &= cand - 1;
                        int o = w * 64 + t;
                        if (iou_gt(b, area, s_box[o], s_area[o], iou_thr)) bits |= 1ull << t;
                    }
                }
                s_mask[tid][w] = bits;
            }
        }
        __syncthreads();
        if (dbgc && tid == 0) { long long tt = clock64(); dbgc[1] += tt - t_mark; t_mark = tt; }
        if (tid < 32) {
            // lane w (< 8) owns word w of the live set; walk the live boxes in order (shared memory only)
            const int lane = tid;
            unsigned long long live = lane < kNmsChunk / 64 ? s_alive[lane] : 0ull;
            int nnew = 0;
            while (true) {
                unsigned nz = __ballot_sync(0xffffffffu, live != 0ull);
                if (nz == 0) break;
                int w = __ffs(nz) - 1;
                unsigned long long lw = __shfl_sync(0xffffffffu, live, w);
                int i = w * 64 + __ffsll((long long)lw) - 1;
                if (lane < kNmsChunk / 64) {
                    live &= ~s_mask[i][lane];
                    if (lane == w) live &= ~(1ull << (i & 63));
                }
                if (lane == 0) s_keptlist[nnew] = (unsigned short)i;
                ++nnew;
            }
            if (lane == 0) s_new = nnew;
        }
        __syncthreads();
        if (dbgc && tid == 0) { long long tt = clock64(); dbgc[2] += tt - t_mark; t_mark = tt; }
        // publish the newly kept boxes (in order) and insert them into the spatial index, in parallel
        {
            const int nk0 = s_nk, nnew = s_new;
            if (tid < nnew) {
                const int i = s_keptlist[tid];
                const int slot = nk0 + tid;
                float4 kb = s_box[i];
                kept_boxes[slot] = kb;
                kept_keys[slot] = ws.keys[c0 + i];
                if (g.prune) {
                    const float cx = 0.5f * (kb.x + kb.z) - offset, cy = 0.5f * (kb.y + kb.w) - offset;
                    const int l = grid_level(fmaxf(kb.z - kb.x, kb.w - kb.y), g);
                    const float inv = 1.0f / (g.c0 * (float)(1 << l));
                    int cell = g.base[l];
                    if (g.nx[l] * g.ny[l] > 1)
                        cell += clampi((int)floorf(cy * inv * kGridFY), 0, g.ny[l] - 1) * g.nx[l] + clampi((int)floorf(cx * inv * kGridFX), 0, g.nx[l] - 1);
                    kept_next[slot] = atomicExch(&heads[cell], slot);
                }
            }
            __syncthreads();
            if (tid == 0) s_nk = nk0 + nnew;
        }
        __syncthreads();
        if (dbgc && tid == 0) { long long tt = clock64(); dbgc[3] += tt - t_mark; t_mark = tt; dbgc[4] += 1; }
    }
    if (dbgc && tid == 0) { dbgc[5] = s1 - s0; dbgc[6] = s_nk; }
    if (tid == 0) ws.seg_kept[seg] = s_nk;
}

// ------------------------------------------------------------------------------------------------
// Parallel exact greedy NMS ("dependency rounds").  Greedy NMS is the lexicographically-first maximal
// independent set of the conflict graph (edge = same class and IoU > thr) under the score order.  A box
// is kept iff all its earlier conflicting neighbours are suppressed, suppressed iff one of them is kept:
// with the per-box predecessor lists built once (spatial hash over ALL candidates, cells sorted by
// priority), every round decides all boxes whose predecessors are decided, for all images and classes
// at once.  With scores unrelated to position the number of rounds is O(log^2 n); the result is exactly
// the sequential one, whatever the class skew.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float seg_offset(const DetWs& ws, int n, int cls, int nms_mode) {
    const int ncand = ws.n_cand[n];
    bool trick = nms_mode == HN_NMS_TRICK || (nms_mode == HN_NMS_AUTO_CUDA && (long long)ncand * 4 <= 100000) ||
                 (nms_mode == HN_NMS_AUTO_CPU && (long long)ncand * 4 <= 4000);
    return trick ? __fmul_rn((float)cls, __fadd_rn(ordered_to_float(ws.max_coord[n]), 1.0f)) : 0.0f;
}
__device__ __forceinline__ int grid_cell(const float4& b, float offset, const GridGeom& g) {
    const float cx = 0.5f * (b.x + b.z) - offset, cy = 0.5f * (b.y + b.w) - offset;
    const int l = grid_level(fmaxf(b.z - b.x, b.w - b.y), g);
    const float inv = 1.0f / (g.c0 * (float)(1 << l));
    int cell = g.base[l];
    if (g.nx[l] * g.ny[l] > 1)
        cell += clampi((int)floorf(cy * inv * kGridFY), 0, g.ny[l] - 1) * g.nx[l] + clampi((int)floorf(cx * inv * kGridFX), 0, g.nx[l] - 1);
    return cell;
}

__global__ void hn_nms2_prepare_kernel(DetWs ws, long long NA, int A, int nms_mode, GridGeom g) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NA) return;
    const uint64_t key = ws.keys[i];
    ws.cval[i] = (int)i;
    ws.kflag[i] = 0;
    ws.status[i] = 2;
    ws.npred[i] = 0;
    ws.nfor[i] = 0;
    if (key == ~0ull) { ws.ckey[i] = 0xFFFFFFFFu; return; }
    const int seg = (int)(key >> kClsShift), n = seg / kMaxCls, cls = seg % kMaxCls;
    const int a = (int)(key & ((1u << kAnchorBits) - 1));
    const float off = seg_offset(ws, n, cls, nms_mode);
    float4 b = ws.boxes[(long long)n * A + a];
    if (off != 0.0f) {
        b.x = __fadd_rn(b.x, off); b.y = __fadd_rn(b.y, off); b.z = __fadd_rn(b.z, off); b.w = __fadd_rn(b.w, off);
    }
    ws.sbox[i] = b;
    const uint32_t ck = (uint32_t)seg * kCellStride + (uint32_t)grid_cell(b, off, g);
    ws.ckey[i] = ck;
    atomicAdd(ws.cell_cnt + ck, 1);
}

__global__ void hn_nms2_cells_kernel(DetWs ws, long long NA) {
    long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= NA) return;
    const uint32_t ck = ws.ckey[q];
    if (ck == 0xFFFFFFFFu) return;
    ws.cbox[q] = ws.sbox[ws.cval[q]];
}

// Conflict-graph construction, one thread per candidate, candidates taken in CELL order.  Cells of one grid row
// are consecutive in cell order, so the members of the cells x_lo..x_hi of a window row form ONE contiguous index
// range.  A thread first collects the ranges of all its window rows (independent loads), then walks them as one
// flattened sequence: the 32 candidates of a warp are neighbours in the grid, so their walks have similar lengths
// and touch the same cache lines (measured: the earlier 4-lanes-per-candidate version spent 3/4 of its issue
// slots on idle lanes and per-range loop overhead, ranges being ~10 entries long).
// Every conflicting pair is discovered ONCE, from its smaller box: a candidate scans only its own size level and
// the coarser ones (few, large cells) and records the edge at the later box of the pair, i.e. in its own
// predecessor list or -- atomically -- in the other box's.
static constexpr int kBuildMaxRanges = 24;
__global__ void __launch_bounds__(256, 6) hn_nms2_build_kernel(DetWs ws, long long NA, int nms_mode, float iou_thr, GridGeom g) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= NA || ws.ckey[slot] == 0xFFFFFFFFu) return;
    const long long i = ws.cval[slot];
    const uint64_t key = ws.keys[i];
    const int seg = (int)(key >> kClsShift), n = seg / kMaxCls;
    const float offset = seg_offset(ws, n, seg % kMaxCls, nms_mode);
    const float4 b = ws.sbox[i];
    const float area = box_area(b);
    const float wj = b.z - b.x, hj = b.w - b.y;
    const float cx = 0.5f * (b.x + b.z) - offset, cy = 0.5f * (b.y + b.w) - offset;
    const int t = grid_level(fmaxf(wj, hj), g);
    const int l_hi = min(g.nlev - 1, t + g.delta);
    int scanned = 0, edges = 0;
    // The box's own predecessor list grows from the front with plain stores (only this thread writes there); edges
    // owed to ANOTHER box take an atomic slot counted down from the far end of that box's list.  A returning atomic
    // in the scan loop stalls the warp for an L2 round trip, and with ~10 % of the entries hitting, some lane of the
    // warp hits in nearly every iteration.
    int n_own = 0;
    int* my_preds = ws.preds + i * kMaxPreds;
    auto visit = [&](int q, bool upper) {
        const float4 kb = ws.cbox[q];
        ++scanned;
        if (iou_gt(kb, box_area(kb), b, area, iou_thr)) {  // only then look up the priority
            const int m = ws.cval[q];
            ++edges;
            if (m < i) {
                if (n_own < kMaxPreds) my_preds[n_own] = m;
                ++n_own;
            } else if (m > i && upper) {  // same level: the other box records it itself
                const int pos = atomicAdd(ws.nfor + m, 1);
                if (pos < kMaxPreds) ws.preds[(long long)m * kMaxPreds + (kMaxPreds - 1 - pos)] = (int)i;
            }
        }
    };
    // pass 1: the non-empty index range of every window row
    int r_beg[kBuildMaxRanges], r_end[kBuildMaxRanges];
    int nr = 0, n_same = 0;  // ranges [0, n_same) belong to the candidate's own level
    for (int l = t; l <= l_hi; ++l) {
        const float cl = g.c0 * (float)(1 << l);
        const float inv = 1.0f / cl;
        const int nx = g.nx[l], ny = g.ny[l];
        int x_lo = 0, x_hi = 0, y_lo = 0, y_hi = 0;
        if (nx * ny > 1) {
            const float rx = g.rfac * (0.5f * fmaxf(wj, 0.0f) + cl) + 1.0f, ry = g.rfac * (0.5f * fmaxf(hj, 0.0f) + cl) + 1.0f;
            x_lo = clampi((int)floorf((cx - rx) * inv * kGridFX), 0, nx - 1);
            x_hi = clampi((int)floorf((cx + rx) * inv * kGridFX), 0, nx - 1);
            y_lo = clampi((int)floorf((cy - ry) * inv * kGridFY), 0, ny - 1);
            y_hi = clampi((int)floorf((cy + ry) * inv * kGridFY), 0, ny - 1);
        }
        for (int yy = y_lo; yy <= y_hi; ++yy) {
            const uint32_t k0 = (uint32_t)seg * kCellStride + (uint32_t)(g.base[l] + yy * nx + x_lo);
            const int qb = ws.cell_begin[k0], qe = ws.cell_begin[k0 + (x_hi - x_lo) + 1];
            if (qb == qe) continue;
            if (nr < kBuildMaxRanges) {
                r_beg[nr] = qb;
                r_end[nr] = qe;
                ++nr;
            } else {  // (rare) more window rows than range slots: scan this row right away
                for (int q = qb; q < qe; ++q) visit(q, l > t);
            }
        }
        if (l == t) n_same = nr;
    }
    // pass 2: one flattened walk over the ranges
    if (nr > 0) {
        int r = 0, q = r_beg[0], qe = r_end[0];
        while (true) {
            visit(q, r >= n_same);
            if (++q == qe) {
                if (++r == nr) break;
                q = r_beg[r];
                qe = r_end[r];
            }
        }
    }
    ws.npred[i] = n_own;
    if (ws.dbg) {
        atomicAdd(reinterpret_cast<unsigned long long*>(ws.dbg) + 0, (unsigned long long)scanned);
        atomicAdd(reinterpret_cast<unsigned long long*>(ws.dbg) + 1, (unsigned long long)edges);
    }
}

// after all edges are in place: boxes without predecessors are kept outright
__global__ void __launch_bounds__(256) hn_nms2_seed_kernel(DetWs ws, long long NA) {
    __shared__ int s_cnt[8], s_base;
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool und = false;
    if (i < NA && ws.keys[i] != ~0ull) {
        int own = ws.npred[i], fgn = ws.nfor[i];
        if (own + fgn > kMaxPreds) {  // the two ends of the list ran into each other: the sequential kernel redoes the image
            ws.overflow[(int)(ws.keys[i] >> kClsShift) / kMaxCls] = 1;
            own = min(own, kMaxPreds);
            fgn = min(fgn, kMaxPreds - own);
            ws.npred[i] = own;
            ws.nfor[i] = fgn;
        }
        const int np = own + fgn;
        ws.status[i] = np == 0 ? 1 : 0;
        und = np != 0;
    }
    // undecided boxes go on the first worklist: ONE atomic per block (they all hit the same counter, which the L2
    // serialises), slots from a block-wide prefix; order is irrelevant
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned m = __ballot_sync(0xffffffffu, und);
    if (lane == 0) s_cnt[wid] = __popc(m);
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int w = 0; w < 8; ++w) { const int c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
        s_base = tot ? atomicAdd(ws.changed + 0, tot) : 0;
    }
    __syncthreads();
    if (und) ws.wl_a[s_base + s_cnt[wid] + __popc(m & ((1u << lane) - 1u))] = (int)i;
}

// Rounds over a shrinking worklist: every round decides the boxes whose predecessors are all decided and
// carries the rest over; the loop ends when the worklist is empty (the earliest undecided box of every
// segment always gets decided, so every round makes progress).
__global__ void __launch_bounds__(256) hn_nms2_rounds_kernel(DetWs ws, long long NA, int g_rounds_passes) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ int s_cnt[8], s_base;
    // three rotating worklist counters: [r % 3] = size of the current list, [(r+1) % 3] = the list being built,
    // [(r+2) % 3] = reset now for the round after next -- so one grid barrier per round is enough
    volatile int* cnt = ws.changed;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long tid0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    int* cur = ws.wl_a;
    int* nxt = ws.wl_b;
    int round = 0;
    while (true) {
        const int c_cur = round % 3, c_nxt = (round + 1) % 3, c_clr = (round + 2) % 3;
        const int n_cur = cnt[c_cur];
        if (n_cur == 0) break;
        if (tid0 == 0) ws.changed[c_clr] = 0;
        const long long n_pad = ((long long)n_cur + 255) & ~255ll;  // whole CTAs iterate together (block-wide append below)
        // Short lists are bound by the grid barrier and by load latency, not by work: walk them several times per
        // round, so that chains of dependent boxes advance more than one link per barrier (a later pass sees what
        // other threads decided meanwhile; stale reads only delay a decision, they never change it).
        const int passes = n_cur < 400000 ? g_rounds_passes : 1;
        for (int pass = 0; pass < passes; ++pass)
        for (long long w = tid0; w < n_pad; w += stride) {
            bool carry = false;
            int i = 0;
            if (w < n_cur) i = __ldcg(cur + w);
            if (w < n_cur && (pass == 0 || __ldcg(ws.status + i) == 0)) {
                const int own = ws.npred[i], fgn = ws.nfor[i];
                const int4* my = reinterpret_cast<const int4*>(ws.preds + (long long)i * kMaxPreds);
                bool kept_pred = false, all_supp = true;
                // four predecessors at a time: their status loads are independent (the loop is latency-bound).
                // own part preds[0 .. own), then the foreign part preds[kMaxPreds - fgn .. kMaxPreds)
                auto look = [&](int first, int count) {
                    const int k0 = first & ~3;  // aligned int4 reads; entries outside [first, first+count) are skipped
                    for (int k = k0; k < first + count && !kept_pred; k += 4) {
                        const int4 pr = __ldcg(my + (k >> 2));
                        const int lo = first, hi = first + count;
                        const unsigned char s0 = (k + 0 >= lo && k + 0 < hi) ? __ldcg(ws.status + pr.x) : (unsigned char)2;
                        const unsigned char s1 = (k + 1 >= lo && k + 1 < hi) ? __ldcg(ws.status + pr.y) : (unsigned char)2;
                        const unsigned char s2 = (k + 2 >= lo && k + 2 < hi) ? __ldcg(ws.status + pr.z) : (unsigned char)2;
                        const unsigned char s3 = (k + 3 >= lo && k + 3 < hi) ? __ldcg(ws.status + pr.w) : (unsigned char)2;
                        kept_pred = s0 == 1 || s1 == 1 || s2 == 1 || s3 == 1;
                        if (s0 == 0 || s1 == 0 || s2 == 0 || s3 == 0) all_supp = false;
                    }
                };
                look(0, own);
                if (fgn > 0) look(kMaxPreds - fgn, fgn);
                if (kept_pred) ws.status[i] = 2;
                else if (all_supp) ws.status[i] = 1;
                else carry = pass == passes - 1;
            }
            // append to the next worklist: one atomic per CTA (every append hits the same counter)
            const unsigned m = __ballot_sync(0xffffffffu, carry);
            const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
            if (lane == 0) s_cnt[wid] = __popc(m);
            __syncthreads();
            if (threadIdx.x == 0) {
                int tot = 0;
                for (int k = 0; k < 8; ++k) { const int c = s_cnt[k]; s_cnt[k] = tot; tot += c; }
                s_base = tot ? atomicAdd(ws.changed + c_nxt, tot) : 0;
            }
            __syncthreads();
            if (carry) nxt[s_base + s_cnt[wid] + __popc(m & ((1u << lane) - 1u))] = i;
            __syncthreads();  // s_cnt / s_base are reused by the next iteration
        }
        int* t = cur; cur = nxt; nxt = t;
        ++round;
        grid.sync();
    }
    if (tid0 == 0) ws.changed[4] = round;  // diagnostics: number of rounds
    for (long long i = tid0; i < NA; i += stride) ws.kflag[i] = __ldcg(ws.status + i) == 1 ? 1 : 0;
}

__global__ void hn_nms2_compact_kernel(DetWs ws, long long NA) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NA) return;
    const uint64_t key = ws.keys[i];
    if (key == ~0ull) return;
    const int seg = (int)(key >> kClsShift), n = seg / kMaxCls;
    if (ws.overflow[n]) return;  // the sequential kernel owns this image
    const int s0 = ws.seg_start[seg], s1 = ws.seg_end[seg];
    const int base = ws.kscan[s0];
    const int f = ws.kflag[i];
    if (f) ws.kept_keys[s0 + ws.kscan[i] - base] = key;
    if (i == s1 - 1) ws.seg_kept[seg] = ws.kscan[i] + f - base;
}

// final order = score descending, ties by ascending candidate index, across classes: the rank of a
// kept box is the number of kept boxes (all classes) whose (~score, anchor) key is smaller
__global__ void hn_det_gather_kernel(DetWs ws, int A, float* __restrict__ out_boxes, float* __restrict__ out_scores,
                                     int64_t* __restrict__ out_class, int* __restrict__ out_count) {
    const int seg = blockIdx.x;
    const int n = seg / kMaxCls, cls = seg % kMaxCls;
    const int nk = ws.seg_kept[seg];
    const uint64_t low_mask = (1ull << kClsShift) - 1;
    if (cls == 0 && threadIdx.x == 0 && blockIdx.y == 0) {
        int tot = 0;
        for (int c = 0; c < kMaxCls; ++c) tot += ws.seg_kept[n * kMaxCls + c];
        out_count[n] = tot;
    }
    const uint64_t* mine = ws.kept_keys + ws.seg_start[seg];
    // blockIdx.y strides over the segment too: with skewed classes one segment holds a third of an image's boxes
    for (int i = blockIdx.y * blockDim.x + threadIdx.x; i < nk; i += gridDim.y * blockDim.x) {
        uint64_t key = mine[i] & low_mask;
        int rank = i;
        for (int c = 0; c < kMaxCls; ++c) {
            if (c == cls) continue;
            int oseg = n * kMaxCls + c;
            int cnt = ws.seg_kept[oseg];
            if (cnt == 0) continue;
            const uint64_t* other = ws.kept_keys + ws.seg_start[oseg];
            int lo = 0, hi = cnt;  // first element with key >= mine
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if ((other[mid] & low_mask) < key) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        int a = (int)(key & ((1u << kAnchorBits) - 1));
        long long src = (long long)n * A + a, dst = (long long)n * A + rank;
        *reinterpret_cast<float4*>(out_boxes + dst * 4) = ws.boxes[src];
        out_scores[dst] = ws.scores[src];
        out_class[dst] = cls;
    }
}

__global__ void hn_det_init_kernel(DetWs ws, int N) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < N) {
        ws.n_cand[i] = 0;
        ws.max_coord[i] = float_to_ordered(-INFINITY);
        ws.overflow[i] = 0;
    }
    if (i < 8) ws.changed[i] = 0;
    if (i < N * kMaxCls) {
        ws.seg_start[i] = 0;
        ws.seg_end[i] = 0;
        ws.seg_kept[i] = 0;
    }
}

__global__ void hn_copy_i32_kernel(const int* in, int* out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i];
}

static void* g_det_dbg = nullptr;
static int g_det_force_sequential = 0;
extern "C" void hn_det_force_sequential(int on) { g_det_force_sequential = on; }
static int g_det_rounds_ctas_per_sm = 2;
static int g_det_rounds_passes = 3;
extern "C" void hn_det_set_rounds_passes(int n) { g_det_rounds_passes = n < 1 ? 1 : n; }
extern "C" void hn_det_set_rounds_ctas_per_sm(int n) { g_det_rounds_ctas_per_sm = n < 1 ? 1 : n; }
extern "C" void hn_det_set_debug_buffer(void* p) { g_det_dbg = p; }

extern "C" int hn_det_decode_nms(const hn_det_desc* d, void* stream) {
    HN_REQUIRE(d != nullptr, "det: null desc");
    HN_REQUIRE(d->N >= 0 && d->N <= 255 && d->A >= 0 && d->A < (1 << kAnchorBits) && d->ncls >= 1 && d->ncls <= kMaxCls,
               "det: N<=255, A<2^20, ncls<=16 required (N=%d A=%d ncls=%d)", d->N, d->A, d->ncls);
    HN_REQUIRE(d->classification && d->out_boxes && d->out_scores && d->out_class && d->out_count, "det: null pointer");
    HN_REQUIRE(d->pre_boxes || (d->anchors && d->regression), "det: need anchors+regression or pre_boxes");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    if (d->N == 0) return HN_OK;
    if (d->A == 0) {
        HN_CHECK_CUDA(cudaMemsetAsync(d->out_count, 0, sizeof(int) * d->N, s));
        if (d->out_cand) HN_CHECK_CUDA(cudaMemsetAsync(d->out_cand, 0, sizeof(int) * d->N, s));
        return HN_OK;
    }
    HN_REQUIRE(d->workspace && d->workspace_bytes >= hn_det_workspace_bytes(d->N, d->A), "det: workspace too small");
    DetWs ws;
    det_layout(d->N, d->A, d->workspace, &ws);
    ws.dbg = reinterpret_cast<long long*>(g_det_dbg);
    const long long NA = (long long)d->N * d->A;
    HN_CHECK_CUDA(cudaMemsetAsync(ws.heads, 0xFF, (size_t)d->N * kMaxCls * kGridHeads * 4, s));
    hn_det_init_kernel<<<hn_cdiv(d->N * kMaxCls, 256), 256, 0, s>>>(ws, d->N);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_det_decode_kernel<<<hn_cdiv(NA, 256), 256, 0, s>>>(d->anchors, d->regression, d->classification, d->pre_boxes, d->N,
                                                         d->A, d->ncls, (float)(d->img_w - 1), (float)(d->img_h - 1),
                                                         d->conf_thres, ws);
    HN_CHECK_CUDA(cudaGetLastError());
    cub::DoubleBuffer<uint64_t> db(ws.keys, ws.keys_alt);
    size_t tmp = ws.cub_bytes;
    // bits [0, 20) hold the anchor index: candidates are generated in anchor order and the radix sort is stable, so ties on
    // (image, class, score) already come out anchor-ascending -- 44 key bits = 6 passes instead of 8
    HN_CHECK_CUDA(cub::DeviceRadixSort::SortKeys(ws.cub_tmp, tmp, db, (int)NA, kScoreShift, 64, s));
    ws.keys = db.Current();
    ws.keys_alt = db.Alternate();
    hn_det_segments_kernel<<<hn_cdiv(NA, 256), 256, 0, s>>>(ws.keys, NA, ws.seg_start, ws.seg_end);
    HN_CHECK_CUDA(cudaGetLastError());
    GridGeom geom = make_grid(d->img_h, d->img_w, d->iou_thres);
    const bool parallel = geom.prune && !g_det_force_sequential;
    if (parallel) {
        const long long T = (long long)d->N * kMaxCls * kCellStride + 1;
        HN_CHECK_CUDA(cudaMemsetAsync(ws.cell_cnt, 0, (size_t)T * 4, s));
        hn_nms2_prepare_kernel<<<hn_cdiv(NA, 256), 256, 0, s>>>(ws, NA, d->A, d->nms_mode, geom);
        HN_CHECK_CUDA(cudaGetLastError());
        cub::DoubleBuffer<uint32_t> dk(ws.ckey, ws.ckey_alt);
        cub::DoubleBuffer<int> dv(ws.cval, ws.cval_alt);
        size_t tb = ws.cub2_bytes;
        // key = segment * kCellStride + cell: only the bits a valid key can have set are sorted (batch 32: 24 bits = 3 passes
        // instead of 4); the all-ones key of a non-candidate reads as (a segment >= the last one, cell 32767 > any real cell)
        // in those bits and still lands behind every valid key
        int seg_bits = 1;
        while ((1 << seg_bits) < d->N * kMaxCls) ++seg_bits;
        HN_CHECK_CUDA(cub::DeviceRadixSort::SortPairs(ws.cub2_tmp, tb, dk, dv, (int)NA, 0, 15 + seg_bits, s));
        ws.ckey = dk.Current(); ws.ckey_alt = dk.Alternate();
        ws.cval = dv.Current(); ws.cval_alt = dv.Alternate();
        hn_nms2_cells_kernel<<<hn_cdiv(NA, 256), 256, 0, s>>>(ws, NA);
        HN_CHECK_CUDA(cudaGetLastError());
        tb = ws.cub2_bytes;
        HN_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(ws.cub2_tmp, tb, ws.cell_cnt, ws.cell_begin, (int)T, s));
        hn_nms2_build_kernel<<<hn_cdiv(NA, 256), 256, 0, s>>>(ws, NA, d->nms_mode, d->iou_thres, geom);
        HN_CHECK_CUDA(cudaGetLastError());
        hn_nms2_seed_kernel<<<hn_cdiv(NA, 256), 256, 0, s>>>(ws, NA);
        HN_CHECK_CUDA(cudaGetLastError());
        {
            static int blocks_per_sm = 0;
            if (blocks_per_sm == 0)
                HN_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, hn_nms2_rounds_kernel, 256, 0));
            int sms = hn_device_sm_count();
            // few, fat CTAs: the loop is dominated by grid-wide barriers, whose cost grows with the CTA count
            // and a cooperative grid must be co-resident as a whole: with g_det_rounds_ctas_per_sm = 1 it leaves room on
            // every SM for a persistent conv CTA of a concurrently running branch of the plan
            const int per_sm = blocks_per_sm > g_det_rounds_ctas_per_sm ? g_det_rounds_ctas_per_sm : (blocks_per_sm > 0 ? blocks_per_sm : 1);
            long long want = (NA + 255) / 256, cap = (long long)sms * per_sm;
            dim3 grid((unsigned)(want < cap ? want : cap));
            long long na = NA;
            int passes = g_det_rounds_passes;
            void* args[] = {&ws, &na, &passes};
            HN_CHECK_CUDA(cudaLaunchCooperativeKernel((void*)hn_nms2_rounds_kernel, grid, dim3(256), args, 0, s));
        }
        tb = ws.cub2_bytes;
        HN_CHECK_CUDA(cub::DeviceScan::ExclusiveSum(ws.cub2_tmp, tb, ws.kflag, ws.kscan, (int)NA, s));
        hn_nms2_compact_kernel<<<hn_cdiv(NA, 256), 256, 0, s>>>(ws, NA);
        HN_CHECK_CUDA(cudaGetLastError());
    }
    // sequential kernel: whole problem when pruning is impossible (thr ~ 0), otherwise only flagged images
    hn_det_nms_kernel<<<d->N * kMaxCls, kNmsChunk, 0, s>>>(ws, d->A, d->iou_thres, d->nms_mode, geom, parallel ? 1 : 0);
    HN_CHECK_CUDA(cudaGetLastError());
    hn_det_gather_kernel<<<dim3(d->N * kMaxCls, 8), 256, 0, s>>>(ws, d->A, d->out_boxes, d->out_scores, d->out_class, d->out_count);
    HN_CHECK_CUDA(cudaGetLastError());
    if (d->out_cand) {
        hn_copy_i32_kernel<<<hn_cdiv(d->N, 256), 256, 0, s>>>(ws.n_cand, d->out_cand, d->N);
        HN_CHECK_CUDA(cudaGetLastError());
    }
    return HN_OK;
}

// ------------------------------------------------------------------------------------------------
// lane decode + NMS: one CTA per image
// ------------------------------------------------------------------------------------------------
struct LaneParams {
    const float* cls;
    const float* loc;
    int N, fh, fw, ppl, cls_is_prob, use_mean;
    float conf_thres, nms_thres;
    double step_w, interval, ppa;
    float x_lim_up, x_lim_down;
    float* ws_x;    // [N][na][ppl] x by absolute position
    int* ws_i;      // [N][na][4]: anchor, start, end, order
    float* ws_prob; // [N][na]
    int* out_count;
    int* out_meta;
    float* out_prob;
    float* out_x;
    int* out_cand;
};

extern "C" int64_t hn_lane_workspace_bytes(int32_t N, int32_t n_anchor, int32_t ppl) {
    if (N <= 0 || n_anchor <= 0 || ppl <= 0) return 256;
    return (int64_t)(align_up((size_t)N * n_anchor * ppl * 4) + align_up((size_t)N * n_anchor * 16) +
                     align_up((size_t)N * n_anchor * 4) + align_up((size_t)N * n_anchor * 4));
}

__device__ __forceinline__ float lane_dist(const float* x1, int s1, int e1, const float* x2, int s2, int e2, int use_mean,
                                           bool* no_overlap) {
    // calc_err_dis_with_pos (lane_codec_utils.py:487-515): x arrays are indexed by absolute position
    int ms = max(s1, s2), me = min(e1, e2);
    *no_overlap = (me <= ms) || (ms < 0) || (me < 1);
    if (*no_overlap) return 0.0f;
    float dis = 0.0f;
    for (int i = ms; i < me; ++i) dis = __fadd_rn(dis, fabsf(__fsub_rn(x1[i], x2[i])));
    dis = __fdiv_rn(dis, (float)(me - ms));
    if (use_mean) return dis;
    float ds = fabsf(__fsub_rn(x1[ms], x2[ms]));
    dis = fmaxf(dis, ds);
    float de = fabsf(__fsub_rn(x1[me - 1], x2[me - 1]));
    dis = fmaxf(dis, de);
    return dis;
}

__global__ void __launch_bounds__(256) hn_lane_kernel(const __grid_constant__ LaneParams p) {
    extern __shared__ int s_dyn[];
    const int na = p.fh * p.fw;
    int* s_order = s_dyn;            // [na] sorted position -> candidate slot
    int* s_supp = s_dyn + na;        // [na] suppressed flag per sorted position
    __shared__ int s_ncand, s_nkeep;
    const int n = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
    const int L = 2 * p.ppl + 2;
    float* wx = p.ws_x + (long long)n * na * p.ppl;
    int* wi = p.ws_i + (long long)n * na * 4;
    float* wp = p.ws_prob + (long long)n * na;
    if (tid == 0) { s_ncand = 0; s_nkeep = 0; }
    __syncthreads();

    // ---- decode: one warp per anchor (lane_codec.py:139-217) ----
    for (int a = warp; a < na; a += nwarp) {
        const float* c = p.cls + ((long long)n * na + a) * 2;
        float prob;
        if (p.cls_is_prob) {
            prob = c[1];
        } else {
            float m = fmaxf(c[0], c[1]);
            float e0 = expf(c[0] - m), e1 = expf(c[1] - m);
            prob = e1 / (e0 + e1);
        }
        if (prob < p.conf_thres) continue;
        const int h = a / p.fw, w = a - h * p.fw;
        const int y_pos = (int)((double)(p.fh - 1 - h) * p.ppa);
        const float cx = (float)(((double)w + 0.5) * p.step_w);
        const float itv = (float)p.interval;
        const float* loc = p.loc + ((long long)n * na + a) * L;
        const float end_up = loc[p.ppl + 1], end_down = loc[p.ppl];
        // up branch: first failing index
        int n_up = p.ppl;
        for (int base = 0; base < p.ppl; base += 32) {
            int i = base + lane;
            bool fail = true;
            float x = 0.0f;
            if (i < p.ppl) {
                x = __fadd_rn(cx, __fmul_rn(loc[p.ppl + 2 + i], itv));
                fail = ((float)i >= end_up) || (y_pos + i >= p.ppl) || (x < 0.0f) || (x >= p.x_lim_up);
            }
            unsigned m = __ballot_sync(0xffffffffu, fail);
            if (m) { n_up = base + __ffs(m) - 1; break; }
        }
        int n_down = y_pos;
        for (int base = 0; base < y_pos; base += 32) {
            int i = base + lane;
            bool fail = true;
            if (i < y_pos) {
                float x = __fadd_rn(cx, __fmul_rn(loc[i], itv));
                fail = ((float)i >= end_down) || (y_pos - 1 - i < 0) || (x < 0.0f) || (x >= p.x_lim_down);
            }
            unsigned m = __ballot_sync(0xffffffffu, fail);
            if (m) { n_down = base + __ffs(m) - 1; break; }
        }
        if (n_up + n_down < 2) continue;
        int slot = 0;
        if (lane == 0) slot = atomicAdd(&s_ncand, 1);
        slot = __shfl_sync(0xffffffffu, slot, 0);
        const int start = y_pos - n_down, end = y_pos + n_up;
        float* xs = wx + (long long)slot * p.ppl;
        for (int pos = start + lane; pos < end; pos += 32) {
            float x = pos < y_pos ? __fadd_rn(cx, __fmul_rn(loc[y_pos - 1 - pos], itv))
                                  : __fadd_rn(cx, __fmul_rn(loc[p.ppl + 2 + pos - y_pos], itv));
            xs[pos] = x;
        }
        if (lane == 0) {
            wi[slot * 4 + 0] = a;
            wi[slot * 4 + 1] = start;
            wi[slot * 4 + 2] = end;
            wp[slot] = prob;
        }
    }
    __syncthreads();
    const int nc = s_ncand;
    // ---- stable sort by prob descending (sorted(lane_set) with Lane.__lt__, lane_codec_utils.py:62-64):
    //      rank = number of candidates that come first; ties keep (h, w) scan order = anchor order
    for (int i = tid; i < nc; i += blockDim.x) {
        float pi = wp[i];
        int ai = wi[i * 4];
        int rank = 0;
        for (int j = 0; j < nc; ++j) {
            float pj = wp[j];
            rank += (pj > pi) || (pj == pi && wi[j * 4] < ai);
        }
        s_order[rank] = i;
        s_supp[rank] = 0;
    }
    __syncthreads();
    // ---- greedy NMS (nms_with_pos, lane_codec_utils.py:518-542) ----
    const bool far_suppresses = 10e6 <= (double)p.nms_thres;
    for (int i = 0; i < nc; ++i) {
        if (s_supp[i]) continue;  // uniform: s_supp only changes between barriers
        const int si = s_order[i];
        if (tid == 0) {
            int k = s_nkeep++;
            int* om = p.out_meta + ((long long)n * na + k) * 4;
            om[0] = wi[si * 4]; om[1] = wi[si * 4 + 1]; om[2] = wi[si * 4 + 2]; om[3] = wi[si * 4 + 2] - wi[si * 4 + 1];
            p.out_prob[(long long)n * na + k] = wp[si];
        }
        const float* xi = wx + (long long)si * p.ppl;
        const int st_i = wi[si * 4 + 1], en_i = wi[si * 4 + 2];
        for (int t = i + 1 + tid; t < nc; t += blockDim.x) {
            if (s_supp[t]) continue;
            const int stt = s_order[t];
            bool no_ov;
            float dis = lane_dist(xi, st_i, en_i, wx + (long long)stt * p.ppl, wi[stt * 4 + 1], wi[stt * 4 + 2], p.use_mean,
                                  &no_ov);
            bool sup = no_ov ? far_suppresses : (dis <= p.nms_thres);
            if (sup) s_supp[t] = 1;
        }
        __syncthreads();
    }
    __syncthreads();
    // ---- emit x of kept lanes in list order (bottom -> top) ----
    const int nk = s_nkeep;
    for (int k = warp; k < nk; k += nwarp) {
        const int* om = p.out_meta + ((long long)n * na + k) * 4;
        // find the slot again through the anchor id
        int a = om[0], start = om[1], cnt = om[3];
        int slot = -1;
        for (int j = lane; j < nc; j += 32)
            if (wi[j * 4] == a) slot = j;
        for (int o = 16; o > 0; o >>= 1) slot = max(slot, __shfl_xor_sync(0xffffffffu, slot, o));
        const float* xs = wx + (long long)slot * p.ppl;
        float* ox = p.out_x + ((long long)n * na + k) * p.ppl;
        for (int q = lane; q < cnt; q += 32) ox[q] = xs[start + q];
    }
    if (tid == 0) {
        p.out_count[n] = nk;
        if (p.out_cand) p.out_cand[n] = nc;
    }
}

extern "C" int hn_lane_decode_nms(const hn_lane_desc* d, void* stream) {
    HN_REQUIRE(d != nullptr, "lane: null desc");
    HN_REQUIRE(d->N >= 0 && d->fh >= 1 && d->fw >= 1 && d->ppl >= 1, "lane: bad sizes");
    if (d->N == 0) return HN_OK;
    HN_REQUIRE(d->cls && d->loc && d->workspace && d->out_count && d->out_meta && d->out_prob && d->out_x, "lane: null pointer");
    const int na = d->fh * d->fw;
    HN_REQUIRE(na <= 8192, "lane: at most 8192 anchors per image");
    LaneParams p;
    p.cls = d->cls; p.loc = d->loc;
    p.N = d->N; p.fh = d->fh; p.fw = d->fw; p.ppl = d->ppl; p.cls_is_prob = d->cls_is_prob; p.use_mean = d->use_mean;
    p.conf_thres = d->conf_thres; p.nms_thres = d->nms_thres;
    p.step_w = d->step_w; p.interval = d->interval; p.ppa = d->points_per_anchor;
    p.x_lim_up = d->input_width;
    p.x_lim_down = (float)((double)d->input_width + (double)d->margin_width);
    uint8_t* w = reinterpret_cast<uint8_t*>(d->workspace);
    p.ws_x = reinterpret_cast<float*>(w);
    w += align_up((size_t)d->N * na * d->ppl * 4);
    p.ws_i = reinterpret_cast<int*>(w);
    w += align_up((size_t)d->N * na * 16);
    p.ws_prob = reinterpret_cast<float*>(w);
    p.out_count = d->out_count; p.out_meta = d->out_meta; p.out_prob = d->out_prob; p.out_x = d->out_x; p.out_cand = d->out_cand;
    size_t smem = (size_t)na * 2 * sizeof(int);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        HN_CHECK_CUDA(cudaFuncSetAttribute(hn_lane_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
    }
    hn_lane_kernel<<<d->N, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(p);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}


// ------------------------------------------------------------------------------------------------
// host tails of the decoders, on the device
// ------------------------------------------------------------------------------------------------
// DetectionHeader.invert_affine (head_detect/detection.py:218-230): rois[:, [0, 2]] /= (new_w / old_w), rois[:, [1, 3]] /=
// (new_h / old_h) -- numpy float32 arrays divided by a Python float: the quotient of the two ints is formed in double, cast
// to float32, and the division runs in fp32 (NEP 50).  `scale` holds that float32 divisor per image: [N][2] = (x, y).
__global__ void hn_det_invert_affine_kernel(float* __restrict__ boxes, const int* __restrict__ count, int N, int A, const float* __restrict__ scale) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)N * A) return;
    const int n = (int)(i / A), k = (int)(i - (long long)n * A);
    if (k >= count[n]) return;
    float4 b = reinterpret_cast<float4*>(boxes)[i];
    const float sx = scale[n * 2], sy = scale[n * 2 + 1];
    b.x = __fdiv_rn(b.x, sx); b.z = __fdiv_rn(b.z, sx);
    b.y = __fdiv_rn(b.y, sy); b.w = __fdiv_rn(b.w, sy);
    reinterpret_cast<float4*>(boxes)[i] = b;
}
extern "C" int hn_det_invert_affine(float* boxes, const int32_t* count, int32_t N, int32_t A, const float* scale_xy, void* stream) {
    HN_REQUIRE(boxes && count && scale_xy && N >= 0 && A >= 0, "invert_affine: bad arguments");
    const long long total = (long long)N * A;
    if (total == 0) return HN_OK;
    hn_det_invert_affine_kernel<<<hn_cdiv(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(boxes, count, N, A, scale_xy);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}

// LaneHeader.scale_to_org (lanedetect.py:118-124 -> lane_codec_utils.py:185-282): for every kept lane the ordering keys of
// LaneWithCrossK (slope k and x where the extension of its two bottom points crosses y = net_height - 1; lane_codec_utils.py:128-182)
// and the points scaled to the original frame.  Arithmetic as Python / numpy perform it: x is numpy float32, y a Python float
// (double); float32 (op) Python-float runs in fp32 with the double operand rounded to fp32 first; y * sy stays in double.
// keys [N][na][4] = (cross_x, k, first x, last x); xo fp32 [N][na][ppl]; yo fp64 [N][na][ppl].
__global__ void hn_lane_scale_kernel(const int* __restrict__ count, const int* __restrict__ meta, const float* __restrict__ xs, int na, int ppl,
                                     double input_height, double interval, double cross_y, double sx, double sy, float* __restrict__ keys,
                                     float* __restrict__ xo, double* __restrict__ yo) {
    const int n = blockIdx.y, k = blockIdx.x;
    if (k >= count[n]) return;
    const int* m = meta + ((long long)n * na + k) * 4;
    const int start = m[1], npts = m[3];
    const float* x = xs + ((long long)n * na + k) * ppl;
    const float fsx = (float)sx;
    for (int q = threadIdx.x; q < npts; q += blockDim.x) {
        const double y = input_height - 1.0 - (double)(start + q) * interval;
        xo[((long long)n * na + k) * ppl + q] = __fmul_rn(x[q], fsx);
        yo[((long long)n * na + k) * ppl + q] = y * sy;
    }
    if (threadIdx.x == 0) {
        float* key = keys + ((long long)n * na + k) * 4;
        const double y0 = input_height - 1.0 - (double)start * interval, y1 = input_height - 1.0 - (double)(start + 1) * interval;
        // k = (p1.x - p0.x) / (p1.y - p0.y); cross_x = kk * y + (p0.x - kk * p0.y) with kk = (p0.x - p1.x) / (p0.y - p1.y)
        const float kslope = __fdiv_rn(__fsub_rn(x[1], x[0]), (float)(y1 - y0));
        float cross = -1.0f;
        if (fabs(y0 - y1) >= 1e-6) {
            const float kk = __fdiv_rn(__fsub_rn(x[0], x[1]), (float)(y0 - y1));
            const float b = __fsub_rn(x[0], __fmul_rn(kk, (float)y0));
            cross = __fadd_rn(__fmul_rn(kk, (float)cross_y), b);
        }
        key[0] = cross; key[1] = kslope; key[2] = x[0]; key[3] = x[npts - 1];
    }
}
extern "C" int hn_lane_scale_to_org(const int32_t* count, const int32_t* meta, const float* xs, int32_t N, int32_t n_anchor, int32_t ppl,
                                    double input_height, double interval, double cross_y, double scale_x, double scale_y, float* keys, float* x_out,
                                    double* y_out, void* stream) {
    HN_REQUIRE(count && meta && xs && keys && x_out && y_out && N >= 0 && n_anchor >= 1 && ppl >= 2, "lane scale: bad arguments");
    if (N == 0) return HN_OK;
    hn_lane_scale_kernel<<<dim3((unsigned)n_anchor, (unsigned)N), 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(count, meta, xs, n_anchor, ppl, input_height,
                                                                                                                  interval, cross_y, scale_x, scale_y, keys, x_out, y_out);
    HN_CHECK_CUDA(cudaGetLastError());
    return HN_OK;
}
