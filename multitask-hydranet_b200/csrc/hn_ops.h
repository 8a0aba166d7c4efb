// Internal (non-ABI) declarations shared between translation units.
#pragma once
#include "hn_common.cuh"

struct alignas(64) ConvParams {
    CUtensorMap tmA[HN_MAX_SRC];
    CUtensorMap tmArun[HN_MAX_SRC];  // the same views with a (TH + 2)-row box: one load feeds a run of dy taps
    int run_max, a_stage_bytes;      // taps per ring stage (1 or 3), bytes of a stage's A region
    CUtensorMap tmB;
    CUtensorMap tmBpart;  // B tile slice of bn/cluster rows (cluster multicast)
    int cluster;          // CTAs per cluster sharing the weight tile (1, 2 or 4)
    int m_groups;         // ceil(m_tiles / cluster)
    CUtensorMap tmO;  // bf16 output view (TMA store path)
    int n_staging;    // 0: direct stores; 1/2: shared-memory staging buffers of 128 rows x 64 channels
    int flat, TH, TW, n_img, H, W, tiles_x, tiles_y, flat_hw, flat_m;
    int num_taps, cout, bn, stages, tmem_cols;
    int m_tiles, n_tiles, acc_stride;
    unsigned long long div_m_groups, div_per_img, div_tiles_x, div_flat_hw;  // hn_fastdiv magics
    int tr_off;    // byte offset (from the aligned shared-memory base) of the fp32 transpose scratch, 0 = none
    int tw_shift;                                                            // log2(TW)
    long long* dbg;  // optional per-CTA timestamps (globaltimer ns), 16 slots per CTA
    const float* bias;
    int act, epi;
    void* out;
    int out_fp32;
    long long osn, osy, osx;
    int oscale, ooy, oox, halo;
    const bf16* res;
    long long rsn, rsy, rsx;
    int res_relu, grouped;
    uint8_t* out2;
    int n_cls;
    int n_groups, group_addr;
    int group_end[HN_MAX_GROUPS], group_hw[HN_MAX_GROUPS];
    long long group_out_base[HN_MAX_GROUPS];
    const float* group_scale;
    const float* group_shift;
    int gstride;  // floats between consecutive groups in group_scale / group_shift
    hn_tap taps[HN_MAX_TAPS];
};


struct ConvLaunch {
    ConvParams prm;
    dim3 grid;
    size_t smem;
    int cluster;
    int per_sm;  // CTAs of this launch that fit on one SM (1 or 2)
};

#ifdef __CUDACC__
// device-side view with 32-bit element strides: every offset inside one activation tensor fits an int (checked in
// check_view), and 64-bit address arithmetic would otherwise dominate the instruction count of these kernels
struct View {
    const bf16* ptr;
    int N, H, W, C;
    int sn, sy, sx;
};
static inline View to_view(const hn_view& v) {
    View r;
    r.ptr = reinterpret_cast<const bf16*>(v.ptr);
    r.N = v.N; r.H = v.H; r.W = v.W; r.C = v.C;
    r.sn = (int)v.stride_n; r.sy = (int)v.stride_y; r.sx = (int)v.stride_x;
    return r;
}
static inline int check_view(const hn_view& v, const char* what) {
    HN_REQUIRE(v.ptr != nullptr, "%s: null view", what);
    {
        long long span = (long long)(v.N > 0 ? v.N - 1 : 0) * v.stride_n + (long long)(v.H > 0 ? v.H - 1 : 0) * v.stride_y +
                         (long long)(v.W > 0 ? v.W - 1 : 0) * v.stride_x + v.C;
        HN_REQUIRE(span < 0x7fffffffLL && v.stride_n >= 0 && v.stride_y >= 0 && v.stride_x >= 0,
                   "%s: view spans more than 2^31 elements", what);
    }
    HN_REQUIRE((reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0 && v.C % 8 == 0 && v.stride_x % 8 == 0 && v.stride_y % 8 == 0 &&
                   v.stride_n % 8 == 0,
               "%s: view must be 16-byte aligned with C and strides in multiples of 8 (C=%d)", what, v.C);
    return HN_OK;
}

__device__ __forceinline__ void load8(const bf16* p, float (&f)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 a = hn_unpack_bf16x2(u.x), b = hn_unpack_bf16x2(u.y), c = hn_unpack_bf16x2(u.z), d = hn_unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(bf16* p, const float (&f)[8]) {
    uint4 u;
    u.x = hn_pack_bf16x2(f[0], f[1]); u.y = hn_pack_bf16x2(f[2], f[3]);
    u.z = hn_pack_bf16x2(f[4], f[5]); u.w = hn_pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
__device__ __forceinline__ const bf16* vptr(const View& v, int n, int y, int x, int c) {
    return v.ptr + (n * v.sn + y * v.sy + x * v.sx + c);
}

#endif  // __CUDACC__

int hn_conv_prepare(const hn_conv_desc* d, ConvLaunch* L);
int hn_conv_launch(const ConvLaunch* L, cudaStream_t stream);
int hn_det_num_launches(const hn_det_desc* d);
