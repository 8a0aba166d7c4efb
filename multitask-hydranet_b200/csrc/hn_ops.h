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

int hn_conv_prepare(const hn_conv_desc* d, ConvLaunch* L);
int hn_conv_launch(const ConvLaunch* L, cudaStream_t stream);
int hn_det_num_launches(const hn_det_desc* d);
