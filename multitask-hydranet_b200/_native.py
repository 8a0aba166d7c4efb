"""ctypes binding of ``libhydranet_b200.so`` (the C ABI declared in ``include/hydranet_b200.h``).

The library is loaded eagerly and loudly: if it is missing the import raises -- there is no CPU or
PyTorch fallback behind any entry point.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhydranet_b200.so")

HN_MAX_SRC = 6
HN_MAX_TAPS = 96
HN_MAX_GROUPS = 8

ACT_NONE, ACT_RELU, ACT_SWISH, ACT_ELU, ACT_SIGMOID = range(5)
HALO_NONE, HALO_REFLECT, HALO_REPLICATE = range(3)
EPI_STD, EPI_SEGOUT = range(2)
IN_SAME, IN_UP2, IN_POOL = range(3)
POOL_ZERO_RB, POOL_NEGINF = range(2)
NMS_AUTO_CUDA, NMS_AUTO_CPU, NMS_TRICK, NMS_VANILLA = range(4)


class View(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32),
                ("stride_n", C.c_int64), ("stride_y", C.c_int64), ("stride_x", C.c_int64)]


class Tap(C.Structure):
    _fields_ = [("src", C.c_int8), ("dy", C.c_int8), ("dx", C.c_int8), ("rsv0", C.c_int8),
                ("c0", C.c_int16), ("rsv1", C.c_int16)]


class ConvDesc(C.Structure):
    _fields_ = [("src", View * HN_MAX_SRC), ("n_src", C.c_int32), ("weight", C.c_void_p), ("w_rows", C.c_int32),
                ("num_taps", C.c_int32), ("taps", Tap * HN_MAX_TAPS), ("flat", C.c_int32),
                ("tile_h", C.c_int32), ("tile_w", C.c_int32),
                ("n_img", C.c_int32), ("out_h", C.c_int32), ("out_w", C.c_int32), ("flat_hw", C.c_int32),
                ("cout", C.c_int32), ("bn", C.c_int32), ("stages", C.c_int32),
                ("bias", C.c_void_p), ("act", C.c_int32), ("epi", C.c_int32),
                ("out", C.c_void_p), ("out_fp32", C.c_int32),
                ("out_stride_n", C.c_int64), ("out_stride_y", C.c_int64), ("out_stride_x", C.c_int64),
                ("out_scale", C.c_int32), ("out_oy", C.c_int32), ("out_ox", C.c_int32), ("halo", C.c_int32),
                ("res", C.c_void_p), ("res_stride_n", C.c_int64), ("res_stride_y", C.c_int64), ("res_stride_x", C.c_int64),
                ("res_relu", C.c_int32), ("grouped", C.c_int32), ("out2", C.c_void_p), ("n_cls", C.c_int32),
                ("n_groups", C.c_int32), ("group_end", C.c_int32 * HN_MAX_GROUPS), ("group_scale", C.c_void_p),
                ("group_shift", C.c_void_p), ("group_addr", C.c_int32), ("group_hw", C.c_int32 * HN_MAX_GROUPS),
                ("group_out_base", C.c_int64 * HN_MAX_GROUPS)]


class StemDesc(C.Structure):
    _fields_ = [("x", C.c_void_p), ("N", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("w", C.c_void_p), ("b", C.c_void_p), ("out", View), ("no_relu", C.c_int32)]


class NodeDesc(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("in_", View * 3), ("mode", C.c_int32 * 3), ("w", C.c_float * 3),
                ("swish", C.c_int32), ("dw", C.c_void_p), ("out", View)]


class DwMultiDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("in_", View * HN_MAX_GROUPS), ("out", View * HN_MAX_GROUPS), ("dw", C.c_void_p)]


class PoolDesc(C.Structure):
    _fields_ = [("in_", View), ("out", View), ("mode", C.c_int32)]


class LaneFuseDesc(C.Structure):
    _fields_ = [("p3", View), ("p4", View), ("p5", View), ("p6", View), ("out", View), ("stride", C.c_int32)]


class SePoolDesc(C.Structure):
    _fields_ = [("x", View), ("pix_per_block", C.c_int32), ("partial", C.c_void_p), ("counter", C.c_void_p), ("mean", C.c_void_p),
                ("S", C.c_int32), ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("gate", C.c_void_p)]


class GconvSeDesc(C.Structure):
    _fields_ = [("vin", View), ("weight", C.c_void_p), ("bias", C.c_void_p), ("se", SePoolDesc)]


class SegLossDesc(C.Structure):
    _fields_ = [("logits", C.c_void_p), ("target", C.c_void_p), ("weight", C.c_void_p), ("N", C.c_int32), ("C", C.c_int32),
                ("HW", C.c_int64), ("k", C.c_int64), ("ignore_index", C.c_int32), ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
                ("loss", C.c_void_p), ("gout", C.c_void_p), ("dlogits", C.c_void_p)]


class SeScaleDesc(C.Structure):
    _fields_ = [("x", View), ("scale", C.c_void_p)]


class PreprocessDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("N", C.c_int32), ("src_h", C.c_int32), ("src_w", C.c_int32),
                ("src_pitch", C.c_int64), ("src_stride", C.c_int64), ("dst", C.c_void_p),
                ("dst_h", C.c_int32), ("dst_w", C.c_int32)]


class DetDesc(C.Structure):
    _fields_ = [("anchors", C.c_void_p), ("regression", C.c_void_p), ("classification", C.c_void_p),
                ("N", C.c_int32), ("A", C.c_int32), ("ncls", C.c_int32), ("img_h", C.c_int32), ("img_w", C.c_int32),
                ("conf_thres", C.c_float), ("iou_thres", C.c_float), ("nms_mode", C.c_int32),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_int64),
                ("out_boxes", C.c_void_p), ("out_scores", C.c_void_p), ("out_class", C.c_void_p),
                ("out_count", C.c_void_p), ("out_cand", C.c_void_p), ("pre_boxes", C.c_void_p)]


class LaneDesc(C.Structure):
    _fields_ = [("cls", C.c_void_p), ("loc", C.c_void_p),
                ("N", C.c_int32), ("fh", C.c_int32), ("fw", C.c_int32), ("ppl", C.c_int32),
                ("cls_is_prob", C.c_int32), ("conf_thres", C.c_float), ("nms_thres", C.c_float), ("use_mean", C.c_int32),
                ("step_w", C.c_double), ("interval", C.c_double), ("points_per_anchor", C.c_double),
                ("input_width", C.c_float), ("margin_width", C.c_float), ("workspace", C.c_void_p),
                ("out_count", C.c_void_p), ("out_meta", C.c_void_p), ("out_prob", C.c_void_p), ("out_x", C.c_void_p),
                ("out_cand", C.c_void_p)]



# ---- training step (include/hydranet_b200.h, "Training step") ----
HN_MAX_SEG = 8
RS_UP2, RS_POOL_ZERO, RS_POOL_NEGINF = range(3)


class Mat(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int32), ("ld", C.c_int64)]


class BnDesc(C.Structure):
    _fields_ = [("z", Mat), ("n_seg", C.c_int32), ("seg_end", C.c_int64 * HN_MAX_SEG),
                ("gamma", C.c_void_p * HN_MAX_SEG), ("beta", C.c_void_p * HN_MAX_SEG),
                ("running_mean", C.c_void_p * HN_MAX_SEG), ("running_var", C.c_void_p * HN_MAX_SEG),
                ("eps", C.c_float), ("momentum", C.c_float), ("stats", C.c_void_p), ("act", C.c_int32),
                ("res", Mat), ("y", Mat), ("scratch", C.c_void_p), ("scratch_bytes", C.c_int64),
                ("dy", Mat), ("dz", Mat), ("dres", Mat),
                ("dgamma", C.c_void_p * HN_MAX_SEG), ("dbeta", C.c_void_p * HN_MAX_SEG)]


class ActBwdDesc(C.Structure):
    _fields_ = [("dy", Mat), ("ref", Mat), ("dz", Mat), ("act", C.c_int32), ("n_scaled", C.c_int32),
                ("scaled", Mat * 3), ("w", C.c_void_p)]


class WsumDesc(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("in_", Mat * 3), ("w", C.c_void_p), ("s", Mat), ("a", Mat)]


class ResampleDesc(C.Structure):
    _fields_ = [("mode", C.c_int32), ("x", View), ("y", View), ("dy", View), ("dx", View)]


class SegGatherDesc(C.Structure):
    _fields_ = [("low", View), ("skip", View), ("out", View), ("dlow", View), ("dskip", View)]


class HeadGradDesc(C.Structure):
    _fields_ = [("dout", C.c_void_p), ("out", C.c_void_p), ("act", C.c_int32), ("cols_valid", C.c_int32),
                ("stride_n", C.c_int64), ("stride_pix", C.c_int64), ("stride_c", C.c_int64), ("rows_per_img", C.c_int64),
                ("n_groups", C.c_int32), ("group_end", C.c_int64 * HN_MAX_GROUPS), ("group_hw", C.c_int64 * HN_MAX_GROUPS),
                ("group_out_base", C.c_int64 * HN_MAX_GROUPS), ("dz", Mat)]


class SeFcDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("C", C.c_int32), ("S", C.c_int32), ("mean", C.c_void_p),
                ("w1", C.c_void_p), ("b1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p),
                ("h", C.c_void_p), ("gate", C.c_void_p), ("dgate", C.c_void_p), ("dmean", C.c_void_p),
                ("dw1", C.c_void_p), ("db1", C.c_void_p), ("dw2", C.c_void_p), ("db2", C.c_void_p), ("tmp", C.c_void_p)]


class PackEntry(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("dst_ld", C.c_int64), ("rows", C.c_int32), ("rows_pad", C.c_int32),
                ("cols", C.c_int32), ("s_r", C.c_int64), ("s_c", C.c_int64), ("grouped", C.c_int32), ("rsv", C.c_int32)]


class WgradDesc(C.Structure):
    _fields_ = [("dy", View), ("src", View * HN_MAX_SRC), ("n_src", C.c_int32), ("num_taps", C.c_int32),
                ("taps", Tap * HN_MAX_TAPS), ("tap_off", C.c_int64 * HN_MAX_TAPS), ("tap_cin", C.c_int32 * HN_MAX_TAPS),
                ("flat", C.c_int32), ("tile_h", C.c_int32), ("tile_w", C.c_int32), ("cout", C.c_int32),
                ("s_co", C.c_int64), ("s_ci", C.c_int64), ("grouped", C.c_int32), ("dw", C.c_void_p)]


class AdamTensor(C.Structure):
    _fields_ = [("p", C.c_void_p), ("g", C.c_void_p), ("m", C.c_void_p), ("v", C.c_void_p), ("n", C.c_int64)]


#: every symbol ``include/hydranet_b200.h`` declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "hn_conv_fwd": (C.c_int, [C.POINTER(ConvDesc), _P]),
    "hn_stem_fwd": (C.c_int, [C.POINTER(StemDesc), _P]),
    "hn_node_fwd": (C.c_int, [C.POINTER(NodeDesc), _P]),
    "hn_dw_multi_fwd": (C.c_int, [C.POINTER(DwMultiDesc), _P]),
    "hn_pool_fwd": (C.c_int, [C.POINTER(PoolDesc), _P]),
    "hn_lanefuse_fwd": (C.c_int, [C.POINTER(LaneFuseDesc), _P]),
    "hn_se_pool_fwd": (C.c_int, [C.POINTER(SePoolDesc), _P]),
    "hn_se_scale_fwd": (C.c_int, [C.POINTER(SeScaleDesc), _P]),
    "hn_se_fused_fwd": (C.c_int, [C.POINTER(SePoolDesc), _P]),
    "hn_se_fused_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "hn_gconv_se_fwd": (C.c_int, [C.POINTER(GconvSeDesc), _P]),
    "hn_gconv_se_supported": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.c_int32]),
    "hn_se_fused_set_debug": (None, [_P]),
    "hn_stem_set_mma": (None, [C.c_int]),
    "hn_preprocess_fwd": (C.c_int, [C.POINTER(PreprocessDesc), _P]),
    "hn_seg_argmax": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int64, _P, _P, _P]),
    "hn_u8_to_i64": (C.c_int, [_P, _P, C.c_int64, _P]),
    "hn_det_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "hn_det_decode_nms": (C.c_int, [C.POINTER(DetDesc), _P]),
    "hn_lane_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int32, C.c_int32]),
    "hn_lane_decode_nms": (C.c_int, [C.POINTER(LaneDesc), _P]),
    "hn_bn_train_fwd": (C.c_int, [C.POINTER(BnDesc), _P]),
    "hn_bn_train_bwd": (C.c_int, [C.POINTER(BnDesc), _P]),
    "hn_col_reduce": (C.c_int, [C.POINTER(Mat), C.POINTER(Mat), C.c_int32, C.c_int64, _P, _P, C.c_float, _P, C.c_int64, _P]),
    "hn_act_bwd": (C.c_int, [C.POINTER(ActBwdDesc), _P]),
    "hn_wsum_swish_fwd": (C.c_int, [C.POINTER(WsumDesc), _P]),
    "hn_resample_fwd": (C.c_int, [C.POINTER(ResampleDesc), _P]),
    "hn_resample_bwd": (C.c_int, [C.POINTER(ResampleDesc), _P]),
    "hn_seggather_fwd": (C.c_int, [C.POINTER(SegGatherDesc), _P]),
    "hn_seggather_bwd": (C.c_int, [C.POINTER(SegGatherDesc), _P]),
    "hn_head_grad": (C.c_int, [C.POINTER(HeadGradDesc), _P]),
    "hn_dw_wgrad": (C.c_int, [C.POINTER(View), C.POINTER(View), _P, C.c_int32, _P, C.c_int64, _P]),
    "hn_stem_wgrad": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(View), _P, _P, C.c_int64, _P]),
    "hn_se_fc_fwd": (C.c_int, [C.POINTER(SeFcDesc), _P]),
    "hn_se_fc_bwd": (C.c_int, [C.POINTER(SeFcDesc), _P]),
    "hn_se_apply": (C.c_int, [C.POINTER(Mat), _P, _P, C.c_int64, C.POINTER(Mat), _P]),
    "hn_pack_weights": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "hn_conv_wgrad": (C.c_int, [C.POINTER(WgradDesc), _P]),
    "hn_seg_loss_workspace_bytes": (C.c_int64, [C.c_int32, C.c_int64]),
    "hn_seg_loss_fwd": (C.c_int, [C.POINTER(SegLossDesc), _P]),
    "hn_seg_loss_bwd": (C.c_int, [C.POINTER(SegLossDesc), _P]),
    "hn_lane_loss_workspace_bytes": (C.c_int64, [C.c_int32]),
    "hn_lane_loss": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, C.c_int64, _P, _P, _P, _P]),
    "hn_det_loss": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_float, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "hn_adam_step": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32,
                               C.c_float, _P, _P]),
    "hn_det_invert_affine": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P]),
    "hn_lane_scale_to_org": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double,
                                       _P, _P, _P, _P]),
    "hn_plan_create": (C.c_int, [C.POINTER(_P)]),
    "hn_plan_destroy": (C.c_int, [_P]),
    "hn_plan_add_conv": (C.c_int, [_P, C.POINTER(ConvDesc)]),
    "hn_plan_add_wait": (C.c_int, [_P, C.c_int]),
    "hn_plan_add_stem": (C.c_int, [_P, C.POINTER(StemDesc)]),
    "hn_plan_add_node": (C.c_int, [_P, C.POINTER(NodeDesc)]),
    "hn_plan_add_dw_multi": (C.c_int, [_P, C.POINTER(DwMultiDesc)]),
    "hn_plan_add_pool": (C.c_int, [_P, C.POINTER(PoolDesc)]),
    "hn_plan_add_lanefuse": (C.c_int, [_P, C.POINTER(LaneFuseDesc)]),
    "hn_plan_add_se_pool": (C.c_int, [_P, C.POINTER(SePoolDesc)]),
    "hn_plan_add_se_scale": (C.c_int, [_P, C.POINTER(SeScaleDesc)]),
    "hn_plan_add_se_fused": (C.c_int, [_P, C.POINTER(SePoolDesc)]),
    "hn_plan_add_gconv_se": (C.c_int, [_P, C.POINTER(GconvSeDesc)]),
    "hn_plan_add_det": (C.c_int, [_P, C.POINTER(DetDesc)]),
    "hn_plan_add_lane": (C.c_int, [_P, C.POINTER(LaneDesc)]),
    "hn_plan_size": (C.c_int, [_P]),
    "hn_plan_num_launches": (C.c_int, [_P]),
    "hn_plan_run": (C.c_int, [_P, _P]),
    "hn_plan_run_range": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "hn_plan_set_branch": (C.c_int, [_P, C.c_int]),
    "hn_plan_graph_capture": (C.c_int, [_P, _P]),
    "hn_plan_graph_launch": (C.c_int, [_P, _P]),
    "hn_conv_set_debug_buffer": (None, [_P]),
    "hn_se_set_split_fc": (None, [C.c_int]),
    "hn_se_pool_num_launches": (C.c_int, [C.POINTER(SePoolDesc)]),
    "hn_conv_set_cluster": (None, [C.c_int]),
    "hn_conv_set_tap_runs": (None, [C.c_int]),
    "hn_set_pdl": (None, [C.c_int]),
    "hn_det_set_rounds_passes": (None, [C.c_int]),
    "hn_plan_set_branch_priority": (None, [C.c_int]),
    "hn_det_set_rounds_ctas_per_sm": (None, [C.c_int]),
    "hn_conv_set_pair_min_bn": (None, [C.c_int]),
    "hn_det_set_debug_buffer": (None, [_P]),
    "hn_det_force_sequential": (None, [C.c_int]),
    "hn_version": (C.c_int, []),
    "hn_last_error": (C.c_char_p, []),
    "hn_device_sm_count": (C.c_int, []),
}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "hydranet_b200: native library %s is missing. Build it with `python -c 'import __graft_entry__ as g; "
        "g.build()'` (or ./build.sh). There is no CPU / PyTorch fallback for this path." % LIB_PATH)

lib = C.CDLL(LIB_PATH)
for _name, (_res, _args) in SYMBOLS.items():
    _fn = getattr(lib, _name)  # AttributeError here == header / library mismatch
    _fn.restype = _res
    _fn.argtypes = _args
# tuning knobs for A/B measurements (tools/, bench.py); the defaults are the shipped policy
if os.environ.get("HN_TAP_RUNS"):
    lib.hn_conv_set_tap_runs(int(os.environ["HN_TAP_RUNS"]))
if os.environ.get("HN_PDL"):
    lib.hn_set_pdl(int(os.environ["HN_PDL"]))
if os.environ.get("HN_PAIR_MIN_BN"):
    lib.hn_conv_set_pair_min_bn(int(os.environ["HN_PAIR_MIN_BN"]))
if os.environ.get("HN_ROUNDS_CTAS"):
    lib.hn_det_set_rounds_ctas_per_sm(int(os.environ["HN_ROUNDS_CTAS"]))
if os.environ.get("HN_BRANCH_PRIO"):
    lib.hn_plan_set_branch_priority(int(os.environ["HN_BRANCH_PRIO"]))
if os.environ.get("HN_ROUNDS_PASSES"):
    lib.hn_det_set_rounds_passes(int(os.environ["HN_ROUNDS_PASSES"]))
if os.environ.get("HN_STEM_MMA"):
    lib.hn_stem_set_mma(int(os.environ["HN_STEM_MMA"]))
if os.environ.get("HN_SE_SPLIT_FC"):
    lib.hn_se_set_split_fc(int(os.environ["HN_SE_SPLIT_FC"]))
if os.environ.get("HN_CLUSTER"):
    lib.hn_conv_set_cluster(int(os.environ["HN_CLUSTER"]))


class NativeError(RuntimeError):
    pass


def check(rc):
    """Raise on a non-zero status, mirroring the reference's Python exceptions."""
    if rc != 0:
        raise NativeError("hydranet_b200 native call failed (code %d): %s" % (rc, lib.hn_last_error().decode()))


def view(t, N, H, W, Cc, sn, sy, sx, offset=0):
    """hn_view over a bf16 tensor's storage, ``offset`` in elements from ``t.data_ptr()``."""
    return View(t.data_ptr() + 2 * offset, N, H, W, Cc, sn, sy, sx)
