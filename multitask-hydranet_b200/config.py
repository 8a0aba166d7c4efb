"""Model configurations, restating the keys ``HydraNet.__init__`` consumes (model/model.py:33-145)
with the values of model/cfgs/hydranet_joint_{big,small}_backbone.yml.  A cfg loaded with
``yaml.safe_load`` from the reference's own YAML files works unchanged."""
import copy

_BASE = {
    "train": {"train_detect": True, "train_seg": True, "train_lane": True, "lr": 0.00001, "weight_decay": 0.00000001,
              "batch_size_train": 4, "epoch": 30, "use_distribute": False},
    "dataloader": {"network_input_width": 640, "network_input_height": 640},
    "backbone": {"initial_width": 24, "slope": 36, "quantized_param": 2.5, "network_depth": 30, "bottleneck_ratio": 1,
                 "group_width": 8, "stride": 2, "se_ratio": 4, "fpn_num_filters": 112, "fpn_cell_repeats": 3,
                 "conv_channel_coef": [64, 152, 376, 936]},
    "detection": {"num_classes": 9, "fpn_num_filters_detect": 112, "aspect_ratios_factor": [1.4, 0.7],
                  "scales_factor": [0.0, 0.333, 0.667], "box_class_repeats": 3, "pyramid_levels": 5, "anchor_scale": 2.0,
                  "class_list": ["__background__", "roadtext", "pedestrian", "guidearrow", "traffic", "obstacle", "vehicle_wheel",
                                 "roadsign", "vehicle", "vehicle_light"],
                  "loss_cls_weight": 1.0, "loss_reg_weight": 50.0, "detection_weight": 1.0},
    "segment": {"class_list": ["__background__", "road_area", "marking_area", "marking_general_area", "marking_pavement_area"],
                "class_weight": [0.1, 0.5, 1.0, 5.0, 5.0], "channel_dimension_seg_encode": [24, 112, 112, 112],
                "channel_dimension_seg_decode": [64, 128, 256, 512], "use_top_k": True, "top_k_ratio": 0.3,
                "use_focal": False, "use_lovasz": False, "segment_weight": 5.0},
    "lane": {"anchor_stride": 32, "interval": 8, "anchor_lane_num": 1, "interpolate": True, "scale_invariance": True,
             "base_channel": 448, "num_classes": 2, "conf_thres": 0.8, "nms_thres": 100,
             "loss_cls_pos_weight": 1.0, "loss_cls_neg_weight": 1.0, "loss_loc_weight": 1.0, "lane_weight": 1.0},
}


def big_cfg(width=640, height=640):
    c = copy.deepcopy(_BASE)
    c["dataloader"].update(network_input_width=width, network_input_height=height)
    return c


def small_cfg(width=640, height=640):
    c = big_cfg(width, height)
    c["backbone"].update(network_depth=16, fpn_cell_repeats=2, conv_channel_coef=[64, 152, 376])
    c["segment"].update(use_top_k=False, use_focal=True)
    return c
