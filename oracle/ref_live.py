"""ORACLE -- test infrastructure only.  Import shim for the LIVE reference (this container only:
/root/reference does not exist on the GPU box).  Recipe from SURVEY.md Appendix B."""
import os
import sys
import types

REF_ROOT = "/root/reference/model"


def available():
    return os.path.isdir(REF_ROOT)


def import_reference():
    if not available():
        raise RuntimeError("live reference not present")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    if "webcolors" not in sys.modules:  # display.py:1 imports it; drawing only
        wc = types.ModuleType("webcolors")
        wc.name_to_rgb = lambda name: types.SimpleNamespace(red=0, green=0, blue=0)
        sys.modules["webcolors"] = wc
    import torch
    if not torch.cuda.is_available():  # model.py:123 calls .cuda() at construction
        torch.nn.Module.cuda = lambda self, *a, **k: self
    saved = sys.modules.pop("model", None)
    try:
        import model as ref_model
        from head_lane.lane_codec import LaneCodec as RefLaneCodec
    finally:
        if saved is not None:
            sys.modules["model"] = saved
    return ref_model, RefLaneCodec
