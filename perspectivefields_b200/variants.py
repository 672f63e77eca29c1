"""Model zoo and per-variant configuration of the inference path (host side).

Mirrors ``perspective2d.perspectivefields.model_zoo`` (perspectivefields.py:86-118: same version names, checkpoint
URLs, ``param`` flags and descriptions) and the inference-relevant fields of the yacs defaults + the five yaml files
(config/config.py:4-78, config/*.yaml).  The yaml files themselves are not needed: every field the path reads is here.
"""

_HUB = "https://huggingface.co/spaces/jinlinyi/PerspectiveFields/resolve/main/models/"

model_zoo = {
    "Paramnet-360Cities-edina-centered": {
        "weights": _HUB + "paramnet_360cities_edina_rpf.pth",
        "config_file": "paramnet_360cities_edina_rpf.yaml",
        "param": True,
        "description": "Trained on 360cities and EDINA dataset. Assumes centered principal point. Predicts roll, pitch and fov.",
    },
    "Paramnet-360Cities-edina-uncentered": {
        "weights": _HUB + "paramnet_360cities_edina_rpfpp.pth",
        "config_file": "paramnet_360cities_edina_rpfpp.yaml",
        "param": True,
        "description": "Trained on 360cities and EDINA dataset. Predicts roll, pitch, fov and principal point.",
    },
    "PersNet-360Cities": {
        "weights": _HUB + "cvpr2023.pth",
        "config_file": "cvpr2023.yaml",
        "param": False,
        "description": "Trained on 360cities. Predicts perspective fields.",
    },
    "PersNet_Paramnet-GSV-uncentered": {
        "weights": _HUB + "paramnet_gsv_rpfpp.pth",
        "config_file": "paramnet_gsv_rpfpp.yaml",
        "param": True,
        "description": "Trained on GSV. Predicts roll, pitch, fov and principal point.",
    },
    "PersNet_Paramnet-GSV-centered": {
        "weights": _HUB + "paramnet_gsv_rpf.pth",
        "config_file": "paramnet_gsv_rpf.yaml",
        "param": True,
        "description": "Trained on GSV. Assumes centered principal point. Predicts roll, pitch and fov.",
    },
}

_REG = dict(gravity="regression", latitude="regression", gravity_classes=2, latitude_classes=1)
_CENTERED = dict(param_net="ParamNet", predict_params=("roll", "pitch", "vfov"), recover_rpf=True, recover_pp=False, input_size=64)
_UNCENTERED = dict(param_net="ParamNetConvNextRegress", predict_params=("roll", "pitch", "general_vfov", "rel_cx", "rel_cy"),
                   recover_rpf=True, recover_pp=True, input_size=64)

VARIANTS = {
    "Paramnet-360Cities-edina-centered": dict(_REG, **_CENTERED),
    "Paramnet-360Cities-edina-uncentered": dict(_REG, **_UNCENTERED),
    "PersNet-360Cities": dict(gravity="classification", latitude="classification", gravity_classes=73, latitude_classes=180,
                              param_net=None, predict_params=(), recover_rpf=False, recover_pp=False, input_size=320),
    "PersNet_Paramnet-GSV-uncentered": dict(_REG, **_UNCENTERED),
    "PersNet_Paramnet-GSV-centered": dict(_REG, **_CENTERED),
}

PIXEL_MEAN = (103.53, 116.28, 123.675)   # config.py:77 (B, G, R)
PIXEL_STD = (1.0, 1.0, 1.0)              # config.py:78
RESIZE = (320, 320)                      # DATALOADER.RESIZE in every yaml
INPUT_FORMAT = "BGR"                     # config.py:12, no yaml overrides it

MIT_DIMS = (64, 128, 320, 512)
MIT_HEADS = (1, 2, 5, 8)
MIT_DEPTHS = (3, 4, 18, 3)
MIT_SR = (8, 4, 2, 1)
CNX_DIMS = (96, 192, 384, 768)
CNX_DEPTHS = (3, 3, 9, 3)
HEAD_EMBED = 768


class CfgNode(dict):
    """Minimal attribute-dict stand-in for the yacs node callers read (``model.cfg.MODEL.RECOVER_PP`` ...)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def make_cfg(version):
    v = VARIANTS[version]
    model = CfgNode(
        GRAVITY_ON=True, LATITUDE_ON=True, RECOVER_RPF=v["recover_rpf"], RECOVER_PP=v["recover_pp"],
        BACKBONE=CfgNode(NAME="mitb3"), PERSFORMER_HEADS=CfgNode(NAME="StandardPersformerHeads"), WEIGHTS="",
        GRAVITY_DECODER=CfgNode(NAME="GravityDecoder", LOSS_TYPE=v["gravity"], NUM_CLASSES=73, IGNORE_VALUE=72, LOSS_WEIGHT=1.0),
        LATITUDE_DECODER=CfgNode(NAME="LatitudeDecoder", LOSS_TYPE=v["latitude"], NUM_CLASSES=v["latitude_classes"], IGNORE_VALUE=-1, LOSS_WEIGHT=1.0),
        PARAM_DECODER=CfgNode(NAME=v["param_net"] or "ParamNet", LOSS_TYPE="regression", PREDICT_PARAMS=list(v["predict_params"]),
                              INPUT_SIZE=v["input_size"]),
        PIXEL_MEAN=list(PIXEL_MEAN), PIXEL_STD=list(PIXEL_STD), FREEZE=[])
    return CfgNode(VIS_PERIOD=100, DEBUG_ON=False, INPUT=CfgNode(FORMAT=INPUT_FORMAT, ONLINE_CROP=False),
                   DATALOADER=CfgNode(RESIZE=list(RESIZE)), MODEL=model)
