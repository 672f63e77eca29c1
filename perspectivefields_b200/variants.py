"""Model zoo and per-variant configuration of the inference path (host side).

Mirrors ``perspective2d.perspectivefields.model_zoo`` (perspectivefields.py:86-118: same version names, checkpoint
URLs, ``param`` flags and descriptions) and the inference-relevant fields of the yacs defaults + the five yaml files
(config/config.py:4-78, config/*.yaml), which live as data in ``config/defaults.yaml`` / ``config/variants.yaml`` and are
parsed with PyYAML at import time.
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

class CfgNode(dict):
    """Minimal attribute-dict stand-in for the yacs node callers read (``model.cfg.MODEL.RECOVER_PP`` ...)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def _merge(base, over):
    """yacs ``merge_from_file`` semantics for the keys present here: nested dicts merge, leaves are replaced."""
    out = CfgNode()
    for k, v in base.items():
        out[k] = _merge(v, {}) if isinstance(v, dict) else (list(v) if isinstance(v, list) else v)
    for k, v in (over or {}).items():
        out[k] = _merge(out.get(k, {}), v) if isinstance(v, dict) else (list(v) if isinstance(v, list) else v)
    return out


def _load_configs():
    """The configuration is DATA: config/defaults.yaml + the per-variant overrides of config/variants.yaml (PyYAML), the
    counterpart of ``default_conf.merge_from_file(config_path)`` (perspectivefields.py:124-131)."""
    import os

    import yaml

    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "config")
    with open(os.path.join(here, "defaults.yaml")) as f:
        defaults = yaml.safe_load(f)
    with open(os.path.join(here, "variants.yaml")) as f:
        overrides = yaml.safe_load(f)
    return {name: _merge(defaults, overrides[z["config_file"]]) for name, z in model_zoo.items()}


_CFGS = _load_configs()


def make_cfg(version):
    """Frozen-config stand-in for ``PerspectiveFields.cfg`` (a fresh copy per model)."""
    return _merge(_CFGS[version], {})


def _variant(cfg):
    """The fields the engine needs, derived the way the reference's builders read the config: a regression gravity head has
    2 outputs whatever NUM_CLASSES says (gravity_head.py:58-63), a regression latitude head 1 (latitude_head.py:49-53); the
    ParamNet exists iff RECOVER_RPF or RECOVER_PP (perspectivefields.py:140-144) and its class is PARAM_DECODER.NAME."""
    m = cfg.MODEL
    g, l = m.GRAVITY_DECODER, m.LATITUDE_DECODER
    has_pn = bool(m.RECOVER_RPF or m.RECOVER_PP)
    return dict(gravity=g.LOSS_TYPE, latitude=l.LOSS_TYPE,
                gravity_classes=2 if g.LOSS_TYPE == "regression" else int(g.NUM_CLASSES),
                latitude_classes=1 if l.LOSS_TYPE == "regression" else int(l.NUM_CLASSES),
                param_net=m.PARAM_DECODER.NAME if has_pn else None,
                predict_params=tuple(m.PARAM_DECODER.PREDICT_PARAMS) if has_pn else (),
                recover_rpf=bool(m.RECOVER_RPF), recover_pp=bool(m.RECOVER_PP), input_size=int(m.PARAM_DECODER.INPUT_SIZE))


VARIANTS = {name: _variant(cfg) for name, cfg in _CFGS.items()}

_ANY = next(iter(_CFGS.values()))
PIXEL_MEAN = tuple(_ANY.MODEL.PIXEL_MEAN)   # (B, G, R); no variant overrides it
PIXEL_STD = tuple(_ANY.MODEL.PIXEL_STD)
RESIZE = tuple(_ANY.DATALOADER.RESIZE)      # every variant: 320 x 320
INPUT_FORMAT = _ANY.INPUT.FORMAT            # "BGR"

MIT_DIMS = (64, 128, 320, 512)
MIT_HEADS = (1, 2, 5, 8)
MIT_DEPTHS = (3, 4, 18, 3)
MIT_SR = (8, 4, 2, 1)
CNX_DIMS = (96, 192, 384, 768)
CNX_DEPTHS = (3, 3, 9, 3)
HEAD_EMBED = 768
