"""The five model variants of the reference zoo, restated as data.  TEST INFRASTRUCTURE (oracle).

Follows perspective2d/perspectivefields.py:86-118 (zoo), config/config.py:4-78 (defaults) and the
five yaml files in perspective2d/config/ (per-variant overrides).  Only fields read on the inference
path are kept.
"""

PIXEL_MEAN = (103.53, 116.28, 123.675)  # config.py:77, B,G,R order
PIXEL_STD = (1.0, 1.0, 1.0)  # config.py:78
NET_H = NET_W = 320  # DATALOADER.RESIZE in every yaml

_REG = dict(gravity="regression", latitude="regression", gravity_classes=2, latitude_classes=1)

VARIANTS = {
    # yaml: paramnet_360cities_edina_rpf.yaml
    "Paramnet-360Cities-edina-centered": dict(
        ckpt="paramnet_360cities_edina_rpf.pth", **_REG,
        param_net="ParamNet", predict_params=("roll", "pitch", "vfov"), recover_pp=False),
    # yaml: paramnet_360cities_edina_rpfpp.yaml
    "Paramnet-360Cities-edina-uncentered": dict(
        ckpt="paramnet_360cities_edina_rpfpp.pth", **_REG,
        param_net="ParamNetConvNextRegress",
        predict_params=("roll", "pitch", "general_vfov", "rel_cx", "rel_cy"), recover_pp=True, input_size=64),
    # yaml: cvpr2023.yaml
    "PersNet-360Cities": dict(
        ckpt="cvpr2023.pth", gravity="classification", latitude="classification",
        gravity_classes=73, latitude_classes=180, param_net=None),
    # yaml: paramnet_gsv_rpfpp.yaml
    "PersNet_Paramnet-GSV-uncentered": dict(
        ckpt="paramnet_gsv_rpfpp.pth", **_REG,
        param_net="ParamNetConvNextRegress",
        predict_params=("roll", "pitch", "general_vfov", "rel_cx", "rel_cy"), recover_pp=True, input_size=64),
    # yaml: paramnet_gsv_rpf.yaml
    "PersNet_Paramnet-GSV-centered": dict(
        ckpt="paramnet_gsv_rpf.pth", **_REG,
        param_net="ParamNet", predict_params=("roll", "pitch", "vfov"), recover_pp=False),
}

# mit_b3 hyper-parameters, mix_transformers.py:511-524
MIT_DIMS = (64, 128, 320, 512)
MIT_HEADS = (1, 2, 5, 8)
MIT_DEPTHS = (3, 4, 18, 3)
MIT_SR = (8, 4, 2, 1)
# ConvNeXt defaults, convnext.py:81-82
CNX_DIMS = (96, 192, 384, 768)
CNX_DEPTHS = (3, 3, 9, 3)
HEAD_EMBED = 768  # gravity_head.py:132
