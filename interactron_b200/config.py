"""Configuration surface: the reference's YAML files load unchanged.

`get_config(path)` parses a reference-style YAML into nested attribute objects with the same
scalar coercion as the reference (reference utils/config_utils.py:9-40: every scalar that
`float()` accepts becomes a number, integral values become `int`, so "1e-3" -> 0.001 and
True -> 1).  `default_config(name)` rebuilds the MODEL section of the four shipped reference
configs (configs/*.yaml) so the package can run where those files are not present; the keys the
hot path reads are NUM_CLASSES, SET_COST_{CLASS,BBOX,GIOU}, NUM_LAYERS, NUM_HEADS, EMBEDDING_DIM,
BLOCK_SIZE, IMG_FEATURE_SIZE, OUTPUT_SIZE, BOX_EMB_SIZE, *_PDROP, ADAPTIVE_LR, WEIGHTS.
"""
import os


def _coerce(value):
    try:
        f = float(value)
    except (TypeError, ValueError):
        return value
    return int(f) if f.is_integer() else f


class Config:
    def __init__(self, **entries):
        for key, value in entries.items():
            setattr(self, key, Config(**value) if isinstance(value, dict) else _coerce(value))

    def dictionarize(self):
        return {k: v.dictionarize() if isinstance(v, Config) else v for k, v in self.__dict__.items()}

    def __repr__(self):
        return f"Config({self.dictionarize()})"


def get_config(path):
    import yaml
    assert os.path.exists(path), "File {} does not exist".format(path)
    with open(path) as f:
        return Config(**yaml.safe_load(f))


_COMMON = dict(NUM_CLASSES=1235, BACKBONE="resnet50", SET_COST_CLASS=1.0, SET_COST_BBOX=5.0,
               SET_COST_GIOU=2.0, TEST_RESOLUTION=300)
_FUSION = dict(NUM_LAYERS=4, NUM_HEADS=8, EMBEDDING_DIM=512, BLOCK_SIZE=2060, IMG_FEATURE_SIZE=256,
               OUTPUT_SIZE=512, BOX_EMB_SIZE=256, EMBEDDING_PDROP=0.1, RESIDUAL_PDROP=0.1,
               ATTENTION_PDROP=0.1)
_MODELS = {
    "single_frame_baseline": dict(TYPE="detr", WEIGHTS="pretrained_weights/detr-dc5.pth",
                                  FROZEN_WEIGHTS="pretrained_weights/detr-dc5.pth", **_COMMON),
    "multi_frame_baseline": dict(TYPE="detr_multiframe", WEIGHTS="pretrained_weights/detr-dc5-backbone.pth",
                                 PREDICT_ACTIONS=False, **_COMMON, **_FUSION),
    "interactron_random": dict(TYPE="interactron_random", WEIGHTS="pretrained_weights/detr-dc5-backbone.pth",
                               PREDICT_ACTIONS=False, ADAPTIVE_LR="1e-3", **_COMMON, **_FUSION),
    "interactron": dict(TYPE="interactron", WEIGHTS="pretrained_weights/detr-dc5-backbone.pth",
                        PREDICT_ACTIONS=True, ADAPTIVE_LR="1e-3", **_COMMON, **_FUSION),
}


def default_config(name, weights=None):
    """MODEL section of reference configs/<name>.yaml; `weights` overrides WEIGHTS
    (e.g. "synthetic" for the seeded random-init weights used by tests and bench)."""
    if name not in _MODELS:
        raise KeyError(f"unknown config {name!r}; choose from {sorted(_MODELS)}")
    d = dict(_MODELS[name])
    if weights is not None:
        d["WEIGHTS"] = weights
    return Config(MODEL=d)
