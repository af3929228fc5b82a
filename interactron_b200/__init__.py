"""interactron_b200 — B200-native (sm_100a) implementation of Interactron's test-time-adaptation
hot path (adapt on a 5-frame episode, re-detect), behind the reference's model interface."""
from .config import Config, default_config, get_config  # noqa: F401
from .models import build_model, detr, detr_multiframe, interactron, interactron_random  # noqa: F401

__version__ = "0.1"
