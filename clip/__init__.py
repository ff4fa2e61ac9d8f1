"""Drop-in replacement for the `clip` package (openai/CLIP) that the reference imports
(`import clip` at CLIP/predict.py:2, CLIP/train.py:8, CLIP_prefix_caption/parse_coco.py:3,
application.py:8): same `load / tokenize / available_models` surface, B200-native arithmetic."""
from .clip import *  # noqa: F401,F403
from .clip import _transform  # noqa: F401
