"""Drop-in for the hot-path part of the reference's ``tf_extended`` package (same re-exports
as reference tf_extended/__init__.py:19-23, minus image.py which is augmentation-side)."""
from .metrics import *  # noqa: F401,F403
from .tensors import *  # noqa: F401,F403
from .bboxes import *  # noqa: F401,F403
from .math import *  # noqa: F401,F403
