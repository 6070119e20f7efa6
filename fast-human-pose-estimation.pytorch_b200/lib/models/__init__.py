# same registry the reference builds in lib/models/__init__.py:15-17 (looked up by cfg.MODEL.NAME)
from . import pose_resnet  # noqa: F401
from . import pose_hrnet  # noqa: F401
from . import hourglass  # noqa: F401
