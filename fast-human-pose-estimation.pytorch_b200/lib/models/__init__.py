# same registry the reference builds in lib/models/__init__.py:15-17 (looked up by cfg.MODEL.NAME); pose_resnet is not
# part of the hot path and stays the reference's own module (fpd_b200.dropin.install() leaves it in place)
from . import hourglass  # noqa: F401
from . import pose_hrnet  # noqa: F401
