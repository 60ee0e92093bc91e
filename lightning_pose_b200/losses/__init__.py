"""Loss stack mirror of ``lightning_pose.losses``."""
from lightning_pose_b200.losses.factory import get_loss_classes, get_loss_factories  # noqa: F401

__all__ = ["get_loss_classes", "get_loss_factories"]
