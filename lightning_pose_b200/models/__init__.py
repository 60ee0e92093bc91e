"""Model-side mirror of ``lightning_pose.models`` (heads + the tracker glue on the hot path)."""
