from .heatmap import HeatmapHead  # noqa: F401
