"""Drop-ins for the reference ``utils`` package around the hot path: metric.py (SegmentationMetric), visualize.py (palettes,
overlay), loss.py (the training criteria) and lr_scheduler.py."""
