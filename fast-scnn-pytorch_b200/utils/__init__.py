"""Drop-in for the hot-path part of the reference ``utils`` package (utils/metric.py)."""
