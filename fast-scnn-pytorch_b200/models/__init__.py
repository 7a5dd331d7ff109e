"""Drop-in for the reference ``models`` package (models/__init__.py:1-3)."""
from .fast_scnn import FastSCNN, get_fast_scnn

__all__ = ['FastSCNN', 'get_fast_scnn']
