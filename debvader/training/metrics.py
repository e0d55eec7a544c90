"""reference module path debvader.training.metrics -> debvader_b200.training.metrics"""
from debvader_b200.training.metrics import *  # noqa: F401,F403
from debvader_b200.training import metrics as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
