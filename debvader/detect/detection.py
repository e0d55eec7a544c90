"""reference module path debvader.detect.detection -> debvader_b200.detect.detection"""
from debvader_b200.detect.detection import *  # noqa: F401,F403
from debvader_b200.detect import detection as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
