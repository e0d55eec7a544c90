"""reference module path debvader.normalize.normalize -> debvader_b200.normalize.normalize"""
from debvader_b200.normalize.normalize import *  # noqa: F401,F403
from debvader_b200.normalize import normalize as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
