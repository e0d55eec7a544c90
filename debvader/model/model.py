"""reference module path debvader.model.model -> debvader_b200.model.model"""
from debvader_b200.model.model import *  # noqa: F401,F403
from debvader_b200.model import model as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
