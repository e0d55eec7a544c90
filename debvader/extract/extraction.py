"""reference module path debvader.extract.extraction -> debvader_b200.extract.extraction"""
from debvader_b200.extract.extraction import *  # noqa: F401,F403
from debvader_b200.extract import extraction as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
