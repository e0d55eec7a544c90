"""reference module path debvader.deblend_cutout.optimization -> debvader_b200.deblend_cutout.optimization"""
from debvader_b200.deblend_cutout.optimization import *  # noqa: F401,F403
from debvader_b200.deblend_cutout import optimization as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
