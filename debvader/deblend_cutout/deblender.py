"""reference module path debvader.deblend_cutout.deblender -> debvader_b200.deblend_cutout.deblender"""
from debvader_b200.deblend_cutout.deblender import *  # noqa: F401,F403
from debvader_b200.deblend_cutout import deblender as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
