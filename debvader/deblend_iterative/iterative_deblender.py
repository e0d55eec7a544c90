"""reference module path debvader.deblend_iterative.iterative_deblender -> debvader_b200.deblend_iterative.iterative_deblender"""
from debvader_b200.deblend_iterative.iterative_deblender import *  # noqa: F401,F403
from debvader_b200.deblend_iterative import iterative_deblender as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
