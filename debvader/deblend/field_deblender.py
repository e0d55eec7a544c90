"""reference module path debvader.deblend.field_deblender -> debvader_b200.deblend.field_deblender"""
from debvader_b200.deblend.field_deblender import *  # noqa: F401,F403
from debvader_b200.deblend import field_deblender as _impl

__all__ = [n for n in dir(_impl) if not n.startswith("_")]
