"""Drop-in shim: the reference's module paths (``debvader.extract.extraction`` ...) backed by
``debvader_b200``.  A user of astrodeepnet/debvader keeps their imports unchanged.
Mirrors src/debvader/__init__.py:1-2, but lazily (no TensorFlow / sep import at package import)."""


def __getattr__(name):
    import debvader_b200

    return getattr(debvader_b200, name)
