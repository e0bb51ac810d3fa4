"""`blt` -- the module name the reference's Python package exports (blt_python/python/blt/__init__.py:12-16:
ByteTokenizer, load_bpe_merges, version, __version__), served by the B200 implementation in `blt_b200`.
`import blt` in a program written for the reference picks this package up when the repository root is on sys.path."""
from blt_b200 import ByteTokenizer, load_bpe_merges, version

__all__ = ["ByteTokenizer", "load_bpe_merges", "version", "__version__"]


def __getattr__(name):
    if name == "__version__":
        return version()
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")
