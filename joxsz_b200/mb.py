"""``mb``: the real ``mbproj2`` when it is importable, else the bundled work-alike."""
try:  # pragma: no cover - mbproj2 is not installable in the build environment
    import mbproj2 as mb
    USING_SHIM = False
except ImportError:
    from . import mbshim as mb
    USING_SHIM = True

__all__ = ["mb", "USING_SHIM"]
