"""Only the file helper JoXSZ calls (``joxsz_main.py:113``); XSPEC itself is out of scope."""
import os


def deleteFile(filename):
    try:
        os.unlink(filename)
    except OSError:
        pass


class XSpecHelper:
    def __init__(self, *a, **k):
        raise RuntimeError("XSPEC is not available; supply count-rate tables to CountRate.setTables()")
