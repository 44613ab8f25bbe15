"""Package logger (same name and level as the reference's ``lycoris.logging`` so trainer scripts
that grab ``logging.getLogger("LyCORIS")`` keep working; reference: lycoris/logging.py:25-52)."""

import logging
import sys
from functools import lru_cache

_LEVEL_COLOURS = {
    logging.DEBUG: "36",
    logging.INFO: "32",
    logging.WARNING: "33",
    logging.ERROR: "31",
    logging.CRITICAL: "37;41",
}


class _Tinted(logging.Formatter):
    def format(self, record):
        tint = _LEVEL_COLOURS.get(record.levelno)
        if tint is None:
            return super().format(record)
        clone = logging.makeLogRecord(record.__dict__)
        clone.levelname = f"\033[0;{tint}m{record.levelname}\033[0m"
        return super().format(clone)


logger = logging.getLogger("LyCORIS")
logger.propagate = False
logger.setLevel(logging.INFO)
if not logger.handlers:
    _h = logging.StreamHandler(sys.stdout)
    _h.setFormatter(_Tinted("%(asctime)s|[%(name)s]-%(levelname)s: %(message)s", "%Y-%m-%d %H:%M:%S"))
    logger.addHandler(_h)


@lru_cache(maxsize=None)
def info_once(msg):
    logger.info(msg)


@lru_cache(maxsize=None)
def warning_once(msg):
    logger.warning(msg)


@lru_cache(maxsize=None)
def error_once(msg):
    logger.error(msg)
