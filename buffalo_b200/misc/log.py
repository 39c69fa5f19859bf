"""Logging facade with buffalo's level constants (buffalo/misc/log.py:7-66); Python logging only --
there is no native logger to keep in sync (the reference shares the level with spdlog, log.cc:3-27)."""
import logging
import time

NOTSET, WARN, INFO, DEBUG, TRACE = 0, 1, 2, 3, 4
_PY = {NOTSET: logging.NOTSET, WARN: logging.WARNING, INFO: logging.INFO, DEBUG: logging.DEBUG, TRACE: logging.DEBUG}
_state = {"level": INFO, "loggers": []}


def set_log_level(lvl):
    _state["level"] = lvl
    for lg in _state["loggers"]:
        lg.setLevel(_PY.get(lvl, logging.DEBUG))


def get_log_level():
    return _state["level"]


def get_logger(name=__file__, no_fileno=False):
    lg = logging.getLogger(name)
    if lg.handlers:
        return lg
    lg.setLevel(_PY.get(_state["level"], logging.DEBUG))
    h = logging.StreamHandler()
    if name == "pbar":
        fmt = logging.Formatter("%(message)s")
    elif no_fileno:
        fmt = logging.Formatter("[%(levelname)-8s] %(asctime)s %(message)s", "%Y-%m-%d %H:%M:%S")
    else:
        fmt = logging.Formatter("[%(levelname)-8s] %(asctime)s [%(filename)s:%(lineno)d] %(message)s", "%Y-%m-%d %H:%M:%S")
    h.setFormatter(fmt)
    lg.addHandler(h)
    lg.propagate = False
    _state["loggers"].append(lg)
    return lg


class ProgressBar(object):
    """Minimal stand-in for the reference's tqdm-backed bar: usable as a context manager with
    update()/refresh(), or as an iterable wrapper (``ProgressBar(level, iterable=...)``)."""

    def __init__(self, level, iterable=None, total=None, desc="", mininterval=30, **kwargs):
        self.enabled = level <= get_log_level()
        self.iterable, self.total, self.desc, self.mininterval = iterable, total, desc, mininterval
        self.n, self.t0, self.last = 0, time.time(), time.time()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def update(self, n=1):
        self.n += n
        now = time.time()
        if self.enabled and now - self.last >= self.mininterval:
            self.last = now
            get_logger("pbar").info("%s %s/%s (%.1fs)" % (self.desc, self.n, self.total, now - self.t0))

    def refresh(self):
        pass

    def __iter__(self):
        for x in self.iterable:
            yield x
            self.update(1)
