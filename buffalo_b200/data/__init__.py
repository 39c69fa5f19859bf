from buffalo_b200.misc import aux

from .base import Data
from .mm import MatrixMarket, MatrixMarketDataReader, MatrixMarketOptions
from .stream import Stream, StreamOptions


def load(opt):
    """buffalo/data/__init__.py:7-18"""
    if isinstance(opt, str):
        opt = aux.Option(opt)
    assert isinstance(opt, (dict, aux.Option)), "opt must be either str, or dict/aux.Option but {}".format(type(opt))
    if opt["type"] == "matrix_market":
        return MatrixMarket(opt)
    if opt["type"] == "stream":
        return Stream(opt)
    raise RuntimeError("Unexpected data.type: {}".format(opt["type"]))
