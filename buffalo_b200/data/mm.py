"""MatrixMarket ingest (buffalo/data/mm.py): text file, scipy sparse or dense 2-D array -> database."""
import os

import numpy as np
import scipy.sparse

from buffalo_b200.data.base import Data, DataOption, DataReader
from buffalo_b200.data import prepro
from buffalo_b200.misc import aux, log


class MatrixMarketOptions(DataOption):
    def get_default_option(self):
        return aux.Option({
            "type": "matrix_market",
            "input": {"main": "", "uid": "", "iid": ""},
            "data": {"internal_data_type": "matrix",
                     "validation": {"name": "sample", "p": 0.01, "max_samples": 500},
                     "batch_mb": 1024, "use_cache": False, "tmp_dir": "/tmp/", "path": "./mm.h5py",
                     "disk_based": False}})                       # mm.py:14-37

    def is_valid_option(self, opt):
        assert super().is_valid_option(opt)
        if not opt["type"] == "matrix_market":
            raise RuntimeError("Invalid data type: %s" % opt["type"])
        if opt["data"]["internal_data_type"] != "matrix":
            raise RuntimeError("MatrixMarket only support internal data type(matrix)")
        for field in ["uid", "iid"]:
            v = opt["input"][field]
            ok = v is None or isinstance(v, (str, list)) or (isinstance(v, np.ndarray) and v.ndim == 1)
            assert ok, f"Not supported data type for MatrixMarketOption.input.{field}: {type(v)}"
        main = opt["input"]["main"]
        ok = isinstance(main, str) or (isinstance(main, np.ndarray) and main.ndim == 2) or scipy.sparse.issparse(main)
        assert ok, f"Not supported data type for MatrixMarketOption.input.main field: {type(main)}"
        return True


class MatrixMarketDataReader(DataReader):
    pass


def _read_ids(spec):
    if spec is None or (isinstance(spec, str) and spec == ""):
        return None
    if isinstance(spec, str):
        with open(spec) as fin:
            return [line.rstrip("\n").strip() for line in fin if line.strip() != ""]
    return [str(x) for x in (spec.tolist() if isinstance(spec, np.ndarray) else spec)]


def _read_mm_text(path):
    """-> (num_rows, num_cols, rows0, cols0, vals) from a coordinate MatrixMarket file (1-based text)."""
    import pandas as pd
    skip = 0
    with open(path) as fin:
        for line in fin:
            skip += 1
            if not line.strip().startswith("%"):
                header = line
                break
    U, I, nnz = map(int, header.split())
    if nnz == 0:
        return U, I, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0, np.float32)
    df = pd.read_csv(path, sep=r"\s+", header=None, skiprows=skip, comment="%", dtype=np.float64, engine="c")
    rows = df[0].to_numpy().astype(np.int64) - 1
    cols = df[1].to_numpy().astype(np.int64) - 1
    vals = df[2].to_numpy().astype(np.float32) if df.shape[1] > 2 else np.ones(len(rows), np.float32)
    return U, I, rows, cols, vals


class MatrixMarket(Data):
    def __init__(self, opt, *args, **kwargs):
        super().__init__(opt, *args, **kwargs)
        self.name = "MatrixMarket"
        self.logger = log.get_logger("MatrixMarket")
        if isinstance(self.value_prepro, prepro.SPPMI):
            raise RuntimeError(f"{self.opt.data.value_prepro.name} does not support MatrixMarket")
        self.data_type = "matrix"
        self.reader = MatrixMarketDataReader(self.opt)

    def _load_triples(self):
        main = self.opt.input.main
        if isinstance(main, str):
            return _read_mm_text(main)
        if isinstance(main, np.ndarray) and main.ndim == 2:
            main = scipy.sparse.csr_matrix(main)
        if scipy.sparse.issparse(main):
            coo = main.tocoo()
            return coo.shape[0], coo.shape[1], coo.row.astype(np.int64), coo.col.astype(np.int64), coo.data.astype(np.float32)
        raise RuntimeError(f"Unexpected data type for MatrixMarketOption.input.main field: {type(main)}")

    def create(self):
        path = self.opt.data.path
        if os.path.isfile(path) and self.opt.data.use_cache:
            self.logger.info("Use cached DB on %s" % path)
            self.open(path)
            return
        self.logger.info("Create the database from matrix market file.")
        U, I, rows, cols, vals = self._load_triples()
        nnz = len(rows)
        vali = None
        vopt = self.opt.data.validation
        if vopt:
            if vopt.name != "sample":
                raise RuntimeError("MatrixMarket supports validation.name == 'sample' only")
            # base.py:225-231: sample line indexes, never the last line
            sz = min(vopt.max_samples, int(nnz * vopt.p))
            idx = np.sort(np.random.choice(nnz - 1, sz, replace=False)) if sz > 0 else np.zeros(0, np.int64)
            mask = np.ones(nnz, bool)
            mask[idx] = False
            vali = dict(method="sample", n=0, indexes=idx, row=rows[idx], col=cols[idx], val=vals[idx])
            rows, cols, vals = rows[mask], cols[mask], vals[mask]
        self._write_database(path, U, I, rows, cols, vals, _read_ids(self.opt.input.uid), _read_ids(self.opt.input.iid), vali)
        self.logger.info("DB built on %s" % path)
