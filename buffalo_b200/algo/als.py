"""ALS trainer: the reference's Python epoch driver (buffalo/algo/als.py) on top of the B200 backend.

Two feeding modes, same results:
  * resident (default when the CSR fits in device memory): both CSR orientations and the factor matrices live
    on the GPU for the whole of train(); one launch set per half-epoch, no host traffic inside the loop;
  * chunked: the reference's own protocol -- BufferedDataMatrix chunks pushed through
    obj.partial_update(start_x, next_x, indptr, keys, vals, axis) (als.py:115-142).
The backend is GPU-only; `accelerator` is accepted and ignored (both values run the sm_100a kernels).
"""
import json
import time

import numpy as np

from buffalo_b200 import data as _data
from buffalo_b200.algo.base import Algo, Serializable
from buffalo_b200.algo.options import ALSOption
from buffalo_b200.backend import CuALS
from buffalo_b200.data.base import Data
from buffalo_b200.data.buffered_data import BufferedDataMatrix
from buffalo_b200.evaluate import Evaluable
from buffalo_b200.misc import aux, log

inited_CUALS = True


class ALS(Algo, ALSOption, Evaluable, Serializable):
    """Collaborative Filtering for Implicit Feedback datasets (Hu, Koren, Volinsky) -- drop-in for buffalo.algo.als.ALS."""

    def __init__(self, opt_path=None, *args, **kwargs):
        Algo.__init__(self, *args, **kwargs)
        ALSOption.__init__(self, *args, **kwargs)
        Evaluable.__init__(self, *args, **kwargs)
        Serializable.__init__(self, *args, **kwargs)
        if opt_path is None:
            opt_path = ALSOption().get_default_option()
        self.logger = log.get_logger("ALS")
        self.opt, self.opt_path = self.get_option(opt_path)
        self.obj = CuALS()
        assert self.obj.init(bytes(self.opt_path, "utf-8")), \
            "cannot parse option file: %s (%s)" % (opt_path, getattr(self.obj, "last_error", ""))
        self.data = None
        data = kwargs.get("data")
        data_opt = kwargs.get("data_opt", self.opt.get("data_opt"))
        if data_opt:
            self.data = _data.load(data_opt)
            self.data.create()
        elif isinstance(data, Data):
            self.data = data
        self.logger.info("ALS(%s)" % json.dumps(self.opt, indent=2))
        if self.data:
            self.logger.info(self.data.show_info())
            assert self.data.data_type in ["matrix"]

    @staticmethod
    def new(path, data_fields=[]):
        return ALS.instantiate(ALSOption, path, data_fields)

    def set_data(self, data):
        assert isinstance(data, Data), "Wrong instance: {}".format(type(data))
        self.data = data

    def normalize(self, group="item"):
        if group == "item" and not self.opt._nrz_Q:
            self.Q = self._normalize(self.Q)
            self.opt._nrz_Q = True
        elif group == "user" and not self.opt._nrz_P:
            self.P = self._normalize(self.P)
            self.opt._nrz_P = True

    def initialize(self):
        super().initialize()
        self.init_factors()

    def init_factors(self):
        assert self.data, "Data is not set"
        self.vdim = self.obj.get_vdim()
        header = self.data.get_header()
        for name, rows in (("P", header["num_users"]), ("Q", header["num_items"])):
            setattr(self, name, None)
            F = np.zeros((rows, self.vdim), dtype=np.float32)
            # abs(N(0, 1/d^2)) (als.py:85-86); drawn at width d so a seed gives the reference's values
            F[:, :self.opt.d] = np.abs(np.random.normal(scale=1.0 / (self.opt.d ** 2), size=(rows, self.opt.d)))
            setattr(self, name, F)
        self.obj.initialize_model(self.P, self.Q)

    # ---- queries (host) -----------------------------------------------------------------------
    def _get_topk_recommendation(self, rows, topk, pool=None):
        topks = super()._get_topk_recommendation(self.P[rows], self.Q, pb=None, Qb=None, pool=pool, topk=topk,
                                                 num_workers=self.opt.num_workers)
        return zip(rows, topks)

    def _get_most_similar_item(self, col, topk, pool):
        return super()._get_most_similar_item(col, topk, self.Q, self.opt._nrz_Q, pool)

    def get_scores(self, row_col_pairs):
        return {(r, c): self.P[r].dot(self.Q[c]) for r, c in row_col_pairs}

    def _get_scores(self, row, col):
        return (self.P[row] * self.Q[col]).sum(axis=1)

    def _get_feature(self, index, group="item"):
        return {"item": self.Q, "user": self.P}[group][index] if group in ("item", "user") else None

    # ---- training -----------------------------------------------------------------------------
    def _get_buffer(self):
        buf = BufferedDataMatrix()
        buf.initialize(self.data)
        return buf

    def _iterate(self, buf, group="rowwise"):
        """The reference protocol: precompute, then one partial_update per chunk (als.py:115-142)."""
        axis = 0 if group == "rowwise" else 1
        t0 = time.time()
        self.obj.precompute(axis)
        nume = deno = 0.0
        buf.set_group(group)
        updated = 0
        for sz in buf.fetch_batch():
            updated += sz
            start_x, next_x, indptr, keys, vals = buf.get()
            n_, d_ = self.obj.partial_update(start_x, next_x, indptr, keys, vals, axis)
            nume += n_
            deno += d_
        self.logger.debug(f"{group} updated: processed({updated}) elapsed({time.time() - t0:0.3f}s)")
        return nume, deno

    def _resident_capable(self):
        if self.opt.get("_b200_resident") is False:
            return False
        try:
            import torch
            free, _ = torch.cuda.mem_get_info()
        except Exception:
            return False
        h = self.data.get_header()
        need = 2 * h["num_nnz"] * 8 + (h["num_users"] + h["num_items"]) * (self.vdim * 4 + 8)
        return need * 1.3 < free

    def _train_resident(self, training_callback):
        import torch
        dev = torch.device("cuda", torch.cuda.current_device())
        h = self.data.get_header()
        U, I = h["num_users"], h["num_items"]
        tP, tQ = torch.from_numpy(self.P).to(dev), torch.from_numpy(self.Q).to(dev)
        self.obj.bind_factors(tP, tQ)
        for axis, G in enumerate(("rowwise", "colwise")):
            grp = self.data.get_group(G)
            n = int(grp["indptr"][-1]) if len(grp["indptr"]) else 0
            t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a, dtype=dt)).to(dev)  # noqa: E731
            self.obj.bind_csr(axis, t(grp["indptr"][:], np.int64), t(grp["key"][:max(n, 1)] if n else np.zeros(1), np.int32),
                              t(grp["val"][:max(n, 1)] if n else np.zeros(1), np.float32))
        loss = torch.zeros(2, dtype=torch.float64, device=dev)

        def sync_back():
            self.P[:], self.Q[:] = tP.cpu().numpy(), tQ.cpu().numpy()

        def one_iteration():
            loss.zero_()
            for axis, rows in ((0, U), (1, I)):
                self.obj.precompute_device(axis)
                self.obj.update_device(axis, 0, rows, loss)
            n_, d_ = loss.cpu().numpy()
            return float(n_), float(d_)
        try:
            return self._epoch_loop(one_iteration, sync_back, training_callback)
        finally:
            sync_back()
            self.obj.initialize_model(self.P, self.Q)   # leave the holder on host-pointer semantics

    def _train_chunked(self, training_callback):
        buf = self._get_buffer()
        lindptr, rindptr, batch_size = buf.get_indptrs()
        self.obj.set_placeholder(lindptr, rindptr, batch_size)

        def one_iteration():
            n1, d1 = self._iterate(buf, group="rowwise")
            n2, d2 = self._iterate(buf, group="colwise")
            return n1 + n2, d1 + d2
        return self._epoch_loop(one_iteration, lambda: None, training_callback)

    def _epoch_loop(self, one_iteration, sync_back, training_callback):
        best_loss, rmse, self.validation_result = float("inf"), None, {}
        t_all = time.time()
        for i in range(self.opt.num_iters):
            t0 = time.time()
            nume, deno = one_iteration()
            train_t = time.time() - t0
            rmse = (nume / (deno + self.opt.eps)) ** 0.5           # als.py:171
            metrics = {"train_loss": rmse}
            if self.opt.validation and self.opt.evaluation_on_learning and self.periodical(self.opt.evaluation_period, i):
                t0 = time.time()
                sync_back()
                self.validation_result = self.get_validation_results()
                vals = " ".join(f"{k}:{v:0.5f}" for k, v in self.validation_result.items())
                self.logger.info(f"Validation: {vals} Elapsed {time.time() - t0:0.3f} secs")
                metrics.update({"val_%s" % k: v for k, v in self.validation_result.items()})
                if callable(training_callback):
                    training_callback(i, metrics)
            self.logger.info("Iteration %d: RMSE %.3f Elapsed %.3f secs" % (i + 1, rmse, train_t))
            if self.opt.save_best:
                sync_back()
            best_loss = self.save_best_only(rmse, best_loss, i)
            if self.early_stopping(rmse):
                break
        self.logger.info(f"elapsed for full epochs: {time.time() - t_all:.2f} sec")
        return rmse

    def train(self, training_callback=None):
        if self.P.shape[1] != self.vdim:      # factors replaced by the user at width d: re-pad
            for name in ("P", "Q"):
                F = getattr(self, name)
                G = np.zeros((F.shape[0], self.vdim), dtype=np.float32)
                G[:, :self.opt.d] = F[:, :self.opt.d]
                setattr(self, name, G)
        self.obj.initialize_model(self.P, self.Q)
        rmse = self._train_resident(training_callback) if self._resident_capable() else self._train_chunked(training_callback)
        if self.opt.d < self.vdim:            # als.py:191-193
            self.P = np.ascontiguousarray(self.P[:, :self.opt.d])
            self.Q = np.ascontiguousarray(self.Q[:, :self.opt.d])
        ret = {"train_loss": rmse}
        ret.update({"val_%s" % k: v for k, v in self.validation_result.items()})
        return ret

    def _get_data(self):
        return super()._get_data() + [("opt", self.opt), ("Q", self.Q), ("P", self.P)]

    def get_evaluation_metrics(self):
        return ["train_loss", "val_rmse", "val_ndcg", "val_map", "val_accuracy", "val_error"]
