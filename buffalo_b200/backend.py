"""Backend holder objects: the `self.obj` of the algo drivers.

``CuALS`` / ``CuSGD`` expose exactly the method set of the reference's Cython holders
(buffalo/algo/_als.pyx:28-63, buffalo/algo/cuda/_als.pyx:25-67, buffalo/algo/_bpr.pyx:34-92,
buffalo/algo/cuda/_bpr.pyx:27-80, buffalo/algo/_warp.pyx:34-92) on top of the C ABI, plus a
device-resident path that takes torch CUDA tensors (PyTorch is used for device memory and
streams only).
"""
import ctypes as C
import json
import os

import numpy as np

from buffalo_b200 import _cabi


def _host(a, dtype, ndim, name):
    # the Cython signatures type-check dtype/ndim and assume C contiguity (_als.pyx:42-63)
    if not isinstance(a, np.ndarray) or a.dtype != dtype or a.ndim != ndim:
        raise ValueError("Buffer dtype/ndim mismatch for %s: expected %s ndim=%d" % (name, np.dtype(dtype), ndim))
    if not a.flags["C_CONTIGUOUS"]:
        raise ValueError("%s must be C-contiguous" % name)
    return a.ctypes.data


def _dev(t, dtype_name, name):
    import torch
    if not isinstance(t, torch.Tensor) or not t.is_cuda or not t.is_contiguous() or str(t.dtype) != "torch." + dtype_name:
        raise ValueError("%s must be a contiguous CUDA tensor of dtype %s" % (name, dtype_name))
    return t.data_ptr()


def _stream_ptr(stream):
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return s.cuda_stream


def _opt_bytes(opt):
    if isinstance(opt, (bytes, bytearray)):
        return bytes(opt), True
    if isinstance(opt, str):
        return opt.encode("utf-8"), True
    return json.dumps(dict(opt)).encode("utf-8"), False


class CuALS(object):
    """ALS backend (CyALS, _als.pyx:28-63; CUDA holder cuda/_als.pyx:25-67)."""

    def __init__(self):
        self._lib = _cabi.lib()
        self._h = self._lib.bfl_als_create()
        if not self._h:
            raise MemoryError("bfl_als_create")
        self._keep = []

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.bfl_als_destroy(h)

    # --- reference method set -------------------------------------------------------------
    def init(self, opt_path):
        """opt_path: bytes/str path of the JSON option file (als.py:43), or a dict."""
        data, is_path = _opt_bytes(opt_path)
        rc = self._lib.bfl_als_init(self._h, data) if is_path else self._lib.bfl_als_init_json(self._h, data)
        if rc == 1:  # BFL_ERR_OPTION: the reference's init returns False (algo.cc:22-34)
            self.last_error = self._lib.bfl_last_error().decode("utf-8", "replace")
            return False
        _cabi.check(rc, "bfl_als_init")
        if is_path:
            with open(data.decode("utf-8")) as fin:
                self._d = int(json.load(fin)["d"])
        else:
            self._d = int(json.loads(data.decode("utf-8"))["d"])
        return True

    def get_vdim(self):
        return self._lib.bfl_als_get_vdim(self._h)

    def initialize_model(self, P, Q):
        pP, pQ = _host(P, np.float32, 2, "P"), _host(Q, np.float32, 2, "Q")
        vdim = self.get_vdim()
        if P.shape[1] != vdim or Q.shape[1] != vdim:
            raise ValueError("factor matrices must have %d columns (get_vdim())" % vdim)
        self._keep = [P, Q]  # native side retains the pointers (als.cc:78-79)
        _cabi.check(self._lib.bfl_als_initialize_model(self._h, pP, P.shape[0], pQ, Q.shape[0]), "initialize_model")

    def set_placeholder(self, lindptr, rindptr, batch_size):
        _cabi.check(self._lib.bfl_als_set_placeholder(self._h, _host(lindptr, np.int64, 1, "lindptr"),
                                                      _host(rindptr, np.int64, 1, "rindptr"), int(batch_size)),
                    "set_placeholder")

    def precompute(self, axis):
        _cabi.check(self._lib.bfl_als_precompute(self._h, int(axis)), "precompute")

    def partial_update(self, start_x, next_x, indptr, keys, vals, axis):
        nume, deno = C.c_double(0.0), C.c_double(0.0)
        _cabi.check(self._lib.bfl_als_partial_update(self._h, int(start_x), int(next_x),
                                                     _host(indptr, np.int64, 1, "indptr"),
                                                     _host(keys, np.int32, 1, "keys"),
                                                     _host(vals, np.float32, 1, "vals"), int(axis),
                                                     C.byref(nume), C.byref(deno)), "partial_update")
        return nume.value, deno.value

    # --- device-resident path ---------------------------------------------------------------
    def bind_factors(self, P, Q):
        """P, Q: torch float32 CUDA tensors [rows, vdim], updated in place."""
        vdim = self.get_vdim()
        assert P.shape[1] == vdim and Q.shape[1] == vdim, "factor tensors need vdim=%d columns" % vdim
        self._keep = [P, Q]
        _cabi.check(self._lib.bfl_als_bind_factors_device(self._h, _dev(P, "float32", "P"), P.shape[0],
                                                          _dev(Q, "float32", "Q"), Q.shape[0]), "bind_factors")

    def bind_csr(self, axis, indptr, keys, vals):
        """indptr int64[rows] END offsets, keys int32[nnz], vals float32[nnz]: torch CUDA tensors."""
        self._keep += [indptr, keys, vals]
        _cabi.check(self._lib.bfl_als_bind_csr_device(self._h, int(axis), _dev(indptr, "int64", "indptr"),
                                                      _dev(keys, "int32", "keys"), _dev(vals, "float32", "vals"),
                                                      indptr.shape[0], keys.shape[0]), "bind_csr")

    def precompute_device(self, axis, stream=None):
        _cabi.check(self._lib.bfl_als_precompute_device(self._h, int(axis), _stream_ptr(stream)), "precompute_device")

    def precompute_rows_device(self, axis, row_begin, row_end, stream=None):
        """Partial Gram over rows [row_begin,row_end) of the opposite factor (to be all-reduced via gram_tensor())."""
        _cabi.check(self._lib.bfl_als_precompute_rows_device(self._h, int(axis), int(row_begin), int(row_end),
                                                             _stream_ptr(stream)), "precompute_rows_device")

    def update_device(self, axis, row_begin, row_end, loss=None, stream=None):
        """loss: optional torch float64 CUDA tensor[2] receiving (+=) numerator, denominator."""
        lp = _dev(loss, "float64", "loss") if loss is not None else None
        _cabi.check(self._lib.bfl_als_update_device(self._h, int(axis), int(row_begin), int(row_end), lp,
                                                    _stream_ptr(stream)), "update_device")

    def set_peer_replicas(self, axis, pointers):
        """pointers: device addresses (ints), valid in this process, of the other ranks' replicas of the matrix
        updated on `axis` (buffalo_b200.parallel.dist.open_peer_replicas); [] switches the fused exchange off."""
        arr = (C.c_void_p * max(len(pointers), 1))(*[int(p) for p in pointers])
        _cabi.check(self._lib.bfl_als_set_peer_replicas(self._h, int(axis), len(pointers), arr), "set_peer_replicas")

    def gram_tensor(self):
        """View of the current d x d Gram matrix as a torch tensor (multi-GPU all-reduce, tests)."""
        import torch
        ptr = self._lib.bfl_als_gram_device_mut(self._h)
        n = self._d * self._d

        class _Arr(object):
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Arr(), device="cuda").view(self._d, self._d)


class CuSGD(object):
    """BPRMF / WARP backend (CyBPRMF _bpr.pyx:34-92, CyWARP _warp.pyx:34-92, CyBPR cuda/_bpr.pyx:27-80)."""

    KIND = {"bpr": 0, "warp": 1}

    def __init__(self, kind):
        self._lib = _cabi.lib()
        self.kind = kind
        self._h = self._lib.bfl_sgd_create(self.KIND[kind])
        if not self._h:
            raise MemoryError("bfl_sgd_create")
        self._keep = []

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.bfl_sgd_destroy(h)

    def init(self, opt_path):
        data, is_path = _opt_bytes(opt_path)
        rc = self._lib.bfl_sgd_init(self._h, data) if is_path else self._lib.bfl_sgd_init_json(self._h, data)
        if rc == 1:
            self.last_error = self._lib.bfl_last_error().decode("utf-8", "replace")
            return False
        _cabi.check(rc, "bfl_sgd_init")
        return True

    def get_vdim(self):
        return self._lib.bfl_sgd_get_vdim(self._h)

    def initialize_model(self, P, Q, Qb, num_nnz, set_gpu=True):
        vdim = self.get_vdim()
        if P.ndim != 2 or Q.ndim != 2 or P.shape[1] != vdim or Q.shape[1] != vdim:
            raise ValueError("P and Q must be [rows, vdim=%d] (got %s, %s): the backend copies rows*vdim floats"
                             % (vdim, P.shape, Q.shape))
        if Qb.shape != (Q.shape[0], 1):
            raise ValueError("Qb must be [%d, 1], got %s" % (Q.shape[0], Qb.shape))
        self._keep = [P, Q, Qb]
        _cabi.check(self._lib.bfl_sgd_initialize_model(self._h, _host(P, np.float32, 2, "P"), P.shape[0],
                                                       _host(Q, np.float32, 2, "Q"), Q.shape[0],
                                                       _host(Qb, np.float32, 2, "Qb"), int(num_nnz)),
                    "initialize_model")

    def set_cumulative_table(self, sampling_table, size):
        _cabi.check(self._lib.bfl_sgd_set_cumulative_table(self._h, _host(sampling_table, np.int64, 1, "cum"),
                                                           int(size)), "set_cumulative_table")

    def set_placeholder(self, indptr, batch_size):
        _cabi.check(self._lib.bfl_sgd_set_placeholder(self._h, _host(indptr, np.int64, 1, "indptr"),
                                                      int(batch_size)), "set_placeholder")

    def launch_workers(self):
        _cabi.check(self._lib.bfl_sgd_launch_workers(self._h), "launch_workers")

    def add_jobs(self, start_x, next_x, indptr, keys):
        _cabi.check(self._lib.bfl_sgd_add_jobs(self._h, int(start_x), int(next_x),
                                               _host(indptr, np.int64, 1, "indptr"),
                                               _host(keys, np.int32, 1, "keys")), "add_jobs")

    def update_parameters(self):
        _cabi.check(self._lib.bfl_sgd_update_parameters(self._h), "update_parameters")

    def wait_until_done(self):
        _cabi.check(self._lib.bfl_sgd_wait_until_done(self._h), "wait_until_done")

    def synchronize(self, device_to_host):
        _cabi.check(self._lib.bfl_sgd_synchronize(self._h, int(bool(device_to_host))), "synchronize")

    def compute_loss(self, users, positives, negatives):
        out = C.c_double(0.0)
        _cabi.check(self._lib.bfl_sgd_compute_loss(self._h, int(users.shape[0]), _host(users, np.int32, 1, "users"),
                                                   _host(positives, np.int32, 1, "positives"),
                                                   _host(negatives, np.int32, 1, "negatives"), C.byref(out)),
                    "compute_loss")
        return out.value

    def join(self):
        out = C.c_double(0.0)
        _cabi.check(self._lib.bfl_sgd_join(self._h, C.byref(out)), "join")
        return out.value

    # --- device-resident path ---------------------------------------------------------------
    def bind_factors(self, P, Q, Qb, num_total_samples):
        self._keep = [P, Q, Qb]
        _cabi.check(self._lib.bfl_sgd_bind_factors_device(self._h, _dev(P, "float32", "P"), P.shape[0],
                                                          _dev(Q, "float32", "Q"), Q.shape[0],
                                                          _dev(Qb, "float32", "Qb"), int(num_total_samples)),
                    "bind_factors")

    def bind_csr(self, indptr, keys):
        self._keep += [indptr, keys]
        _cabi.check(self._lib.bfl_sgd_bind_csr_device(self._h, _dev(indptr, "int64", "indptr"),
                                                      _dev(keys, "int32", "keys"), indptr.shape[0], keys.shape[0]),
                    "bind_csr")

    def add_jobs_device(self, row_begin, row_end, stream=None):
        _cabi.check(self._lib.bfl_sgd_add_jobs_device(self._h, int(row_begin), int(row_end), _stream_ptr(stream)),
                    "add_jobs_device")

    def update_parameters_device(self, stream=None):
        _cabi.check(self._lib.bfl_sgd_update_parameters_device(self._h, _stream_ptr(stream)),
                    "update_parameters_device")

    def sample_device(self, row_begin, row_end, users, pos, neg, stream=None):
        _cabi.check(self._lib.bfl_sgd_sample_device(self._h, int(row_begin), int(row_end),
                                                    _dev(users, "int32", "users"), _dev(pos, "int32", "pos"),
                                                    _dev(neg, "int32", "neg"), _stream_ptr(stream)), "sample_device")

    def apply_triples_device(self, users, pos, neg, lr, stream=None):
        _cabi.check(self._lib.bfl_sgd_apply_triples_device(self._h, _dev(users, "int32", "users"),
                                                           _dev(pos, "int32", "pos"), _dev(neg, "int32", "neg"),
                                                           users.shape[0], float(lr), _stream_ptr(stream)),
                    "apply_triples_device")

    def grad_tensor(self, which, shape):
        import torch
        ptr = self._lib.bfl_sgd_grad_device(self._h, int(which))
        if not ptr:
            return None
        n = int(np.prod(shape))

        class _Arr(object):
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Arr(), device="cuda").view(*shape)

    def count_tensor(self, which, rows):
        """int32[rows] sample counters of per_coordinate_normalize (0: P rows, 1: Q rows); None if not allocated."""
        import torch
        ptr = self._lib.bfl_sgd_count_device(self._h, int(which))
        if not ptr:
            return None

        class _Arr(object):
            __cuda_array_interface__ = {"shape": (int(rows),), "typestr": "<i4", "data": (ptr, False), "version": 2}
        return torch.as_tensor(_Arr(), device="cuda")

    def set_trace(self, trials, negs):
        self._keep += [trials, negs]
        _cabi.check(self._lib.bfl_sgd_set_trace_device(self._h, _dev(trials, "int32", "trials"),
                                                       _dev(negs, "int32", "negs")), "set_trace")

    def epoch(self):
        return self._lib.bfl_sgd_epoch(self._h)

    def current_lr(self):
        return self._lib.bfl_sgd_current_lr(self._h)

    def read_stats(self):
        loss, n = C.c_double(0.0), C.c_int64(0)
        _cabi.check(self._lib.bfl_sgd_read_stats(self._h, C.byref(loss), C.byref(n)), "read_stats")
        return loss.value, n.value


def device_available():
    """True when the CUDA library is loadable and a GPU is visible (used by host-side helpers that have a device
    implementation; the training backends never consult this: they fail loudly without a GPU)."""
    try:
        import torch
        return bool(torch.cuda.is_available()) and os.path.isfile(_cabi.LIB_PATH)
    except Exception:
        return False


def topk_host(queries, items, item_bias, k):
    """k best item indices per query row for scores = queries @ items.T (+ item_bias), best first, computed on the
    device (bfl_topk_host).  queries [nq, d], items [I, d] float32 host arrays; returns int32 [nq, k]."""
    q = np.ascontiguousarray(queries, dtype=np.float32)
    it = np.ascontiguousarray(items, dtype=np.float32)
    if q.ndim == 1:
        q = q.reshape(1, -1)
    d = min(q.shape[1], it.shape[1])
    k = int(min(k, it.shape[0]))
    out = np.empty((q.shape[0], k), dtype=np.int32)
    b = None if item_bias is None else np.ascontiguousarray(np.asarray(item_bias, dtype=np.float32).reshape(-1))
    _cabi.check(_cabi.lib().bfl_topk_host(q.ctypes.data, q.shape[0], q.shape[1], it.ctypes.data, it.shape[0],
                                          it.shape[1], None if b is None else b.ctypes.data, int(d), k,
                                          out.ctypes.data, None), "bfl_topk_host")
    return out


def topk_device(queries, items, item_bias, k, stream=None):
    """Device tensors in, device tensors out: (idx int32 [nq, k], val float32 [nq, k])."""
    import torch
    nq, n_items = queries.shape[0], items.shape[0]
    k = int(min(k, n_items))
    idx = torch.empty((nq, k), dtype=torch.int32, device=queries.device)
    val = torch.empty((nq, k), dtype=torch.float32, device=queries.device)
    _cabi.check(_cabi.lib().bfl_topk_device(_dev(queries, "float32", "queries"), nq, queries.stride(0),
                                            _dev(items, "float32", "items"), n_items, items.stride(0),
                                            None if item_bias is None else _dev(item_bias, "float32", "bias"),
                                            int(min(queries.shape[1], items.shape[1])), k, idx.data_ptr(), val.data_ptr(),
                                            _stream_ptr(stream)), "bfl_topk_device")
    return idx, val


def csr_from_triples_host(major, minor, vals, num_major, num_minor, sort_minor=True):
    """(indptr_end int64, key int32, val float32) of one orientation through the hand-written device radix sort
    (csrc/ingest.cu: bfl_csr_from_triples_host)."""
    mj = np.ascontiguousarray(major, dtype=np.int32)
    mn = np.ascontiguousarray(minor, dtype=np.int32)
    v = np.ascontiguousarray(vals, dtype=np.float32)
    n = len(mj)
    indptr = np.empty(int(num_major), dtype=np.int64)
    key = np.empty(max(n, 1), dtype=np.int32)
    val = np.empty(max(n, 1), dtype=np.float32)
    _cabi.check(_cabi.lib().bfl_csr_from_triples_host(mj.ctypes.data, mn.ctypes.data, v.ctypes.data, n, int(num_major),
                                                      int(max(num_minor, 1)), int(bool(sort_minor)), indptr.ctypes.data,
                                                      key.ctypes.data, val.ctypes.data), "bfl_csr_from_triples_host")
    return indptr, key[:n], val[:n]


def popularity_table_host(keys, n_items, power):
    """int64 cumulative table of count(item)**power (bpr.py:99-111) built on the device."""
    k = np.ascontiguousarray(keys, dtype=np.int32)
    cum = np.empty(int(n_items), dtype=np.int64)
    _cabi.check(_cabi.lib().bfl_popularity_table_host(k.ctypes.data, len(k), int(n_items), int(power), cum.ctypes.data),
                "bfl_popularity_table_host")
    return cum
