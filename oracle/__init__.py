"""CPU oracle for the matrix-factorisation training hot path (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product packages
(``buffalo_b200`` and its ``buffalo`` alias) never do.

PARITY UNPINNED: the reference ships no golden vectors for this path and cannot be
built in this image; see the header of ``buffalo_oracle.c`` and DESIGN.md.

``OracleALS`` / ``OracleSGD`` mirror the method set of the reference's Cython holders
(buffalo/algo/_als.pyx:28-63, _bpr.pyx:34-92, _warp.pyx:34-92) so that parity tests
read like the reference's own call sites (buffalo/algo/als.py:115-142).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libbuffalo_oracle.so")

OPTIMIZER_CODES = {"llt": 0, "ldlt": 1, "manual_cg": 2, "ialspp": 8}
SGD_OPTIMIZERS = {"sgd": 0, "adagrad": 1, "adam": 2}


def build(force=False):
    """Compile the C restatement with oracle/Makefile (gcc + OpenMP)."""
    src = os.path.join(_HERE, "buffalo_oracle.c")
    if force or not os.path.isfile(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _AlsOpt(C.Structure):
    _fields_ = [("d", C.c_int32), ("num_workers", C.c_int32), ("num_cg_max_iters", C.c_int32),
                ("optimizer_code", C.c_int32), ("block_size", C.c_int32), ("adaptive_reg", C.c_int32),
                ("compute_loss", C.c_int32), ("alpha", C.c_float), ("reg_u", C.c_float),
                ("reg_i", C.c_float), ("eps", C.c_float), ("cg_tolerance", C.c_float)]


class _SgdOpt(C.Structure):
    _fields_ = [("d", C.c_int32), ("num_workers", C.c_int32), ("optimizer", C.c_int32),
                ("use_bias", C.c_int32), ("update_i", C.c_int32), ("update_j", C.c_int32),
                ("num_negative_samples", C.c_int32), ("verify_neg", C.c_int32),
                ("uniform_sampling", C.c_int32), ("per_coordinate_normalize", C.c_int32),
                ("max_trials", C.c_int32), ("score_l2", C.c_int32), ("random_seed", C.c_int32),
                ("num_iters", C.c_int32), ("reg_u", C.c_float), ("reg_i", C.c_float),
                ("reg_j", C.c_float), ("reg_b", C.c_float), ("lr", C.c_float), ("min_lr", C.c_float),
                ("beta1", C.c_float), ("beta2_unused", C.c_float), ("threshold", C.c_float)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_bpr_compute_loss.restype = C.c_double
        _lib.orc_warp_compute_loss.restype = C.c_double
        _lib.orc_lr_decay.restype = C.c_double
        _lib.orc_lr_decay.argtypes = [C.c_double] * 4
        _lib.orc_draw_range.restype = C.c_int32
        _lib.orc_draw_range.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32]
    return _lib


def _p(a, typ):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(typ))


def _f32(a):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"], "float32 C-contiguous expected"
    return _p(a, C.c_float)


def _i32(a):
    assert a.dtype == np.int32 and a.flags["C_CONTIGUOUS"]
    return _p(a, C.c_int32)


def _i64(a):
    assert a.dtype == np.int64 and a.flags["C_CONTIGUOUS"]
    return _p(a, C.c_int64)


def als_opt_struct(opt):
    d = int(opt["d"])
    code = OPTIMIZER_CODES.get(opt.get("optimizer", "manual_cg"))
    if code is None:
        raise ValueError("oracle restates optimizers %s only" % sorted(OPTIMIZER_CODES))
    code = lib().orc_als_effective_optimizer(d, code)  # d >= 128 => ialspp (als.cc:46)
    return _AlsOpt(d=d, num_workers=int(opt.get("num_workers", 1)),
                   num_cg_max_iters=int(opt.get("num_cg_max_iters", 3)), optimizer_code=code,
                   block_size=int(opt.get("block_size", 32)), adaptive_reg=int(bool(opt.get("adaptive_reg", False))),
                   compute_loss=int(bool(opt.get("compute_loss_on_training", True))),
                   alpha=float(opt.get("alpha", 8.0)), reg_u=float(opt.get("reg_u", 0.1)),
                   reg_i=float(opt.get("reg_i", 0.1)), eps=float(opt.get("eps", 1e-10)),
                   cg_tolerance=float(opt.get("cg_tolerance", 1e-10)))


class OracleALS(object):
    """Same method set as CyALS (buffalo/algo/_als.pyx:28-63); options as a dict."""

    def __init__(self):
        self.o = None

    def init(self, opt):
        self.o = als_opt_struct(opt)
        self.FF = np.zeros((self.o.d, self.o.d), dtype=np.float32)
        return True

    def initialize_model(self, P, Q):
        self.P, self.Q = P, Q  # kept by reference, mutated in place (als.cc:76-83)

    def precompute(self, axis):
        F = self.Q if axis == 0 else self.P
        lib().orc_als_precompute(_f32(F), C.c_int64(F.shape[0]), self.o.d, _f32(self.FF), self.o.num_workers)

    def partial_update(self, start_x, next_x, indptr, keys, vals, axis):
        nume, deno = C.c_double(0), C.c_double(0)
        rc = lib().orc_als_partial_update(C.byref(self.o), _f32(self.P), C.c_int64(self.P.shape[0]),
                                          _f32(self.Q), C.c_int64(self.Q.shape[0]), _f32(self.FF),
                                          int(start_x), int(next_x), _i64(indptr), _i32(keys), _f32(vals),
                                          int(axis), C.byref(nume), C.byref(deno))
        assert rc == 0
        return nume.value, deno.value


def sgd_opt_struct(opt, warp=False):
    return _SgdOpt(d=int(opt["d"]), num_workers=int(opt.get("num_workers", 1)),
                   optimizer=SGD_OPTIMIZERS[opt.get("optimizer", "adagrad" if warp else "sgd")],
                   use_bias=int(bool(opt.get("use_bias", not warp))), update_i=int(bool(opt.get("update_i", True))),
                   update_j=int(bool(opt.get("update_j", True))),
                   num_negative_samples=int(opt.get("num_negative_samples", 1)),
                   verify_neg=int(bool(opt.get("verify_neg", True))),
                   uniform_sampling=int(float(opt.get("sampling_power", 0.0)) == 0.0),
                   per_coordinate_normalize=int(bool(opt.get("per_coordinate_normalize", False))),
                   max_trials=int(opt.get("max_trials", 500)),
                   score_l2=int(str(opt.get("score_func", "dot")).lower() == "l2"),
                   random_seed=int(opt.get("random_seed", 0)), num_iters=int(opt.get("num_iters", 1)),
                   reg_u=float(opt.get("reg_u", 0.0)), reg_i=float(opt.get("reg_i", 0.0)),
                   reg_j=float(opt.get("reg_j", 0.0)), reg_b=float(opt.get("reg_b", 0.0)),
                   lr=float(opt.get("lr", 0.05)), min_lr=float(opt.get("min_lr", 0.0001)),
                   beta1=float(opt.get("beta1", 0.9)), beta2_unused=float(opt.get("beta2", 0.999)),
                   threshold=float(opt.get("threshold", 1.0)))


class OracleSGD(object):
    """Deterministic single-worker restatement of CBPRMF / CWARP + SGDAlgorithm.

    Method names follow CyBPRMF / CyWARP (buffalo/algo/_bpr.pyx:45-92); the worker
    threads / job queue of the reference are replaced by immediate sequential
    execution inside add_jobs (see buffalo_oracle.c for the documented deviations).
    """

    def __init__(self, warp=False, use_lut=True):
        self.warp = warp
        self.use_lut = use_lut

    def init(self, opt):
        self.opt = dict(opt)
        self.o = sgd_opt_struct(opt, warp=self.warp)
        return True

    def initialize_model(self, P, Q, Qb, num_total_samples):
        self.P, self.Q, self.Qb = P, Q, Qb
        self.iters = 0
        self.epoch = 0
        self.processed = 0.0
        self.total = float(num_total_samples) * self.o.num_iters
        # lr / min_lr / beta1 are doubles in the reference (opt_["lr"].number_value(), algo.cc:267-268,394-396)
        self.lr0 = float(self.opt.get("lr", 0.05))
        self.min_lr = float(self.opt.get("min_lr", 0.0001))
        self.beta1 = float(self.opt.get("beta1", 0.9))
        self.lr = self.lr0
        z = np.zeros_like
        if self.o.optimizer != 0:  # initialize_adam_optimizer (algo.cc:221-254)
            self.gP, self.gQ, self.gQb = z(P), z(Q), z(Qb)
            self.mP, self.mQ, self.mQb = z(P), z(Q), z(Qb)
            self.vP, self.vQ, self.vQb = z(P), z(Q), z(Qb)
        else:
            self.gP = self.gQ = self.gQb = None
        self.cP = np.zeros(P.shape[0], dtype=np.int32)
        self.cQ = np.zeros(Q.shape[0], dtype=np.int32)
        self.cum = None
        self.warp_loss = 0.0
        self.warp_updates = 0

    def set_cumulative_table(self, cum, size):
        self.cum = np.ascontiguousarray(cum, dtype=np.int64)

    def launch_workers(self):
        pass

    def sample(self, start_x, next_x, indptr, keys):
        beg = 0 if start_x == 0 else int(indptr[start_x - 1])
        n = (int(indptr[next_x - 1]) - beg) * self.o.num_negative_samples
        u = np.empty(n, dtype=np.int32)
        p = np.empty(n, dtype=np.int32)
        g = np.empty(n, dtype=np.int32)
        lib().orc_bpr_sample(C.byref(self.o), int(self.Q.shape[0]), int(start_x), int(next_x), _i64(indptr),
                             _i32(keys), _i64(self.cum) if self.cum is not None else None,
                             C.c_uint32(self.epoch), _i32(u), _i32(p), _i32(g))
        return u, p, g

    def apply_triples(self, u, p, g, lr=None):
        lr = self.lr if lr is None else lr
        lib().orc_bpr_update(C.byref(self.o), _f32(self.P), _f32(self.Q), _f32(self.Qb),
                             _f32(self.gP) if self.gP is not None else None,
                             _f32(self.gQ) if self.gQ is not None else None,
                             _f32(self.gQb) if self.gQb is not None else None,
                             _i32(self.cP), _i32(self.cQ), _i32(u), _i32(p), _i32(g), C.c_int64(len(u)),
                             C.c_float(lr), int(self.use_lut))

    def add_jobs(self, start_x, next_x, indptr, keys, trials_out=None, negs_out=None):
        if next_x - start_x == 0:
            return
        # job.alpha = lr_ at job creation (algo.cc:351,359); decay by processed fraction (algo.cc:284-287)
        self.lr = lib().orc_lr_decay(self.lr0, self.min_lr, self.processed, self.total)
        beg = 0 if start_x == 0 else int(indptr[start_x - 1])
        nnz = int(indptr[next_x - 1]) - beg
        if self.warp:
            loss, upd = C.c_double(0), C.c_int64(0)
            lib().orc_warp_accumulate(C.byref(self.o), _f32(self.P), _f32(self.Q), int(self.Q.shape[0]),
                                      _f32(self.gP), _f32(self.gQ), _i32(self.cP), _i32(self.cQ),
                                      int(start_x), int(next_x), _i64(indptr), _i32(keys),
                                      C.c_uint32(self.epoch), C.byref(loss), C.byref(upd),
                                      _i32(trials_out) if trials_out is not None else None,
                                      _i32(negs_out) if negs_out is not None else None)
            self.warp_loss += loss.value
            self.warp_updates += upd.value
        else:
            u, p, g = self.sample(start_x, next_x, indptr, keys)
            self.apply_triples(u, p, g)
        self.processed += nnz

    def update_parameters(self):
        o = self.o
        if o.optimizer != 0:
            for th, g, m, v, c, reg in [(self.P, self.gP, self.mP, self.vP, self.cP, o.reg_u),
                                        (self.Q, self.gQ, self.mQ, self.vQ, self.cQ, o.reg_i)]:
                lib().orc_sgd_apply(o.optimizer, _f32(th), _f32(g), _f32(m), _f32(v), _i32(c),
                                    C.c_int64(th.shape[0]), int(th.shape[1]), C.c_double(reg),
                                    C.c_double(self.lr0), C.c_double(self.beta1), int(self.iters),
                                    int(o.per_coordinate_normalize), int(o.num_workers))
            if o.use_bias and not self.warp:
                # bias grads are normalised by the item counters too (algo.cc:411-413)
                lib().orc_sgd_apply(o.optimizer, _f32(self.Qb), _f32(self.gQb), _f32(self.mQb), _f32(self.vQb),
                                    _i32(self.cQ), C.c_int64(self.Qb.shape[0]), 1, C.c_double(o.reg_b),
                                    C.c_double(self.lr0), C.c_double(self.beta1), int(self.iters),
                                    int(o.per_coordinate_normalize), int(o.num_workers))
            if o.per_coordinate_normalize:
                self.cP[:] = 0
                self.cQ[:] = 0
        if self.warp:  # CWARP::update_parameters (warp.cc:192-201)
            lib().orc_warp_project(_f32(self.Q), C.c_int64(self.Q.shape[0]), o.d, o.num_workers)
            lib().orc_warp_project(_f32(self.P), C.c_int64(self.P.shape[0]), o.d, o.num_workers)
        self.iters += 1
        self.epoch += 1

    def wait_until_done(self):
        pass

    def compute_loss(self, users, positives, negatives):
        n = len(users)
        if self.warp:
            return lib().orc_warp_compute_loss(_f32(self.P), _f32(self.Q), self.o.d, self.o.score_l2,
                                               C.c_double(self.o.threshold), _i32(users), _i32(positives),
                                               _i32(negatives), n)
        return lib().orc_bpr_compute_loss(_f32(self.P), _f32(self.Q), _f32(self.Qb), self.o.d, self.o.use_bias,
                                          _i32(users), _i32(positives), _i32(negatives), n)

    def join(self):
        return 0.0


def draw_range(seed, epoch, idx, t, rng):
    return lib().orc_draw_range(seed, epoch, idx, t, rng)
