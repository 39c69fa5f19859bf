"""Host-side mixins shared by the trainers (buffalo/algo/base.py): id maps, top-k queries, similarity,
early stopping / save-best bookkeeping and the length-prefixed pickle container of Serializable.
Pure host glue around the factor matrices; nothing here is on the GPU hot path."""
import abc
import pickle
import struct

import numpy as np

from buffalo_b200.misc import aux

EPS = 1e-8


class Algo(abc.ABC):
    def __init__(self, *args, **kwargs):
        self._idmanager = aux.Option({"userid": [], "userid_map": {}, "itemid": [], "itemid_map": {},
                                      "userid_mapped": False, "itemid_mapped": False})

    # ---- options / lifecycle ------------------------------------------------------------------
    def get_option(self, opt_path):
        """dict options are dumped to a temp JSON whose PATH goes to the native side (base.py:18-24)."""
        if isinstance(opt_path, (dict, aux.Option)):
            opt_path = self.create_temporary_option_from_dict(opt_path)
        opt = aux.Option(opt_path)
        self.is_valid_option(opt)
        return aux.Option(opt), opt_path

    def initialize(self):
        self._es_round, self._es_min = 0, 987654321
        if self.opt.random_seed:
            np.random.seed(self.opt.random_seed)        # base.py:33-34: unseeded when random_seed == 0

    @abc.abstractmethod
    def normalize(self, group="item"):
        raise NotImplementedError

    def _normalize(self, feat):
        return feat / np.sqrt((feat ** 2).sum(-1) + EPS)[..., np.newaxis]

    def periodical(self, period, current):
        return (not period) or (current + 1) % period == 0

    def save_best_only(self, loss, best_loss, i):
        if self.opt.save_best and best_loss > loss and self.periodical(self.opt.save_period, i):
            self.save(self.opt.model_path)
            return loss
        return best_loss

    def early_stopping(self, loss):
        if self.opt.early_stopping_rounds < 1:
            return False
        self._es_round = self._es_round + 1 if self._es_min < loss else 0
        self._es_min = loss                              # base.py:216-221: last loss, not the running minimum
        if self._es_round >= self.opt.early_stopping_rounds:
            self.logger.info("Reached at early_stopping rounds, stopping train.")
            return True
        return False

    # ---- id maps ------------------------------------------------------------------------------
    def _build_map(self, field, count_key, ids_attr, map_attr, flag):
        names = self.data.get_group("idmap")[field]
        n = self.data.get_header()[count_key]
        ids = [str(i) for i in range(n)] if names.shape[0] == 0 else [b.decode("utf-8", "ignore") for b in names[:]]
        self._idmanager[ids_attr] = ids
        self._idmanager[map_attr] = {v: i for i, v in enumerate(ids)}
        self._idmanager[flag] = True

    def build_itemid_map(self):
        self._build_map("cols", "num_items", "itemids", "itemid_map", "itemid_mapped")

    def build_userid_map(self):
        self._build_map("rows", "num_users", "userids", "userid_map", "userid_mapped")

    def get_index(self, keys, group="item"):
        many = isinstance(keys, list)
        keys = keys if many else [keys]
        if group == "item":
            if not self._idmanager.itemid_mapped:
                self.build_itemid_map()
            idx = [self._idmanager.itemid_map.get(k) for k in keys]
        elif group == "user":
            if not self._idmanager.userid_mapped:
                self.build_userid_map()
            idx = [self._idmanager.userid_map.get(k) for k in keys]
        else:
            idx = []
        return np.array(idx) if many else idx[0]

    def get_index_pool(self, pool, group="item"):
        if isinstance(pool, list):
            pool = np.array([p for p in self.get_index(pool, group) if p is not None])
        elif not isinstance(pool, np.ndarray):
            raise ValueError("Unexpected type for pool: %s" % type(pool))
        return pool

    # ---- queries ------------------------------------------------------------------------------
    def _get_topk_recommendation(self, p, Q, pb, Qb, pool, topk, num_workers):
        if pool is not None:
            Q = Q[pool]
            Qb = Qb[pool] if Qb is not None else None
        from buffalo_b200 import backend
        if backend.device_available() and 0 < topk <= 4096 and Q.shape[0] >= 1:
            # scores and top-k on the device (csrc/topk.cu); pb is constant per query row and does not change the order
            topks = backend.topk_host(p, Q, Qb, topk)
            return topks if pool is None else np.array([pool[t] for t in topks])
        scores = p.dot(Q.T)
        if pb is not None:
            scores += pb
        if Qb is not None:
            scores += Qb.T
        topks = self.get_topk(scores, k=topk, num_threads=num_workers)
        return topks if pool is None else np.array([pool[t] for t in topks])

    def topk_recommendation(self, keys, topk=10, pool=None):
        many = isinstance(keys, list)
        keys = keys if many else [keys]
        if not self._idmanager.userid_mapped:
            self.build_userid_map()
        if not self._idmanager.itemid_mapped:
            self.build_itemid_map()
        if pool is not None:
            pool = self.get_index_pool(pool, group="item")
            if len(pool) == 0:
                return []
        rows = [self._idmanager.userid_map[k] for k in keys if k in self._idmanager.userid_map]
        recs = list(self._get_topk_recommendation(rows, topk, pool))
        if not recs:
            return []
        named = {self._idmanager.userids[r]: [self._idmanager.itemids[v] for v in vv] for r, vv in recs}
        return named if many else next(iter(named.values()))

    def most_similar(self, key, topk=10, group="item", pool=None):
        if group != "item":
            return []
        if not self._idmanager.itemid_mapped:
            self.build_itemid_map()
        is_vec = isinstance(key, np.ndarray)
        q = key if is_vec else self._idmanager.itemid_map.get(key)
        if q is None:
            return []
        if pool is not None:
            pool = self.get_index_pool(pool, group="item")
            if len(pool) == 0:
                return []
        topks, scores = self._get_most_similar_item(q, topk, pool)
        return [(self._idmanager.itemids[k], v) for k, v in zip(topks, scores) if is_vec or k != q]

    def _get_most_similar_item(self, col, topk, Factor, nrz, pool):
        if isinstance(col, np.ndarray):
            q = col
        else:
            topk += 1
            q = Factor[col]
        cand = Factor if pool is None else Factor[pool]
        dot = q.dot(cand.T)
        if not nrz:
            dot = dot / (np.linalg.norm(q) * np.linalg.norm(cand, axis=1) + EPS)
        topks = self.get_topk(dot, k=topk, num_threads=self.opt.num_workers)
        scores = dot[topks]
        if pool is not None:
            topks = np.array([pool[t] for t in topks])
        return topks, scores

    def get_feature(self, name, group="item"):
        index = self.get_index(name, group=group)
        return None if index is None else self._get_feature(index, group)

    @abc.abstractmethod
    def _get_feature(self, index, group="item"):
        raise NotImplementedError

    def get_weighted_feature(self, weights, group="item", min_length=1):
        if isinstance(weights, dict):
            feat = [(self.get_feature(k), w) for k, w in weights.items()]
            feat = [f * w for f, w in feat if f is not None]
        else:
            feat = [f for f in (self.get_feature(k) for k, _ in weights) if f is not None]
        if len(feat) < min_length:
            return None
        feat = np.array(feat, dtype=np.float64).mean(axis=0)
        return (feat / np.linalg.norm(feat) + EPS).astype(np.float32)


class Serializable(abc.ABC):
    """Container: u64 count, then per object u64 name length, name, u64 payload length, pickle (base.py:275-311)."""

    def __init__(self, *args, **kwargs):
        pass

    def _get_data(self):
        return [("_idmanager", self._idmanager)]

    def save(self, path=None, with_itemid_map=True, with_userid_map=True, data_fields=[]):
        path = self.opt.model_path if path is None else path
        if with_itemid_map and not self._idmanager.itemid_mapped:
            self.build_itemid_map()
        if with_userid_map and not self._idmanager.userid_mapped:
            self.build_userid_map()
        items = [(k, v) for k, v in self._get_data() if not data_fields or k in data_fields]
        with open(path, "wb") as fout:
            fout.write(struct.pack("Q", len(items)))
            for name, obj in items:
                bname, blob = name.encode("utf-8"), pickle.dumps(obj, protocol=4)
                fout.write(struct.pack("Q", len(bname)) + bname + struct.pack("Q", len(blob)) + blob)

    def load(self, path, data_fields=[]):
        with open(path, "rb") as fin:
            (count,) = struct.unpack("Q", fin.read(8))
            for _ in range(count):
                (n,) = struct.unpack("Q", fin.read(8))
                name = fin.read(n).decode("utf8")
                (size,) = struct.unpack("Q", fin.read(8))
                if data_fields and name not in data_fields:
                    fin.seek(size, 1)
                    continue
                setattr(self, name, pickle.loads(fin.read(size)))

    @classmethod
    def instantiate(cls, cls_opt, path, data_fields):
        obj = cls(cls_opt().get_default_option())
        obj.load(path, data_fields)
        return obj
