"""Row-aligned chunk feed of a CSR orientation (buffalo/data/buffered_data.py:27-172).

Same chunking rule as the reference: at most limit = batch_mb * 2^20 / 16 / 2 entries per chunk, cut on row
boundaries with bisect_left(indptr, beg + limit); the whole-matrix case is yielded once without copying.
Reference quirk NOT reproduced: its loop stops when next_x + 1 >= max_x and can drop a trailing
single-row chunk; here every row is fed exactly once per pass."""
import bisect

import numpy as np

from buffalo_b200.misc import log


class BufferedData(object):
    def __init__(self):
        self.logger = log.get_logger("BufferedData")


class BufferedDataMatrix(BufferedData):
    def __init__(self):
        super().__init__()
        self.group = "rowwise"
        self.major = {"rowwise": {}, "colwise": {}, "sppmi": {}}

    def initialize(self, data, with_sppmi=False):
        self.data = data
        limit = max(int((data.opt.data.batch_mb * 1024 * 1024) / 16.0), 64)   # 16 B per entry (indptr8,key4,val4)
        header = data.get_header()
        need = 0
        for G in ("rowwise", "colwise"):
            grp = data.get_group(G)
            indptr = np.ascontiguousarray(grp["indptr"][:], dtype=np.int64)
            if len(indptr):
                need = max(need, int(np.max(np.diff(indptr, prepend=0))))
            self.major[G] = {"limit": limit // 2, "start_x": 0, "next_x": 0, "indptr": indptr,
                             "max_x": header["num_users"] if G == "rowwise" else header["num_items"],
                             "keys": None, "vals": None, "sz": 0}
        if need > limit // 2:
            self.logger.warning("Given batch size(%d) is smaller than minimum required batch size(%d) for the data. "
                                "Increasing batch_mb would be helpful for faster traininig.", limit // 2, need)
            for G in ("rowwise", "colwise"):
                self.major[G]["limit"] = need + 1

    def get_indptrs(self):
        return self.major["rowwise"]["indptr"], self.major["colwise"]["indptr"], self.major["rowwise"]["limit"]

    def set_group(self, group):
        assert group in ["rowwise", "colwise", "sppmi"], "Unexpected group: {}".format(group)
        self.group = group

    def fetch_batch(self):
        m = self.major[self.group]
        grp = self.data.get_group(self.group)
        indptr, rows = m["indptr"], m["max_x"]
        total = int(indptr[-1]) if rows else 0
        if total <= m["limit"]:          # whole matrix in one chunk, no copy (buffered_data.py:89-93)
            m["start_x"], m["next_x"] = 0, rows
            if m["keys"] is None or len(m["keys"]) != max(total, 1):
                m["keys"] = np.ascontiguousarray(grp["key"][:total], dtype=np.int32) if total else np.zeros(1, np.int32)
                m["vals"] = np.ascontiguousarray(grp["val"][:total], dtype=np.float32) if total else np.zeros(1, np.float32)
            m["sz"] = total
            yield total
            return
        start = 0
        while start < rows:
            beg = 0 if start == 0 else int(indptr[start - 1])
            nxt = bisect.bisect_left(indptr, beg + m["limit"])
            if nxt == start:
                raise RuntimeError("Need more memory to load the data, cannot load data with buffer size %d that should "
                                   "be at least %d. Increase batch_mb value to deal with this."
                                   % (m["limit"], int(indptr[nxt]) - beg))
            nxt = min(nxt, rows)
            end = int(indptr[nxt - 1])
            m["start_x"], m["next_x"], m["sz"] = start, nxt, end - beg
            m["keys"] = np.ascontiguousarray(grp["key"][beg:end], dtype=np.int32)
            m["vals"] = np.ascontiguousarray(grp["val"][beg:end], dtype=np.float32)
            yield end - beg
            start = nxt
        m["start_x"], m["next_x"] = 0, 0

    def get(self):
        m = self.major[self.group]
        return [m[k] for k in ("start_x", "next_x", "indptr", "keys", "vals")]
