"""Ranking / score metrics on the held-out split (buffalo/evaluate/base.py:44-148): NDCG, MAP,
accuracy (= |topk & gt| / |gt|, :93), AUC, RMSE/error.  Top-k selection uses argpartition instead of the
reference's OpenMP quickselect (buffalo/parallel/_core.hpp:69-86) -- per-evaluation work on <= 500 rows."""
import numpy as np


def topk_indices(scores, k, ordered=True):
    """indices of the k largest entries per row, best first."""
    scores = np.asarray(scores)
    k = min(k, scores.shape[1])
    part = np.argpartition(-scores, k - 1, axis=1)[:, :k]
    if ordered:
        vals = np.take_along_axis(scores, part, axis=1)
        part = np.take_along_axis(part, np.argsort(-vals, axis=1, kind="stable"), axis=1)
    return part.astype(np.int32)


class Evaluable(object):
    def __init__(self, *args, **kwargs):
        pass

    def prepare_evaluation(self):
        if not self.opt.validation or not self.data.has_group("vali"):
            return
        if not hasattr(self.data, "vali_data"):
            self.data._prepare_validation_data()

    def show_validation_results(self):
        res = self.get_validation_results()
        if not res:
            return "No validation results"
        return "Validation results: " + ", ".join(f"{k}: {v:0.5f}" for k, v in res.items())

    def get_validation_results(self):
        if not self.opt.validation or not self.data.has_group("vali"):
            return
        res = {}
        res.update(self._evaluate_ranking_metrics())
        res.update(self._evaluate_score_metrics())
        return res

    def get_topk(self, scores, k, sorted=True, num_threads=4):
        scores = np.asarray(scores)
        single = scores.ndim == 1
        if single:
            scores = scores.reshape(1, -1)
        assert min(k, scores.shape[1]) > 0, f"k({k}) or cols({scores.shape[1]}) should be greater than 0"
        out = topk_indices(scores, k, ordered=sorted)
        return out[0] if single else out

    def _evaluate_ranking_metrics(self):
        if not hasattr(self.data, "vali_data"):
            self.prepare_evaluation()
        v = self.data.vali_data
        batch = self.opt.validation.get("batch", 128)
        topk = self.opt.validation.topk
        gt, rows, seen_of = v["vali_gt"], v["vali_rows"], v["validation_seen"]
        num_items = self.data.get_header()["num_items"]
        if self.opt.validation.eval_samples:
            rows = np.random.choice(rows, size=min(self.opt.validation.eval_samples, len(rows)), replace=False)
        gains = 1.0 / np.log2(np.arange(2, topk + 2))
        ideal = np.cumsum(gains)
        tot = dict(ndcg=0.0, map=0.0, accuracy=0.0, auc=0.0)
        count = 0.0
        for s in range(0, len(rows), batch):
            recs = self._get_topk_recommendation(rows[s:s + batch], topk=topk + v["validation_max_seen_size"])
            for row, cand in recs:
                seen = seen_of.get(row, set())
                if not seen:
                    continue
                ranked = [c for c in cand if c not in seen][:topk]
                truth = gt[row]
                hits = np.array([1.0 if r in truth else 0.0 for r in ranked])
                tot["accuracy"] += len(set(ranked) & truth) / len(truth)
                cum_hits = np.cumsum(hits)
                dcg = float((hits * gains[:len(hits)]).sum())
                ap = float((hits * cum_hits / np.arange(1, len(hits) + 1)).sum())
                n_pos, miss = len(truth), float((1 - hits).sum())
                n_neg = num_items - n_pos
                auc = float(((1 - hits) * cum_hits).sum()) + ((cum_hits[-1] if len(hits) else 0.0) + n_pos) / 2.0 * (n_neg - miss)
                tot["auc"] += auc / (n_pos * n_neg)
                tot["ndcg"] += dcg / ideal[min(n_pos, topk) - 1]
                tot["map"] += ap / min(n_pos, topk)
                count += 1.0
        return {k: val / count for k, val in tot.items()}

    def _evaluate_score_metrics(self):
        if not hasattr(self.data, "vali_data"):
            self.prepare_evaluation()
        v = self.data.vali_data
        err = np.asarray(self._get_scores(v["row"], v["col"]), dtype=np.float64) - v["val"]
        return {"rmse": float(np.sqrt(np.mean(err ** 2))), "error": float(np.mean(np.abs(err)))}
