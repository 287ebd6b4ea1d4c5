"""Result API of the species / genus models — field for field the reference's ``ModelResult``
(models/result.py:7-189): ``hits``, ``num_kmers``, ``sparse_sampling_step``, ``prediction``,
``input_source``, ``misclassified`` and the same JSON."""

from json import dumps
from pathlib import Path


class ModelResult:
    """Hits per subsequence and label, with scores derived on demand."""

    def __init__(
        self,
        model_slug: str,
        hits: dict[str, dict[str, int]],
        num_kmers: dict[str, int],
        sparse_sampling_step: int = 1,
        prediction: str | None = None,
        input_source: str | None = None,
    ):
        if "total" in hits:
            raise ValueError("'total' is a reserved key and cannot be used as a subsequence")
        self.model_slug = model_slug
        self.hits = hits
        self.num_kmers = num_kmers
        self.sparse_sampling_step = sparse_sampling_step
        self.prediction = prediction
        self.input_source = input_source
        self.misclassified = self.hits.pop("misclassified", None)

    def get_total_hits(self) -> dict[str, int]:
        """Hits summed over subsequences; label set and order come from the first subsequence."""
        totals = dict.fromkeys(list(self.hits.values())[0], 0)   # IndexError on an empty result, like the reference
        for per_label in self.hits.values():
            for label, n in per_label.items():
                totals[label] += n
        return totals

    def get_scores(self) -> dict:
        """``round(hits / num_kmers, 2)`` per subsequence and label, plus the 'total' row."""
        scores = {}
        for subsequence, per_label in self.hits.items():
            denom = self.num_kmers[subsequence]
            scores[subsequence] = {label: round(n / denom, 2) for label, n in per_label.items()}
        total_kmers = sum(self.num_kmers.values())
        scores["total"] = {label: round(n / total_kmers, 2) for label, n in self.get_total_hits().items()}
        return scores

    def get_filter_mask(self, label: str, filter_threshold: float) -> dict[str, bool]:
        """Score >= threshold per subsequence, or (threshold == -1) label holds the maximum score."""
        if (filter_threshold < 0 and filter_threshold != -1) or filter_threshold > 1:
            raise ValueError("The filter threshold must be between 0 and 1.")
        scores = self.get_scores()
        del scores["total"]
        if filter_threshold != -1:
            return {sub: sc[label] >= filter_threshold for sub, sc in scores.items()}
        return {sub: sc[label] == max(sc.values()) for sub, sc in scores.items()}

    def get_filtered_subsequence_labels(self, label: str, filter_threshold: float = 0.7) -> list[str]:
        return [sub for sub, keep in self.get_filter_mask(label, filter_threshold).items() if keep]

    def to_dict(self) -> dict:
        res = {
            "model_slug": self.model_slug,
            "sparse_sampling_step": self.sparse_sampling_step,
            "hits": self.hits,
            "scores": self.get_scores(),
            "num_kmers": self.num_kmers,
            "misclassified": self.misclassified,
            "input_source": self.input_source,
        }
        if self.prediction is not None:
            res["prediction"] = self.prediction
        return res

    def save(self, path: Path) -> None:
        path.parent.mkdir(exist_ok=True, parents=True)
        with open(path, "w", encoding="utf-8") as f:
            f.write(dumps(self.to_dict(), indent=4))


class ColumnarModelResult(ModelResult):
    """A ModelResult backed by the count matrix of a batched query.

    Same attributes, methods and JSON as ``ModelResult``; the nested ``hits`` / ``num_kmers`` dictionaries are only
    built when somebody reads them.  ``get_total_hits`` and the total scores work on the matrix, and ``save``
    streams the JSON through the native writer (``xs_result_write_json``), byte for byte what
    ``json.dumps(self.to_dict(), indent=4)`` produces.  Dictionary semantics are kept: a record id that occurs
    twice keeps its first position and its last values (probabilistic_filter_model.py:295-310)."""

    def __init__(self, model_slug, ids, doc_names, doc_keys, doc_include, counts, num_kmers, sparse_sampling_step=1,
                 prediction=None, input_source=None):
        if "total" in ids:
            raise ValueError("'total' is a reserved key and cannot be used as a subsequence")
        self.model_slug = model_slug
        self.sparse_sampling_step = sparse_sampling_step
        self.prediction = prediction
        self.input_source = input_source
        self.misclassified = None
        self._ids = ids
        self._doc_names = doc_names          # index document names (matrix columns)
        self._doc_keys = doc_keys            # dictionary key per column (display-name rewrite applied)
        self._doc_include = doc_include      # bool per column (exclude_ids applied)
        self._counts = counts                # [n_records, n_docs]
        self._nk = num_kmers                 # [n_records]
        self._hits = None
        self._num_kmers = None
        self._emit = None
        if "misclassified" in ids:
            # reserved key of the reference's ModelResult (result.py:29): a record with this id has its hits moved
            # into the `misclassified` field.  Rare: take the dictionary path for such a batch.
            self._materialize()
            self.misclassified = self._hits.pop("misclassified", None)

    # ---- dictionary views, built on demand
    def _emit_index(self):
        """Rows that survive dict insertion: last occurrence of every id, in first-occurrence order."""
        if self._emit is None:
            import numpy as np
            last = {}
            for i, rid in enumerate(self._ids):
                last[rid] = i
            self._emit = np.fromiter(last.values(), dtype=np.uint64, count=len(last))
        return self._emit

    def _materialize(self):
        if self._hits is not None:
            return
        import numpy as np
        from .. import engine
        emit = self._emit_index()
        rows = np.ascontiguousarray(np.asarray(self._counts)[emit.astype(np.int64)]).astype(np.uint32)
        order = engine.result_order_batch(rows)
        keys = np.array(self._doc_keys, dtype=object)
        inc = np.asarray(self._doc_include, dtype=bool)
        sorted_counts = np.take_along_axis(rows, order.astype(np.int64), axis=1)
        hits, num_kmers = {}, {}
        all_in = bool(inc.all())
        for j, i in enumerate(emit.tolist()):
            o = order[j]
            if all_in:
                hits[self._ids[i]] = dict(zip(keys[o].tolist(), sorted_counts[j].tolist()))
            else:
                m = inc[o]
                hits[self._ids[i]] = dict(zip(keys[o][m].tolist(), sorted_counts[j][m].tolist()))
            num_kmers[self._ids[i]] = int(self._nk[i])
        self._hits, self._num_kmers = hits, num_kmers

    @property
    def hits(self):
        self._materialize()
        return self._hits

    @hits.setter
    def hits(self, value):
        self._hits = value

    @property
    def num_kmers(self):
        self._materialize()
        return self._num_kmers

    @num_kmers.setter
    def num_kmers(self, value):
        self._num_kmers = value

    # ---- matrix-side shortcuts
    def get_total_hits(self) -> dict[str, int]:
        if self._hits is not None:
            return super().get_total_hits()
        import numpy as np
        from .. import engine
        emit = self._emit_index().astype(np.int64)
        counts = np.asarray(self._counts)
        totals = counts[emit].sum(axis=0, dtype=np.int64)
        first = engine.CobsIndex.result_order(counts[emit[0]].astype(np.uint32))   # IndexError without records, like the reference
        return {self._doc_keys[d]: int(totals[d]) for d in first.tolist() if self._doc_include[d]}

    def get_filter_mask(self, label: str, filter_threshold: float) -> dict[str, bool]:
        """Same mask as ModelResult.get_filter_mask (result.py:92-123), taken on the matrix: the rounded score
        ``round(hits / num_kmers, 2)`` is evaluated with Python's own ``round`` once per distinct
        (hits, num_kmers) pair instead of once per record and label."""
        if self._hits is not None:
            return super().get_filter_mask(label, filter_threshold)
        if filter_threshold < 0 and not filter_threshold == -1 or filter_threshold > 1:
            raise ValueError("The filter threshold must be between 0 and 1.")
        import numpy as np
        inc = np.asarray(self._doc_include, dtype=bool)
        cols = [d for d, key in enumerate(self._doc_keys) if key == label and inc[d]]
        if not cols:
            raise KeyError(label)
        emit = self._emit_index().astype(np.int64)
        counts = np.asarray(self._counts)
        h = counts[emit, cols[-1]].astype(np.int64)
        n = np.asarray(self._nk)[emit].astype(np.int64)

        def rounded(hits, kmers):      # round(h / n, 2) with Python semantics, one evaluation per distinct pair
            pairs, inverse = np.unique(np.stack([hits, kmers], axis=1), axis=0, return_inverse=True)
            vals = np.array([round(int(a) / int(b), 2) for a, b in pairs], dtype=np.float64)
            return vals[inverse.reshape(-1)]

        if filter_threshold != -1:
            mask = rounded(h, n) >= filter_threshold
        else:
            hmax = counts[emit][:, inc].max(axis=1).astype(np.int64)
            mask = rounded(h, n) == rounded(hmax, n)
        return {self._ids[i]: bool(m) for i, m in zip(emit.tolist(), mask.tolist())}

    def get_total_scores(self) -> dict[str, float]:
        """``get_scores()["total"]`` without the per-record rows."""
        import numpy as np
        total_kmers = int(np.asarray(self._nk)[self._emit_index().astype(np.int64)].sum()) if self._hits is None else sum(self.num_kmers.values())
        return {label: round(h / total_kmers, 2) for label, h in self.get_total_hits().items()}

    def save(self, path: Path) -> None:
        if self._hits is not None or len(self._ids) == 0 or self.misclassified is not None:
            return super().save(path)
        import ctypes as C
        import json
        import numpy as np
        from .. import _abi
        path.parent.mkdir(exist_ok=True, parents=True)
        emit = self._emit_index()
        counts = np.ascontiguousarray(self._counts, dtype=np.uint32)
        nk = np.ascontiguousarray(np.asarray(self._nk)[emit.astype(np.int64)], dtype=np.uint64)

        def blob(strings):
            enc = [json.dumps(s).encode("ascii") for s in strings]
            ends = np.cumsum([len(e) for e in enc], dtype=np.uint64) if enc else np.zeros(0, np.uint64)
            return b"".join(enc), np.ascontiguousarray(ends, dtype=np.uint64)

        rec_blob, rec_end = blob([self._ids[i] for i in emit.tolist()])
        doc_blob, doc_end = blob(self._doc_keys)
        inc = np.ascontiguousarray(self._doc_include, dtype=np.uint8)
        prefix = "{\n    \"model_slug\": %s,\n    \"sparse_sampling_step\": %s,\n" % (json.dumps(self.model_slug), json.dumps(self.sparse_sampling_step))
        suffix = "    \"misclassified\": null,\n    \"input_source\": %s" % json.dumps(self.input_source)
        if self.prediction is not None:
            suffix += ",\n    \"prediction\": %s" % json.dumps(self.prediction)
        suffix += "\n}"
        _abi.check(_abi.lib().xs_result_write_json(
            str(path).encode(), prefix.encode("ascii"), suffix.encode("ascii"), counts.ctypes.data, counts.shape[1],
            emit.ctypes.data, emit.size, rec_blob, rec_end.ctypes.data, nk.ctypes.data, doc_blob, doc_end.ctypes.data, inc.ctypes.data))
