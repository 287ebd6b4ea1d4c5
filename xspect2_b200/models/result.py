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
        totals = dict.fromkeys(next(iter(self.hits.values())), 0)
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
