"""Species model + SVM on the score vector.

API of the reference's ``ProbabilisticFilterSVMModel`` (models/probabilistic_filter_svm_model.py:17-315).
Scoring runs on the GPU through the parent class; the SVM stays on the host with scikit-learn, fit from the
model's ``scores.csv`` exactly as the reference does (:225-274), including its handling of ``exclude_ids``; the
fitted classifier is kept while ``scores.csv`` and the excluded ids do not change (the reference refits per call —
the fit is deterministic, so the predictions are the same).
"""

from __future__ import annotations

import csv
import json
from pathlib import Path

from .probabilistic_filter_model import ProbabilisticFilterModel
from .result import ModelResult


class ProbabilisticFilterSVMModel(ProbabilisticFilterModel):
    def __init__(
        self,
        k: int,
        model_display_name: str,
        author: str | None,
        author_email: str | None,
        model_type: str,
        base_path: Path,
        kernel: str,
        c: float,
        fpr: float = 0.01,
        num_hashes: int = 7,
        training_accessions: dict[str, list[str]] | None = None,
        svm_accessions: dict[str, list[str]] | None = None,
    ) -> None:
        super().__init__(
            k=k,
            model_display_name=model_display_name,
            author=author,
            author_email=author_email,
            model_type=model_type,
            base_path=base_path,
            fpr=fpr,
            num_hashes=num_hashes,
            training_accessions=training_accessions,
        )
        self.kernel = kernel
        self.c = c
        self.svm_accessions = svm_accessions

    def to_dict(self) -> dict:
        return super().to_dict() | {"kernel": self.kernel, "C": self.c, "svm_accessions": self.svm_accessions}

    def set_svm_params(self, kernel: str, c: float) -> None:
        self.kernel = kernel
        self.c = c
        self.save()

    def fit(self, dir_path: Path, svm_path: Path, display_names: dict[str, str] | None = None, svm_step: int = 1,
            training_accessions=None, svm_accessions=None, device: int | None = None) -> None:
        """Build the index (parent ``fit``), then score every file of ``svm_path/<label>/`` and write
        ``<slug>/scores.csv`` — rows ``accession,<total score per sorted document id>,label`` under the header
        ``file,<sorted ids>,label_id`` (reference :112-173)."""
        from ..definitions import fasta_endings, fastq_endings
        super().fit(dir_path, display_names=display_names, training_accessions=training_accessions, device=device)
        self.svm_accessions = svm_accessions
        score_list = []
        for species_folder in sorted(svm_path.iterdir()):
            if not species_folder.is_dir():
                continue
            for file in sorted(species_folder.iterdir()):
                if file.suffix[1:] not in fasta_endings + fastq_endings:
                    continue
                print(f"Calculating {file.name} scores for SVM training...")
                res = ProbabilisticFilterModel.predict(self, file, step=svm_step)
                scores = dict(sorted(res.get_total_scores().items()))
                score_list.append(f"{file.stem},{','.join(str(v) for v in scores.values())},{species_folder.name}")
        keys = sorted(self.display_names.keys())
        score_list.insert(0, f"file,{','.join(keys)},label_id")
        with open(self.base_path / self.slug() / "scores.csv", "w", encoding="utf-8") as file:
            file.write("\n".join(score_list))

    def svm_input(self, res: ModelResult) -> list[list[float]]:
        """The SVM feature row: total scores ordered by sorted label string (reference :212-213)."""
        total = res.get_total_scores() if hasattr(res, "get_total_scores") else res.get_scores()["total"]
        return [list(dict(sorted(total.items())).values())]

    def predict(self, sequence_input, exclude_ids: list[str] = None, step: int = 1, display_name: bool = False,
                validation: bool = False) -> ModelResult:
        res = super().predict(sequence_input, exclude_ids, step, display_name, validation)
        svm_scores = self.svm_input(res)
        svm = self._get_svm(exclude_ids)
        res.prediction = str(svm.predict(svm_scores)[0])     # same fields as the reference's re-wrapped ModelResult
        return res

    def _get_svm(self, exclude_ids):
        """SVC(kernel, C) fit on ``<slug>/scores.csv``.  Bug-compatible with the reference (:240-267): feature
        columns to drop are located in the *unsorted* display_names key order although the CSV columns are
        sorted; rows labelled with an excluded id are dropped."""
        from sklearn.svm import SVC

        # the fit is deterministic in scores.csv, kernel, C and the excluded ids: keep the fitted classifier while
        # those do not change (the reference refits on every predict call, ~70 ms for 360 training rows)
        csv_path = self.base_path / self.slug() / "scores.csv"
        st = csv_path.stat()
        key = (st.st_mtime_ns, st.st_size, self.kernel, self.c, tuple(self.display_names.keys()),
               None if exclude_ids is None else frozenset(exclude_ids))
        cache = self.__dict__.setdefault("_svm_cache", {})
        if key in cache:
            return cache[key]
        svm = SVC(kernel=self.kernel, C=self.c)
        keys = list(self.display_names.keys())
        drop = {i for i, key in enumerate(keys) if exclude_ids is not None and key in exclude_ids}
        x_train, y_train = [], []
        with open(csv_path, "r", encoding="utf-8") as file:
            file.readline()
            for row in csv.reader(file):
                label = row[-1]
                if exclude_ids is not None and label in exclude_ids:
                    continue
                x_train.append([float(v) for i, v in enumerate(row[1:-1]) if i not in drop])
                y_train.append(label)
        svm.fit(x_train, y_train)
        cache.clear()                      # one classifier per model object is enough
        cache[key] = svm
        return svm

    @staticmethod
    def load(path: Path, device: int | None = None) -> "ProbabilisticFilterSVMModel":
        with open(path, "r", encoding="utf-8") as file:
            model_json = json.loads(file.read())
        model = ProbabilisticFilterSVMModel(
            model_json["k"],
            model_json["model_display_name"],
            model_json["author"],
            model_json["author_email"],
            model_json["model_type"],
            path.parent,
            model_json["kernel"],
            model_json["C"],
            fpr=model_json["fpr"],
            num_hashes=model_json["num_hashes"],
            training_accessions=model_json["training_accessions"],
            svm_accessions=model_json["svm_accessions"],
        )
        model.display_names = model_json["display_names"]
        model._open_index(device)
        return model
