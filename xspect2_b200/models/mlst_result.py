"""MLST result container with the reference's JSON layout (models/mlst_result.py:7-62)."""

import json
from pathlib import Path


class MlstResult:
    def __init__(self, scheme_model: str, steps: int, hits: dict[str, list[dict]], input_source: str | None = None):
        self.scheme_model = scheme_model
        self.steps = steps
        self.hits = hits
        self.input_source = input_source

    def get_results(self) -> dict:
        return dict(self.hits.items())

    def to_dict(self) -> dict:
        return {
            "Scheme": self.scheme_model,
            "Steps": self.steps,
            "Results": self.get_results(),
            "Input_source": self.input_source,
        }

    def save(self, output_path: Path | str) -> None:
        output_path = Path(output_path)
        output_path.parent.mkdir(exist_ok=True, parents=True)
        with open(output_path, "w", encoding="utf-8") as file:
            file.write(json.dumps(self.to_dict(), indent=4))
