"""Sequence filtering workflows (filter_sequences.py:12-124 of the reference)."""

from pathlib import Path

from .file_io import filter_sequences, prepare_input_output_paths
from .model_management import get_genus_model_path, get_species_model_path


def _filter(model, label: str, what: str, input_path: Path, output_path: Path, threshold: float,
            classification_output_path: Path | None, sparse_sampling_step: int):
    input_paths, get_output_path = prepare_input_output_paths(input_path)
    for idx, current_path in enumerate(input_paths):
        result = model.predict(current_path, step=sparse_sampling_step)
        result.input_source = current_path.name
        if classification_output_path:
            cls_out = get_output_path(idx, classification_output_path)
            result.save(cls_out)
            print(f"Saved classification results from {current_path.name} as {cls_out.name}")
        included_ids = result.get_filtered_subsequence_labels(label, threshold)
        if not included_ids:
            print(f"No sequences found for the given {what} in {current_path.name}.")
            continue
        filter_output_path = get_output_path(idx, output_path)
        filter_sequences(current_path, filter_output_path, included_ids)
        print(f"Saved filtered sequences from {current_path.name} as {filter_output_path.name}")


def filter_species(
    model_genus: str,
    model_species: str,
    input_path: Path,
    output_path: Path,
    threshold: float,
    classification_output_path: Path | None = None,
    sparse_sampling_step: int = 1,
):
    from .models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel

    model = ProbabilisticFilterSVMModel.load(get_species_model_path(model_genus))
    _filter(model, model_species, "species", input_path, output_path, threshold, classification_output_path, sparse_sampling_step)


def filter_genus(
    model_genus: str,
    input_path: Path,
    output_path: Path,
    threshold: float,
    classification_output_path: Path | None = None,
    sparse_sampling_step: int = 1,
):
    from .models.probabilistic_single_filter_model import ProbabilisticSingleFilterModel

    model = ProbabilisticSingleFilterModel.load(get_genus_model_path(model_genus))
    _filter(model, model_genus, "genus", input_path, output_path, threshold, classification_output_path, sparse_sampling_step)
