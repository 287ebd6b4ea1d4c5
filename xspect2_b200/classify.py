"""Classification workflows (classify.py:12-121 of the reference): load the model by slug, classify every
input file, save one JSON per file."""

from pathlib import Path

from . import model_management as mm
from .file_io import prepare_input_output_paths


def classify_genus(model_genus: str, input_path: Path, output_path: Path, step: int = 1):
    from .models.probabilistic_single_filter_model import ProbabilisticSingleFilterModel

    model = ProbabilisticSingleFilterModel.load(mm.get_genus_model_path(model_genus))
    input_paths, get_output_path = prepare_input_output_paths(input_path)
    for idx, current_path in enumerate(input_paths):
        result = model.predict(current_path, step=step)
        result.input_source = current_path.name
        cls_path = get_output_path(idx, output_path)
        result.save(cls_path)
        print(f"Saved result as {cls_path.name}")


def classify_species(
    model_genus: str,
    input_path: Path,
    output_path: Path,
    step: int = 1,
    display_name: bool = False,
    validation: bool = False,
    exclude_ids: list[str] | None = None,
):
    if mm.is_svm_model(f"{model_genus}-species"):
        from .models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel as ModelClass
    else:
        from .models.probabilistic_filter_model import ProbabilisticFilterModel as ModelClass
    model = ModelClass.load(mm.get_species_model_path(model_genus))
    input_paths, get_output_path = prepare_input_output_paths(input_path)
    for idx, current_path in enumerate(input_paths):
        result = model.predict(current_path, exclude_ids=exclude_ids, step=step, display_name=display_name, validation=validation)
        result.input_source = current_path.name
        cls_path = get_output_path(idx, output_path)
        result.save(cls_path)
        print(f"Saved result as {cls_path.name}")


def classify_mlst(input_path: Path, organism, mlst_scheme, output_path: Path, limit: bool):
    from .models.probabilistic_filter_mlst_model import ProbabilisticFilterMlstSchemeModel

    model = ProbabilisticFilterMlstSchemeModel.load(mm.get_mlst_model_path(organism, mlst_scheme))
    input_paths, get_output_path = prepare_input_output_paths(input_path)
    for idx, current_path in enumerate(input_paths):
        result = model.predict(current_path, step=1, limit=limit)
        result.input_source = current_path.name
        cls_path = get_output_path(idx, output_path)
        result.save(cls_path)
        print(f"Saved result as {cls_path.name}")
