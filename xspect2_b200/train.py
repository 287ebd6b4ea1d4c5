"""Training from local data (train.py:28-184 of the reference): ``<dir>/cobs/<species>/*.fasta`` (and optionally
``<dir>/svm/<species>/*``) -> species model (+SVM), and with ``meta`` a genus Bloom model over all training genomes.
Index and filter construction run on the GPU through the model classes' ``fit``.  The NCBI / PubMLST driven
workflows (`train_from_ncbi`, MLST download) need the network and are not part of this package."""

from pathlib import Path
from tempfile import TemporaryDirectory

from .definitions import get_xspect_model_path
from .file_io import concatenate_metagenome, concatenate_species_fasta_files
from .models.probabilistic_filter_model import ProbabilisticFilterModel
from .models.probabilistic_filter_svm_model import ProbabilisticFilterSVMModel
from .models.probabilistic_single_filter_model import ProbabilisticSingleFilterModel


def train_from_directory(
    display_name: str,
    dir_path: Path,
    meta: bool = False,
    training_accessions: dict[str, list[str]] | None = None,
    svm_accessions: dict[str, list[str]] | None = None,
    svm_step: int = 1,
    translation_dict: dict[str, str] | None = None,
    author: str | None = None,
    author_email: str | None = None,
):
    if not isinstance(display_name, str):
        raise TypeError("display_name must be a string")
    if not isinstance(dir_path, Path):
        raise TypeError("dir must be Path object to a valid directory")
    cobs_training_path = dir_path / "cobs"
    if not cobs_training_path.exists():
        raise ValueError("cobs directory not found")
    cobs_folders = sorted(f for f in cobs_training_path.iterdir() if f.is_dir())
    if len(cobs_folders) == 0:
        raise ValueError("no folders found in cobs directory")
    svm_path = dir_path / "svm"
    if svm_path.exists():
        svm_folders = sorted(f for f in svm_path.iterdir() if f.is_dir())
        if len(svm_folders) != len(cobs_folders):
            raise ValueError("number of svm folders does not match number of cobs folders")
        for cobs_folder, svm_folder in zip(cobs_folders, svm_folders):
            if cobs_folder.name != svm_folder.name:
                raise ValueError("cobs folder and svm folder names do not match")
    else:
        print("SVM directory not found. Model will be trained without SVM.")

    with TemporaryDirectory() as tmp_dir:
        tmp_dir = Path(tmp_dir)
        species_dir = tmp_dir / "species"
        species_dir.mkdir(parents=True, exist_ok=True)
        concatenate_species_fasta_files(cobs_folders, species_dir)
        if svm_path.exists():
            species_model = ProbabilisticFilterSVMModel(
                k=21, model_display_name=display_name, author=author, author_email=author_email, model_type="Species",
                base_path=get_xspect_model_path(), kernel="rbf", c=1.0,
            )
            species_model.fit(species_dir, svm_path, display_names=translation_dict, svm_step=svm_step,
                              training_accessions=training_accessions, svm_accessions=svm_accessions)
        else:
            species_model = ProbabilisticFilterModel(
                k=21, model_display_name=display_name, author=author, author_email=author_email, model_type="Species",
                base_path=get_xspect_model_path(),
            )
            species_model.fit(species_dir, display_names=translation_dict, training_accessions=training_accessions)
        species_model.save()

        if meta:
            meta_fasta = tmp_dir / f"{display_name}.fasta"
            concatenate_metagenome(species_dir, meta_fasta)
            genus_model = ProbabilisticSingleFilterModel(
                k=21, model_display_name=display_name, author=author, author_email=author_email, model_type="Genus",
                base_path=get_xspect_model_path(),
            )
            genus_model.fit(meta_fasta, display_name,
                            training_accessions=(sum(training_accessions.values(), []) if training_accessions else None))
            genus_model.save()
