"""Paths and file-type conventions shared with XspecT (definitions.py:6-110): the same data root
(``~/xspect-data`` or ``./xspect-data``) and sub-directories, so models installed for the reference are
found unchanged."""

from os import getcwd
from pathlib import Path

fasta_endings = ["fasta", "fna", "fa", "ffn", "frn"]
fastq_endings = ["fastq", "fq"]


def get_xspect_root_path() -> Path:
    """``~/xspect-data`` if present, else ``./xspect-data`` if present, else create the former."""
    home_based = Path.home() / "xspect-data"
    if home_based.exists():
        return home_based
    cwd_based = Path(getcwd()) / "xspect-data"
    if cwd_based.exists():
        return cwd_based
    home_based.mkdir(exist_ok=True, parents=True)
    return home_based


def _sub(name: str) -> Path:
    p = get_xspect_root_path() / name
    p.mkdir(exist_ok=True, parents=True)
    return p


def get_xspect_model_path() -> Path:
    return _sub("models")


def get_xspect_upload_path() -> Path:
    return _sub("uploads")


def get_xspect_runs_path() -> Path:
    return _sub("runs")


def get_xspect_mlst_path() -> Path:
    return _sub("mlst")


def get_xspect_misclassification_path() -> Path:
    return _sub("misclassification")
