"""Input fan-out and record iteration (file_io.py:47-79,166-234 of the reference) on this package's reader."""

from pathlib import Path
from typing import Callable, Iterator

from . import seqio
from .definitions import fasta_endings, fastq_endings


def get_record_iterator(file_path: Path) -> Iterator:
    """Record iterator for a fasta or fastq file, chosen by extension; same ValueErrors as the reference."""
    if not isinstance(file_path, Path):
        raise ValueError("Path must be a Path object")
    if not file_path.exists():
        raise ValueError("File does not exist")
    if not file_path.is_file():
        raise ValueError("Path must be a file")
    if file_path.suffix[1:] in fasta_endings:
        return seqio.parse(file_path, "fasta")
    if file_path.suffix[1:] in fastq_endings:
        return seqio.parse(file_path, "fastq")
    raise ValueError("Invalid file format, must be a fasta or fastq file")


def filter_sequences(input_file: Path, output_file: Path, included_ids: list[str]) -> None:
    """Write the records of ``input_file`` whose id is in ``included_ids`` as FASTA."""
    if not included_ids:
        print("No IDs provided, no output file will be created.")
        return
    wanted = set(included_ids)
    with open(output_file, "w", encoding="utf-8") as out_f:
        for record in get_record_iterator(input_file):
            if record.id in wanted:
                seqio.write_fasta(record, out_f)


def prepare_input_output_paths(input_path: Path) -> tuple[list[Path], Callable[[int, Path], Path]]:
    """A file -> itself; a directory -> every fasta/fastq file in it, outputs suffixed ``_<n>``."""
    input_is_dir = input_path.is_dir()
    if input_is_dir:
        input_paths = [p for ending in fasta_endings + fastq_endings for p in input_path.glob(f"*.{ending}")]
    elif input_path.is_file():
        input_paths = [input_path]
    else:
        raise ValueError("Invalid input path")

    def get_output_path(idx: int, output_path: Path) -> Path:
        if input_is_dir:
            return output_path.parent / f"{output_path.stem}_{idx + 1}{output_path.suffix}"
        return output_path

    return input_paths, get_output_path
