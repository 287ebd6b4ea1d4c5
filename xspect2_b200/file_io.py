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
    import ctypes as C

    import numpy as np

    from . import _abi
    from .seqio import SequenceBatch

    get_record_iterator(input_file)          # same path / format errors as the reference
    wanted = set(included_ids)
    ids = SequenceBatch.from_file(input_file).ids
    keep = np.fromiter((rid in wanted for rid in ids), dtype=np.uint8, count=len(ids))
    L = _abi.lib()
    h = C.c_void_p()
    fmt = 2 if input_file.suffix[1:] in fastq_endings else 1
    _abi.check(L.xs_fastx_open(str(input_file).encode(), fmt, C.byref(h)))
    try:
        _abi.check(L.xs_fastx_filter_fasta(h, keep.ctypes.data, str(output_file).encode()))
    finally:
        L.xs_fastx_close(h)


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


def concatenate_species_fasta_files(input_folders: list[Path], output_directory: Path) -> None:
    """One ``<species>.fasta`` per species folder, the folder's fasta files appended one after another
    (file_io.py:82-109 of the reference; files in sorted order here)."""
    for species_folder in input_folders:
        fasta_files = sorted(f for ending in fasta_endings for f in species_folder.glob(f"*.{ending}"))
        if len(fasta_files) == 0:
            raise ValueError(f"no fasta files found in {species_folder}")
        with open(output_directory / f"{species_folder.name}.fasta", "wb") as out:
            for fasta_file in fasta_files:
                out.write(fasta_file.read_bytes())


def concatenate_metagenome(fasta_dir: Path, meta_path: Path) -> None:
    """All fasta files of a directory appended into one file (file_io.py:112-132 of the reference)."""
    fasta_files = sorted(f for ending in fasta_endings for f in fasta_dir.glob(f"*.{ending}"))
    with open(meta_path, "wb") as meta_file:
        for fasta_file in fasta_files:
            meta_file.write(fasta_file.read_bytes())
