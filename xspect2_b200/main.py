"""Command line interface: the reference's command tree for the prediction path
(main.py:17-795): ``all``, ``models list``, ``classify {genus,species,mlst}``, ``filter {genus,species}`` with the
same options.  The training workflows, model download and the web UI are not offered (the model classes' ``fit``
methods build indices on the GPU)."""

from pathlib import Path
from uuid import uuid4

import click

from .model_management import get_available_mlst_schemes, get_models


@click.group()
@click.version_option(package_name=None, version=__import__("xspect2_b200").__version__)
def cli():
    """XspecT CLI (B200 scoring path)."""


def _genus_option(kind: str, help_text: str):
    return click.option("-g", "--genus", "model_genus", help=help_text, type=click.Choice(get_models().get(kind, [])), prompt=True)


_input_option = click.option(
    "-i", "--input-path", help="Path to FASTA or FASTQ file for classification.",
    type=click.Path(exists=True, dir_okay=True, file_okay=True), prompt=True, default=Path("."),
)
_step_option = click.option(
    "--sparse-sampling-step", type=int, default=1,
    help="Sparse sampling step (e. g. only every 500th kmer for '--sparse-sampling-step 500').",
)


@cli.command(name="all", help="Run full classification pipeline: genus filtering, species classification, and MLST (if applicable).")
@_genus_option("Species", "Genus of the model to use.")
@_input_option
@click.option("-o", "--output-dir", type=click.Path(dir_okay=True, file_okay=False), default=None,
              help="Directory for output files (default: auto-generated 'xspect_results_<uuid>' directory).")
@click.option("-t", "--threshold", type=click.FloatRange(0, 1), default=0.7, help="Threshold for genus filtering (default: 0.7).")
@_step_option
@click.option("-n", "--display-names", is_flag=True, help="Includes the display names next to taxonomy-IDs.")
@click.option("-v", "--validation", is_flag=True, help="Detects misclassification for small reads or contigs.")
def all_pipeline(model_genus, input_path, output_dir, threshold, sparse_sampling_step, display_names, validation):
    """Genus filter -> species classification of the kept sequences -> MLST when the prediction is 470.
    Stages hand over through files, as in the reference (main.py:108-145)."""
    import json

    from . import classify, filter_sequences
    from .definitions import fasta_endings, fastq_endings

    run_id = uuid4()
    output_dir = Path(f"xspect_results_{run_id}") if output_dir is None else Path(output_dir)
    output_dir.mkdir(exist_ok=True, parents=True)
    input_path = Path(input_path)
    filtered_dir = output_dir / "filtered_sequences"
    filtered_dir.mkdir(exist_ok=True, parents=True)
    genus_filtered_path = filtered_dir / f"genus_filtered_{run_id}.fasta"
    genus_classification_path = output_dir / f"genus_classification_{run_id}.json"
    species_classification_path = output_dir / f"species_classification_{run_id}.json"

    click.echo(f"Step 1/3: Filtering for genus {model_genus}...")
    filter_sequences.filter_genus(model_genus, input_path, genus_filtered_path, threshold, genus_classification_path,
                                  sparse_sampling_step=sparse_sampling_step)
    filtered_files = [p for e in fasta_endings + fastq_endings for p in filtered_dir.glob(f"*.{e}")]
    if not filtered_files:
        click.echo("No sequences passed the genus filter. Pipeline aborted.")
        return

    click.echo(f"Step 2/3: Classifying species for {len(filtered_files)} filtered file(s)...")
    classify.classify_species(model_genus, filtered_dir, species_classification_path, sparse_sampling_step, display_names,
                              validation, None)

    mlst_needed = False
    for species_result_path in output_dir.glob(f"species_classification_{run_id}*.json"):
        with open(species_result_path, "r", encoding="utf-8") as f:
            if json.load(f).get("prediction") == "470":
                mlst_needed = True
                click.echo(f"Species prediction is 470 (abaumannii) in {species_result_path.name}.")
    if mlst_needed:
        click.echo("Step 3/3: Running MLST classification for abaumannii...")
        mlst_schemes = get_available_mlst_schemes()
        if mlst_schemes.get("abaumannii"):
            mlst_output_path = output_dir / f"mlst_classification_{run_id}.json"
            classify.classify_mlst(filtered_dir, "abaumannii", mlst_schemes["abaumannii"][0], mlst_output_path, False)
            click.echo(f"MLST classification completed: {mlst_output_path.name}")
        else:
            click.echo("Warning: No MLST schemes available for abaumannii. Skipping MLST classification.")
    else:
        click.echo("Step 3/3: Not running MLST classification (organism is not Acinetobacter baumannii).")
    click.echo("\nPipeline completed successfully!")
    click.echo(f"Results saved in: {output_dir}")


@cli.group()
def models():
    """Model management commands."""


@models.command(name="list", help="List all models in the model directory.")
def list_models():
    available_models = get_models()
    if not available_models:
        click.echo("No models found.")
        return
    click.echo("Models found:")
    click.echo("--------------")
    for model_type, names in available_models.items():
        if not names:
            continue
        click.echo(f"  {model_type}:")
        for name in names:
            click.echo(f"    - {name}")


@cli.group(name="classify", help="Classify sequences using XspecT models.")
def classify_seqs():
    """Classification commands."""


@classify_seqs.command(name="genus", help="Classify samples using a genus model.")
@_genus_option("Genus", "Genus of the model to classify.")
@_input_option
@click.option("-o", "--output-path", help="Path to the output file.", type=click.Path(dir_okay=False, file_okay=True),
              default=Path(".") / f"result_{uuid4()}.json")
@_step_option
def classify_genus(model_genus, input_path, output_path, sparse_sampling_step):
    click.echo("Classifying...")
    from . import classify

    classify.classify_genus(model_genus, Path(input_path), Path(output_path), sparse_sampling_step)


@classify_seqs.command(name="species", help="Classify samples using a species model.")
@_genus_option("Species", "Genus of the model to classify.")
@_input_option
@click.option("-o", "--output-path", help="Path to the output file.", type=click.Path(dir_okay=False, file_okay=True),
              default=Path(".") / f"result_{uuid4()}.json")
@_step_option
@click.option("-n", "--display-names", is_flag=True, help="Includes the display names next to taxonomy-IDs.")
@click.option("-v", "--validation", is_flag=True, help="Detects misclassification for small reads or contigs.")
@click.option("--exclude-species", type=str, default=None, help="Comma-separated species IDs to exclude from classification.")
def classify_species(model_genus, input_path, output_path, sparse_sampling_step, display_names, validation, exclude_species):
    click.echo("Classifying...")
    from . import classify

    exclude_ids = [s.strip() for s in exclude_species.split(",")] if exclude_species else None
    classify.classify_species(model_genus, Path(input_path), Path(output_path), sparse_sampling_step, display_names,
                              validation, exclude_ids)


@classify_seqs.command(name="mlst", help="Classify samples using a MLST model.")
@_input_option
@click.option("--organism", help="Underlying organism of the MLST model.", type=click.Choice(list(get_available_mlst_schemes().keys())), prompt=True)
@click.option("--mlst-scheme", type=str, default=None, help="MLST scheme to use.")
@click.option("-o", "--output-path", help="Path to the output file.", type=click.Path(dir_okay=False, file_okay=True),
              default=Path(".") / f"MLST_result_{uuid4()}.json")
@click.option("-l", "--limit", is_flag=True, help="Limit the output to 5 results for each locus.")
def classify_mlst(input_path, organism, mlst_scheme, output_path, limit):
    mlst_schemes = get_available_mlst_schemes()
    if not mlst_scheme:
        mlst_scheme = click.prompt("Please enter the MLST scheme you want to use:", type=click.Choice(mlst_schemes[organism]))
    elif mlst_scheme not in mlst_schemes.get(organism, []):
        raise click.BadParameter(
            f"Scheme '{mlst_scheme}' not found for organism '{organism}'. "
            f"Available schemes: {', '.join(mlst_schemes.get(organism, []))}"
        )
    click.echo("Classifying...")
    from . import classify

    classify.classify_mlst(Path(input_path), organism, mlst_scheme, Path(output_path), limit)


@cli.group(name="filter", help="Filter sequences using XspecT models.")
def filter_seqs():
    """Filter commands."""


@filter_seqs.command(name="genus", help="Filter sequences using a genus model.")
@_genus_option("Species", "Genus of the model to use for filtering.")
@_input_option
@click.option("-o", "--output-path", help="Path to the output file.", type=click.Path(dir_okay=False, file_okay=True),
              default=Path(".") / f"genus_filtered_{uuid4()}.fasta")
@click.option("--classification-output-path", type=click.Path(dir_okay=False, file_okay=True), default=None,
              help="Optional path to save the classification results.")
@click.option("-t", "--threshold", type=click.FloatRange(0, 1), default=0.7, help="Threshold for filtering (default: 0.7).")
@_step_option
def filter_genus(model_genus, input_path, output_path, classification_output_path, threshold, sparse_sampling_step):
    click.echo("Filtering...")
    from . import filter_sequences

    filter_sequences.filter_genus(
        model_genus, Path(input_path), Path(output_path), threshold,
        Path(classification_output_path) if classification_output_path else None,
        sparse_sampling_step=sparse_sampling_step,
    )


@filter_seqs.command(name="species", help="Filter sequences using a species model.")
@_genus_option("Species", "Genus of the model to use for filtering.")
@click.option("-s", "--species", "model_species", default=None, help="Species to filter for.")
@_input_option
@click.option("-o", "--output-path", help="Path to the output file.", type=click.Path(dir_okay=False, file_okay=True),
              default=Path(".") / f"species_filtered_{uuid4()}.fasta")
@click.option("--classification-output-path", type=click.Path(dir_okay=False, file_okay=True), default=None,
              help="Optional path to save the classification results.")
@click.option("-t", "--threshold", type=float, default=0.7,
              help="Threshold for filtering (default: 0.7). Use -1 to filter for the highest scoring species.")
@_step_option
def filter_species(model_genus, model_species, input_path, output_path, threshold, classification_output_path, sparse_sampling_step):
    if threshold != -1 and (threshold < 0 or threshold > 1):
        raise click.BadParameter("Threshold must be between 0 and 1, or -1 for filtering by the highest scoring species.")
    from .model_management import get_model_metadata

    available = get_model_metadata(f"{model_genus}-species")["display_names"]
    available = {sid: name.replace(f"{model_genus} ", "") for sid, name in available.items()}
    if not model_species:
        model_species = click.prompt(f"Please enter the species name: {model_genus}",
                                     type=click.Choice(sorted(available.values()), case_sensitive=False))
    if model_species not in available.values():
        raise click.BadParameter(f"Species '{model_species}' not found in the {model_genus} species model.")
    species_id = [sid for sid, name in available.items() if name.lower() == model_species.lower()][0]
    click.echo("Filtering...")
    from . import filter_sequences

    filter_sequences.filter_species(
        model_genus, species_id, Path(input_path), Path(output_path), threshold,
        Path(classification_output_path) if classification_output_path else None,
        sparse_sampling_step=sparse_sampling_step,
    )


if __name__ == "__main__":
    cli()
